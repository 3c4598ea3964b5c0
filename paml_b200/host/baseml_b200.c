/* baseml_b200.c -- baseml with its likelihood function on the GPU.
 * The reference program is compiled from where it lies (-I<reference>/src), unmodified and textually included;
 * only the hot-path entry points are renamed away and re-defined by plfun_b200.inc:
 *   lfun / lfundG   function-like macros rename the reference DEFINITIONS (and direct calls, of which there are none
 *                   outside lfundG's own file) while the bare names in `com.plfun = lfun` / `== lfun` bind to ours;
 *   fx_r            is only ever called directly, always as fx_r(x, ..): the macro below pastes on the first
 *                   argument, so the reference's definition/prototype `fx_r(double x[], int np)` becomes fx_r_ref and
 *                   every call `fx_r(x, ..)` becomes fx_r_gpu (any other spelling fails to compile, loudly);
 *   PatternWeight   (the step upstream of the path) the same way: `PatternWeight(void)` -> PatternWeight_ref, the call
 *                   `PatternWeight()` -> PatternWeight_b200;
 *   eigenQREV       lives in tools.c (a separate object): only the CALLS made from this translation unit are renamed, to
 *                   eigenQREV_b200 (device Jacobi for codon-sized matrices), which can still call the original. */
double lfun(double x[], int np);
double lfundG(double x[], int np);
int fx_r_gpu(double x[], int np);
int PatternWeight_b200(void);
int eigenQREV_b200(double Q[], double pi[], int n, double Root[], double U[], double V[], double spacesqrtpi[]);
#define lfun(a, b)   lfun_ref(a, b)
#define lfundG(a, b) lfundG_ref(a, b)
#define fx_r(a, b)   PBH_FXR_##a, b)
#define PBH_FXR_x      fx_r_gpu(x
#define PBH_FXR_double fx_r_ref(double
#define PatternWeight(a) PBH_PW_##a
#define eigenQREV(a, b, c, d, e, f, g) eigenQREV_b200(a, b, c, d, e, f, g)
#define PBH_PW_void PatternWeight_ref(void)
#define PBH_PW_     PatternWeight_b200()
#include "baseml.c"
#undef lfun
#undef lfundG
#undef fx_r
#undef PatternWeight
#undef eigenQREV
#include "plfun_b200.inc"
