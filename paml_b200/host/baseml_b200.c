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
 *                   eigenQREV_b200 (device Jacobi for codon-sized matrices), which can still call the original;
 *   minbranches, updateconP, lfunt, lfuntdd, lfunt_SiteClass, lfuntdd_SiteClass   (`method = 1`) by the same first-argument
 *                   rule: the reference's minbranches keeps its control flow (as minbranches_ref) while the data-touching
 *                   calls inside it go to conditional vectors resident in HBM. */
double lfun(double x[], int np);
double lfundG(double x[], int np);
int fx_r_gpu(double x[], int np);
int PatternWeight_b200(void);
int eigenQREV_b200(double Q[], double pi[], int n, double Root[], double U[], double V[], double spacesqrtpi[]);
int updateconP_disp(double x[], int inode);
double minbranches_b200(double xcom[], int np);
double minbranches(double xcom[], int np);     /* the bare name (passed to ming2 in minB2, src/treesub.c:7968) is ours too */
int lfunt_disp(double t, int a, int b, double xcom[], double *l, double space[]);
int lfuntdd_disp(double t, int a, int b, double xcom[], double *l, double *dl, double *ddl, double space[]);
int lfunt_SiteClass_disp(double t, int a, int b, double xcom[], double *l, double space[]);
int lfuntdd_SiteClass_disp(double t, int a, int b, double xcom[], double *l, double *dl, double *ddl, double space[]);
int ConditionalPNode_disp(int inode, int igene, double x[]);
void SetSeed_b200(int seed, int PrintSeed);
#define lfun(a, b)   lfun_ref(a, b)
#define lfundG(a, b) lfundG_ref(a, b)
#define fx_r(a, b)   PBH_FXR_##a, b)
#define PBH_FXR_x      fx_r_gpu(x
#define PBH_FXR_double fx_r_ref(double
#define PatternWeight(a) PBH_PW_##a
#define eigenQREV(a, b, c, d, e, f, g) eigenQREV_b200(a, b, c, d, e, f, g)
#define SetSeed(a, b) SetSeed_b200(a, b)        /* tests only: PB200_TEST_SEED gives this program the reference run's starting values */
/* method = 1: minbranches(xcom, ..), updateconP(x, ..), lfunt*(t, ..) calls -> ours; the definitions -> *_ref */
#define minbranches(a, b) PBH_MB_##a, b)
#define PBH_MB_xcom   minbranches_b200(xcom
#define PBH_MB_double minbranches_ref(double
#define updateconP(a, b) PBH_UC_##a, b)
#define PBH_UC_x      updateconP_disp(x
#define PBH_UC_double updateconP_ref(double
#define lfunt(a, b, c, d, e, f) PBH_LT_##a, b, c, d, e, f)
#define PBH_LT_t      lfunt_disp(t
#define PBH_LT_double lfunt_ref(double
#define lfuntdd(a, b, c, d, e, f, g, h) PBH_LD_##a, b, c, d, e, f, g, h)
#define PBH_LD_t      lfuntdd_disp(t
#define PBH_LD_double lfuntdd_ref(double
#define lfunt_SiteClass(a, b, c, d, e, f) PBH_LTS_##a, b, c, d, e, f)
#define PBH_LTS_t      lfunt_SiteClass_disp(t
#define PBH_LTS_double lfunt_SiteClass_ref(double
#define lfuntdd_SiteClass(a, b, c, d, e, f, g, h) PBH_LDS_##a, b, c, d, e, f, g, h)
#define PBH_LDS_t      lfuntdd_SiteClass_disp(t
#define PBH_LDS_double lfuntdd_SiteClass_ref(double
/* ConditionalPNode: whole-tree calls `ConditionalPNode(tree.root, ..)` -> ConditionalPNode_disp; the definition, its own recursion
   `ConditionalPNode(nodes[..].sons[i], ..)` and updateconP's `ConditionalPNode(inode, ..)` -> ConditionalPNode_ref */
#define ConditionalPNode(a, b, c) PBH_CPN_##a, b, c)
#define PBH_CPN_int   ConditionalPNode_ref(int
#define PBH_CPN_nodes ConditionalPNode_ref(nodes
#define PBH_CPN_inode ConditionalPNode_ref(inode
#define PBH_CPN_tree  ConditionalPNode_disp(tree
#define PBH_PW_void PatternWeight_ref(void)
#define PBH_PW_     PatternWeight_b200()
#include "baseml.c"
#undef lfun
#undef lfundG
#undef fx_r
#undef PatternWeight
#undef eigenQREV
#undef SetSeed
#undef minbranches
#undef updateconP
#undef lfunt
#undef lfuntdd
#undef lfunt_SiteClass
#undef lfuntdd_SiteClass
#undef ConditionalPNode
#include "plfun_b200.inc"
