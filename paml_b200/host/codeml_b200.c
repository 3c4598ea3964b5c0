/* codeml_b200.c -- codeml with its likelihood function on the GPU.
 * The reference program is compiled from where it lies (-I<reference>/src), unmodified and textually included;
 * only the two hot-path entry points are renamed away and re-defined by plfun_b200.inc. */
double lfun(double x[], int np);
double lfundG(double x[], int np);
#define lfun(a, b)   lfun_ref(a, b)
#define lfundG(a, b) lfundG_ref(a, b)
#include "codeml.c"
#undef lfun
#undef lfundG
#include "plfun_b200.inc"
