/* TEST INFRASTRUCTURE ONLY: reference baseml (unmodified, textually included from the reference tree at
   compile time; nothing is copied into this repo) with lfun/lfundG interposed as in SURVEY.md 8(b). */
double lfun(double x[], int np);
double lfundG(double x[], int np);
#define lfun(a, b)   lfun_ref(a, b)
#define lfundG(a, b) lfundG_ref(a, b)
#include "baseml.c"
#undef lfun
#undef lfundG
#include "ref_hook.inc"
