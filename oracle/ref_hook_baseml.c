/* TEST INFRASTRUCTURE ONLY: reference baseml (unmodified, textually included from the reference tree at
   compile time; nothing is copied into this repo) with lfun/lfundG interposed as in SURVEY.md 8(b). */
double lfun(double x[], int np);
double lfundG(double x[], int np);
#define lfun(a, b)   lfun_ref(a, b)
#define lfundG(a, b) lfundG_ref(a, b)
/* lfuntdd(t, ..) / lfuntdd_SiteClass(t, ..) calls (from minbranches) go through the hook, which can dump a BRANCH fixture
   (PBX_BRANCH_DUMP); the definitions keep their bodies as *_ref */
int lfuntdd_hook(double t, int a, int b, double xcom[], double *l, double *dl, double *ddl, double space[]);
int lfuntdd_SiteClass_hook(double t, int a, int b, double xcom[], double *l, double *dl, double *ddl, double space[]);
#define lfuntdd(a, b, c, d, e, f, g, h) PBX_LD_##a, b, c, d, e, f, g, h)
#define PBX_LD_t      lfuntdd_hook(t
#define PBX_LD_double lfuntdd_ref(double
#define lfuntdd_SiteClass(a, b, c, d, e, f, g, h) PBX_LDS_##a, b, c, d, e, f, g, h)
#define PBX_LDS_t      lfuntdd_SiteClass_hook(t
#define PBX_LDS_double lfuntdd_SiteClass_ref(double
#include "baseml.c"
#undef lfun
#undef lfundG
#undef lfuntdd
#undef lfuntdd_SiteClass
#include "ref_hook.inc"
