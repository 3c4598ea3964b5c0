/* mcmctree_b200.c -- mcmctree with its exact sequence likelihood (usedata = 1) on the GPU.
 *
 * The reference program is compiled from where it lies, unmodified and textually included; the DEFINITIONS of lfun / lfundG in
 * the treesub.c it includes (src/treesub.c:7764, 7608) are renamed away by function-like macros, while the bare names in
 * `com.plfun = lfun` / `com.plfun = lfundG` (src/mcmctree.c:272-275) bind to the ones defined below.  Everything else -- the
 * control-file parser, the priors on times and rates, every MCMC proposal, the accept / reject bookkeeping -- is the reference's.
 *
 * What the reference does inside `lnL = -com.plfun(NULL, -1)` (lnpD_locus, src/mcmctree.c:1130-1171): its own ConditionalPNode
 * (src/mcmctree.c:1036-1087) walks the gene tree of the CURRENT locus (UseLocus, src/mcmctree.c:887-950, has pointed nodes / com.z /
 * com.fpatt / com.pi / com.kappa / com.alpha at it), recomputing only the nodes whose com.oldconP flag is clear, with P(t) from
 * GetPMatBranch2 (src/mcmctree.c:1010-1033: PMatK80 for JC69 / K80, PMatTN93 for F81 / F84 / HKY85) and, under discrete-gamma
 * rates, once per rate class with t * rK (fx_r).  Here every call is ONE evaluation of the whole locus through the C ABI:
 * all branches' P(t), all patterns, all classes.  The saved / alternative conP copies the reference keeps for rejected
 * proposals are not needed (the value does not depend on them); they are left alone.
 *   - one context per locus (data and gene-tree topology are uploaded once; the topology never changes, only nodes[].branch);
 *   - JC69 / K80: PB200_KIND_K80 with {kappa};  F81 / F84 / HKY85: the spectral form eigenTN93 leaves in Root / Cijk at every
 *     UseLocus (src/mcmctree.c:932), PB200_KIND_CIJK -- P(t) agrees with PMatTN93 to rounding;
 *   - the reference's bookkeeping around its own conditional vectors is switched off where it costs: UseLocus(locus, 1, ..) and
 *     AcceptRejectLocus copy ALL conditional vectors of the locus between the two halves of com.conPin[] at every proposal
 *     (src/mcmctree.c:911-912, 967-968: sG * 4 * npatt * (ns - 1) doubles -- 119 MB for 40 species x 19 000 patterns x 5 classes,
 *     more than the whole GPU evaluation costs).  Nothing reads those vectors here, so the textually included program's memcpy
 *     is a macro that skips copies INTO that buffer and is the C library's memcpy for everything else;
 *   - no node scaling (the reference has it commented out, src/mcmctree.c:937-944);
 *   - errors abort like zerror; there is no CPU fallback.
 * Environment: PB200_DEVICE = device ordinal (default 0); PB200_VERBOSE = 1 prints one line per upload. */
#include <stdio.h>
#include <string.h>
static void *pbm_memcpy(void *dst, const void *src, size_t n);
double lfun(double x[], int np);
double lfundG(double x[], int np);
#define lfun(a, b)   lfun_ref(a, b)
#define lfundG(a, b) lfundG_ref(a, b)
#define memcpy(d, s, n) pbm_memcpy(d, s, n)
#include "mcmctree.c"
#undef lfun
#undef lfundG
#undef memcpy
#include "paml_b200.h"

typedef struct {
   pb200_ctx *ctx;
   const double *fpatt;          /* identity of the locus: data.fpatt[locus] */
   int ns, npatt, nnode, root, cleandata;
   int *branch_set;
   double *branch_t;
} pbm_locus_t;
static pbm_locus_t pbm_loci[NGENE];
static long pbm_nupload = 0;

static void *pbm_memcpy(void *dst, const void *src, size_t n)
{
   /* copies into the reference's own conditional-vector store (usedata = 1: com.conPin[0] .. + 2 com.sconP) are dead here */
   if (mcmc.usedata == 1 && com.conPin[0] && com.sconP
       && (char*)dst >= (char*)com.conPin[0] && (char*)dst < (char*)com.conPin[0] + 2 * com.sconP)
      return dst;
   return memcpy(dst, src, n);
}

static void pbm_die(const char *what)
{
   fprintf(stderr, "\npaml_b200 (mcmctree): %s: %s\n", what, pb200_last_error());
   exit(-1);
}

static pbm_locus_t *pbm_locus(void)
{
   int locus, i, j, ms = 0;
   long long h;
   pbm_locus_t *L;
   const char *e;

   for (locus = 0; locus < data.ngene; locus++)
      if (data.fpatt[locus] == com.fpatt) break;
   if (locus == data.ngene) { fprintf(stderr, "\npaml_b200 (mcmctree): com.fpatt is not a locus of data.fpatt[]\n"); exit(-1); }
   L = &pbm_loci[locus];
   if (L->ctx && L->ns == com.ns && L->npatt == com.npatt && L->nnode == tree.nnode && L->root == tree.root && L->cleandata == com.cleandata)
      return L;

   /* first use of this locus: upload its patterns and its gene tree */
   if (com.ncode != 4) { fprintf(stderr, "\npaml_b200 (mcmctree): usedata = 1 is for nucleotides (ncode = %d)\n", com.ncode); exit(-1); }
   if (com.NnodeScale) { fprintf(stderr, "\npaml_b200 (mcmctree): node scaling is not part of mcmctree's likelihood\n"); exit(-1); }
   if (!L->ctx) {
      e = getenv("PB200_DEVICE");
      if (pb200_create(&L->ctx, e ? atoi(e) : 0)) pbm_die("cannot create GPU context (this build has no CPU fallback)");
   }
   {
      unsigned char *z = (unsigned char*)malloc((size_t)com.ns * com.npatt), nchara[256], *charamap = (unsigned char*)malloc(256 * 64);
      int *nson = (int*)malloc(tree.nnode * sizeof(int)), *sons;
      if (!z || !charamap || !nson) zerror("oom");
      for (i = 0; i < com.ns; i++)
         for (h = 0; h < com.npatt; h++) z[(size_t)i * com.npatt + h] = (unsigned char)com.z[i][h];
      for (i = 0; i < 256; i++) {
         nchara[i] = (unsigned char)nChara[i];
         for (j = 0; j < 64; j++) charamap[i * 64 + j] = (unsigned char)CharaMap[i][j];
      }
      if (pb200_set_data(L->ctx, com.ncode, com.ns, com.npatt, z, com.fpatt, com.cleandata, nchara, charamap)) pbm_die("pb200_set_data");
      for (i = 0; i < tree.nnode; i++) if (nodes[i].nson > ms) ms = nodes[i].nson;
      sons = (int*)malloc((size_t)tree.nnode * ms * sizeof(int));
      if (!sons) zerror("oom");
      for (i = 0; i < tree.nnode; i++) {
         nson[i] = nodes[i].nson;
         for (j = 0; j < ms; j++) sons[i * ms + j] = (j < nodes[i].nson ? nodes[i].sons[j] : -1);
      }
      if (pb200_set_tree(L->ctx, tree.nnode, tree.root, nson, sons, ms, NULL)) pbm_die("pb200_set_tree");
      free(z); free(charamap); free(nson); free(sons);
   }
   L->fpatt = com.fpatt; L->ns = com.ns; L->npatt = com.npatt; L->nnode = tree.nnode; L->root = tree.root; L->cleandata = com.cleandata;
   free(L->branch_set); free(L->branch_t);
   L->branch_set = (int*)calloc((size_t)NCATG * tree.nnode, sizeof(int));
   L->branch_t = (double*)malloc((size_t)NCATG * tree.nnode * sizeof(double));
   if (!L->branch_set || !L->branch_t) zerror("oom");
   pbm_nupload++;
   e = getenv("PB200_VERBOSE");
   if (e && atoi(e)) fprintf(stderr, "paml_b200 (mcmctree): locus %d uploaded (%d taxa, %d patterns)\n", locus + 1, com.ns, com.npatt);
   return L;
}

/* One evaluation of the current locus: lfun (one rate for all sites) or lfundG (discrete gamma, classes rK / freqK). */
static double pbm_eval(int is_dG)
{
   pbm_locus_t *L = pbm_locus();
   int K = (is_dG ? com.ncatG : 1), ir, i, kind, nr;
   long long stride;
   double set[8 + 16 * 8], one = 1, v = 0;
   const double *freqK = (is_dG ? com.freqK : &one);

   NFunCall++;
   for (ir = 0; ir < K; ir++)
      for (i = 0; i < tree.nnode; i++) {
         double t = (i == tree.root ? 0 : nodes[i].branch * (is_dG ? com.rK[ir] : 1));      /* src/mcmctree.c:1051 */
         if (t < 0) {
            printf("\nt =%12.6f ratesite = %.6f", nodes[i].branch, is_dG ? com.rK[ir] : 1.0);
            zerror("blength < 0");
         }
         L->branch_t[ir * tree.nnode + i] = t;
      }
   if (com.model <= K80) {            /* PMatK80(PMat, t, com.kappa), src/mcmctree.c:1019-1020 (JC69: kappa = 1) */
      kind = PB200_KIND_K80; nr = 0; stride = 1;
      set[0] = com.kappa;
   }
   else if (com.model <= TN93) {      /* PMatTN93 with the rates of src/mcmctree.c:1021-1030 = the spectral form of eigenTN93 */
      kind = PB200_KIND_CIJK; nr = nR; stride = nr + 16 * nr;
      for (i = 0; i < nr; i++) set[i] = Root[i];
      for (i = 0; i < 16 * nr; i++) set[nr + i] = Cijk[i];
   }
   else { fprintf(stderr, "\npaml_b200 (mcmctree): model %d is not one of JC69 .. HKY85\n", com.model); exit(-1); }
   if (pb200_eval(L->ctx, is_dG ? PB200_SEM_LFUNDG : PB200_SEM_LFUN, K, freqK, kind, nr, 1, stride, set, L->branch_set, L->branch_t,
                  com.pi, &v))
      pbm_die("pb200_eval");
   return v;
}

double lfun(double x[], int np)   { (void)x; (void)np; return pbm_eval(0); }
double lfundG(double x[], int np) { (void)x; (void)np; return pbm_eval(1); }
