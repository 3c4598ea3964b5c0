/* paml_oracle.c -- TEST INFRASTRUCTURE ONLY (see paml_oracle.h).
 *
 * Scalar CPU restatement of the reference algorithm for the hot path.  Every function names the
 * reference lines it follows (paths relative to the reference tree's src/).  It is written to be
 * read next to the reference, keeps its loop/summation ORDER (so results agree to the last few ulps),
 * but shares no code with it: data come in as flat arrays (the fixture / C-ABI layout), not globals.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "paml_oracle.h"

/* tools.c:516-546 PMatUVRoot: P = I + sum_k U[:,k] expm1(t*Root[k]) V[k,:]; P<0 -> 0; t<1e-100 -> I */
void orc_pmat_uvroot(double *P, double t, int n, const double *U, const double *V, const double *Root)
{
   int i, j, k;
   if (t < 1e-100) {
      for (i = 0; i < n * n; i++) P[i] = 0;
      for (i = 0; i < n; i++) P[i * n + i] = 1;
      return;
   }
   for (i = 0; i < n * n; i++) P[i] = 0;
   for (k = 0; k < n; k++) {
      double e = expm1(t * Root[k]);
      for (i = 0; i < n; i++) {
         double ue = U[i * n + k] * e;
         for (j = 0; j < n; j++) P[i * n + j] += ue * V[k * n + j];
      }
   }
   for (i = 0; i < n; i++) P[i * n + i] += 1;
   for (i = 0; i < n * n; i++) if (P[i] < 0) P[i] = 0;
}

/* baseml.c:1572-1589 PMatCijk: P_ij = delta_ij + sum_k Cijk[i][j][k] expm1(t Root[k]) with exptm1[0] = 0 */
void orc_pmat_cijk(double *P, double t, int n, int nr, const double *Cijk, const double *Root)
{
   int i, j, k;
   double e[8] = {0};
   for (k = 1; k < nr; k++) e[k] = expm1(t * Root[k]);
   for (i = 0; i < n; i++) {
      for (j = 0; j < n; j++) {
         double s = 0;
         for (k = 0; k < nr; k++) s += Cijk[i * n * nr + j * nr + k] * e[k];
         P[i * n + j] = s;
      }
      P[i * n + i] += 1;
   }
}

/* tools.c:578-606 PMatK80 (JC69 when kappa == 1); state order T C A G */
void orc_pmat_k80(double *P, double t, double kappa)
{
   int i, j;
   double e1 = expm1(-4 * t / (kappa + 2)), e2;
   if (fabs(kappa - 1) < 1e-20) {
      for (i = 0; i < 4; i++) for (j = 0; j < 4; j++) P[i * 4 + j] = (i == j ? 1. + 3 / 4.0 * e1 : -e1 / 4);
      return;
   }
   e2 = expm1(-2 * t * (kappa + 1) / (kappa + 2));
   for (i = 0; i < 4; i++)
      for (j = 0; j < 4; j++) {
         if (i == j) P[i * 4 + j] = 1 + (e1 + 2 * e2) / 4;
         else if ((i ^ j) == 1) P[i * 4 + j] = (e1 - 2 * e2) / 4;   /* T<->C, A<->G transitions */
         else P[i * 4 + j] = -e1 / 4;
      }
}

/* codeml.c:3585-3594 PMatJC69like */
void orc_pmat_jc69like(double *P, double t, int n)
{
   int i;
   double pii = 1. / n + (1. - 1. / n) * exp(-n / (n - 1.) * t);
   double pij = (1. - pii) / (n - 1.);
   for (i = 0; i < n * n; i++) P[i] = pij;
   for (i = 0; i < n; i++) P[i * n + i] = pii;
}

/* treesub.c:7523-7527 (UNREST branch of baseml's GetPMatBranch): P = matexp(Q t) with 7 Taylor terms and 5
   squarings, tools.c:4879-4919: B = Qt/2^5; S = I + B + B^2/2! + ... + B^7/7!; P = S^(2^5).
   No clamping of negative entries and no small-t shortcut on this path. */
void orc_pmat_qmatexp(double *P, double t, int n, const double *Q)
{
   enum { NTAYLOR = 7, NSQUARE = 5 };
   double B[64 * 64], S[64 * 64], pw[64 * 64], nx[64 * 64], factor = 1, s;
   int i, j, k, term, sq, nn = n * n;
   for (i = 0; i < nn; i++) { B[i] = Q[i] * t * (1.0 / (1 << NSQUARE)); S[i] = B[i]; pw[i] = B[i]; }
   for (i = 0; i < n; i++) S[i * n + i] += 1;
   for (term = 2; term <= NTAYLOR; term++) {
      for (i = 0; i < n; i++) for (j = 0; j < n; j++) {
         for (k = 0, s = 0; k < n; k++) s += pw[i * n + k] * B[k * n + j];
         nx[i * n + j] = s;
      }
      factor /= term;
      for (i = 0; i < nn; i++) { pw[i] = nx[i]; S[i] += nx[i] * factor; }
   }
   for (sq = 0; sq < NSQUARE; sq++) {
      for (i = 0; i < n; i++) for (j = 0; j < n; j++) {
         for (k = 0, s = 0; k < n; k++) s += S[i * n + k] * S[k * n + j];
         nx[i * n + j] = s;
      }
      for (i = 0; i < nn; i++) S[i] = nx[i];
   }
   for (i = 0; i < nn; i++) P[i] = S[i];
}

void orc_pmat(double *P, double t, int n, int kind, int nr, const double *set)
{
   switch (kind) {
   case ORC_KIND_QMATEXP:  orc_pmat_qmatexp(P, t, n, set); break;
   case ORC_KIND_UVROOT:   orc_pmat_uvroot(P, t, n, set + n, set + n + (size_t)n * n, set); break;
   case ORC_KIND_CIJK:     orc_pmat_cijk(P, t, n, nr, set + nr, set); break;
   case ORC_KIND_K80:      orc_pmat_k80(P, t, set[0]); break;
   case ORC_KIND_JC69LIKE: orc_pmat_jc69like(P, t, n); break;
   }
}

/* treesub.c:7178-7198 SetNodeScale: post-order count of tips below; flag when > every, reset to 1 */
static int set_scale_rec(int inode, int root, const int *nson, const int *sons, int maxsons, int every,
                         unsigned char *flags, int *count)
{
   int i, d = 0;
   for (i = 0; i < nson[inode]; i++) {
      int ison = sons[inode * maxsons + i];
      d += (nson[ison] ? set_scale_rec(ison, root, nson, sons, maxsons, every, flags, count) : 1);
   }
   if (inode != root && d > every) { flags[inode] = 1; d = 1; (*count)++; }
   return d;
}
int orc_set_node_scale(int nnode, int root, const int *nson, const int *sons, int maxsons, int every,
                       unsigned char *flags)
{
   int count = 0;
   memset(flags, 0, nnode);
   set_scale_rec(root, root, nson, sons, maxsons, every, flags, &count);
   return count;
}

typedef struct {
   int n, ns, nnode, root, maxsons, cleandata, kind, nr;
   long long N, set_stride;
   const int *nson, *sons;
   const unsigned char *scale_flags, *z, *nchara, *charamap;
   const double *set_data;
   const int *bset;        /* this class: [nnode] */
   const double *bt;       /* this class: [nnode] */
   double *conP;           /* [nnode-ns][N][n] */
   double *scaleF;         /* [nscale][N] */
   double *PMat;
} orc_ctx;

/* treesub.c:7200-7230 NodeScale */
static void node_scale(orc_ctx *c, int inode)
{
   int n = c->n, j, k = 0;
   long long h;
   double *cp = c->conP + (size_t)(inode - c->ns) * c->N * n;
   for (j = 0; j < c->nnode; j++) { if (j == inode) break; else if (c->scale_flags[j]) k++; }
   for (h = 0; h < c->N; h++) {
      double t = 0;
      for (j = 0; j < n; j++) if (cp[h * n + j] > t) t = cp[h * n + j];
      if (t < 1e-300) {
         for (j = 0; j < n; j++) cp[h * n + j] = 1;
         c->scaleF[(size_t)k * c->N + h] = -800;
      }
      else {
         for (j = 0; j < n; j++) cp[h * n + j] /= t;
         c->scaleF[(size_t)k * c->N + h] = log(t);
      }
   }
}

/* codeml.c:3526-3582 == baseml.c:1517-1570 ConditionalPNode (internal inode only; one gene) */
static void cond_p_node(orc_ctx *c, int inode)
{
   int n = c->n, i, j, k;
   long long h, N = c->N;
   double *cp, t;

   for (i = 0; i < c->nson[inode]; i++) {
      int ison = c->sons[inode * c->maxsons + i];
      if (c->nson[ison] > 0) cond_p_node(c, ison);
   }
   cp = c->conP + (size_t)(inode - c->ns) * N * n;
   for (h = 0; h < N * n; h++) cp[h] = 1;

   for (i = 0; i < c->nson[inode]; i++) {
      int ison = c->sons[inode * c->maxsons + i];
      orc_pmat(c->PMat, c->bt[ison], n, c->kind, c->nr, c->set_data + (size_t)c->bset[ison] * c->set_stride);
      if (c->nson[ison] < 1 && c->cleandata) {              /* tip && clean */
         const unsigned char *zs = c->z + (size_t)ison * N;
         for (h = 0; h < N; h++)
            for (j = 0; j < n; j++) cp[h * n + j] *= c->PMat[j * n + zs[h]];
      }
      else if (c->nson[ison] < 1) {                          /* tip & unclean */
         const unsigned char *zs = c->z + (size_t)ison * N;
         for (h = 0; h < N; h++)
            for (j = 0; j < n; j++) {
               for (k = 0, t = 0; k < c->nchara[zs[h]]; k++) t += c->PMat[j * n + c->charamap[zs[h] * 64 + k]];
               cp[h * n + j] *= t;
            }
      }
      else {                                                 /* internal node */
         const double *cs = c->conP + (size_t)(ison - c->ns) * N * n;
         for (h = 0; h < N; h++)
            for (j = 0; j < n; j++) {
               for (k = 0, t = 0; k < n; k++) t += c->PMat[j * n + k] * cs[h * n + k];
               cp[h * n + j] *= t;
            }
      }
   }
   if (c->scale_flags && c->scale_flags[inode]) node_scale(c, inode);
}

int orc_lnl(int n, int ns, int nnode, int root, int maxsons, long long npatt, int cleandata,
            const int *nson, const int *sons, const unsigned char *scale_flags,
            const unsigned char *z, const double *fpatt,
            const unsigned char *nchara, const unsigned char *charamap,
            int semantics, int nclass, const double *freqK,
            int set_kind, int nr, int nset, long long set_stride, const double *set_data,
            const int *branch_set, const double *branch_t, const double *pi,
            double *neg_lnL, double *lnf, double *fhK_out, double *conP_out)
{
   orc_ctx c;
   int i, ir, k, nscale = 0, K = nclass, it;
   long long h, N = npatt;
   double lnL = 0, fh, t, *fhK;
   size_t sconP = (size_t)(nnode - ns) * N * n;

   (void)nset;
   memset(&c, 0, sizeof(c));
   c.n = n; c.ns = ns; c.nnode = nnode; c.root = root; c.maxsons = maxsons; c.cleandata = cleandata;
   c.kind = set_kind; c.nr = nr; c.N = N; c.set_stride = set_stride;
   c.nson = nson; c.sons = sons; c.scale_flags = scale_flags; c.z = z; c.nchara = nchara; c.charamap = charamap;
   c.set_data = set_data;
   if (scale_flags) for (i = 0; i < nnode; i++) nscale += (scale_flags[i] != 0);
   c.conP = (double*)malloc(sconP * sizeof(double));
   c.scaleF = (double*)calloc((size_t)(nscale ? nscale : 1) * N, sizeof(double));
   c.PMat = (double*)malloc((size_t)n * n * sizeof(double));
   fhK = (double*)malloc((size_t)K * N * sizeof(double));
   if (!c.conP || !c.scaleF || !c.PMat || !fhK) return -1;

   if (semantics == ORC_SEM_LFUN) {
      /* treesub.c:7764-7807 lfun */
      c.bset = branch_set; c.bt = branch_t;
      cond_p_node(&c, root);
      if (conP_out) memcpy(conP_out, c.conP, sconP * sizeof(double));
      for (h = 0; h < N; h++) {
         const double *cr = c.conP + (size_t)(root - ns) * N * n + h * n;
         if (fpatt[h] <= 0) { if (lnf) lnf[h] = 0; continue; }
         for (i = 0, fh = 0; i < n; i++) fh += pi[i] * cr[i];
         if (fh <= 0) fh = 1e-80;
         fh = log(fh);
         for (k = 0; k < nscale; k++) fh += c.scaleF[(size_t)k * N + h];
         lnL -= fh * fpatt[h];
         if (lnf) lnf[h] = fh;
      }
   }
   else {
      /* treesub.c:7696-7761 fx_r */
      for (ir = 0; ir < K; ir++) {
         c.bset = branch_set + (size_t)ir * nnode; c.bt = branch_t + (size_t)ir * nnode;
         cond_p_node(&c, root);
         if (conP_out) memcpy(conP_out + (size_t)ir * sconP, c.conP, sconP * sizeof(double));
         for (h = 0; h < N; h++) {
            const double *cr = c.conP + (size_t)(root - ns) * N * n + h * n;
            if (fpatt[h] <= 0) { fhK[(size_t)ir * N + h] = 0; continue; }
            for (i = 0, fh = 0; i < n; i++) fh += pi[i] * cr[i];
            if (fh <= 0) fh = 1e-300;
            if (!nscale) fhK[(size_t)ir * N + h] = fh;
            else {
               fhK[(size_t)ir * N + h] = log(fh);
               for (k = 0; k < nscale; k++) fhK[(size_t)ir * N + h] += c.scaleF[(size_t)k * N + h];
            }
         }
      }
      /* treesub.c:7608-7660 lfundG */
      for (h = 0; h < N; h++) {
         if (fpatt[h] <= 0) { if (lnf) lnf[h] = 0; continue; }
         if (nscale) {
            for (ir = 1, it = 0; ir < K; ir++) if (fhK[(size_t)ir * N + h] > fhK[(size_t)it * N + h]) it = ir;
            t = fhK[(size_t)it * N + h];
            for (ir = 0, fh = 0; ir < K; ir++) fh += freqK[ir] * exp(fhK[(size_t)ir * N + h] - t);
            fh = t + log(fh);
         }
         else {
            for (ir = 0, fh = 0; ir < K; ir++) fh += freqK[ir] * fhK[(size_t)ir * N + h];
            if (fh <= 0) fh = 1e-300;
            fh = log(fh);
         }
         lnL -= fh * fpatt[h];
         if (lnf) lnf[h] = fh;
      }
      if (fhK_out) memcpy(fhK_out, fhK, (size_t)K * N * sizeof(double));
   }
   *neg_lnL = lnL;
   free(c.conP); free(c.scaleF); free(c.PMat); free(fhK);
   return 0;
}
