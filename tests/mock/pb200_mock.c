/* pb200_mock.c -- TEST INFRASTRUCTURE ONLY (tests/).  Never built into, linked with or shipped beside the product.
 *
 * A stand-in for libpaml_b200.so that answers the C ABI of include/paml_b200.h on the CPU with the oracle
 * (oracle/paml_oracle.c) and plain loops, so that the HOST side of the drop-in -- paml_b200/host/plfun_b200.inc and
 * pb200_gather.inc: model mapping, gene units, fx_r, eigen queue + memo, PatternWeight bookkeeping, the minbranches
 * dispatch -- can be exercised by `pytest -m "not gpu"` on a machine without a GPU.  tests/mock/Makefile links the SAME
 * wrapper translation units against this file into tests/mock/_build/{codeml,baseml}_mock; tests/test_host_logic_cpu.py runs
 * them beside the unmodified reference.  It checks the shim, not the kernels: those are checked on the GPU
 * (tests/test_*_gpu.py).  The product has no CPU path (paml_b200/ never references this directory).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "paml_b200.h"
#include "paml_oracle.h"

struct pb200_ctx {
   int n, ns, cleandata, nnode, root, maxsons, have_tree;
   long long npatt;
   unsigned char *z, nchara[256], charamap[256 * 64], *scale_flags;
   double *fpatt;
   int *nson, *sons;
   /* last evaluation */
   int nclass, pending;
   double neg_lnL, *lnf, *fhK;
   /* resident conditional vectors */
   int cnclass, cnscale, ckind, cnr, cnset;
   long long cstride;
   double *cond, *cscale, *csets;
   long long launches;
};

static char g_err[512] = "";
static int fail(const char *msg) { snprintf(g_err, sizeof(g_err), "mock: %s", msg); return 1; }
const char *pb200_last_error(void) { return g_err; }
const char *pb200_version(void) { return "paml_b200 MOCK (tests only; oracle on the CPU) sm_100a-ABI"; }

int pb200_create(pb200_ctx **out, int device)
{
   (void)device;
   *out = (pb200_ctx*)calloc(1, sizeof(pb200_ctx));
   return *out ? 0 : fail("oom");
}
void pb200_destroy(pb200_ctx *c)
{
   if (!c) return;
   free(c->z); free(c->fpatt); free(c->nson); free(c->sons); free(c->scale_flags); free(c->lnf); free(c->fhK);
   free(c->cond); free(c->cscale); free(c->csets); free(c);
}

int pb200_set_data(pb200_ctx *c, int n, int ns, long long npatt, const unsigned char *z, const double *fpatt, int cleandata,
                   const unsigned char *nchara, const unsigned char *charamap)
{
   int i;
   free(c->z); free(c->fpatt);
   c->n = n; c->ns = ns; c->npatt = npatt; c->cleandata = cleandata;
   c->z = (unsigned char*)malloc((size_t)ns * npatt); memcpy(c->z, z, (size_t)ns * npatt);
   c->fpatt = (double*)malloc(npatt * sizeof(double)); memcpy(c->fpatt, fpatt, npatt * sizeof(double));
   memset(c->nchara, 0, 256); memset(c->charamap, 0, 256 * 64);
   if (nchara && charamap) { memcpy(c->nchara, nchara, 256); memcpy(c->charamap, charamap, 256 * 64); }
   else for (i = 0; i < n; i++) { c->nchara[i] = 1; c->charamap[i * 64] = (unsigned char)i; }
   c->cnclass = 0;
   return 0;
}

int pb200_set_tree(pb200_ctx *c, int nnode, int root, const int *nson, const int *sons, int maxsons, const unsigned char *scale_flags)
{
   free(c->nson); free(c->sons); free(c->scale_flags);
   if (nnode != c->nnode) c->cnclass = 0;
   c->nnode = nnode; c->root = root; c->maxsons = maxsons; c->have_tree = 1;
   c->nson = (int*)malloc(nnode * sizeof(int)); memcpy(c->nson, nson, nnode * sizeof(int));
   c->sons = (int*)malloc((size_t)nnode * maxsons * sizeof(int)); memcpy(c->sons, sons, (size_t)nnode * maxsons * sizeof(int));
   c->scale_flags = (unsigned char*)calloc(nnode, 1);
   if (scale_flags) memcpy(c->scale_flags, scale_flags, nnode);
   return 0;
}

int pb200_default_scale_flags(int nnode, int root, const int *nson, const int *sons, int maxsons, int every, unsigned char *flags)
{
   return orc_set_node_scale(nnode, root, nson, sons, maxsons, every, flags);
}

int pb200_eval_begin(pb200_ctx *c, int semantics, int nclass, const double *freqK, int set_kind, int nr, int nset, long long set_stride,
                     const double *set_data, const int *branch_set, const double *branch_t, const double *pi)
{
   int i, any = 0;
   if (!c->have_tree || !c->z) return fail("set_data and set_tree must be called first");
   free(c->lnf); free(c->fhK);
   c->lnf = (double*)calloc(c->npatt, sizeof(double));
   c->fhK = (double*)calloc((size_t)nclass * c->npatt, sizeof(double));
   for (i = 0; i < c->nnode; i++) any |= c->scale_flags[i];
   if (orc_lnl(c->n, c->ns, c->nnode, c->root, c->maxsons, c->npatt, c->cleandata, c->nson, c->sons, any ? c->scale_flags : NULL,
               c->z, c->fpatt, c->nchara, c->charamap, semantics, nclass, freqK, set_kind, nr, nset, set_stride, set_data,
               branch_set, branch_t, pi, &c->neg_lnL, c->lnf, c->fhK, NULL))
      return fail("oracle out of memory");
   c->nclass = nclass; c->pending = 1; c->launches += 3;
   return 0;
}
int pb200_eval_end(pb200_ctx *c, double *neg_lnL)
{
   if (!c->pending) return fail("no evaluation in flight");
   c->pending = 0; *neg_lnL = c->neg_lnL;
   return 0;
}
int pb200_eval(pb200_ctx *c, int semantics, int nclass, const double *freqK, int set_kind, int nr, int nset, long long set_stride,
               const double *set_data, const int *branch_set, const double *branch_t, const double *pi, double *neg_lnL)
{
   if (pb200_eval_begin(c, semantics, nclass, freqK, set_kind, nr, nset, set_stride, set_data, branch_set, branch_t, pi)) return 1;
   return pb200_eval_end(c, neg_lnL);
}
int pb200_get_lnf(pb200_ctx *c, double *lnf) { memcpy(lnf, c->lnf, c->npatt * sizeof(double)); return 0; }
int pb200_get_fhK(pb200_ctx *c, double *fhK) { memcpy(fhK, c->fhK, (size_t)c->nclass * c->npatt * sizeof(double)); return 0; }

int pb200_pmat(pb200_ctx *c, int n, int set_kind, int nr, int nset, long long set_stride, const double *set_data, int nmat,
               const int *mat_set, const double *mat_t, double *P_out)
{
   int m;
   (void)c; (void)nset;
   for (m = 0; m < nmat; m++) orc_pmat(P_out + (size_t)m * n * n, mat_t[m], n, set_kind, nr, set_data + (size_t)mat_set[m] * set_stride);
   return 0;
}

/* eigen-solution of a reversible rate matrix: cyclic Jacobi on the symmetrised matrix (test code) */
int pb200_eigen_qrev(pb200_ctx *c, int n, int njob, const double *Q, const double *pi, double *Root, double *U, double *V, int *sweeps)
{
   typedef int (*eigfn)(double*, double*, int, double*, double*, double*, double*);
   static eigfn host = NULL;
   static int looked = 0;
   int job, i, j, k, p, q, sw;
   (void)c;
   if (!looked) { looked = 1; host = (eigfn)dlsym(RTLD_DEFAULT, "eigenQREV"); }   /* the mock drop-in exports the reference's solver */
   for (job = 0; job < njob && host; job++) {               /* fast path for the host-logic tests: Householder + QL of the reference */
      double Qc[64 * 64], pc[64], sp[64];
      memcpy(Qc, Q + (size_t)job * n * n, (size_t)n * n * sizeof(double)); memcpy(pc, pi + (size_t)job * n, n * sizeof(double));
      host(Qc, pc, n, Root + (size_t)job * n, U + (size_t)job * n * n, V + (size_t)job * n * n, sp);
      if (sweeps) sweeps[job] = 1;
   }
   if (host) return 0;
   for (job = 0; job < njob; job++) {
      const double *Qj = Q + (size_t)job * n * n, *pj = pi + (size_t)job * n;
      double *R = Root + (size_t)job * n, *Uj = U + (size_t)job * n * n, *Vj = V + (size_t)job * n * n;
      int idx[64], m = 0, rank[64];
      double ps[64], A[64 * 64], E[64 * 64];
      for (j = 0; j < n; j++) if (pj[j] > 1e-100) { idx[m] = j; ps[m] = sqrt(pj[j]); m++; }
      for (i = 0; i < m; i++) for (j = 0; j < m; j++) {
         int hi = i > j ? i : j, lo = i > j ? j : i;
         A[i * m + j] = (i == j ? Qj[idx[i] * n + idx[i]] : Qj[idx[hi] * n + idx[lo]] * ps[hi] / ps[lo]);
         E[i * m + j] = (i == j);
      }
      for (sw = 0; sw < 60; sw++) {
         double off = 0, tot = 0;
         for (i = 0; i < m; i++) for (j = 0; j < m; j++) { tot += A[i * m + j] * A[i * m + j]; if (i != j) off += A[i * m + j] * A[i * m + j]; }
         if (off <= 1e-30 * tot) break;
         for (p = 0; p < m - 1; p++) for (q = p + 1; q < m; q++) {
            double apq = A[p * m + q], d, t, cs, sn;
            if (apq == 0) continue;
            d = A[q * m + q] - A[p * m + p];
            t = 2 * apq / (fabs(d) + sqrt(d * d + 4 * apq * apq)); if (d < 0) t = -t;
            cs = 1 / sqrt(t * t + 1); sn = t * cs;
            for (k = 0; k < m; k++) { double x = A[k * m + p], y = A[k * m + q]; A[k * m + p] = cs * x - sn * y; A[k * m + q] = sn * x + cs * y; }
            for (k = 0; k < m; k++) { double x = A[p * m + k], y = A[q * m + k]; A[p * m + k] = cs * x - sn * y; A[q * m + k] = sn * x + cs * y; }
            for (k = 0; k < m; k++) { double x = E[k * m + p], y = E[k * m + q]; E[k * m + p] = cs * x - sn * y; E[k * m + q] = sn * x + cs * y; }
         }
      }
      if (sweeps) sweeps[job] = sw;
      for (i = 0; i < m; i++) { rank[i] = 0; for (j = 0; j < m; j++) if (A[j * m + j] > A[i * m + i] || (A[j * m + j] == A[i * m + i] && j < i)) rank[i]++; }
      for (i = 0; i < n; i++) { R[i] = 0; for (j = 0; j < n; j++) Uj[i * n + j] = Vj[i * n + j] = (i == j); }
      for (i = 0; i < m; i++) R[idx[rank[i]]] = A[i * m + i];
      for (i = 0; i < m; i++) for (k = 0; k < m; k++) {
         int col = idx[rank[k]];
         Uj[idx[i] * n + col] = E[i * m + k] / ps[i];
         Vj[col * n + idx[i]] = E[i * m + k] * ps[i];
      }
   }
   return 0;
}

/* ---- resident conditional vectors: the definitions of pb200_cond_update / pb200_branch_lnl, as plain loops ---- */
int pb200_cond_update(pb200_ctx *c, int nclass, int set_kind, int nr, int nset, long long set_stride, const double *set_data,
                      const int *branch_set, const double *branch_t, int nupd, const int *upd_node, const int *upd_nson,
                      const int *upd_sons, int maxsons, const unsigned char *scale_flags)
{
   int n = c->n, ns = c->ns, nnode = c->nnode, nint = nnode - ns, nscale = 0, sk[8192], u, s, ir, i, j, k;
   long long N = c->npatt, h;
   double *P;
   if (nnode > 8192) return fail("tree too large for the mock");
   P = (double*)malloc((size_t)n * n * sizeof(double));
   for (i = 0; i < nnode; i++) { sk[i] = -1; if (scale_flags && scale_flags[i]) sk[i] = nscale++; }
   if (c->cnclass != nclass || c->cnscale != nscale || !c->cond) {
      free(c->cond); free(c->cscale);
      c->cond = (double*)calloc((size_t)nclass * nint * N * n, sizeof(double));
      c->cscale = (double*)calloc((size_t)nclass * (nscale ? nscale : 1) * N, sizeof(double));
      c->cnclass = nclass; c->cnscale = nscale;
   }
   free(c->csets);
   c->csets = (double*)malloc((size_t)(nset * set_stride + 1) * sizeof(double));
   if (set_stride) memcpy(c->csets, set_data, (size_t)nset * set_stride * sizeof(double));
   c->ckind = set_kind; c->cnr = nr; c->cnset = nset; c->cstride = set_stride;
   for (u = 0; u < nupd; u++) {
      int node = upd_node[u];
      for (ir = 0; ir < nclass; ir++) {
         double *cp = c->cond + ((size_t)ir * nint + (node - ns)) * N * n;
         for (h = 0; h < N * n; h++) cp[h] = 1;
         for (s = 0; s < upd_nson[u]; s++) {
            int son = upd_sons[u * maxsons + s];
            orc_pmat(P, branch_t[ir * nnode + son], n, set_kind, nr, set_data + (size_t)branch_set[ir * nnode + son] * set_stride);
            for (h = 0; h < N; h++)
               for (j = 0; j < n; j++) {
                  double t = 0;
                  if (son < ns) {
                     int code = c->z[(size_t)son * N + h];
                     for (k = 0; k < c->nchara[code]; k++) t += P[j * n + c->charamap[code * 64 + k]];
                  }
                  else {
                     const double *cs = c->cond + ((size_t)ir * nint + (son - ns)) * N * n + h * n;
                     for (k = 0; k < n; k++) t += P[j * n + k] * cs[k];
                  }
                  cp[h * n + j] *= t;
               }
         }
         if (sk[node] >= 0)
            for (h = 0; h < N; h++) {
               double t = 0, *f = c->cscale + ((size_t)ir * nscale + sk[node]) * N + h;
               for (j = 0; j < n; j++) if (cp[h * n + j] > t) t = cp[h * n + j];
               if (t < 1e-300) { for (j = 0; j < n; j++) cp[h * n + j] = 1; *f = -800; }
               else { for (j = 0; j < n; j++) cp[h * n + j] /= t; *f = log(t); }
            }
      }
   }
   free(P);
   c->launches += nupd + 1;
   return 0;
}

int pb200_branch_lnl(pb200_ctx *c, int nclass, const double *freqK, int a, int b, double t, const double *mult, const int *set_of_class,
                     const double *pi, int semantics, int want_dd, double out[3])
{
   int n = c->n, ns = c->ns, nint = c->nnode - ns, nscale = c->cnscale, ir, i, j, k, nroot;
   long long N = c->npatt, h;
   double *P = (double*)calloc((size_t)3 * n * n, sizeof(double)), *dP = P + n * n, *ddP = dP + n * n;
   double *fh = (double*)calloc(3 * N, sizeof(double)), *dfh = fh + N, *ddfh = dfh + N, *Sh = (double*)calloc(N, sizeof(double));
   if (!c->cond || c->cnclass != nclass) return fail("pb200_cond_update must come first");
   if (c->ckind != PB200_KIND_UVROOT && c->ckind != PB200_KIND_CIJK) return fail("needs UVROOT or CIJK sets");
   for (h = 0; h < N && nscale; h++)
      for (ir = 0; ir < nclass; ir++) {
         double y = 0;
         for (k = 0; k < nscale; k++) y += c->cscale[((size_t)ir * nscale + k) * N + h];
         if (ir == 0 || y > Sh[h]) Sh[h] = y;
      }
   for (ir = 0; ir < nclass; ir++) {
      const double *set = c->csets + (size_t)set_of_class[ir] * c->cstride, *Root = set;
      const double *A = c->cond + ((size_t)ir * nint + (a - ns)) * N * n;
      nroot = (c->ckind == PB200_KIND_UVROOT ? n : c->cnr);
      memset(P, 0, (size_t)3 * n * n * sizeof(double));
      for (k = 0; k < nroot; k++) {                      /* src/treesub.c:8264-8288 */
         double mu = mult[ir] * Root[k], ex = (k ? exp(t * mu) : 1);
         for (i = 0; i < n; i++) for (j = 0; j < n; j++) {
            double cijk = (c->ckind == PB200_KIND_UVROOT ? set[n + i * n + k] * set[n + n * n + k * n + j] : set[c->cnr + (i * n + j) * c->cnr + k]);
            P[i * n + j] += cijk * ex;
            if (k) { dP[i * n + j] += cijk * ex * mu; ddP[i * n + j] += cijk * ex * mu * mu; }
         }
      }
      for (h = 0; h < N; h++) {
         double pK = 1;
         int n1 = n, code = 0;
         if (semantics == 1) {
            pK = freqK[ir];
            if (nscale) { double y = 0; for (k = 0; k < nscale; k++) y += c->cscale[((size_t)ir * nscale + k) * N + h]; pK *= exp(y - Sh[h]); }
         }
         if (b < ns) { code = c->z[(size_t)b * N + h]; n1 = c->nchara[code]; }
         for (i = 0; i < n1; i++) {
            int xb = (b < ns ? c->charamap[code * 64 + i] : i);
            double piqi = (b < ns ? pK * pi[xb] : pK * pi[i] * c->cond[((size_t)ir * nint + (b - ns)) * N * n + h * n + i]);
            double pq = 0, dpq = 0, ddpq = 0;
            for (j = 0; j < n; j++) { pq += P[xb * n + j] * A[h * n + j]; dpq += dP[xb * n + j] * A[h * n + j]; ddpq += ddP[xb * n + j] * A[h * n + j]; }
            fh[h] += piqi * pq; dfh[h] += piqi * dpq; ddfh[h] += piqi * ddpq;
         }
      }
   }
   out[0] = out[1] = out[2] = 0;
   for (h = 0; h < N; h++) {
      double fp = c->fpatt[h];
      if (fp == 0) continue;
      out[0] -= ((semantics == 0 && !want_dd && fh[h] < 0) ? -500 : log(fh[h])) * fp;
      if (nscale) out[0] -= Sh[h] * fp;
      if (want_dd) { out[1] -= dfh[h] / fh[h] * fp; out[2] -= (fh[h] * ddfh[h] - dfh[h] * dfh[h]) / (fh[h] * fh[h]) * fp; }
   }
   free(P); free(fh); free(Sh);
   c->launches += 3;
   return 0;
}

int pb200_cond_get(pb200_ctx *c, int node, double *cond, double *scaleF)
{
   int ir, nint = c->nnode - c->ns;
   size_t per = (size_t)c->npatt * c->n;
   if (!c->cond) return fail("nothing resident");
   for (ir = 0; cond && ir < c->cnclass; ir++) memcpy(cond + ir * per, c->cond + ((size_t)ir * nint + (node - c->ns)) * per, per * sizeof(double));
   if (scaleF && c->cnscale) memcpy(scaleF, c->cscale, (size_t)c->cnclass * c->cnscale * c->npatt * sizeof(double));
   return 0;
}

/* ---- entry points that make no sense without a GPU ---- */
int pb200_nccl_unique_id(char id[128]) { (void)id; return fail("no NCCL in the mock"); }
int pb200_comm_init(pb200_ctx *c, int nranks, int rank, const char id[128]) { (void)c; (void)nranks; (void)rank; (void)id; return fail("no NCCL in the mock"); }
int pb200_p2p_local_handle(pb200_ctx *c, char handle[64]) { (void)c; (void)handle; return fail("no peer memory in the mock"); }
int pb200_p2p_init(pb200_ctx *c, int nranks, int rank, const char *handles) { (void)c; (void)nranks; (void)rank; (void)handles; return fail("no peer memory in the mock"); }
int pb200_p2p_enable(pb200_ctx *c, int enable) { (void)c; (void)enable; return fail("no peer memory in the mock"); }
int pb200_set_timing(pb200_ctx *c, int enable) { (void)c; (void)enable; return 0; }
int pb200_last_timing(pb200_ctx *c, double ms[5]) { (void)c; memset(ms, 0, 5 * sizeof(double)); return 0; }
/* the branch iteration (src/treesub.c:8087-8108) as a plain host loop over the mock's own pb200_branch_lnl */
int pb200_branch_newton(pb200_ctx *c, int nclass, const double *freqK, int a, int b, double t, double l0, const double *mult,
                        const int *set_of_class, const double *pi, int semantics, int first_cycle, double e, const double tb[2],
                        double out[4], int *nevals)
{
   double t0 = t, l = l0, dl = 0, ddl = -1, p, step, o[3], smallv = 1e-20;
   int icycleb, i, nev = 0;
   for (icycleb = 0; icycleb < 10; icycleb++) {
      if (pb200_branch_lnl(c, nclass, freqK, a, b, t, mult, set_of_class, pi, semantics, 1, o)) return 1;
      nev++; dl = o[1]; ddl = o[2];
      p = -dl / fabs(ddl);
      if (fabs(p) < smallv) step = 0;
      else if (p < 0)       step = (1 < (tb[0] - t0) / p ? 1 : (tb[0] - t0) / p);
      else                  step = (1 < (tb[1] - t0) / p ? 1 : (tb[1] - t0) / p);
      if (first_cycle && step != 1 && step != 0) step *= 0.99;
      for (i = 0; step > smallv; i++, step /= 4) {
         t = t0 + step * p;
         if (pb200_branch_lnl(c, nclass, freqK, a, b, t, mult, set_of_class, pi, semantics, 0, o)) return 1;
         nev++; l = o[0];
         if (l < l0) break;
      }
      if (step <= smallv) { t = t0; l = l0; break; }
      if (fabs(t - t0) < e * fabs(1 + t) && fabs(l - l0) < e) break;
      t0 = t; l0 = l;
   }
   out[0] = t; out[1] = l; out[2] = dl; out[3] = ddl;
   if (nevals) *nevals = nev;
   return 0;
}

long long pb200_launch_count(pb200_ctx *c) { return c ? c->launches : -1; }
const char *pb200_last_kernel(pb200_ctx *c) { return "mock (CPU oracle, tests only)"; }
int pb200_plan_tree(int nnode, int ns, int root, const int *nson, const int *sons, int maxsons, const unsigned char *scale_flags, int *prog, int cap, int *nslots)
{ (void)nnode; (void)ns; (void)root; (void)nson; (void)sons; (void)maxsons; (void)scale_flags; (void)prog; (void)cap; (void)nslots; return -1; }
int pb200_plan_tree_ps(int nnode, int ns, int root, const int *nson, const int *sons, int maxsons, const unsigned char *scale_flags, int *prog, int cap, int *nslots)
{ (void)nnode; (void)ns; (void)root; (void)nson; (void)sons; (void)maxsons; (void)scale_flags; (void)prog; (void)cap; (void)nslots; return -1; }
int pb200_plan_tree4(int nnode, int ns, int root, const int *nson, const int *sons, int maxsons, const unsigned char *scale_flags, int *ops, int cap_ops,
                     int *tip_order, int cap_tips, int *ntip_out, int *nslots)
{ (void)nnode; (void)ns; (void)root; (void)nson; (void)sons; (void)maxsons; (void)scale_flags; (void)ops; (void)cap_ops; (void)tip_order; (void)cap_tips; (void)ntip_out; (void)nslots; return -1; }
int pb200_measure_fp64_peak(pb200_ctx *c, double *a, double *b) { (void)c; (void)a; (void)b; return fail("no GPU in the mock"); }
int pb200_debug_set_program(pb200_ctx *c, const int *prog, int nints, int nslots, int nscale) { (void)c; (void)prog; (void)nints; (void)nslots; (void)nscale; return fail("no GPU in the mock"); }
int pb200_bench_resident(pb200_ctx *c, int steps, long long flush_bytes, double *a, double *b, double *d, double *e) { (void)c; (void)steps; (void)flush_bytes; (void)a; (void)b; (void)d; (void)e; return fail("no GPU in the mock"); }
int pb200_bench_e2e(pb200_ctx *c, int steps, long long flush_bytes, int semantics, int nclass, const double *freqK, int set_kind, int nr, int nset,
                    long long set_stride, const double *set_data, const int *branch_set, const double *branch_t, const double *pi, double *sec_step, double *neg_lnL)
{ (void)c; (void)steps; (void)flush_bytes; (void)semantics; (void)nclass; (void)freqK; (void)set_kind; (void)nr; (void)nset; (void)set_stride; (void)set_data; (void)branch_set; (void)branch_t; (void)pi; (void)sec_step; (void)neg_lnL; return fail("no GPU in the mock"); }
int pb200_flush_l2(pb200_ctx *c, long long bytes) { (void)c; (void)bytes; return 0; }
