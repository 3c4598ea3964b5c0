/* paml_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, IEEE double) of PAML's tree-likelihood hot path, used ONLY as
 * the checker in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline "port" leg.
 * The product (libpaml_b200.so) never includes, links or calls anything in oracle/.
 * Pinned against the reference itself: tests/test_oracle_golden.py checks it on every fixture in
 * tests/golden/ (dumped from the unmodified reference at the examples' known-answer MLEs).
 */
#ifndef PAML_ORACLE_H
#define PAML_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_KIND_UVROOT = 0, ORC_KIND_CIJK = 1, ORC_KIND_K80 = 2, ORC_KIND_JC69LIKE = 3, ORC_KIND_QMATEXP = 4 };
enum { ORC_SEM_LFUN = 0, ORC_SEM_LFUNDG = 1 };

/* P(t) routines.  P is row-major n x n, P[i*n+j] = Pr(i -> j | t). */
void orc_pmat_uvroot(double *P, double t, int n, const double *U, const double *V, const double *Root);
void orc_pmat_cijk(double *P, double t, int n, int nr, const double *Cijk, const double *Root);
void orc_pmat_k80(double *P, double t, double kappa);
void orc_pmat_jc69like(double *P, double t, int n);
void orc_pmat_qmatexp(double *P, double t, int n, const double *Q);   /* set = Q[n*n] (UNREST) */
/* dispatch on kind; set = one spectral set laid out as in the fixtures/C-ABI (see include/paml_b200.h) */
void orc_pmat(double *P, double t, int n, int kind, int nr, const double *set);

/* SetNodeScale (treesub.c:7178-7198): flags[nnode] out, returns NnodeScale. every = 100/15/50. */
int orc_set_node_scale(int nnode, int root, const int *nson, const int *sons, int maxsons, int every,
                       unsigned char *flags);

/* One full likelihood evaluation = lfun (sem 0) or fx_r + lfundG (sem 1).
 * Returns 0 on success. Outputs: *neg_lnL; lnf[npatt] (per-pattern log f_h, may be NULL);
 * fhK[nclass*npatt] (may be NULL; meaningful for sem 1); conP_out (may be NULL):
 * [nclass][nnode-ns][npatt][n] conditional vectors of the internal nodes (after scaling). */
int orc_lnl(int n, int ns, int nnode, int root, int maxsons, long long npatt, int cleandata,
            const int *nson, const int *sons, const unsigned char *scale_flags,
            const unsigned char *z, const double *fpatt,
            const unsigned char *nchara, const unsigned char *charamap,
            int semantics, int nclass, const double *freqK,
            int set_kind, int nr, int nset, long long set_stride, const double *set_data,
            const int *branch_set, const double *branch_t, const double *pi,
            double *neg_lnL, double *lnf, double *fhK, double *conP_out);

#ifdef __cplusplus
}
#endif
#endif
