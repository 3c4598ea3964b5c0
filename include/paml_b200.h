/* paml_b200.h -- C ABI of libpaml_b200.so: the B200 (sm_100a) implementation of PAML's tree-likelihood
 * hot path.  Plain C types only (pointers, sizes); no torch / C++ types cross this boundary.
 *
 * The reference (abacus-gene/paml 4.10.10) has no plugin or FFI layer: its seam is the function pointer
 * `double (*com.plfun)(double x[], int np)` (src/codeml.c:125, src/baseml.c:70) bound to `lfun`
 * (src/treesub.c:7764) or `lfundG` (src/treesub.c:7608), plus the globals those functions read.  The entry
 * points below are what a C binding for that seam needs; each names the reference state it replaces.
 * paml_b200/host/plfun_b200.inc is that binding (interposed lfun/lfundG calling these functions);
 * INTEGRATION.md shows how it is compiled against an unmodified reference tree.
 *
 * Conventions: every function returns 0 on success, non-zero on error (pb200_last_error() gives the
 * message; the reference itself has no error returns, its `zerror` prints and exits, src/tools.c:1204).
 * All floating point is IEEE double, as in the reference.  A context owns one GPU and one CUDA stream;
 * calls on one context must come from one thread at a time (the reference is single-threaded).
 */
#ifndef PAML_B200_H
#define PAML_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb200_ctx pb200_ctx;

/* Spectral-set kinds: which reference P(t) routine a set feeds.  Layout of ONE set (set_stride doubles):
 *   PB200_KIND_UVROOT   Root[n], U[n*n], V[n*n] (row-major)   PMatUVRoot, src/tools.c:516-546
 *   PB200_KIND_CIJK     Root[nr], Cijk[n*n*nr]                PMatCijk,   src/baseml.c:1572-1589
 *   PB200_KIND_K80      kappa                                 PMatK80,    src/tools.c:578-606
 *   PB200_KIND_JC69LIKE (nothing)                             PMatJC69like, src/codeml.c:3585-3594
 *   PB200_KIND_QMATEXP  Q[n*n] (row-major rate matrix, n<=16) matexp(Q t, 7, 5), src/treesub.c:7523-7527,
 *                                                             src/tools.c:4879-4919 (UNREST / UNRESTu)   */
enum { PB200_KIND_UVROOT = 0, PB200_KIND_CIJK = 1, PB200_KIND_K80 = 2, PB200_KIND_JC69LIKE = 3, PB200_KIND_QMATEXP = 4 };

/* Which reference function's arithmetic pb200_eval reproduces:
 *   PB200_SEM_LFUN    lfun   (src/treesub.c:7764-7807): one class, f_h <= 0 -> 1e-80
 *   PB200_SEM_LFUNDG  fx_r + lfundG (src/treesub.c:7696-7761, 7608-7660): class mixture with freqK,
 *                     log-sum-exp form when scaling nodes exist, f_h <= 0 -> 1e-300                      */
enum { PB200_SEM_LFUN = 0, PB200_SEM_LFUNDG = 1 };

/* Create / destroy a context bound to CUDA device `device`.  Fails loudly (non-zero + message) when no
 * usable sm_100 GPU is present: there is no CPU fallback. */
int pb200_create(pb200_ctx **out, int device);
void pb200_destroy(pb200_ctx *ctx);
/* Message of the last failing call on this thread ("" if none). */
const char *pb200_last_error(void);
/* Library / build description string (arch, version). */
const char *pb200_version(void);

/* Replaces com.z / com.fpatt / com.cleandata / nChara / CharaMap (src/codeml.c:110, src/treesub.c:1116-1286,
 * src/tools.c:20): uploads this context's site patterns.
 *   n        com.ncode (<= 64)           ns     number of tips
 *   npatt    patterns held by THIS context (a shard when several GPUs split the alignment)
 *   z        [ns][npatt] coded tip states, 0..n-1 resolved, >= n index nchara/charamap
 *   fpatt    [npatt] pattern weights     cleandata  com.cleandata
 *   nchara   [256], charamap [256][64]   ambiguity sets (may be NULL when every z < n)                    */
int pb200_set_data(pb200_ctx *ctx, int n, int ns, long long npatt, const unsigned char *z,
                   const double *fpatt, int cleandata, const unsigned char *nchara,
                   const unsigned char *charamap);

/* Replaces nodes[].nson/sons + tree.root (src/codeml.c:138-147) and com.nodeScale (src/treesub.c:7141-7198):
 * the topology and the nodes at which conditional vectors are rescaled.  Tips are nodes 0..ns-1.
 *   sons        [nnode][maxsons], unused entries ignored
 *   scale_flags [nnode] or NULL (no scaling); the root entry is ignored                                  */
int pb200_set_tree(pb200_ctx *ctx, int nnode, int root, const int *nson, const int *sons, int maxsons,
                   const unsigned char *scale_flags);

/* Restatement of SetNodeScale (src/treesub.c:7178-7198) for callers that do not have the reference's flags:
 * writes flags[nnode], returns the number of scaling nodes.  every = 100 (baseml), 15 (codon), 50 (aa).   */
int pb200_default_scale_flags(int nnode, int root, const int *nson, const int *sons, int maxsons,
                              int every, unsigned char *flags);

/* One likelihood evaluation = one call of com.plfun after SetParameters(x) (src/treesub.c:7764 / 7608):
 * batched P(t) for every (class, branch), Felsenstein pruning of every pattern of this context, root
 * reduction with pi, class mixture, weighting by fpatt, deterministic sum; if a communicator was set up
 * (pb200_comm_init) one ncclAllReduce of the scalar.  Host buffers in, one double out.
 *   nclass, freqK[nclass]     com.ncatG / com.freqK (1 / {1} for PB200_SEM_LFUN)
 *   set_kind, nr, nset, set_stride, set_data[nset][set_stride]   spectral sets, layout above
 *   branch_set[nclass][nnode] which set branch (class, node) uses (root entry ignored)
 *   branch_t  [nclass][nnode] its t, all factors applied (rate class, rgene / branch rate, Qfactor)
 *   pi[n]                     com.pi
 *   neg_lnL                   out: minus log-likelihood, the value com.plfun returns                     */
int pb200_eval(pb200_ctx *ctx, int semantics, int nclass, const double *freqK, int set_kind, int nr,
               int nset, long long set_stride, const double *set_data, const int *branch_set,
               const double *branch_t, const double *pi, double *neg_lnL);

/* The same evaluation split in two, so that one host thread can drive several GPUs (one context each, each
 * holding a shard of the patterns): _begin packs and enqueues everything on the context's stream and returns
 * without waiting, _end waits and delivers this context's -lnL (the global one if a communicator is set). */
int pb200_eval_begin(pb200_ctx *ctx, int semantics, int nclass, const double *freqK, int set_kind, int nr,
                     int nset, long long set_stride, const double *set_data, const int *branch_set,
                     const double *branch_t, const double *pi);
int pb200_eval_end(pb200_ctx *ctx, double *neg_lnL);

/* Per-pattern outputs of the last pb200_eval on this context (this context's shard only):
 *   lnf[npatt]          log f_h, what the reference stores in dfsites[] / prints to `lnf`
 *                       (src/treesub.c:7801-7803, 7655-7656)
 *   fhK[nclass][npatt]  com.fhK as left by fx_r (src/treesub.c:7743-7750)                                */
int pb200_get_lnf(pb200_ctx *ctx, double *lnf);
int pb200_get_fhK(pb200_ctx *ctx, double *fhK);

/* The batched P(t) kernel on its own (what GetPMatBranch computes, src/treesub.c:7503-7592):
 * nmat matrices, matrix m from set mat_set[m] at mat_t[m]; P_out[nmat][n][n] row-major on the host.      */
int pb200_pmat(pb200_ctx *ctx, int n, int set_kind, int nr, int nset, long long set_stride,
               const double *set_data, int nmat, const int *mat_set, const double *mat_t, double *P_out);

/* Replaces eigenQREV (src/tools.c:5023-5110; eigenRealSym = Householder + QL, src/tools.c:5116-5309), the spectral
 * decomposition eigenQcodon / eigenQaa / eigenQREVbase need once per parameter set:
 * Q = S diag(pi) reversible, Q = U diag(Root) V with U V = I, Root in decreasing order, states with pi <= 1e-100
 * dropped and re-inserted as identity like the reference.  njob matrices are solved by njob thread blocks of one
 * launch (parallel-order two-sided Jacobi in shared memory).  The factors differ from the reference's by the signs of
 * the eigenvectors (and by any rotation inside a repeated root); P(t) = U exp(Root t) V is the same to rounding.
 *   Q[njob][n*n] row-major, pi[njob][n]  ->  Root[njob][n], U[njob][n*n], V[njob][n*n]; sweeps[njob] (may be NULL) */
int pb200_eigen_qrev(pb200_ctx *ctx, int n, int njob, const double *Q, const double *pi, double *Root, double *U,
                     double *V, int *sweeps);

/* The whole Newton iteration of ONE branch on the device: the reference's "iterating a branch" loop of minbranches
 * (src/treesub.c:8087-8108) -- lfuntdd at t, the direction p = -dl/|ddl| bounded to tb[], step quartering on lfunt until -lnL
 * drops below l0, the convergence test with e, up to 10 rounds -- as a conditional WHILE graph whose body is the evaluation
 * kernels plus a one-block step kernel.  One host->device copy, one graph launch and one device->host copy per branch.
 * PRE: as pb200_branch_lnl.  t: current branch length; l0: current -lnL; first_cycle: icycle == 0 (border avoidance, :8098).
 * out[4] = { t, -lnL(t), dl, ddl } as the loop leaves them (dl, ddl of the last lfuntdd: g_minbranches / varb_minbranches);
 * *nevals (may be NULL) = evaluations done.  Same arithmetic as a host-driven loop over pb200_branch_lnl.               */
int pb200_branch_newton(pb200_ctx *ctx, int nclass, const double *freqK, int a, int b, double t, double l0,
                        const double *mult, const int *set_of_class, const double *pi, int semantics, int first_cycle,
                        double e, const double tb[2], double out[4], int *nevals);
/* The one-branch-at-a-time optimiser (`method = 1`: minbranches, src/treesub.c:8039-8123) needs what the evaluation
 * kernels never write: every internal node's conditional vectors, RESIDENT in HBM ([class][internal node][pattern]
 * [state], the layout of com.conP, src/treesub.c:3518-3530).
 *
 * pb200_cond_update replaces updateconP / ConditionalPNode with com.oldconP (src/treesub.c:7982-8036, src/codeml.c:
 * 3526-3582): it recomputes the listed internal nodes, in the order given (sons before fathers), under the CURRENT
 * rooting described by the son lists, from P(t) of the branches they read; nodes not listed keep their vectors.
 *   nclass               1, or com.ncatG when com.conPSiteClass
 *   set_kind .. branch_t as in pb200_eval, for the current rooting (entry of node i = the branch above i)
 *   upd_node[nupd], upd_nson[nupd], upd_sons[nupd][maxsons]   nodes to recompute and their sons (tips: ids < ns)
 *   scale_flags[nnode]   com.nodeScale or NULL: NodeScale (src/treesub.c:7200-7230) is applied at flagged nodes
 * The spectral sets stay resident for the pb200_branch_lnl calls that follow.
 *
 * pb200_branch_lnl replaces lfunt / lfuntdd / lfunt_SiteClass / lfuntdd_SiteClass (src/treesub.c:8127-8543) for the
 * branch b -> a at length t (the tree rooted at b, a internal, b a tip or internal):
 *   f_h = sum_classes pK sum_i pi_i L_b(i) sum_j P_ij L_a(j),  P = sum_k C_ijk exp(t mult Root_k)  (and dP/dt, d2P/dt2)
 *   mult[nclass]         rgene * _rateSite * Qfactor_NS_branch, set_of_class[nclass] the spectral set of branch a per class
 *   semantics            0 = lfunt/lfuntdd (one class), 1 = the _SiteClass forms (freqK, per-class scale bookkeeping)
 *   out[3]               -lnL, d(-lnL)/dt, d2(-lnL)/dt2 (the last two only when want_dd)
 * pb200_cond_get copies one node's resident vectors / the scale factors back (tests).                              */
int pb200_cond_update(pb200_ctx *ctx, int nclass, int set_kind, int nr, int nset, long long set_stride,
                      const double *set_data, const int *branch_set, const double *branch_t, int nupd,
                      const int *upd_node, const int *upd_nson, const int *upd_sons, int maxsons,
                      const unsigned char *scale_flags);
int pb200_branch_lnl(pb200_ctx *ctx, int nclass, const double *freqK, int a, int b, double t, const double *mult,
                     const int *set_of_class, const double *pi, int semantics, int want_dd, double out[3]);
int pb200_cond_get(pb200_ctx *ctx, int node, double *cond, double *scaleF);

/* Replaces PatternWeight (src/treesub.c:1386-1516), the step that turns an alignment into the site patterns the
 * likelihood path consumes: the ls columns are collapsed, gene by gene, into their distinct patterns in ascending
 * strcmp order of (character + 1), each represented by the first site showing it -- the reference's order, counts and
 * site map, bit for bit, in O(ls ns + npatt log npatt) instead of the reference's O(npatt^2) insertion (host code:
 * bytes only, runs once per data set).
 *   ns, width        sequences; bytes per site and sequence (3 for codon data before encoding, else 1)
 *   ls               sites (codons for codon data);   z[ns] row pointers, row j = ls*width bytes (com.z)
 *   gene[ls]         gene label of every site (com.pose on entry) or NULL;   ngene  com.ngene
 *   out: *npatt; posG[ngene+1] (com.posG); lgene[ngene] sites per gene (may be NULL); pose[ls] site -> pattern
 *        (com.pose on return); fpatt[ls] (first *npatt entries used, com.fpatt); rep[ls] (first *npatt entries: the
 *        site whose column represents each pattern -- the caller gathers com.z from it)                           */
int pb200_compress_patterns(int ns, int width, long long ls, const unsigned char *const *z, const int *gene,
                            int ngene, long long *npatt, long long *posG, long long *lgene, int *pose,
                            double *fpatt, long long *rep);

/* Multi-GPU: site patterns are sharded, one context (process) per GPU, and each evaluation
 * ends with ONE ncclAllReduce(sum) of the scalar -lnL.  Rank 0 makes an id, the launcher broadcasts the
 * 128 bytes, every rank calls pb200_comm_init.  NCCL is dlopen'ed at the first of these calls.           */
int pb200_nccl_unique_id(char id[128]);
int pb200_comm_init(pb200_ctx *ctx, int nranks, int rank, const char id[128]);

/* The same all-reduce fused into the reduction kernel over NVLink peer memory (no separate collective launch):
 * every rank exports a small exchange buffer (CUDA IPC handle, 64 bytes), the launcher all-gathers the handles,
 * every rank maps its peers' buffers.  From then on the last block of the mixture kernel stores this rank's
 * partial -lnL into every peer's buffer, waits for theirs and adds them in rank order (bit-identical on all
 * ranks).  Replaces the NCCL call when initialised; one process per GPU, GPUs connected by NVLink/NVSwitch. */
int pb200_p2p_local_handle(pb200_ctx *ctx, char handle[64]);
int pb200_p2p_init(pb200_ctx *ctx, int nranks, int rank, const char *handles /* [nranks][64] */);
/* Switch between the fused peer-memory all-reduce (1) and the NCCL call (0); all ranks must agree. */
int pb200_p2p_enable(pb200_ctx *ctx, int enable);

/* Timing / introspection for bench.py and the tests. */
/* Device time (ms, CUDA events on the context's stream) of the last eval: [0] total, [1] P(t) kernel,
 * [2] pruning kernel, [3] mixture/reduction kernel, [4] all-reduce.  Only filled when timing is enabled. */
int pb200_set_timing(pb200_ctx *ctx, int enable);
int pb200_last_timing(pb200_ctx *ctx, double ms[5]);
/* Number of kernel launches this context has issued since creation. */
long long pb200_launch_count(pb200_ctx *ctx);
/* Name, template shape and launch geometry of the pruning kernel the last evaluation used ("" before the first). */
const char *pb200_last_kernel(pb200_ctx *ctx);
/* Host-side planner on its own (no GPU needed): the pruning "program" the kernel walks for this tree.
 * Writes up to cap ints to prog; returns the number of ints needed (negative on error) and the number
 * of on-chip slots (conditional-vector tiles live at once) in *nslots.                                   */
int pb200_plan_tree(int nnode, int ns, int root, const int *nson, const int *sons, int maxsons,
                    const unsigned char *scale_flags, int *prog, int cap, int *nslots);
/* The same for the P-stationary kernel of the 61..64-state shape (csrc/plan.h, PlanPS; csrc/prune_ps.cuh). */
int pb200_plan_tree_ps(int nnode, int ns, int root, const int *nson, const int *sons, int maxsons,
                       const unsigned char *scale_flags, int *prog, int cap, int *nslots);
/* The same for the 4-state kernel's macro-op program (csrc/plan.h, Plan4): ops[2 ints per macro-op] and the order
 * in which the walk consumes the tips.  Returns the number of op ints (negative on error); *ntip_out = tips.  */
int pb200_plan_tree4(int nnode, int ns, int root, const int *nson, const int *sons, int maxsons,
                     const unsigned char *scale_flags, int *ops, int cap_ops, int *tip_order, int cap_tips,
                     int *ntip_out, int *nslots);
/* FP64 peak micro-benchmark on the context's GPU: tensor (DMMA m8n8k4) and vector (DFMA) TFLOP/s.
 * MEASURED_PEAKS.json has no FP64 figure; bench.py uses this as the FP64 roofline denominator.          */
int pb200_measure_fp64_peak(pb200_ctx *ctx, double *dmma_tflops, double *dfma_tflops);
/* Kernel-tuning aid: replace the pruning program of the current tree by a hand-made one (same format as
 * pb200_plan_tree writes), e.g. only tip children or only internal children, to time the phases of the
 * pruning kernel in isolation.  Results of evaluations after this call are meaningless.                    */
int pb200_debug_set_program(pb200_ctx *ctx, const int *prog, int nints, int nslots, int nscale);
/* Device-resident throughput: re-run the kernels of the last pb200_eval (parameters, data and tree already in
 * HBM) `steps` times back to back on the context's stream, optionally writing flush_bytes of scratch before
 * each step to evict L2, and report per-step device times (ms, CUDA events on that stream): whole step,
 * P(t) kernel, pruning kernel, mixture kernel (+ all-reduce when a communicator is set).                   */
int pb200_bench_resident(pb200_ctx *ctx, int steps, long long flush_bytes, double *ms_step, double *ms_pt,
                         double *ms_prune, double *ms_mix);
/* End-to-end throughput of the reference-facing call: `steps` times { [flush L2], synchronise, pb200_eval(host
 * buffers...) } with each pb200_eval timed by the host clock (pack + H2D + kernels + [all-reduce] + D2H + sync);
 * sec_step[steps] out.  The same as calling pb200_eval in a loop from C, as the drop-in shim does.         */
int pb200_bench_e2e(pb200_ctx *ctx, int steps, long long flush_bytes, int semantics, int nclass,
                    const double *freqK, int set_kind, int nr, int nset, long long set_stride,
                    const double *set_data, const int *branch_set, const double *branch_t, const double *pi,
                    double *sec_step, double *neg_lnL);
/* Write `bytes` of device memory (L2 flush between timed iterations). */
int pb200_flush_l2(pb200_ctx *ctx, long long bytes);

#ifdef __cplusplus
}
#endif
#endif /* PAML_B200_H */
