// cond.cuh -- conditional-likelihood vectors RESIDENT in HBM, for the one-branch-at-a-time optimiser.
//
// The evaluation kernels (kernels.cuh, prune4v2.cuh) never materialise internal-node vectors.  minbranches
// (src/treesub.c:8039-8123, `method = 1`) needs them: it reroots the tree at every branch b->a in turn, refreshes only
// the vectors on the path between two consecutive roots (ReRootTree src/treespace.c:236, updateconP src/treesub.c:7982,
// ConditionalPNode src/codeml.c:3526 with com.oldconP), and runs Newton steps on that branch's length from
// f_h(t) = sum_i pi_i L_b(i) sum_j P_ij(t) L_a(j) and its first two derivatives (lfunt / lfuntdd and their _SiteClass
// forms, src/treesub.c:8127-8543).  Here:
//   cond_node_kernel   one internal node: L[h][j] = prod_sons sum_k P_son[j][k] L_son[h][k]  (tips: column / ambiguity
//                      sums of P), optional NodeScale (src/treesub.c:7200-7230); reads the sons' vectors from HBM, writes
//                      the node's -- layout [class][internal node][pattern][state], the reference's own.
//   pt3_kernel         P, dP/dt, d2P/dt2 of the branch under the Newton iteration, per site class, from the spectral set
//                      exactly as lfuntdd forms them (no expm1, root 0 skipped in the derivatives).
//   branch_kernel      per pattern f, f', f'' summed over site classes (with the per-class scale bookkeeping of
//                      lfuntdd_SiteClass), then -sum fpatt log f etc.; deterministic block partials + cond_sum_kernel.
// These are streaming kernels on the FP64 vector pipe (2 n^2 flops per pattern and son against 16 n bytes): at n = 61
// compute-bound, at n = 4 HBM-bound.  A thread owns one pattern and 8 consecutive states; the matrices sit TRANSPOSED in
// shared memory so that those 8 entries of a column are four 128-bit broadcast loads per 8 FMAs (the first version read
// two shared operands per multiply-add and was bound by the shared-memory pipe: ncu, profiles/).  No DMMA yet.
#pragma once
#include <cuda_runtime.h>

namespace pb200 {

constexpr int COND_TP = 32;          // patterns per block (one per lane)
constexpr int COND_THREADS = 256;    // 8 warps: warp w owns the states / rows 8w .. 8w+7
constexpr int COND_TMAX = 8;         // most tiles per block (COND_TMAX * COND_TP = COND_THREADS patterns)
constexpr int COND_LDP = 64;         // leading dimension of the transposed matrices in shared memory (n <= 64, zero padded)

struct CondNodeParams {
    int n, nclass, nnode, ns, ncodes, nson, node, scale_k, nscale;
    int tiles;                       // tiles of COND_TP patterns per block: a son's matrix is staged once for all of them
    long long npatt, zstride;
    const int* sons;                 // [nson] node ids (device)
    const double* P;                 // [nclass][nnode][n*n] P(t) of every branch (by child node), COLUMN-major: [k*n + j] = P[j][k]
    const unsigned char* z;          // [ns][zstride] compact codes
    const unsigned char* amb_n;      // [ncodes - n]
    const unsigned char* amb_map;    // [ncodes - n][64]
    double* cond;                    // [nclass][nnode - ns][npatt][n]
    double* scaleF;                  // [nclass][nscale][npatt]
};

__global__ void __launch_bounds__(COND_THREADS) cond_node_kernel(const CondNodeParams p)
{
    extern __shared__ __align__(16) double csm[];
    const int n = p.n, LDT = n | 1, T = p.tiles;
    double* Pt = csm;                          // [n][COND_LDP] the son's P(t) TRANSPOSED: Pt[k][j] = P[j][k], columns >= n zero
    double* tin = Pt + n * COND_LDP;           // [TP][LDT] a son's vectors, one tile
    double* prod = tin + COND_TP * LDT;        // [T][TP][LDT] this node's vectors (running products over the sons)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, cls = blockIdx.y;
    const long long hb = (long long)blockIdx.x * T * COND_TP;        // the block's first pattern: it owns T tiles of TP patterns
    const size_t nint = (size_t)(p.nnode - p.ns);
    double* cbase = p.cond + (size_t)cls * nint * p.npatt * n;
    const int j0 = 8 * w;                      // warp w owns the parent states 8w .. 8w+7 (one pattern per lane)

    for (int s = 0; s < p.nson; s++) {
        const int son = p.sons[s];
        const double* Pm = p.P + ((size_t)cls * p.nnode + son) * n * n;     // column-major: Pm[k*n + j] = P[j][k]
        __syncthreads();                                        // the previous son's matrix fully consumed
        for (int e = tid; e < n * COND_LDP; e += COND_THREADS) { const int k = e / COND_LDP, j = e - k * COND_LDP; Pt[e] = (j < n) ? Pm[k * n + j] : 0.0; }
        if (son < p.ns) __syncthreads();                        // (an internal son synchronises in the tile loop)
        for (int tt = 0; tt < T; tt++) {                        // the matrix stays in shared memory for T tiles of patterns
            const long long h0 = hb + (long long)tt * COND_TP;
            if (h0 >= p.npatt) break;
            const int nh = (int)min((long long)COND_TP, p.npatt - h0);
            if (son >= p.ns) {
                __syncthreads();                                // the previous tile's vectors fully consumed
                const double* src = cbase + ((size_t)(son - p.ns) * p.npatt + h0) * n;
                for (int e = tid; e < nh * n; e += COND_THREADS) { const int hh = e / n, k = e - hh * n; tin[hh * LDT + k] = src[e]; }
                __syncthreads();
            }
            if (lane < nh) {
                double t[8];
                if (son < p.ns) {                               // tip: src/codeml.c:3555-3567
                    const int code = p.z[(size_t)son * p.zstride + h0 + lane];
                    if (code < n) {
#pragma unroll
                        for (int jj = 0; jj < 8; jj++) t[jj] = Pt[code * COND_LDP + j0 + jj];
                    }
                    else {
                        const unsigned char* mp = p.amb_map + (size_t)(code - n) * 64;
                        const int cnt = p.amb_n[code - n];
#pragma unroll
                        for (int jj = 0; jj < 8; jj++) t[jj] = 0.0;
                        for (int k = 0; k < cnt; k++) {
                            const double* q = Pt + mp[k] * COND_LDP + j0;
#pragma unroll
                            for (int jj = 0; jj < 8; jj++) t[jj] += q[jj];
                        }
                    }
                }
                else {                                          // internal son: src/codeml.c:3568-3575
                    const double* v = tin + lane * LDT;
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) t[jj] = 0.0;
                    for (int k = 0; k < n; k++) {               // one per-lane load + four 128-bit broadcast loads per 8 FMAs
                        const double vk = v[k];
                        const double2* q = reinterpret_cast<const double2*>(Pt + k * COND_LDP + j0);
                        const double2 a = q[0], b = q[1], c = q[2], d = q[3];
                        t[0] = fma(a.x, vk, t[0]); t[1] = fma(a.y, vk, t[1]); t[2] = fma(b.x, vk, t[2]); t[3] = fma(b.y, vk, t[3]);
                        t[4] = fma(c.x, vk, t[4]); t[5] = fma(c.y, vk, t[5]); t[6] = fma(d.x, vk, t[6]); t[7] = fma(d.y, vk, t[7]);
                    }
                }
                double* o = prod + ((size_t)tt * COND_TP + lane) * LDT + j0;      // this thread's own 8 entries: no race
#pragma unroll
                for (int jj = 0; jj < 8; jj++)
                    if (j0 + jj < n) o[jj] = (s == 0) ? t[jj] : o[jj] * t[jj];   // 1 * first son (src/codeml.c:3539-3540)
            }
        }
    }
    __syncthreads();
    for (int tt = 0; tt < T; tt++) {
        const long long h0 = hb + (long long)tt * COND_TP;
        if (h0 >= p.npatt) break;
        const int nh = (int)min((long long)COND_TP, p.npatt - h0);
        double* tout = prod + (size_t)tt * COND_TP * LDT;
        if (p.scale_k >= 0) {                                   // NodeScale, src/treesub.c:7214-7228
            if (tid < nh) {
                double* v = tout + tid * LDT;
                double t = 0.0;
                for (int j = 0; j < n; j++) if (v[j] > t) t = v[j];
                double f;
                if (t < 1e-300) { for (int j = 0; j < n; j++) v[j] = 1.0; f = -800.0; }
                else { for (int j = 0; j < n; j++) v[j] = v[j] / t; f = log(t); }
                p.scaleF[((size_t)cls * p.nscale + p.scale_k) * p.npatt + h0 + tid] = f;
            }
            __syncthreads();
        }
        double* dst = cbase + ((size_t)(p.node - p.ns) * p.npatt + h0) * n;
        for (int e = tid; e < nh * n; e += COND_THREADS) { const int hh = e / n, k = e - hh * n; dst[e] = tout[hh * LDT + k]; }
    }
}

// ---- the Newton iteration of one branch, kept on the device (pb200_branch_newton) ----
// State of the reference's "iterating a branch" loop (src/treesub.c:8087-8108): the three kernels of one evaluation read what
// to evaluate from it (t_eval, want_dd, a, b), newton_step_kernel sums their partials and advances it.
struct NewtonState {
    double t_eval;             // branch length of the next evaluation
    double t, t0, l, l0, dl, ddl, p, step;
    double e, tb0, tb1;        // convergence threshold, bounds of t
    int want_dd;               // next evaluation: 1 = lfuntdd (value + derivatives), 0 = lfunt (value)
    int icycleb, first_cycle, done, nevals, max_evals;
    int a, b;                  // the branch b -> a
};

// ---- P, P', P'' of the branch being optimised ----
struct Pt3Params {
    int n, kind, nr, nclass;
    long long set_stride;
    const double* sets;       // resident spectral sets
    const int* set_of_class;  // [nclass]
    const double* mult;       // [nclass] rgene * _rateSite * Qfactor (src/treesub.c:8435-8439)
    double t;
    const NewtonState* st;    // device-side iteration: t is st->t_eval
    double* out;              // [nclass][3][n*n], each matrix COLUMN-major: [j*n + i] = P[i][j]
};

__global__ void __launch_bounds__(256) pt3_kernel(const Pt3Params p)
{
    __shared__ double ex[64], mu[64];
    const int n = p.n, cls = blockIdx.x, tid = threadIdx.x;
    const double* set = p.sets + (size_t)p.set_of_class[cls] * p.set_stride;
    const int nroot = (p.kind == KIND_UVROOT) ? n : p.nr;
    const double* Root = set;
    const double tt = p.st ? p.st->t_eval : p.t;
    if (tid < nroot) {
        const double m = __dmul_rn(p.mult[cls], Root[tid]);     // "multiply", src/treesub.c:8266
        mu[tid] = m;
        ex[tid] = tid ? exp(__dmul_rn(tt, m)) : 1.0;            // expt = 1 for the first root (src/treesub.c:8265-8267)
    }
    __syncthreads();
    double* o = p.out + (size_t)cls * 3 * n * n;
    for (int e = tid; e < n * n; e += 256) {
        const int i = e / n, j = e - i * n;
        double P = 0.0, dP = 0.0, ddP = 0.0;
        if (p.kind == KIND_UVROOT) {
            const double *U = set + n, *V = set + n + (size_t)n * n;
            for (int k = 0; k < nroot; k++) {
                const double uv = __dmul_rn(__dmul_rn(U[i * n + k], ex[k]), V[k * n + j]);   // uexpt * V[k][j]
                P = __dadd_rn(P, uv);
                if (k) {
                    const double d1 = __dmul_rn(uv, mu[k]);
                    dP = __dadd_rn(dP, d1);
                    ddP = __dadd_rn(ddP, __dmul_rn(d1, mu[k]));
                }
            }
        }
        else {                                                  // Cijk form, src/treesub.c:8280-8287
            const double* C = set + p.nr + ((size_t)i * n + j) * p.nr;
            for (int k = 0; k < nroot; k++) {
                const double ce = __dmul_rn(C[k], ex[k]);
                P = __dadd_rn(P, ce);
                if (k) {
                    const double d1 = __dmul_rn(ce, mu[k]);
                    dP = __dadd_rn(dP, d1);
                    ddP = __dadd_rn(ddP, __dmul_rn(d1, mu[k]));
                }
            }
        }
        o[j * n + i] = P; o[n * n + j * n + i] = dP; o[2 * n * n + j * n + i] = ddP;
    }
}

// ---- f, f', f'' per pattern, summed over classes; -lnL and its two derivatives in t ----
enum { BR_SEM_LFUNT = 0, BR_SEM_SITECLASS = 1 };

struct BranchParams {
    int n, nclass, nnode, ns, ncodes, a, b, nscale, semantics, want_dd;
    int tiles;                 // tiles of COND_TP patterns per block (<= COND_TMAX): P, P', P'' are staged once per class for all of them
    long long npatt, zstride;
    const double* P3;          // [nclass][3][n*n]
    const double* cond;        // [nclass][nint][npatt][n]
    const double* scaleF;      // [nclass][nscale][npatt]
    const double* freqK;       // [nclass]
    const double* pi;          // [n]
    const double* fpatt;       // [npatt]
    const unsigned char* z;
    const unsigned char* amb_n;
    const unsigned char* amb_map;
    double* partial;           // [gridDim.x][3]
    const NewtonState* st;     // device-side iteration: a, b, want_dd come from the state
};

__global__ void __launch_bounds__(COND_THREADS) branch_kernel(const BranchParams p0)
{
    extern __shared__ __align__(16) double bsm[];
    BranchParams p = p0;
    if (p.st) { p.a = p.st->a; p.b = p.st->b; p.want_dd = p.st->want_dd; }
    const int n = p.n, LDT = n | 1, nm = p.want_dd ? 3 : 1, T = p.tiles;
    double* P3 = bsm;                           // [3][n][COND_LDP] TRANSPOSED: P3[m][j][i] = (d^m P)[i][j], rows i >= n zero
    double* tA = P3 + 3 * n * COND_LDP;         // [TP][LDT]
    double* tB = tA + COND_TP * LDT;            // [TP][LDT]
    double* red = tB + COND_TP * LDT;           // [3][8][TP]   (reused as [3][256] for the final block sum)
    double* Sh = red + 3 * 8 * COND_TP;         // [COND_TMAX][TP]
    double* facc = Sh + COND_TMAX * COND_TP;    // [3][COND_TMAX][TP] f, f', f'' of the block's patterns, summed over the classes
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long hb = (long long)blockIdx.x * T * COND_TP;      // the block owns T tiles of TP patterns
    const size_t nint = (size_t)(p.nnode - p.ns);
    const bool btip = p.b < p.ns;
    const int i0 = 8 * w;                       // warp w owns the rows (states of b) 8w .. 8w+7

    // scale bookkeeping of lfuntdd_SiteClass (src/treesub.c:8437-8449): Sh = largest per-class sum of scale factors
    if (tid < T * COND_TP) {
        double best = 0.0;
        const long long h = hb + tid;
        if (p.nscale > 0 && h < p.npatt) {
            for (int ir = 0; ir < p.nclass; ir++) {
                double y = 0.0;
                for (int k = 0; k < p.nscale; k++) y += p.scaleF[((size_t)ir * p.nscale + k) * p.npatt + h];
                if (ir == 0 || y > best) best = y;
            }
        }
        Sh[tid] = best;
        facc[tid] = 0.0; facc[COND_TMAX * COND_TP + tid] = 0.0; facc[2 * COND_TMAX * COND_TP + tid] = 0.0;
    }
    for (int cls = 0; cls < p.nclass; cls++) {
        __syncthreads();                                        // the previous class's matrices fully consumed
        const double* Pm = p.P3 + (size_t)cls * 3 * n * n;
        for (int e = tid; e < nm * n * COND_LDP; e += COND_THREADS) {
            const int m = e / (n * COND_LDP), r = e - m * n * COND_LDP, j = r / COND_LDP, i = r - j * COND_LDP;
            P3[e] = (i < n) ? Pm[(size_t)m * n * n + j * n + i] : 0.0;
        }
        const double* cb = p.cond + (size_t)cls * nint * p.npatt * n;
        for (int tt = 0; tt < T; tt++) {                        // the three matrices stay in shared memory for T tiles
            const long long h0 = hb + (long long)tt * COND_TP, h = h0 + lane;
            if (h0 >= p.npatt) break;
            const int nh = (int)min((long long)COND_TP, p.npatt - h0);
            __syncthreads();                                    // previous tile's vectors and partial sums fully consumed
            {
                const double* src = cb + ((size_t)(p.a - p.ns) * p.npatt + h0) * n;
                for (int e = tid; e < nh * n; e += COND_THREADS) { const int hh = e / n, k = e - hh * n; tA[hh * LDT + k] = src[e]; }
            }
            if (!btip) {
                const double* src = cb + ((size_t)(p.b - p.ns) * p.npatt + h0) * n;
                for (int e = tid; e < nh * n; e += COND_THREADS) { const int hh = e / n, k = e - hh * n; tB[hh * LDT + k] = src[e]; }
            }
            __syncthreads();
            double f = 0.0, df = 0.0, ddf = 0.0;
            if (lane < nh) {
                double pK = 1.0;
                if (p.semantics == BR_SEM_SITECLASS) {
                    pK = p.freqK[cls];
                    if (p.nscale > 0) {
                        double y = 0.0;
                        for (int k = 0; k < p.nscale; k++) y += p.scaleF[((size_t)cls * p.nscale + k) * p.npatt + h];
                        pK = pK * exp(y - Sh[tt * COND_TP + lane]);
                    }
                }
                const double* va = tA + lane * LDT;
                if (btip) {                                     // b is a tip: the states its code allows (src/treesub.c:8251-8256)
                    const int code = p.z[(size_t)p.b * p.zstride + h];
                    int n1 = 1;
                    const unsigned char* mp = nullptr;
                    if (code >= n) { n1 = p.amb_n[code - n]; mp = p.amb_map + (size_t)(code - n) * 64; }
                    for (int ii = w; ii < n1; ii += 8) {
                        const int xb = (code < n) ? code : mp[ii];
                        const double piqi = pK * p.pi[xb];
                        double pq = 0.0, dpq = 0.0, ddpq = 0.0;
                        for (int j = 0; j < n; j++) {
                            const double aj = va[j];
                            pq = fma(P3[j * COND_LDP + xb], aj, pq);
                            if (p.want_dd) {
                                dpq = fma(P3[(n + j) * COND_LDP + xb], aj, dpq);
                                ddpq = fma(P3[(2 * n + j) * COND_LDP + xb], aj, ddpq);
                            }
                        }
                        f += piqi * pq; df += piqi * dpq; ddf += piqi * ddpq;
                    }
                }
                else {                                          // b internal: all n rows, 8 per warp
                    double pq[8], dpq[8], ddpq[8];
#pragma unroll
                    for (int ii = 0; ii < 8; ii++) { pq[ii] = 0.0; dpq[ii] = 0.0; ddpq[ii] = 0.0; }
                    for (int j = 0; j < n; j++) {               // one per-lane load + 128-bit broadcast loads per 8 (24) FMAs
                        const double aj = va[j];
                        const double2* q = reinterpret_cast<const double2*>(P3 + j * COND_LDP + i0);
#pragma unroll
                        for (int u = 0; u < 4; u++) { const double2 x = q[u]; pq[2 * u] = fma(x.x, aj, pq[2 * u]); pq[2 * u + 1] = fma(x.y, aj, pq[2 * u + 1]); }
                        if (p.want_dd) {
                            const double2* q1 = reinterpret_cast<const double2*>(P3 + (n + j) * COND_LDP + i0);
                            const double2* q2 = reinterpret_cast<const double2*>(P3 + (2 * n + j) * COND_LDP + i0);
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const double2 x = q1[u], y = q2[u];
                                dpq[2 * u] = fma(x.x, aj, dpq[2 * u]); dpq[2 * u + 1] = fma(x.y, aj, dpq[2 * u + 1]);
                                ddpq[2 * u] = fma(y.x, aj, ddpq[2 * u]); ddpq[2 * u + 1] = fma(y.y, aj, ddpq[2 * u + 1]);
                            }
                        }
                    }
#pragma unroll
                    for (int ii = 0; ii < 8; ii++) {
                        const int i = i0 + ii;
                        if (i < n) {
                            const double piqi = pK * p.pi[i] * tB[lane * LDT + i];
                            f += piqi * pq[ii]; df += piqi * dpq[ii]; ddf += piqi * ddpq[ii];
                        }
                    }
                }
            }
            red[(0 * 8 + w) * COND_TP + lane] = f;
            red[(1 * 8 + w) * COND_TP + lane] = df;
            red[(2 * 8 + w) * COND_TP + lane] = ddf;
            __syncthreads();
            if (w == 0) {                                       // fixed order over the 8 warps, then over the classes
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int ww = 0; ww < 8; ww++) {
                    a0 += red[(0 * 8 + ww) * COND_TP + lane];
                    a1 += red[(1 * 8 + ww) * COND_TP + lane];
                    a2 += red[(2 * 8 + ww) * COND_TP + lane];
                }
                facc[tt * COND_TP + lane] += a0;
                facc[(COND_TMAX + tt) * COND_TP + lane] += a1;
                facc[(2 * COND_TMAX + tt) * COND_TP + lane] += a2;
            }
        }
    }
    __syncthreads();
    // per pattern -lnL and its derivatives (src/treesub.c:8531-8540), then a fixed-order block sum
    double l = 0.0, dl = 0.0, ddl = 0.0;
    if (tid < T * COND_TP) {
        const long long h = hb + tid;
        if (h < p.npatt) {
            const double fp = p.fpatt[h];
            if (fp != 0.0) {
                const double fh = facc[tid], dfh = facc[COND_TMAX * COND_TP + tid], ddfh = facc[2 * COND_TMAX * COND_TP + tid];
                const double lf = (p.semantics == BR_SEM_LFUNT && !p.want_dd && fh < 0) ? -500.0 : log(fh);   // src/treesub.c:8192-8193
                l = -(lf * fp);
                if (p.nscale > 0) l -= Sh[tid] * fp;
                if (p.want_dd) {
                    dl = -(dfh / fh * fp);
                    ddl = -((fh * ddfh - dfh * dfh) / (fh * fh) * fp);
                }
            }
        }
    }
    red[tid] = l; red[256 + tid] = dl; red[512 + tid] = ddl;    // red has 3 * 8 * TP = 768 doubles
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { red[tid] += red[tid + o]; red[256 + tid] += red[256 + tid + o]; red[512 + tid] += red[512 + tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        p.partial[(size_t)blockIdx.x * 3 + 0] = red[0];
        p.partial[(size_t)blockIdx.x * 3 + 1] = red[256];
        p.partial[(size_t)blockIdx.x * 3 + 2] = red[512];
    }
}

// fixed-order sum of the block partials: out[3]
__global__ void __launch_bounds__(256) cond_sum_kernel(const double* partial, int nblk, double* out)
{
    __shared__ double s[3][256];
    const int tid = threadIdx.x;
    double a[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < nblk; i += 256)
        for (int m = 0; m < 3; m++) a[m] += partial[(size_t)i * 3 + m];
    for (int m = 0; m < 3; m++) s[m][tid] = a[m];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) for (int m = 0; m < 3; m++) s[m][tid] += s[m][tid + o];
        __syncthreads();
    }
    if (tid < 3) out[tid] = s[tid][0];
}

// Sum of the block / warp partials in a fixed order, then one step of the reference's branch iteration (src/treesub.c:
// 8087-8108): after a lfuntdd evaluation the Newton direction and the first trial step, after a lfunt evaluation accept /
// quarter the step / give up, the convergence test, the next round.  Sets the graph's loop condition to 0 when the branch is done.
__global__ void __launch_bounds__(256) newton_step_kernel(const double* partial, int nblk, NewtonState* st, cudaGraphConditionalHandle handle)
{
    __shared__ double s[3][256];
    const int tid = threadIdx.x;
    double a[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < nblk; i += 256)
        for (int m = 0; m < 3; m++) a[m] += partial[(size_t)i * 3 + m];
    for (int m = 0; m < 3; m++) s[m][tid] = a[m];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) for (int m = 0; m < 3; m++) s[m][tid] += s[m][tid + o];
        __syncthreads();
    }
    if (tid != 0) return;
    const double smallv = 1e-20;
    bool try_step = false, done = false;
    st->nevals++;
    if (st->want_dd) {                                          // src/treesub.c:8088-8100
        st->dl = s[1][0]; st->ddl = s[2][0];
        const double p = -st->dl / fabs(st->ddl);
        double step;
        if (fabs(p) < smallv) step = 0;
        else if (p < 0)       step = fmin(1.0, (st->tb0 - st->t0) / p);
        else                  step = fmin(1.0, (st->tb1 - st->t0) / p);
        if (st->first_cycle && step != 1 && step != 0) step *= 0.99;      // avoid border
        st->p = p; st->step = step;
        try_step = true;
    }
    else {                                                      // one trial of the step-halving loop, src/treesub.c:8101-8106
        st->l = s[0][0];
        if (st->l < st->l0) {
            if (fabs(st->t - st->t0) < st->e * fabs(1 + st->t) && fabs(st->l - st->l0) < st->e) done = true;      // :8108
            else {
                st->t0 = st->t; st->l0 = st->l;
                if (++st->icycleb < 10) { st->t_eval = st->t; st->want_dd = 1; }
                else done = true;
            }
        }
        else { st->step /= 4; try_step = true; }
    }
    if (try_step) {
        // (unfused multiply-add: the host compiler of the reference does not contract t0 + step * p either)
        if (st->step > smallv) { st->t = __dadd_rn(st->t0, __dmul_rn(st->step, st->p)); st->t_eval = st->t; st->want_dd = 0; }
        else { st->t = st->t0; st->l = st->l0; done = true; }   // :8107
    }
    if (st->nevals >= st->max_evals) done = true;
    if (done) { st->done = 1; cudaGraphSetConditional(handle, 0); }
}

}  // namespace pb200
