// kernels.cuh -- hand-written sm_100a kernels of the PAML tree-likelihood hot path.
//
//   pt_kernel      batched P(t) for every (site class, branch): replaces PMatUVRoot (src/tools.c:516-546),
//                  PMatCijk (src/baseml.c:1572-1589), PMatK80 (src/tools.c:578-606), PMatJC69like
//                  (src/codeml.c:3585-3594) and the Taylor + squaring matexp of the UNREST models (src/tools.c:4879-4919)
//                  as selected by GetPMatBranch (src/treesub.c:7503-7592).
//   prune_kernel   Felsenstein pruning of whole trees, replaces ConditionalPNode (src/codeml.c:3526-3582,
//                  src/baseml.c:1517-1570) + NodeScale (src/treesub.c:7200-7230) + the root pi-dot of
//                  lfun / fx_r (src/treesub.c:7780-7798, 7726-7750).
//   mix_kernel     class mixture, weighting and deterministic sum: the h-loops of lfun
//                  (src/treesub.c:7778-7804) and lfundG (src/treesub.c:7630-7657).
//
// All arithmetic is IEEE double.  The state x state contraction at internal nodes runs on the FP64 tensor
// path (DMMA, mma.sync.m8n8k4.f64 -- the only FP64 MMA shape sm_100a executes natively; tcgen05 has no
// FP64 kind).  See DESIGN.md for layouts and rooflines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "plan.h"

namespace pb200 {

enum { KIND_UVROOT = 0, KIND_CIJK = 1, KIND_K80 = 2, KIND_JC69LIKE = 3, KIND_QMATEXP = 4 };
enum { SEM_LFUN = 0, SEM_LFUNDG = 1 };
enum { PT_MODE_AUTO = -1, PT_MODE_TIP = 0, PT_MODE_FRAG = 1, PT_MODE_PLAIN = 2, PT_MODE_NUC4 = 3, PT_MODE_PLAIN_T = 4 };

// Physical <-> logical state order inside a conditional vector.  The DMMA C fragment gives lane (g,q) the
// physical columns 8nt+2q+e (e = 0,1); the A fragment of k-step ks wants column 4ks+q in that lane.  Storing
// logical state j = 8nt + 4e + q at physical column 8nt + 2q + e makes the two coincide: the accumulators of
// one DMMA are, register for register, the A operands of the next (k-step 2nt+e), with no shuffle and no
// shared-memory round trip.  P(t) tiles, tip rows, slots and pi all use the physical order.
__host__ __device__ __forceinline__ int phys_to_logical(int p) { return (p & ~7) | ((p & 1) << 2) | ((p >> 1) & 3); }

// ------------------------------------------------------------------------------------------------
// P(t)
// ------------------------------------------------------------------------------------------------
struct PtParams {
    int n;              // states
    int ld;             // n rounded up to a multiple of 4 (= 4*KS): leading dimension in shared memory
    int NT, KS;         // n-tiles (8 output states) / k-steps (4 input states) of the pruning kernel
    int ncodes;         // n + number of ambiguity codes in use (rows of a tip matrix)
    int ns, nnode;      // matrix m belongs to node m % nnode; tips are nodes < ns
    int kind, nr;
    long long set_stride;
    const double* sets; // [nset][set_stride]
    const int* mat_set; // [nmat] (-1 = skip)
    const double* mat_t;// [nmat]
    const unsigned char* amb_n;    // [ncodes - n]
    const unsigned char* amb_map;  // [ncodes - n][64]
    double* out;        // [nmat][pstride]
    long long pstride;
    int mode;           // PT_MODE_*
    int logical;        // 1: tip rows / B fragments in LOGICAL state order (cond_dmma.cuh) instead of the physical order of prune_kernel
};

// grid = (matrices, row blocks): CTA (m, rb) computes rows [16rb, 16rb+16) of matrix m with 256 threads
// (thread = one row x four columns) and writes the part of the consumer's layout that those rows feed.
// Shared memory: ld*ld (V) + 16*ld (scaled U rows) + ld (expm1 terms) doubles for UVROOT, 16*ld + 3*ld*ld for
// QMATEXP, 16*ld otherwise.
__global__ void __launch_bounds__(256) pt_kernel(const PtParams p)
{
    extern __shared__ __align__(16) double sm[];
    const int m = blockIdx.x, rb = blockIdx.y, tid = threadIdx.x, n = p.n, ld = p.ld;
    const int iset = p.mat_set[m];
    if (iset < 0) return;                                   // the root has no branch
    const int i0 = 16 * rb;
    if (i0 >= ld) return;
    const double t = p.mat_t[m];
    const double* set = p.sets + (size_t)iset * p.set_stride;
    double* Prow = sm;                                      // [16][ld] rows i0.. of the result, P[i][j]
    const int ty = tid >> 4, tx = tid & 15;                 // row i0+ty, columns tx, tx+16, tx+32, tx+48

    if (p.kind == KIND_UVROOT && !(t < 1e-100)) {
        // P = I + sum_k U[:,k] expm1(t Root[k]) V[k,:]   (src/tools.c:529-534)
        const double *Root = set, *U = set + n, *V = set + n + (size_t)n * n;
        double* Ue = sm;                 // [16][ld]   rows i0.. of U (scaled by expm1 on the fly; becomes Prow)
        double* Vs = sm + 16 * ld;       // [ld][ld]
        double* ex = Vs + ld * ld;       // [ld]
        // one pass of independent global loads (V, this block's rows of U, the expm1 terms), one barrier.  All of a thread's
        // loads are issued before the first is stored (up to 16 + 4 in flight): on small data this kernel is pure latency, and a
        // load-store-load-store loop pays the memory round trip sixteen times (14 -> 8 us for the 13 matrices of config 2).
        {
            double vv[16], uu[4];
            const double root_k = (tid < n) ? Root[tid] : 0.0;
#pragma unroll
            for (int it = 0; it < 16; it++) {
                const int e = tid + 256 * it, k = e / ld, j = e - k * ld;
                vv[it] = (e < ld * ld && k < n && j < n) ? V[k * n + j] : 0.0;
            }
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int e = tid + 256 * it, r = e / ld, k = e - r * ld, i = i0 + r;
                uu[it] = (e < 16 * ld && i < n && k < n) ? U[i * n + k] : 0.0;
            }
            if (tid < ld) ex[tid] = (tid < n) ? expm1(t * root_k) : 0.0;
#pragma unroll
            for (int it = 0; it < 16; it++) { const int e = tid + 256 * it; if (e < ld * ld) Vs[e] = vv[it]; }
#pragma unroll
            for (int it = 0; it < 4; it++) { const int e = tid + 256 * it; if (e < 16 * ld) Ue[e] = uu[it]; }
        }
        __syncthreads();
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const double* urow = Ue + ty * ld;
#pragma unroll 4
        for (int k = 0; k < n; k++) {
            const double u = urow[k] * ex[k];
#pragma unroll
            for (int c = 0; c < 4; c++) { int j = tx + 16 * c; if (j < ld) acc[c] = fma(u, Vs[k * ld + j], acc[c]); }
        }
        __syncthreads();                 // all reads of Ue done before it is overwritten as Prow
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int i = i0 + ty, j = tx + 16 * c;
            if (j < ld) {
                double v = acc[c] + (i == j ? 1.0 : 0.0);
                if (v < 0) v = 0;                           // src/tools.c:536-537
                Prow[ty * ld + j] = (i < n && j < n) ? v : 0.0;
            }
        }
    }
    else if (p.kind == KIND_QMATEXP) {
        // UNREST models (src/treesub.c:7523-7527): P = matexp(Q t, 7 Taylor terms, 5 squarings), src/tools.c:4879-4919:
        // B = Q t / 2^5, S = I + B + B^2/2! + ... + B^7/7!, P = S^(2^5).  One thread per element (n <= 16, one row
        // block), products and sums in the reference's order and unfused; no clamping on this path.
        double* B = sm + 16 * ld;        // [n][n]
        double* S = B + ld * ld;
        double* pw = S + ld * ld;        // B^(term-1)
        const int i = tid / n, j = tid - i * n;
        const bool on = tid < n * n;
        double sv = 0.0, factor = 1.0;
        if (on) {
            const double b = __dmul_rn(__dmul_rn(set[tid], t), 1.0 / 32);
            B[tid] = b; pw[tid] = b;
            sv = (i == j) ? __dadd_rn(b, 1.0) : b;
        }
        __syncthreads();
        for (int term = 2; term <= 7; term++) {
            double s = 0.0;
            if (on) for (int k = 0; k < n; k++) s = __dadd_rn(s, __dmul_rn(pw[i * n + k], B[k * n + j]));
            __syncthreads();             // every read of pw done before it is replaced
            factor /= term;
            if (on) { pw[tid] = s; sv = __dadd_rn(sv, __dmul_rn(s, factor)); }
            __syncthreads();
        }
        if (on) S[tid] = sv;
        __syncthreads();
        for (int sq = 0; sq < 5; sq++) {
            double s = 0.0;
            if (on) for (int k = 0; k < n; k++) s = __dadd_rn(s, __dmul_rn(S[i * n + k], S[k * n + j]));
            __syncthreads();
            if (on) S[tid] = s;
            __syncthreads();
        }
        for (int e = tid; e < 16 * ld; e += 256) {
            const int r = e / ld, c = e - r * ld;
            Prow[e] = (r < n && c < n) ? S[r * n + c] : 0.0;
        }
    }
    else {
        double e1 = 0, e2 = 0, ex[8];
        if (p.kind == KIND_K80) {                           // src/tools.c:588-603
            double kappa = set[0];
            e1 = expm1(-4 * t / (kappa + 2));
            e2 = expm1(-2 * t * (kappa + 1) / (kappa + 2));
        }
        else if (p.kind == KIND_CIJK) {                     // src/baseml.c:1580
            ex[0] = 0;
            for (int k = 1; k < p.nr && k < 8; k++) ex[k] = expm1(t * set[k]);
        }
        else if (p.kind == KIND_JC69LIKE) {                 // src/codeml.c:3588-3589
            e1 = 1. / n + (1. - 1. / n) * exp(-n / (n - 1.) * t);
            e2 = (1. - e1) / (n - 1.);
        }
        for (int e = tid; e < 16 * ld; e += 256) {
            int r = e / ld, j = e - r * ld, i = i0 + r;
            double v = 0;
            if (i < n && j < n) {
                if (p.kind == KIND_UVROOT) v = (i == j);    // t < 1e-100 -> identity (src/tools.c:525-528)
                else if (p.kind == KIND_K80) {
                    double kappa = set[0];
                    if (fabs(kappa - 1) < 1e-20) v = (i == j ? 1. + 3 / 4.0 * e1 : -e1 / 4);
                    else if (i == j)             v = 1 + (e1 + 2 * e2) / 4;
                    else if ((i ^ j) == 1)       v = (e1 - 2 * e2) / 4;
                    else                         v = -e1 / 4;
                }
                else if (p.kind == KIND_CIJK) {
                    const double* C = set + p.nr + ((size_t)i * n + j) * p.nr;
                    for (int k = 0; k < p.nr; k++) v += C[k] * ex[k];
                    if (i == j) v += 1;
                }
                else v = (i == j ? e1 : e2);
            }
            Prow[e] = v;
        }
    }
    __syncthreads();

    // write this row block's share of the matrix in the layout its consumer wants (Prow[r][k] = P[i0+r][k])
    double* out = p.out + (size_t)m * p.pstride;
    const int node = m % p.nnode;
    int mode = (p.mode != PT_MODE_AUTO) ? p.mode : (node < p.ns ? PT_MODE_TIP : PT_MODE_FRAG);
    const int NP = 8 * p.NT;
    if (mode == PT_MODE_NUC4) {
        // 4-state register kernel (prune4.cuh): internal branch = plain row-major P, tip branch = PT[code][j]
        if (node >= p.ns) mode = PT_MODE_PLAIN;
        else {
            for (int e = tid; e < p.ncodes * 4; e += 256) {
                int code = e >> 2, j = e & 3;
                double v = 0;
                if (code < n) v = Prow[j * ld + code];
                else {
                    const unsigned char* mp = p.amb_map + (size_t)(code - n) * 64;
                    int cnt = p.amb_n[code - n];
                    for (int k = 0; k < cnt; k++) v += Prow[j * ld + mp[k]];
                }
                out[e] = v;
            }
            return;
        }
    }
    if (mode == PT_MODE_PLAIN_T) {                           // column-major (cond.cuh): out[k*n + j] = P[j][k]
        for (int e = tid; e < 16 * n; e += 256) { int r = e / n, k = e - r * n; if (i0 + r < n) out[(size_t)k * n + (i0 + r)] = Prow[r * ld + k]; }
    }
    else if (mode == PT_MODE_PLAIN) {
        for (int e = tid; e < 16 * n; e += 256) { int r = e / n, j = e - r * n; if (i0 + r < n) out[(size_t)(i0 + r) * n + j] = Prow[r * ld + j]; }
    }
    else if (mode == PT_MODE_TIP) {
        // PT[code][phys(j)] = sum_{k in code} P[j][k]: what a tip in state `code` contributes to parent state j
        // (src/codeml.c:3555-3567); padded parent states j >= n get 0.  This block owns the parent states i0..i0+15.
        for (int e = tid; e < p.ncodes * 16; e += 256) {
            int code = e >> 4, r = e & 15, j = i0 + r;
            if (j >= NP) continue;
            double v = 0;
            if (j < n) {
                if (code < n) v = Prow[r * ld + code];
                else {
                    const unsigned char* mp = p.amb_map + (size_t)(code - n) * 64;
                    int cnt = p.amb_n[code - n];
                    for (int k = 0; k < cnt; k++) v += Prow[r * ld + mp[k]];
                }
            }
            // physical position of logical state j (inverse of phys_to_logical)
            int ph = p.logical ? j : ((j & ~7) | ((j & 3) << 1) | ((j >> 2) & 1));
            out[(size_t)code * NP + ph] = v;
        }
        if (rb == 0)                                         // padded parent states beyond ld (none when NP == ld)
            for (int e = tid; e < p.ncodes * NP; e += 256) { int ph = e % NP; if ((p.logical ? ph : phys_to_logical(ph)) >= ld) out[e] = 0.0; }
    }
    else {
        // DMMA B-fragment order, two n-tiles per 16-byte element so that one 128-bit load feeds two DMMAs:
        // element (ks, pr, lane, e) = P[j][4ks + lane%4], j = logical state of physical column 8*(2pr+e) + lane/4
        const int NTP = (p.NT + 1) / 2;
        const int total = p.KS * NTP * 64;
        for (int e0 = tid; e0 < total; e0 += 256) {
            int e = e0 & 1, lane = (e0 >> 1) & 31, f = e0 >> 6, pr = f % NTP, ks = f / NTP;
            int nt = 2 * pr + e;
            int j = p.logical ? (8 * nt + (lane >> 2)) : phys_to_logical(8 * nt + (lane >> 2)), k = 4 * ks + (lane & 3);
            bool mine = (j >= i0 && j < i0 + 16);
            bool pad = (nt >= p.NT || j >= ld);               // entries no row block owns: zero them from block 0
            if (mine) out[e0] = (nt < p.NT && j < n && k < n) ? Prow[(j - i0) * ld + k] : 0.0;
            else if (pad && rb == 0) out[e0] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pruning
// ------------------------------------------------------------------------------------------------
// The pruning program of the tree in use: every warp reads it uniformly, op after op, on its critical path,
// so it lives in constant memory (uniform loads through the constant cache) when it fits.
#define PB200_CONST_PROG_INTS 6144
__constant__ int c_prog[PB200_CONST_PROG_INTS];
// The 4x4 P(t) of the internal branches of the 4-state kernel (prune4v2.cuh, PCONST): warp-uniform operands, read through
// the constant cache into uniform registers (LDCU) and fed to DFMA as uniform operands -- they never pass the LSU.
// c_prog + c_P4 = 24 KB + 40 KB of the 64 KB constant bank.
#define PB200_CONST_P4_DOUBLES 5120
__constant__ double c_P4[PB200_CONST_P4_DOUBLES];

struct PruneParams {
    const int* prog;            // plan.h program in global memory (used when it does not fit c_prog)
    int prog_in_const;          // 1: read c_prog instead (this kernel's program always sits at the start of c_prog: use_const_prog)
    const unsigned char* z;     // [ns][zstride] compact codes
    long long zstride, npatt, ntiles, nitems;
    const double* Pbuf;         // [nclass][nnode][pstride]
    long long pstride;
    int nnode, nclass, nslots, n, semantics, nscale;
    const double* pi;           // [8*NT] zero padded
    double* fhK;                // [nclass][npatt]
    unsigned int* tile_counter; // dynamic work distribution (reset by mix_kernel)
    int warp_smem_bytes;        // shared memory per warp: slots + tip-code ring
    unsigned int stagger_ns;    // start delay between the warps that share an SM sub-partition (see kernel)
    int stagger_ctas;           // > 0: CTAs of 4 warps, this many per SM: the delay goes by the CTA's slot (blockIdx.x / SMs) instead
    int pipe_lock;              // n > 0: at most n warps of a sub-partition inside a DMMA block at once (tuning aid)
    int static_items;           // 1: items dealt round-robin to warps; 0: atomic work counter
    int prefetch_b;             // 1: prefetch.global.L1 the next op's B-fragment tiles one op ahead
    // lfun fused in (one class, PB200_SEM_LFUN): the root op also writes log f_h, weights it and the kernel ends
    // with the deterministic sum -- no separate mixture kernel for this case
    int fuse_lfun;
    const double* fpatt;        // [npatt]
    double* lnf;                // [npatt]
    double* item_partial;       // [nitems]: -sum_h fpatt_h log f_h of each work item (fixed order inside an item)
    unsigned int* done_counter; // CTAs finished (zero on entry; the last CTA resets it and tile_counter)
    double* result;             // [1]
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
// tip gathers stream through: read-only path, no L1 allocation (L1 is kept for the B fragments)
__device__ __forceinline__ double2 ld_gather(const double2* ptr)
{
#ifdef PB200_GATHER_L1
    return __ldg(ptr);
#else
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(ptr));
    return v;
#endif
}
// B-fragment tiles are re-read by every warp for every item: ask L1 to keep them (evict-last) while the tip gathers
// stream past (no-allocate)
__device__ __forceinline__ double2 ld_bfrag(const double2* ptr)
{
#ifdef PB200_BFRAG_EVICT_LAST
    double2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(ptr));
    return v;
#else
    return __ldg(ptr);
#endif
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

template <int NT, int KS, int MT>
struct PruneCfg {
    static constexpr int NP = 8 * NT;                       // padded output states
    static constexpr int KP = 4 * KS;                       // padded input states
    static constexpr int NTP = (NT + 1) / 2;                // n-tile pairs in a B-fragment element
    // Slot rows (physical column order): stride % 16 == 8 doubles, and rows with bit 1 of (row % 8) set keep
    // each (even, odd) column pair swapped; that makes BOTH the 8-byte A-fragment loads and the 16-byte
    // C-fragment stores bank-conflict free.
    static constexpr int STRIDE = ((NP + 7) / 16) * 16 + 8;
    static constexpr int ROWS = 8 * MT;                     // patterns per warp
    static constexpr int SLOT_DOUBLES = ROWS * STRIDE;
    static constexpr int RING = 8;                          // tips whose codes are prefetched ahead
    static constexpr int RING_BYTES = RING * ROWS;
    static_assert(STRIDE >= NP, "a row holds all physical columns");
};

// One warp owns ROWS site patterns of one site class and walks the whole tree for them; conditional
// vectors stay in registers (the node being built, C-fragment layout) and in the warp's private shared
// memory slots (finished children, [pattern][state] rows).  No block-level synchronisation at all.
// Tip codes are streamed through a small per-warp cp.async ring a few tips ahead of their use, so the
// HBM latency of the alignment never sits on the critical path.
template <int NT, int KS, int MT, int WARPS, bool CP>
__global__ void __launch_bounds__(WARPS * 32, 1) prune_kernel(const PruneParams p)
{
    using C = PruneCfg<NT, KS, MT>;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int xsw = (g >> 1) & 1;                           // pair-swap swizzle of this lane's rows
    // first 64 bytes: one ticket lock per SM sub-partition (warp % 4)
    unsigned int* lock_next = reinterpret_cast<unsigned int*>(smraw) + (warp & 3);
    volatile unsigned int* lock_serving = reinterpret_cast<unsigned int*>(smraw) + 8 + (warp & 3);
    double* slots = reinterpret_cast<double*>(smraw + 64 + (size_t)warp * p.warp_smem_bytes);
    unsigned char* ring = reinterpret_cast<unsigned char*>(slots + (size_t)p.nslots * C::SLOT_DOUBLES);
    if (threadIdx.x < 16) reinterpret_cast<unsigned int*>(smraw)[threadIdx.x] = 0u;
    __syncthreads();
    // The four (or two) warps of a sub-partition share one DMMA pipe.  Left to the hardware's round-robin they
    // advance through a DMMA block at a quarter of the speed each and leave it together -- and then all sit in
    // their tip-gather / store phases at once while the pipe idles (a convoy).  Taking turns instead (one warp owns
    // the pipe for a whole block, FIFO order) keeps the phases of the warps apart: while one contracts, the others
    // gather, store or queue.
    auto pipe_acquire = [&]() {
        if (!p.pipe_lock) return;
        if (lane == 0) {
            const unsigned int ticket = atomicAdd(lock_next, 1u);
            // pipe_lock = how many warps of the sub-partition may be inside a DMMA block at once
            while ((int)(ticket - *lock_serving) >= p.pipe_lock) __nanosleep(64);
        }
        __syncwarp();
    };
    auto pipe_release = [&]() {
        if (!p.pipe_lock) return;
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<unsigned int*>(lock_serving), 1u);
    };
    const long long N = p.npatt;
    // program words: constant memory (uniform loads) when the program fits there, else read-only global loads; an
    // explicit branch per access keeps the two address spaces apart (a selected pointer would turn generic)
    // (CP is a template parameter so that each instantiation holds exactly one kind of load)
    auto PW = [&](int i) -> int { if (CP) return c_prog[i]; else return __ldg(p.prog + i); };
    const int nops = PW(0), ntip = PW(2);
    const int ops0 = PW(3);

    // Warps that share a sub-partition (same warp % 4) share one DMMA pipe.  Started together they stay in
    // lockstep -- all gathering tips at once, then all queueing on the tensor pipe -- and the phases simply add
    // up.  Starting them one pipe-phase apart lets one warp's tip gathers/stores run under another's DMMAs.
    if (p.stagger_ns) {
        // (4-warp CTAs: the warps that share a sub-partition belong to different CTAs; the first wave is dealt breadth-first over
        //  the SMs, so blockIdx.x / gridDim-per-wave is the CTA's slot on its SM)
        unsigned long long wait_ns = (unsigned long long)(p.stagger_ctas > 0 ? (blockIdx.x / (gridDim.x / p.stagger_ctas)) % p.stagger_ctas : (warp >> 2)) * p.stagger_ns;
        while (wait_ns > 0) { unsigned int chunk = wait_ns > 500000ull ? 500000u : (unsigned int)wait_ns; __nanosleep(chunk); wait_ns -= chunk; }
    }

    // Items (pattern tile x class) are dealt round-robin over all warps of the grid: every item costs the same, so a
    // static deal balances as well as a work counter and keeps its atomic round trip off the critical path.
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long warp_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long item = p.static_items ? warp_global : -1;;) {
        if (!p.static_items) {
            unsigned int item_u = 0;
            if (lane == 0) item_u = atomicAdd(p.tile_counter, 1u);
            item = __shfl_sync(0xffffffffu, item_u, 0);
        }
        if (item >= p.nitems) break;
        const int cls = (int)(item / p.ntiles);
        const long long h0 = (item - (long long)cls * p.ntiles) * C::ROWS;
        const double* Pcls = p.Pbuf + (size_t)cls * p.nnode * p.pstride;
        const unsigned char* zt = p.z + h0;
        double lnsc[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) lnsc[mt] = 0.0;

        // prologue of the tip-code ring: the first RING tips of the walk
#pragma unroll
        for (int i = 0; i < C::RING; i++) {
            if (i < ntip && lane < C::ROWS / 4) cp_async4(ring + i * C::ROWS + 4 * lane, zt + (size_t)PW(PROG_HDR + i) * p.zstride + 4 * lane);
            cp_async_commit();
        }
        int ti = 0;                                          // tips consumed so far

        int pc = ops0;
        double acc[MT][NT][2];                               // the node being built; survives into the next op (OUT_REG)
        for (int op = 0; op < nops; op++) {
            const int nchild = PW(pc), out_slot = PW(pc + 1), sidx = PW(pc + 2), flags = PW(pc + 4), ntip_op = PW(pc + 5);
            pc += OP_HDR;
            double tmp[MT][NT][2];
            int c = 0;
            bool started = false;

            if (p.prefetch_b && op + 1 < nops) {
                // warm L1 with the B-fragment tiles the NEXT op will contract with, one op ahead of their use
                const int nx = pc + nchild * CH_REC;             // next op's header
                const int nnc = PW(nx);
                for (int k = 0; k < nnc; k++) {
                    const int kd = PW(nx + OP_HDR + k * CH_REC);
                    if (kd == CH_DIRECT || kd == CH_INTERNAL) {
                        const char* t0 = reinterpret_cast<const char*>(Pcls + (size_t)PW(nx + OP_HDR + k * CH_REC + 2) * p.pstride);
                        const int bytes = KS * C::NTP * 64 * (int)sizeof(double);
                        for (int off = lane * 128; off < bytes; off += 32 * 128)
                            asm volatile("prefetch.global.L1 [%0];" :: "l"(t0 + off));
                    }
                }
            }

            if (flags & OPF_HAS_DIRECT) {
                // Child 0 was finished by the previous op and is still in `acc`.  Thanks to the physical column
                // order (see phys_to_logical) acc[mt][nt][e] IS the A-fragment element of k-step 2nt+e.
                const double2* B = reinterpret_cast<const double2*>(Pcls + (size_t)PW(pc + 2) * p.pstride) + lane;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) tmp[mt][nt][0] = tmp[mt][nt][1] = 0.0;
                pipe_acquire();
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    double a[MT], b[2 * C::NTP];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a[mt] = acc[mt][ks >> 1][ks & 1];
#pragma unroll
                    for (int pr = 0; pr < C::NTP; pr++) {
                        double2 v = ld_bfrag(B + (ks * C::NTP + pr) * 32);
                        b[2 * pr] = v.x; b[2 * pr + 1] = v.y;
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) dmma884(tmp[mt][nt][0], tmp[mt][nt][1], a[mt], b[nt]);
                }
                pipe_release();
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) { acc[mt][nt][0] = tmp[mt][nt][0]; acc[mt][nt][1] = tmp[mt][nt][1]; }
                started = true;
                c = 1;
                pc += CH_REC;
            }

            if (ntip_op > 0) {
                // Tips (src/codeml.c:3555-3567): acc[h][j] (*)= PT_a[z_a,h][j] (* PT_b[z_b,h][j]).  The gathers of up to
                // two tips are issued back to back so their latencies overlap.
                const bool two = (ntip_op > 1);
                cp_async_wait<C::RING - 2>();
                __syncwarp();
                const unsigned char* zr0 = ring + (ti % C::RING) * C::ROWS;
                const unsigned char* zr1 = ring + ((ti + 1) % C::RING) * C::ROWS;
                int code0[MT], code1[MT];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) { code0[mt] = zr0[8 * mt + g]; code1[mt] = two ? zr1[8 * mt + g] : 0; }
                __syncwarp();
                for (int k = 0; k < (two ? 2 : 1); k++) {   // refill the ring entries just consumed
                    const int nx = ti + k + C::RING;
                    if (nx < ntip && lane < C::ROWS / 4)
                        cp_async4(ring + ((ti + k) % C::RING) * C::ROWS + 4 * lane, zt + (size_t)PW(PROG_HDR + nx) * p.zstride + 4 * lane);
                    cp_async_commit();
                }
                ti += two ? 2 : 1;
                const double* Pm0 = Pcls + (size_t)PW(pc + 2) * p.pstride;
                const double* Pm1 = Pcls + (size_t)PW(pc + (two ? CH_REC + 2 : 2)) * p.pstride;
                if (!started) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double2* row = reinterpret_cast<const double2*>(Pm0 + (size_t)code0[mt] * C::NP);
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) { double2 v = ld_gather(row + nt * 4 + q); acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y; }
                    }
                }
                else {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double2* row = reinterpret_cast<const double2*>(Pm0 + (size_t)code0[mt] * C::NP);
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) { double2 v = ld_gather(row + nt * 4 + q); acc[mt][nt][0] *= v.x; acc[mt][nt][1] *= v.y; }
                    }
                }
                if (two) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double2* row = reinterpret_cast<const double2*>(Pm1 + (size_t)code1[mt] * C::NP);
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) { double2 v = ld_gather(row + nt * 4 + q); tmp[mt][nt][0] = v.x; tmp[mt][nt][1] = v.y; }
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) { acc[mt][nt][0] *= tmp[mt][nt][0]; acc[mt][nt][1] *= tmp[mt][nt][1]; }
                }
                started = true;
                const int used = two ? 2 : 1;
                c += used;
                pc += used * CH_REC;
            }
            if (!started) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 1.0;
            }

            for (; c < nchild; c++) {
                const int kind = PW(pc), idx = PW(pc + 1), node = PW(pc + 2);
                pc += CH_REC;
                const double* Pm = Pcls + (size_t)node * p.pstride;
                if (kind == CH_TIP) {
                    // third and later tips of a multifurcation: one at a time
                    cp_async_wait<C::RING - 1>();
                    __syncwarp();
                    const unsigned char* zr = ring + (ti % C::RING) * C::ROWS;
                    int code[MT];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) code[mt] = zr[8 * mt + g];
                    __syncwarp();
                    {
                        const int nx = ti + C::RING;
                        if (nx < ntip && lane < C::ROWS / 4)
                            cp_async4(ring + (ti % C::RING) * C::ROWS + 4 * lane, zt + (size_t)PW(PROG_HDR + nx) * p.zstride + 4 * lane);
                        cp_async_commit();
                    }
                    ti++;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double2* row = reinterpret_cast<const double2*>(Pm + (size_t)code[mt] * C::NP);
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            double2 v = ld_gather(row + nt * 4 + q);
                            acc[mt][nt][0] *= v.x; acc[mt][nt][1] *= v.y;
                        }
                    }
                }
                else if (kind == CH_INTERNAL) {
                    // a parked child: acc[h][j] *= sum_k P[j][k] L_child[h][k]   (src/codeml.c:3568-3575) on DMMA
                    const double* S = slots + (size_t)idx * C::SLOT_DOUBLES;
                    const double2* B = reinterpret_cast<const double2*>(Pm) + lane;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) tmp[mt][nt][0] = tmp[mt][nt][1] = 0.0;
                    pipe_acquire();
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
                        double a[MT], b[2 * C::NTP];
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) a[mt] = S[(8 * mt + g) * C::STRIDE + ((8 * (ks >> 1) + 2 * q + (ks & 1)) ^ xsw)];
#pragma unroll
                        for (int pr = 0; pr < C::NTP; pr++) {
                            double2 v = ld_bfrag(B + (ks * C::NTP + pr) * 32);
                            b[2 * pr] = v.x; b[2 * pr + 1] = v.y;
                        }
#pragma unroll
                        for (int nt = 0; nt < NT; nt++)
#pragma unroll
                            for (int mt = 0; mt < MT; mt++) dmma884(tmp[mt][nt][0], tmp[mt][nt][1], a[mt], b[nt]);
                    }
                    pipe_release();
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) { acc[mt][nt][0] *= tmp[mt][nt][0]; acc[mt][nt][1] *= tmp[mt][nt][1]; }
                }
                else {
                    // running product of a folded multifurcation, stored earlier in C-fragment positions
                    const double* S = slots + (size_t)idx * C::SLOT_DOUBLES;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            double2 v = *reinterpret_cast<const double2*>(S + (8 * mt + g) * C::STRIDE + 8 * nt + 2 * q);
                            acc[mt][nt][0] *= xsw ? v.y : v.x; acc[mt][nt][1] *= xsw ? v.x : v.y;
                        }
                }
            }

            if (sidx >= 0) {
                // NodeScale (src/treesub.c:7213-7227): divide by the per-pattern maximum, remember its log
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    double mx = 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) { mx = fmax(mx, acc[mt][nt][0]); mx = fmax(mx, acc[mt][nt][1]); }
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    if (mx < 1e-300) {
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            acc[mt][nt][0] = (8 * nt + q     < p.n) ? 1.0 : 0.0;     // logical state of (nt, q, e) is 8nt+4e+q
                            acc[mt][nt][1] = (8 * nt + 4 + q < p.n) ? 1.0 : 0.0;
                        }
                        lnsc[mt] += -800.0;
                    }
                    else {
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) { acc[mt][nt][0] /= mx; acc[mt][nt][1] /= mx; }
                        lnsc[mt] += log(mx);
                    }
                }
            }

            if (out_slot == OUT_REG) {
                // stays in registers: the next op consumes it as its CH_DIRECT child
            }
            else if (out_slot >= 0) {
                double* S = slots + (size_t)out_slot * C::SLOT_DOUBLES;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
                        *reinterpret_cast<double2*>(S + (8 * mt + g) * C::STRIDE + 8 * nt + 2 * q) =
                            xsw ? make_double2(acc[mt][nt][1], acc[mt][nt][0]) : make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                __syncwarp();
            }
            else {
                // root: f_h = sum_i pi_i L_root[h][i]  (src/treesub.c:7780-7781, 7728-7729), then the
                // guards and scale factors of lfun (7782-7798) or fx_r (7731-7750)
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    double fh = 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        double2 w = __ldg(reinterpret_cast<const double2*>(p.pi) + nt * 4 + q);
                        fh = fma(w.x, acc[mt][nt][0], fh);
                        fh = fma(w.y, acc[mt][nt][1], fh);
                    }
                    fh += __shfl_xor_sync(0xffffffffu, fh, 1);
                    fh += __shfl_xor_sync(0xffffffffu, fh, 2);
                    long long h = h0 + 8 * mt + g;
                    if (q == 0 && h < N) {
                        double v;
                        if (p.semantics == SEM_LFUN) {
                            if (fh <= 0) fh = 1e-80;
                            v = log(fh) + lnsc[mt];
                        }
                        else {
                            if (fh <= 0) fh = 1e-300;
                            v = p.nscale ? log(fh) + lnsc[mt] : fh;
                        }
                        p.fhK[(size_t)cls * N + h] = v;
                        if (p.fuse_lfun) {                         // src/treesub.c:7779, 7800-7801
                            const double w = p.fpatt[h];
                            const bool on = (w > 0);
                            p.lnf[h] = on ? v : 0.0;
                            fh = on ? -v * w : 0.0;
                        }
                    }
                    else fh = 0.0;
                    if (p.fuse_lfun) {
                        // sum of the tile's rows in a fixed order (lanes with q == 0 hold rows g = 0..7)
                        fh += __shfl_xor_sync(0xffffffffu, fh, 4);
                        fh += __shfl_xor_sync(0xffffffffu, fh, 8);
                        fh += __shfl_xor_sync(0xffffffffu, fh, 16);
                        if (mt == 0) lnsc[0] = fh; else lnsc[0] += fh;   // lnsc is dead after the root: reuse as the item's sum
                    }
                }
                if (p.fuse_lfun && lane == 0) p.item_partial[item] = lnsc[0];
            }
        }
        cp_async_wait<0>();
        __syncwarp();
        if (p.static_items) item += warp_total;
    }

    if (p.fuse_lfun) {
        // The last CTA to finish adds the per-item sums in item order: a fixed-shape reduction, bit-reproducible
        // from call to call whatever the scheduling was (the optimiser differentiates -lnL numerically).
        __shared__ double red[32];
        __shared__ bool last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            last = (atomicAdd(p.done_counter, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (last) {
            __threadfence();
            double t = 0.0;
            for (long long i = threadIdx.x; i < p.nitems; i += blockDim.x) t += p.item_partial[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) red[warp] = t;
            __syncthreads();
            if (threadIdx.x == 0) {
                double r = 0.0;
                for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += red[w];
                p.result[0] = r;
                *p.done_counter = 0u;
                *p.tile_counter = 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mixture + weighted deterministic sum
// ------------------------------------------------------------------------------------------------
struct MixParams {
    const double* fhK;      // [nclass][npatt]
    const double* fpatt;    // [npatt]
    const double* freqK;    // [nclass]
    double* lnf;            // [npatt] out
    double* partial;        // [gridDim.x] scratch
    unsigned int* counter;  // zero-initialised, reset by the last block
    unsigned int* tile_counter;  // pruning kernel's work counter: reset here for the next evaluation
    double* result;         // [1] out: -lnL of this shard
    long long npatt;
    int nclass, semantics, nscale;
    // fused all-reduce over NVLink peer memory (nranks > 1): every rank owns an exchange buffer
    // xchg[2 slots][nranks][2 words] that all peers can write; peer_xchg[r] = rank r's buffer as mapped here
    unsigned long long* const* peer_xchg;
    int nranks, rank;
    unsigned long long seq;   // evaluation number (> 0), doubles as the arrival flag
    unsigned long long timeout_ns;   // a peer that has not arrived after this long is reported in result[1] (1 + its rank)
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Block b sums patterns [b*chunk, (b+1)*chunk) in a fixed order; the last block to finish adds the block
// partials in index order.  No floating-point atomics anywhere, so -lnL(x) is reproducible call to call
// (the optimiser differentiates it numerically, src/tools.c:6561-6590).
__global__ void __launch_bounds__(256) mix_kernel(const MixParams p)
{
    __shared__ double red[256];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const long long N = p.npatt;
    double s = 0.0;
    for (long long h = (long long)blockIdx.x * 256 + tid; h < N; h += (long long)gridDim.x * 256) {
        const double w = p.fpatt[h], f0 = p.fhK[h];           // both requested before either is looked at
        double fh;
        if (w <= 0) { p.lnf[h] = 0.0; continue; }           // src/treesub.c:7779, 7631
        if (p.semantics == SEM_LFUN) fh = f0;
        else if (p.nscale) {                                 // src/treesub.c:7632-7638
            int it = 0;
            for (int ir = 1; ir < p.nclass; ir++) if (p.fhK[(size_t)ir * N + h] > p.fhK[(size_t)it * N + h]) it = ir;
            double t = p.fhK[(size_t)it * N + h];
            fh = 0.0;
            for (int ir = 0; ir < p.nclass; ir++) fh += p.freqK[ir] * exp(p.fhK[(size_t)ir * N + h] - t);
            fh = t + log(fh);
        }
        else {                                               // src/treesub.c:7640-7652
            fh = 0.0;
            for (int ir = 0; ir < p.nclass; ir++) fh += p.freqK[ir] * p.fhK[(size_t)ir * N + h];
            if (fh <= 0) fh = 1e-300;
            fh = log(fh);
        }
        p.lnf[h] = fh;
        s -= fh * w;
    }
    // fixed-shape block sum: butterfly inside each warp, then the 8 warp sums in order
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double b = 0.0;
        for (int w = 0; w < 8; w++) b += red[w];
        red[0] = b;
        p.partial[blockIdx.x] = red[0];
        __threadfence();
        unsigned int done = atomicAdd(p.counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t = 0.0;
        for (int b = tid; b < (int)gridDim.x; b += 256) t += p.partial[b];
        red[tid] = t;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
        if (tid == 0) { *p.counter = 0u; *p.tile_counter = 0u; }
        if (p.nranks <= 1) { if (tid == 0) p.result[0] = red[0]; }
        else {
            // All-reduce of the scalar, fused into this kernel: thread r < nranks stores (value, seq) straight into
            // rank r's exchange buffer over NVLink, then waits for rank r's own contribution to land here; the sum is
            // taken in rank order, so every rank gets the same bits.  Slots alternate with the evaluation number: a
            // peer can only be one evaluation ahead, so two slots are enough.
            const int slot = (int)(p.seq & 1ull);
            __shared__ double vals[64];
            if (tid < p.nranks) {
                volatile unsigned long long* dst = p.peer_xchg[tid] + ((size_t)slot * p.nranks + p.rank) * 2;
                dst[0] = (unsigned long long)__double_as_longlong(red[0]);
                __threadfence_system();
                dst[1] = p.seq;
                volatile unsigned long long* src = p.peer_xchg[p.rank] + ((size_t)slot * p.nranks + tid) * 2;
                // bounded wait: a peer that died (or never launched) must not hang every GPU inside this kernel
                const unsigned long long t0 = global_timer_ns();
                bool arrived = true;
                unsigned int spins = 0;
                while (src[1] != p.seq) {
                    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > p.timeout_ns) { arrived = false; break; }
                }
                __threadfence_system();
                vals[tid] = arrived ? __longlong_as_double((long long)src[0]) : __longlong_as_double(0x7ff8000000000000ll);
                if (!arrived) p.result[1] = (double)(1 + tid);      // read by pb200_eval_end: fails loudly, naming the rank
            }
            __syncthreads();
            if (tid == 0) {
                double r = 0.0;
                for (int k = 0; k < p.nranks; k++) r += vals[k];
                p.result[0] = r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 peak micro-benchmarks (roofline denominators) and L2 flush
// ------------------------------------------------------------------------------------------------
__global__ void peak_dmma_kernel(double* out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void peak_dfma_kernel(double* out, int iters, double a, double b)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void flush_kernel(double* buf, long long n, double v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) buf[i] = v;
}

}  // namespace pb200
