// prune_ps.cuh -- the P-stationary, warp-specialised form of the pruning kernel for the 61..64-state (codon) shape.
//
// Same arithmetic as prune_kernel (ConditionalPNode src/codeml.c:3526-3582, NodeScale src/treesub.c:7200-7230, the root pi-dot of
// lfun / fx_r src/treesub.c:7780-7798, 7726-7750), other work split.  prune_kernel gives a warp 8 patterns and lets it walk the whole
// tree alone: every warp re-reads every P(t) tile for its 8 patterns and spends 44 % of its time outside the DMMA blocks (tip gathers,
// products, slot stores).  Here a CTA of 6 warps (two per SM, out of phase) owns a contiguous share of the patterns and works through
// it in CHUNKS of 8*MT patterns of one site class, with two kinds of warps:
//   * 4 CONTRACTION warps; warp w owns the 16 OUTPUT states 16w..16w+15 of every node.  For a child's contraction the warp holds its
//     16 x 64 slab of P(t) in registers as B fragments of mma.m8n8k4 (two per lane and k-step: one 16-byte load) and streams the
//     child's vectors of the whole chunk from shared memory as A fragments (one LDS.64 per TWO DMMAs, conflict-free at a row stride
//     of 68 doubles).  The finished node goes to a shared-memory "D slot" ([pattern][state]; the C fragments of lane (g, q) are
//     pattern 8mt+g, states 16w+8nt+2q, +1: 16-byte stores), one named barrier among these 4 warps per node.  They never gather a tip.
//   * 2 GATHER warps turn the tip children of a node into one vector per pattern -- the product of the tips' rows PT[code][.] -- in a
//     "G slot" (src/codeml.c:3555-3567), running ahead of the contraction warps: G slots are double-banked by chunk parity and
//     handed over through full / empty mbarriers, tip codes are staged one chunk ahead with cp.async.
// Everything is in LOGICAL state order (pt_kernel, logical = 1).  Programs: plan.h (PlanPS).
// Use (pb200_eval_begin): splitting a tile's STATES over four warps makes the walk of one 8-pattern tile four times shorter than
// prune_kernel's, so this kernel is the default when there is little work (small alignments: every tile gets its own CTA and an
// evaluation is that walk's latency); with enough tiles to fill the SMs several times prune_kernel's sixteen independent warps per
// SM overlap better and it is the default (profiles/r2_small_latency.txt).
#pragma once
#include "kernels.cuh"
#include "prune_tma.cuh"     // mbarrier helpers

namespace pb200 {

struct PrunePsParams {
    const int* prog;            // plan.h PlanPS program in global memory (used when it does not fit c_prog)
    int prog_base;              // offset of the program in c_prog
    const unsigned char* z;     // [ns][zstride] compact codes
    long long zstride, npatt;
    long long ntile;            // 8-pattern tiles per class
    long long nunits;           // nclass * ntile: the work of the launch, dealt to the CTAs in equal contiguous shares
    const double* Pbuf;         // [nclass][nnode][pstride]: tip rows PT[code][64] / B fragments, logical state order
    long long pstride;
    int nnode, nclass, n, semantics, nscale;
    int nbanks;                 // 1 or 2 banks of G slots (2: the gather warps may work a whole chunk ahead)
    int ntip;                   // tips per walk (rows of the staged code table)
    const double* pi;           // [64] zero padded, logical order
    double* fhK;                // [nclass][npatt]
};

enum { PS_NWC = 4, PS_NGW = 2, PS_THREADS = 32 * (PS_NWC + PS_NGW), PS_GT = 32 * PS_NGW };

template <int MT>
struct PsCfg {
    static constexpr int CH = 8 * MT;                       // patterns per chunk
    static constexpr int STRIDE = 68;                       // doubles per slot row: A-fragment loads hit 16 distinct 8-byte banks per half-warp
    static constexpr int SLOT_DOUBLES = CH * STRIDE;
    static constexpr int TPP = PS_GT / CH;                  // gather threads per pattern
    static constexpr int NPC = 32 / TPP;                    // 16-byte pieces of a 64-state row per gather thread
    static_assert(TPP >= 1 && NPC <= 16, "gather threads hold two rows' pieces in registers");
    __host__ __device__ static size_t bar_bytes(int nG, int nbanks) { return ((size_t)2 * nG * nbanks * 8 + 15) & ~(size_t)15; }
    static size_t smem_bytes(int nD, int nG, int nbanks, int ntip, int nsteps)
    {
        return bar_bytes(nG, nbanks) + ((size_t)(nD + nG * nbanks) * SLOT_DOUBLES + 2 * (PS_NWC * CH + CH)) * sizeof(double) + (((size_t)2 * ntip * CH + 15) & ~(size_t)15) +
               (size_t)nsteps * PS_STEP_INTS * sizeof(int);
    }
};

// asm volatile: the refill of the slab registers must stay where it is written, between the DMMAs (which are asm volatile too)
__device__ __forceinline__ double2 ld_bslab(const double2* ptr)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(ptr));
    return v;
}
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned saddr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];\n" : "=d"(v) : "r"(saddr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// A fragments of one k-step for HM pattern tiles, pinned in program order (asm volatile)
template <int HM, int KS>
__device__ __forceinline__ void ps_load_a(unsigned sa, double (&a)[HM])
{
    a[0] = lds_f64<(0 * 8 * 68 + 4 * KS) * 8>(sa);
    if (HM > 1) a[1 % HM] = lds_f64<(1 * 8 * 68 + 4 * KS) * 8>(sa);
    if (HM > 2) a[2 % HM] = lds_f64<(2 * 8 * 68 + 4 * KS) * 8>(sa);
    if (HM > 3) a[3 % HM] = lds_f64<(3 * 8 * 68 + 4 * KS) * 8>(sa);
    static_assert(HM <= 4, "ps_load_a covers four pattern tiles");
}
// k-steps KS..15 of a contraction over HM pattern tiles: load the A fragments of k-step KS+1, issue the DMMAs of k-step KS (both
// n-tiles), and -- in the LAST pass over the slab (REFILL) -- refill the first half of the slab registers with the NEXT contraction's
// matrix once k-step 7 is issued
template <int HM, int KS, bool REFILL>
__device__ __forceinline__ void ps_kloop(unsigned sa, int cnt, double (&tmp)[HM][2][2], double (&a)[HM], double (&an)[HM], double2 (&b)[16],
                                         const double2* Pnext)
{
    if constexpr (KS < 16) {
        if constexpr (KS < 15) ps_load_a<HM, KS + 1>(sa, an);
#pragma unroll
        for (int m = 0; m < HM; m++)
            if (m < cnt) {
                dmma884(tmp[m][0][0], tmp[m][0][1], a[m], b[KS].x);
                dmma884(tmp[m][1][0], tmp[m][1][1], a[m], b[KS].y);
            }
        if constexpr (KS == 7 && REFILL) {
            if (Pnext) {
#pragma unroll
                for (int k2 = 0; k2 < 8; k2++) b[k2] = ld_bslab(Pnext + k2 * 128);
            }
        }
        ps_kloop<HM, KS + 1, REFILL>(sa, cnt, tmp, an, a, b, Pnext);     // a / an swap roles
    }
}

template <int MT, bool CP>
__global__ void __launch_bounds__(PS_THREADS, 2) prune_ps_kernel(const PrunePsParams p)
{
    using C = PsCfg<MT>;
    constexpr int CH = C::CH, STRIDE = C::STRIDE;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    auto PW = [&](int i) -> int { if (CP) return c_prog[i]; else return __ldg(p.prog + i); };      // absolute positions: see prune_kernel
    const int cb = CP ? p.prog_base : 0, whdr = cb + PS_HDR;
    const int n_gops = PW(cb), n_steps = PW(cb + 1), nG = PW(cb + 2), nD = PW(cb + 3), off_g = cb + PW(cb + 4), off_d = cb + PW(cb + 5);
    const int nGS = nG * p.nbanks;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smraw);      // [nGS]: gather warps -> contraction warps
    unsigned long long* empty = full + nGS;                                         // [nGS]: contraction warps -> gather warps
    double* dslots = reinterpret_cast<double*>(smraw + C::bar_bytes(nG, p.nbanks));
    double* gslots = dslots + (size_t)nD * C::SLOT_DOUBLES;
    double* red = gslots + (size_t)nGS * C::SLOT_DOUBLES;            // [2][4 warps][CH]: per-warp maxima / pi-dot partials
    double* lnsc = red + 2 * PS_NWC * CH;                            // [2][CH]: accumulated log scale factors of the chunk
    unsigned char* codes = reinterpret_cast<unsigned char*>(lnsc + 2 * CH);   // [2][ntip][CH]: tip codes of this chunk and the next
    int* stab = reinterpret_cast<int*>(codes + (((size_t)2 * p.ntip * CH + 15) & ~(size_t)15));   // [n_steps][8]: the contraction warps' program
    for (int i = tid; i < n_steps * PS_STEP_INTS; i += PS_THREADS) stab[i] = PW(off_d + i);
    if (tid == 0) {
        for (int s = 0; s < nGS; s++) { mbar_init(full + s, PS_NGW); mbar_init(empty + s, PS_NWC); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const long long N = p.npatt, T = p.ntile;
    // this CTA's share of the (class, 8-pattern tile) units; both kinds of warps derive the same chunk sequence from it
    long long u = (long long)blockIdx.x * p.nunits / gridDim.x;
    const long long u_end = (long long)(blockIdx.x + 1) * p.nunits / gridDim.x;
    auto chunk_of = [&](long long uu, int& cls, long long& h0) -> int {
        cls = (int)(uu / T);
        const long long m0 = uu - (long long)cls * T;
        h0 = m0 * 8;
        long long cnt = u_end - uu;
        if (cnt > T - m0) cnt = T - m0;
        if (cnt > MT) cnt = MT;
        return (int)cnt;
    };

    if (w >= PS_NWC) {
        // ---------------------------------------------------------------- gather warps
        const int gt = tid - 32 * PS_NWC, pat = gt / C::TPP, jj = gt % C::TPP;
        auto stage_codes = [&](int buf, long long h0) {
            // tip codes of a chunk, [tip in walk order][pattern]; 4 patterns per cp.async (rows are 128-byte aligned, h0 % 8 == 0)
            unsigned char* dst = codes + (size_t)buf * p.ntip * CH;
            int pcg = off_g, t = 0;
            for (int go = 0; go < n_gops; go++) {
                const int nt = PW(pcg + 2);
                for (int k = 0; k < nt; k++, t++) {
                    const unsigned char* src = p.z + (size_t)PW(pcg + PS_GOP_HDR + k) * p.zstride + h0;
                    for (int e = gt; e < CH / 4; e += PS_GT) cp_async4(dst + t * CH + 4 * e, src + 4 * e);
                }
                pcg += PS_GOP_HDR + nt;
            }
            cp_async_commit();
        };
        if (u < u_end) { int c0; long long h00; chunk_of(u, c0, h00); stage_codes(0, h00); }
        for (int it = 0; u < u_end; it++) {
            int cls; long long h0;
            const int cnt = chunk_of(u, cls, h0);
            u += cnt;
            cp_async_wait<0>();
            bar_sync_named(2, PS_GT);                           // this chunk's codes (all gather threads' copies) are visible, and every
                                                              // gather thread is done with the previous chunk's table
            if (u < u_end) { int c1; long long h1; chunk_of(u, c1, h1); stage_codes((it + 1) & 1, h1); }
            const unsigned char* cd = codes + (size_t)(it & 1) * p.ntip * CH + pat;
            const bool live = (h0 + pat < N);
            const double* Pcls = p.Pbuf + (size_t)cls * p.nnode * p.pstride;
            const int bank = it % p.nbanks, jt = it / p.nbanks;
            int pcg = off_g, t = 0;
            for (int go = 0; go < n_gops; go++) {
                const int gslot = PW(pcg), k = PW(pcg + 1), nt = PW(pcg + 2);
                const int S = bank * nG + gslot;
                const int nw = jt * PW(whdr + gslot) + k;           // how many times this slot has been written before
                const int c0 = live ? cd[t * CH] : 0, c1 = (live && nt > 1) ? cd[(t + 1) * CH] : 0;
                const double2* r0 = reinterpret_cast<const double2*>(Pcls + (size_t)PW(pcg + PS_GOP_HDR) * p.pstride + (size_t)c0 * 64) + jj;
                double2 a[C::NPC];
#pragma unroll
                for (int i = 0; i < C::NPC; i++) a[i] = ld_gather(r0 + i * C::TPP);
                if (nt > 1) {
                    const double2* r1 = reinterpret_cast<const double2*>(Pcls + (size_t)PW(pcg + PS_GOP_HDR + 1) * p.pstride + (size_t)c1 * 64) + jj;
                    double2 b[C::NPC];
#pragma unroll
                    for (int i = 0; i < C::NPC; i++) b[i] = ld_gather(r1 + i * C::TPP);
#pragma unroll
                    for (int i = 0; i < C::NPC; i++) { a[i].x *= b[i].x; a[i].y *= b[i].y; }
                }
                for (int kk = 2; kk < nt; kk++) {            // third and later tips of a multifurcation
                    const int ck = live ? cd[(t + kk) * CH] : 0;
                    const double2* rk = reinterpret_cast<const double2*>(Pcls + (size_t)PW(pcg + PS_GOP_HDR + kk) * p.pstride + (size_t)ck * 64) + jj;
                    double2 b[C::NPC];
#pragma unroll
                    for (int i = 0; i < C::NPC; i++) b[i] = ld_gather(rk + i * C::TPP);
#pragma unroll
                    for (int i = 0; i < C::NPC; i++) { a[i].x *= b[i].x; a[i].y *= b[i].y; }
                }
                if (nw >= 1) mbar_wait(empty + S, (unsigned)((nw - 1) & 1));     // the readers of the previous contents are done
                double2* dst = reinterpret_cast<double2*>(gslots + (size_t)S * C::SLOT_DOUBLES + pat * STRIDE) + jj;
#pragma unroll
                for (int i = 0; i < C::NPC; i++) dst[i * C::TPP] = a[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(full + S);
                t += nt;
                pcg += PS_GOP_HDR + nt;
            }
        }
        cp_async_wait<0>();
        return;
    }

    // -------------------------------------------------------------------- contraction warps
    const int g = lane >> 2, q = lane & 3;
    // pi of this lane's four output states 16w + 8nt + 2q + e
    const double2 piw0 = __ldg(reinterpret_cast<const double2*>(p.pi) + 8 * w + q), piw1 = __ldg(reinterpret_cast<const double2*>(p.pi) + 8 * w + 4 + q);
    // this lane's elements of a B-fragment tile (pt_kernel, PT_MODE_FRAG, logical): (ks, pr = w, lane, e = 0, 1) = the n-tiles 2w, 2w+1
    const int boff = (w * 32 + lane) * 2;
    // The warp's slab of P(t) lives in b[16] (two n-tiles per element).  It is refilled ON THE FLY with the slab of the NEXT contraction
    // (next child, next node or next chunk): the first half as soon as k-steps 0..7 of the running contraction are issued, the second
    // half after its last k-step -- so a contraction never waits for its matrix, and no second set of registers is needed.
    const bool contracts = (stab[0] & PS_KIND_MASK) != PS_NONE;
    double2 b[16];
    if (contracts && u < u_end) {
        int c0; long long h00; chunk_of(u, c0, h00);
        const double2* Pm = reinterpret_cast<const double2*>(p.Pbuf + ((size_t)c0 * p.nnode + stab[2]) * p.pstride + boff);
#pragma unroll
        for (int ks = 0; ks < 16; ks++) b[ks] = __ldg(Pm + ks * 128);
    }
    for (int it = 0; u < u_end; it++) {
        int cls; long long h0;
        const int cnt = chunk_of(u, cls, h0);
        u += cnt;
        const double* Pcls = p.Pbuf + (size_t)cls * p.nnode * p.pstride;
        const double* Pcls_next = nullptr;                     // class of the next chunk (nullptr: this is the last one)
        if (u < u_end) { int c1; long long h1; chunk_of(u, c1, h1); Pcls_next = p.Pbuf + (size_t)c1 * p.nnode * p.pstride; }
        const int bank = it % p.nbanks, jt = it / p.nbanks;
        double* red_c = red + (it & 1) * PS_NWC * CH;          // `red` and `lnsc` alternate with the chunk: the root's last readers never
        double* lnsc_c = lnsc + (it & 1) * CH;                 // hold up the next chunk
        if (tid < CH) lnsc_c[tid] = 0.0;
        double acc[MT][2][2];
        for (int step = 0; step < n_steps; step++) {
            const int4 s0 = *reinterpret_cast<const int4*>(stab + step * PS_STEP_INTS);
            const int kind = s0.x & PS_KIND_MASK;
            // internal child (src/codeml.c:3568-3575): acc[h][j] (*)= sum_k P[j][k] L_child[h][k] on the FP64 tensor path
            if (kind != PS_NONE) {
                const double2* Pnext = nullptr;                // the slab of the next contraction
                {
                    const double* Pn = Pcls;
                    int nstep = step + 1;
                    if (nstep == n_steps) { nstep = 0; Pn = Pcls_next; }
                    if (Pn) Pnext = reinterpret_cast<const double2*>(Pn + (size_t)stab[nstep * PS_STEP_INTS + 2] * p.pstride + boff);
                }
                const double* S;
                int GS = -1;
                if (kind == PS_G) {
                    GS = bank * nG + s0.y;
                    mbar_wait(full + GS, (unsigned)((jt * (s0.w >> 16) + (s0.w & 0xffff)) & 1));
                    S = gslots + (size_t)GS * C::SLOT_DOUBLES + g * STRIDE + q;
                }
                else S = dslots + (size_t)s0.y * C::SLOT_DOUBLES + g * STRIDE + q;
                // the A fragments of k-step ks+1 are loaded while the DMMAs of k-step ks issue; each feeds both n-tiles of the warp
                const bool first = (s0.x & PS_STEP_FIRST) != 0;
                // two passes of MT/2 pattern tiles each: half the accumulator registers, which leaves the scheduler room to keep
                // the A-fragment loads a k-step ahead of their DMMAs
                constexpr int HM = MT / 2;
                const unsigned sa = smem_u32(S);
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    double tmp[HM][2][2], a[HM], an[HM];
#pragma unroll
                    for (int m = 0; m < HM; m++) tmp[m][0][0] = tmp[m][0][1] = tmp[m][1][0] = tmp[m][1][1] = 0.0;
                    const unsigned sh = sa + half * HM * 8 * STRIDE * 8;
                    ps_load_a<HM, 0>(sh, a);
                    if (half == 0) ps_kloop<HM, 0, false>(sh, cnt, tmp, a, an, b, Pnext);
                    else           ps_kloop<HM, 0, true>(sh, cnt - HM, tmp, a, an, b, Pnext);
                    if (first) {
#pragma unroll
                        for (int m = 0; m < HM; m++)
#pragma unroll
                            for (int nt = 0; nt < 2; nt++) { acc[half * HM + m][nt][0] = tmp[m][nt][0]; acc[half * HM + m][nt][1] = tmp[m][nt][1]; }
                    }
                    else {
#pragma unroll
                        for (int m = 0; m < HM; m++)
#pragma unroll
                            for (int nt = 0; nt < 2; nt++) { acc[half * HM + m][nt][0] *= tmp[m][nt][0]; acc[half * HM + m][nt][1] *= tmp[m][nt][1]; }
                    }
                }
                if (Pnext) {
#pragma unroll
                    for (int k2 = 8; k2 < 16; k2++) b[k2] = ld_bslab(Pnext + k2 * 128);
                }
                if (GS >= 0) {
                    // this warp is done with the G slot: hand it back to the gather warps right away
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + GS);
                }
            }
            if (!(s0.x & PS_STEP_LAST)) continue;

            // ---- the node is complete: its own tips, scaling, output
            const int4 s1 = *reinterpret_cast<const int4*>(stab + step * PS_STEP_INTS + 4);
            const int out_slot = s1.x, sidx = s1.y, tp = s1.z;
            if (tp >= 0) {
                // the node's own tips: their product, made by the gather warps
                const int GS = bank * nG + tp;
                mbar_wait(full + GS, (unsigned)((jt * (s1.w >> 16) + (s1.w & 0xffff)) & 1));
                const double* S = gslots + (size_t)GS * C::SLOT_DOUBLES + g * STRIDE + 16 * w + 2 * q;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < 2; nt++) {
                        const double2 v = *reinterpret_cast<const double2*>(S + mt * 8 * STRIDE + 8 * nt);
                        if (kind == PS_NONE) { acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y; }
                        else { acc[mt][nt][0] *= v.x; acc[mt][nt][1] *= v.y; }
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + GS);
            }

            if (sidx >= 0) {
                // NodeScale (src/treesub.c:7213-7227): divide by the per-pattern maximum over ALL states (4 warps), remember its log
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    double mx = fmax(fmax(acc[mt][0][0], acc[mt][0][1]), fmax(acc[mt][1][0], acc[mt][1][1]));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    if (q == 0) red_c[w * CH + 8 * mt + g] = mx;
                }
                bar_sync_named(1, 32 * PS_NWC);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    double mx = 0.0;
#pragma unroll
                    for (int ww = 0; ww < PS_NWC; ww++) mx = fmax(mx, red_c[ww * CH + 8 * mt + g]);
                    if (mx < 1e-300) {
#pragma unroll
                        for (int nt = 0; nt < 2; nt++) {
                            acc[mt][nt][0] = (16 * w + 8 * nt + 2 * q     < p.n) ? 1.0 : 0.0;
                            acc[mt][nt][1] = (16 * w + 8 * nt + 2 * q + 1 < p.n) ? 1.0 : 0.0;
                        }
                        if (w == 0 && q == 0) lnsc_c[8 * mt + g] += -800.0;
                    }
                    else {
#pragma unroll
                        for (int nt = 0; nt < 2; nt++) { acc[mt][nt][0] /= mx; acc[mt][nt][1] /= mx; }
                        if (w == 0 && q == 0) lnsc_c[8 * mt + g] += log(mx);
                    }
                }
            }

            if (out_slot >= 0) {
                double* S = dslots + (size_t)out_slot * C::SLOT_DOUBLES + g * STRIDE + 16 * w + 2 * q;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < 2; nt++)
                        *reinterpret_cast<double2*>(S + mt * 8 * STRIDE + 8 * nt) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                bar_sync_named(1, 32 * PS_NWC);     // the parent reads all states of all patterns; also fences this op's reads of `red` and of its children
            }
            else {
                // root: f_h = sum_i pi_i L_root[h][i] (src/treesub.c:7780-7781, 7728-7729): partial over this warp's 16 states, then the
                // 4 warps in order; guards and scale factors of lfun (7782-7798) or fx_r (7731-7750)
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    double fh = piw0.x * acc[mt][0][0];
                    fh = fma(piw0.y, acc[mt][0][1], fh);
                    fh = fma(piw1.x, acc[mt][1][0], fh);
                    fh = fma(piw1.y, acc[mt][1][1], fh);
                    fh += __shfl_xor_sync(0xffffffffu, fh, 1);
                    fh += __shfl_xor_sync(0xffffffffu, fh, 2);
                    if (q == 0) red_c[w * CH + 8 * mt + g] = fh;
                }
                bar_sync_named(1, 32 * PS_NWC);
                if (tid < 8 * cnt) {
                    double fh = 0.0;
#pragma unroll
                    for (int ww = 0; ww < PS_NWC; ww++) fh += red_c[ww * CH + tid];
                    const long long h = h0 + tid;
                    if (h < N) {
                        double v;
                        if (p.semantics == SEM_LFUN) {
                            if (fh <= 0) fh = 1e-80;
                            v = log(fh) + lnsc_c[tid];
                        }
                        else {
                            if (fh <= 0) fh = 1e-300;
                            v = p.nscale ? log(fh) + lnsc_c[tid] : fh;
                        }
                        p.fhK[(size_t)cls * N + h] = v;
                    }
                }
                // no second barrier: the next chunk works on the other copies of `red` / `lnsc`, and by the time these come round
                // again (two chunks on) every warp has passed a barrier behind this point
            }
        }
    }
}

}  // namespace pb200
