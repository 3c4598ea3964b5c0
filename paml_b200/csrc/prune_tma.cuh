// prune_tma.cuh -- the pruning kernel with the per-branch P(t) tiles staged in shared memory by TMA.
//
// Same arithmetic and program walk as prune_kernel (kernels.cuh), different data movement:
//   * one CTA owns a contiguous range of site patterns; its warps each take ROWS of them per chunk and walk the
//     tree together (not in lockstep: a warp may run up to NBUF-1 matrices ahead of the slowest one);
//   * every P(t) matrix the walk consumes -- B-fragment tile for an internal child, transposed rows for a tip --
//     is brought from L2 into a ring of NBUF shared-memory buffers by ONE bulk async copy (cp.async.bulk, the
//     1-D TMA; UBLKCP in SASS) issued by an elected lane, completion signalled on an mbarrier ("full"); a warp
//     that is done with a buffer arrives on its "empty" mbarrier, and the producer refills it with the matrix
//     NBUF positions further down the walk.  The loads of the next matrices therefore overlap the DMMA work on
//     the current one, and nothing on a warp's critical path touches L1/L2: B fragments and tip rows come from
//     shared memory, A fragments from registers (direct children) or the warp's private slot.
// HBM/L2 traffic per CTA chunk = (matrices per walk) x (tile bytes), shared by all warps of the CTA.
#pragma once
#include "kernels.cuh"

namespace pb200 {

struct PruneTmaParams {
    PruneParams base;
    int nbuf;                   // ring depth
    int buf_bytes;              // bytes per P(t) tile (multiple of 128)
    long long rows_per_cta;     // patterns owned by one CTA (multiple of 8)
    int bar_bytes;              // shared memory reserved for the mbarriers (multiple of 128)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(void* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, void* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NT, int KS, int MT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) prune_tma_kernel(const PruneTmaParams pp)
{
    using C = PruneCfg<NT, KS, MT>;
    const PruneParams& p = pp.base;
    extern __shared__ __align__(128) unsigned char smtma[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int xsw = (g >> 1) & 1;
    const int NBUF = pp.nbuf;
    unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(smtma);
    unsigned long long* empty_bar = full_bar + NBUF;
    unsigned char* bufs = smtma + pp.bar_bytes;
    unsigned char* warp_base = bufs + (size_t)NBUF * pp.buf_bytes + (size_t)warp * p.warp_smem_bytes;
    double* slots = reinterpret_cast<double*>(warp_base);
    unsigned char* ring = warp_base + (size_t)p.nslots * C::SLOT_DOUBLES * sizeof(double);
    const long long N = p.npatt;

    const int* prog = p.prog_in_const ? c_prog : p.prog;
    const int nops = prog[0], ntip = prog[2], nmat = prog[4];
    const int* tip_order = prog + PROG_HDR;
    const int* matseq = prog + prog[5];
    const int* ops = prog + prog[3];

    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; b++) { mbar_init(&full_bar[b], 1); mbar_init(&empty_bar[b], (unsigned)nw); }
        *reinterpret_cast<unsigned int*>(empty_bar + NBUF) = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    const long long r0 = (long long)blockIdx.x * pp.rows_per_cta;
    const long long r1 = (r0 + pp.rows_per_cta < N) ? r0 + pp.rows_per_cta : N;
    if (r0 >= N) return;
    const int chunk_rows = nw * C::ROWS;
    const int nchunks = (int)((r1 - r0 + chunk_rows - 1) / chunk_rows);
    const long long nitems = (long long)nchunks * p.nclass;
    const long long total_loads = nitems * nmat;
    long long s = 0;                    // matrices consumed so far by this warp
    unsigned int* issued_ctr = reinterpret_cast<unsigned int*>(empty_bar + NBUF);   // loads issued so far (CTA-wide)

    // Producer duty is shared: whichever warp passes by (lane 0 only) issues the next load as soon as its ring
    // buffer is free -- in particular the warp whose release() frees a buffer refills it at once.  The claim is an
    // atomicCAS on a shared counter, so every load is issued exactly once and in order.
    auto try_produce = [&]() {
        for (;;) {
            const unsigned int i = *reinterpret_cast<volatile unsigned int*>(issued_ctr);
            if ((long long)i >= total_loads) return;
            const int b = (int)(i % NBUF);
            const unsigned int use = i / NBUF;
            if (use > 0 && !mbar_try_wait(&empty_bar[b], (use - 1) & 1)) return;       // consumers of the previous use not done
            if (atomicCAS(issued_ctr, i, i + 1) != i) continue;                            // somebody else took it
            const long long it = (long long)i / nmat;
            const int ci = (int)((long long)i - it * nmat);
            const int cls_i = (int)(it / nchunks);
            const double* src = p.Pbuf + ((size_t)cls_i * p.nnode + matseq[ci]) * p.pstride;
            mbar_arrive_expect_tx(&full_bar[b], (unsigned)pp.buf_bytes);
            tma_load_1d(bufs + (size_t)b * pp.buf_bytes, src, (unsigned)pp.buf_bytes, &full_bar[b]);
        }
    };
    // a warp takes matrix s: wait until it has landed (helping to issue loads meanwhile); returns its address
    auto acquire = [&]() -> const unsigned char* {
        void* fb = &full_bar[s % NBUF];
        const unsigned parity = (unsigned)((s / NBUF) & 1);
        if (lane == 0) { while (!mbar_try_wait(fb, parity)) try_produce(); }
        __syncwarp();
        mbar_wait(fb, parity);
        return bufs + (size_t)(s % NBUF) * pp.buf_bytes;
    };
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) { mbar_arrive(&empty_bar[s % NBUF]); try_produce(); }
        s++;
    };

    for (long long item = 0; item < nitems; item++) {
        const int cls = (int)(item / nchunks);
        const int chunk = (int)(item - (long long)cls * nchunks);
        const long long h0 = r0 + (long long)chunk * chunk_rows + (long long)warp * C::ROWS;
        const bool active = (h0 < r1);
        if (!active) {
            // nothing to compute in this chunk, but keep the buffer ring moving: take and give back every matrix
            for (int m = 0; m < nmat; m++) { acquire(); release(); }
            continue;
        }
        const unsigned char* zt = p.z + h0;
        double lnsc[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) lnsc[mt] = 0.0;

#pragma unroll
        for (int i = 0; i < C::RING; i++) {
            if (i < ntip && lane < C::ROWS / 4) cp_async4(ring + i * C::ROWS + 4 * lane, zt + (size_t)tip_order[i] * p.zstride + 4 * lane);
            cp_async_commit();
        }
        int ti = 0;

        const int* pc = ops;
        double acc[MT][NT][2];
        for (int op = 0; op < nops; op++) {
            const int nchild = pc[0], out_slot = pc[1], sidx = pc[2], flags = pc[4];
            pc += OP_HDR;
            double tmp[MT][NT][2];
            int c = 0;
            bool started = false;

            if (flags & OPF_HAS_DIRECT) {
                // child 0 is still in `acc`; physical column order makes acc[mt][nt][e] the A element of k-step 2nt+e
                const double2* B = reinterpret_cast<const double2*>(acquire()) + lane;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) tmp[mt][nt][0] = tmp[mt][nt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    double a[MT], b[2 * C::NTP];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) a[mt] = acc[mt][ks >> 1][ks & 1];
#pragma unroll
                    for (int pr = 0; pr < C::NTP; pr++) {
                        double2 v = B[(ks * C::NTP + pr) * 32];
                        b[2 * pr] = v.x; b[2 * pr + 1] = v.y;
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; nt++)
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) dmma884(tmp[mt][nt][0], tmp[mt][nt][1], a[mt], b[nt]);
                }
                release();
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) { acc[mt][nt][0] = tmp[mt][nt][0]; acc[mt][nt][1] = tmp[mt][nt][1]; }
                started = true;
                c = 1;
                pc += CH_REC;
            }

            for (; c < nchild; c++) {
                const int kind = pc[0], idx = pc[1];
                pc += CH_REC;
                if (kind == CH_TIP) {
                    // acc[h][j] (*)= PT[z_h][j]  (src/codeml.c:3555-3567), rows gathered from the shared-memory tile
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
                            cp_async4(ring + (ti % C::RING) * C::ROWS + 4 * lane, zt + (size_t)tip_order[nx] * p.zstride + 4 * lane);
                        cp_async_commit();
                    }
                    ti++;
                    const unsigned char* Pm = acquire();
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        const double2* row = reinterpret_cast<const double2*>(Pm + (size_t)code[mt] * C::NP * sizeof(double));
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            double2 v = row[nt * 4 + q];
                            if (started) { acc[mt][nt][0] *= v.x; acc[mt][nt][1] *= v.y; }
                            else { acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y; }
                        }
                    }
                    release();
                    started = true;
                }
                else if (kind == CH_INTERNAL) {
                    // a parked child: A fragments from the warp's slot, B fragments from the shared-memory tile
                    const double* S = slots + (size_t)idx * C::SLOT_DOUBLES;
                    const double2* B = reinterpret_cast<const double2*>(acquire()) + lane;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) tmp[mt][nt][0] = tmp[mt][nt][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) {
                        double a[MT], b[2 * C::NTP];
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) a[mt] = S[(8 * mt + g) * C::STRIDE + ((8 * (ks >> 1) + 2 * q + (ks & 1)) ^ xsw)];
#pragma unroll
                        for (int pr = 0; pr < C::NTP; pr++) {
                            double2 v = B[(ks * C::NTP + pr) * 32];
                            b[2 * pr] = v.x; b[2 * pr + 1] = v.y;
                        }
#pragma unroll
                        for (int nt = 0; nt < NT; nt++)
#pragma unroll
                            for (int mt = 0; mt < MT; mt++) dmma884(tmp[mt][nt][0], tmp[mt][nt][1], a[mt], b[nt]);
                    }
                    release();
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            if (started) { acc[mt][nt][0] *= tmp[mt][nt][0]; acc[mt][nt][1] *= tmp[mt][nt][1]; }
                            else { acc[mt][nt][0] = tmp[mt][nt][0]; acc[mt][nt][1] = tmp[mt][nt][1]; }
                        }
                    started = true;
                }
                else {
                    // running product of a folded multifurcation (no P(t) involved)
                    const double* S = slots + (size_t)idx * C::SLOT_DOUBLES;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            double2 v = *reinterpret_cast<const double2*>(S + (8 * mt + g) * C::STRIDE + 8 * nt + 2 * q);
                            const double v0 = xsw ? v.y : v.x, v1 = xsw ? v.x : v.y;
                            if (started) { acc[mt][nt][0] *= v0; acc[mt][nt][1] *= v1; }
                            else { acc[mt][nt][0] = v0; acc[mt][nt][1] = v1; }
                        }
                    started = true;
                }
            }

            if (sidx >= 0) {
                // NodeScale (src/treesub.c:7213-7227)
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
                            acc[mt][nt][0] = (8 * nt + q     < p.n) ? 1.0 : 0.0;
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
                // root (src/treesub.c:7780-7798, 7726-7750)
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
                    if (q == 0 && h < r1) {
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
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

}  // namespace pb200
