// cond_dmma.cuh -- the kernels under `method = 1` (cond.cuh) on the FP64 tensor path, for alphabets with more than 4 states.
//
//   cond_node_dmma_kernel   one internal node's conditional vectors from its sons' (ConditionalPNode src/codeml.c:3526-3582 for
//                           one node, NodeScale src/treesub.c:7200-7230): a warp owns 8 site patterns of one class; an internal
//                           son's vectors are read from HBM straight into DMMA A fragments (lane (g, q) of k-step ks holds
//                           L_son[h0 + g][4 ks + q]), contracted with the son's P(t) in B-fragment order (pt_kernel, logical
//                           state order), and the sons' contributions are multiplied in the C-fragment registers; tips are row
//                           gathers from PT[code][j] as in the evaluation kernel.  Vectors stay in the reference's layout
//                           [class][internal node][pattern][state].
//   branch_dmma_kernel      f, f', f'' of lfunt / lfuntdd (+ _SiteClass), src/treesub.c:8127-8543: per class three contractions
//                           L_a . [P | P' | P'']^T on DMMA, then the dot with pi_i L_b(i) in C-fragment layout, the per-class scale
//                           bookkeeping of lfuntdd_SiteClass, -sum fpatt log f and its derivatives; one partial triple per warp
//                           in a fixed tile -> warp assignment, summed in order by cond_sum_kernel (bit-reproducible).
// Both are streaming kernels: 2 n^2 (6 n^2) flops per pattern and son against 8 n bytes read -- at n = 61 just on the compute
// side of the ridge (7.6 flop/byte x 6.5 TB/s = 50 TFLOP/s > 37 DMMA peak), so they run near both roofs.
#pragma once
#include "kernels.cuh"
#include "cond.cuh"

namespace pb200 {

struct CondDmmaParams {
    int n, nclass, nnode, ns, nson, node, scale_k, nscale;
    long long npatt, zstride, pstride, ntiles;
    const int* sons;                 // [nson] node ids (device)
    const double* P;                 // [nclass][nnode][pstride]: tip branch PT[code][8 NT] rows, internal branch B fragments (logical order)
    const unsigned char* z;          // [ns][zstride] compact codes
    double* cond;                    // [nclass][nnode - ns][npatt][n]
    double* scaleF;                  // [nclass][nscale][npatt]
};

template <int NT, int KS>
__global__ void __launch_bounds__(128, 4) cond_node_dmma_kernel(const CondDmmaParams p)
{
    constexpr int NP = 8 * NT, NTP = (NT + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int n = p.n;
    const long long N = p.npatt;
    const size_t nint = (size_t)(p.nnode - p.ns);
    const long long nitems = p.ntiles * p.nclass;
    for (long long item = (long long)blockIdx.x * 4 + warp; item < nitems; item += (long long)gridDim.x * 4) {
        const int cls = (int)(item / p.ntiles);
        const long long h = (item - (long long)cls * p.ntiles) * 8 + g;
        const long long hc = h < N ? h : N - 1;                 // dead rows walk a valid pattern, nothing is written
        double acc[NT][2];
        for (int s = 0; s < p.nson; s++) {
            const int son = __ldg(p.sons + s);
            const double* Pm = p.P + ((size_t)cls * p.nnode + son) * p.pstride;
            double t[NT][2];
            if (son < p.ns) {                                   // tip: src/codeml.c:3555-3567 (ambiguity sums are in the rows)
                const int code = p.z[(size_t)son * p.zstride + hc];
                const double2* row = reinterpret_cast<const double2*>(Pm + (size_t)code * NP);
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { const double2 v = __ldg(row + nt * 4 + q); t[nt][0] = v.x; t[nt][1] = v.y; }
            }
            else {                                              // internal son: src/codeml.c:3568-3575 on DMMA
                const double* src = p.cond + (((size_t)cls * nint + (son - p.ns)) * N + hc) * n;
                double a[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ks++) { const int k = 4 * ks + q; a[ks] = (k < n) ? __ldg(src + k) : 0.0; }
#pragma unroll
                for (int nt = 0; nt < NT; nt++) t[nt][0] = t[nt][1] = 0.0;
                const double2* B = reinterpret_cast<const double2*>(Pm) + lane;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    double b[2 * NTP];
#pragma unroll
                    for (int pr = 0; pr < NTP; pr++) { const double2 v = __ldg(B + (ks * NTP + pr) * 32); b[2 * pr] = v.x; b[2 * pr + 1] = v.y; }
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) dmma884(t[nt][0], t[nt][1], a[ks], b[nt]);
                }
            }
            if (s == 0) {
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { acc[nt][0] = t[nt][0]; acc[nt][1] = t[nt][1]; }
            }
            else {
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { acc[nt][0] *= t[nt][0]; acc[nt][1] *= t[nt][1]; }
            }
        }
        if (p.scale_k >= 0) {                                   // NodeScale, src/treesub.c:7214-7228 (padded columns are 0)
            double mx = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) { mx = fmax(mx, acc[nt][0]); mx = fmax(mx, acc[nt][1]); }
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            double f;
            if (mx < 1e-300) {
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { acc[nt][0] = 1.0; acc[nt][1] = 1.0; }
                f = -800.0;
            }
            else {
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { acc[nt][0] /= mx; acc[nt][1] /= mx; }
                f = log(mx);
            }
            if (q == 0 && h < N) p.scaleF[((size_t)cls * p.nscale + p.scale_k) * N + h] = f;
        }
        if (h < N) {
            double* dst = p.cond + (((size_t)cls * nint + (p.node - p.ns)) * N + h) * n;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const int j = 8 * nt + 2 * q;
                if (j < n) dst[j] = acc[nt][0];
                if (j + 1 < n) dst[j + 1] = acc[nt][1];
            }
        }
    }
}

struct BranchDmmaParams {
    int n, nclass, nnode, ns, a, b, nscale, semantics, want_dd;
    long long npatt, zstride, ntiles, p3stride;
    const double* P3;          // [nclass][3][p3stride]: P, P', P'' of the branch as B fragments (logical order): D[h][i] = sum_j L_a[h][j] M[i][j]
    const double* cond;        // [nclass][nint][npatt][n]
    const double* scaleF;      // [nclass][nscale][npatt]
    const double* freqK;       // [nclass]
    const double* pi;          // [n]
    const double* fpatt;       // [npatt]
    const unsigned char* z;
    const unsigned char* amb_n;
    const unsigned char* amb_map;
    double* partial;           // [gridDim.x * 4][3]: one triple per warp
    const NewtonState* st;     // device-side iteration: a, b, want_dd come from the state
};

template <int NT, int KS>
__global__ void __launch_bounds__(128, 3) branch_dmma_kernel(const BranchDmmaParams p0)
{
    constexpr int NTP = (NT + 1) / 2;
    BranchDmmaParams p = p0;
    if (p.st) { p.a = p.st->a; p.b = p.st->b; p.want_dd = p.st->want_dd; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int n = p.n, nm = p.want_dd ? 3 : 1;
    const long long N = p.npatt;
    const size_t nint = (size_t)(p.nnode - p.ns);
    const bool btip = p.b < p.ns;
    double sl = 0.0, sdl = 0.0, sddl = 0.0;                     // this lane's running -lnL, d/dt, d2/dt2 (rows g of its tiles; q == 0 lanes)
    for (long long tile = (long long)blockIdx.x * 4 + warp; tile < p.ntiles; tile += (long long)gridDim.x * 4) {
        const long long h = tile * 8 + g;
        const long long hc = h < N ? h : N - 1;
        // scale bookkeeping of lfuntdd_SiteClass (src/treesub.c:8437-8449): Sh = largest per-class sum of scale factors
        double Sh = 0.0;
        if (p.nscale > 0)
            for (int ir = 0; ir < p.nclass; ir++) {
                double y = 0.0;
                for (int k = 0; k < p.nscale; k++) y += p.scaleF[((size_t)ir * p.nscale + k) * N + hc];
                if (ir == 0 || y > Sh) Sh = y;
            }
        unsigned long long allowed = 0ull;                       // b is a tip: the states its code allows (src/treesub.c:8251-8256)
        if (btip) {
            const int code = p.z[(size_t)p.b * p.zstride + hc];
            if (code < n) allowed = 1ull << code;
            else {
                const unsigned char* mp = p.amb_map + (size_t)(code - n) * 64;
                const int cnt = p.amb_n[code - n];
                for (int k = 0; k < cnt; k++) allowed |= 1ull << mp[k];
            }
        }
        double fm[3] = {0.0, 0.0, 0.0};
        for (int cls = 0; cls < p.nclass; cls++) {
            double pK = 1.0;
            if (p.semantics == BR_SEM_SITECLASS) {
                pK = p.freqK[cls];
                if (p.nscale > 0) {
                    double y = 0.0;
                    for (int k = 0; k < p.nscale; k++) y += p.scaleF[((size_t)cls * p.nscale + k) * N + hc];
                    pK = pK * exp(y - Sh);
                }
            }
            // w[nt][e] = pi_i L_b[h][i] for i = 8 nt + 2 q + e (C-fragment positions), 0 on the padding
            double w[NT][2];
            if (btip) {
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
#pragma unroll
                    for (int e = 0; e < 2; e++) { const int i = 8 * nt + 2 * q + e; w[nt][e] = (i < n && ((allowed >> i) & 1ull)) ? __ldg(p.pi + i) : 0.0; }
            }
            else {
                const double* vb = p.cond + (((size_t)cls * nint + (p.b - p.ns)) * N + hc) * n;
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
#pragma unroll
                    for (int e = 0; e < 2; e++) { const int i = 8 * nt + 2 * q + e; w[nt][e] = (i < n) ? __ldg(p.pi + i) * __ldg(vb + i) : 0.0; }
            }
            const double* va = p.cond + (((size_t)cls * nint + (p.a - p.ns)) * N + hc) * n;
            double a[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) { const int k = 4 * ks + q; a[ks] = (k < n) ? __ldg(va + k) : 0.0; }
            for (int m = 0; m < nm; m++) {
                double t[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) t[nt][0] = t[nt][1] = 0.0;
                const double2* B = reinterpret_cast<const double2*>(p.P3 + ((size_t)cls * 3 + m) * p.p3stride) + lane;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    double b[2 * NTP];
#pragma unroll
                    for (int pr = 0; pr < NTP; pr++) { const double2 v = __ldg(B + (ks * NTP + pr) * 32); b[2 * pr] = v.x; b[2 * pr + 1] = v.y; }
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) dmma884(t[nt][0], t[nt][1], a[ks], b[nt]);
                }
                double sdot = 0.0;                               // fixed order over the row's states held by this lane, then the quad
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { sdot = fma(w[nt][0], t[nt][0], sdot); sdot = fma(w[nt][1], t[nt][1], sdot); }
                sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
                sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
                fm[m] += pK * sdot;
            }
        }
        if (q == 0 && h < N) {                                   // per pattern -lnL and its derivatives (src/treesub.c:8531-8540)
            const double fp = p.fpatt[h];
            if (fp != 0.0) {
                const double fh = fm[0], dfh = fm[1], ddfh = fm[2];
                const double lf = (p.semantics == BR_SEM_LFUNT && !p.want_dd && fh < 0) ? -500.0 : log(fh);   // src/treesub.c:8192-8193
                double l = -(lf * fp);
                if (p.nscale > 0) l -= Sh * fp;
                sl += l;
                if (p.want_dd) {
                    sdl += -(dfh / fh * fp);
                    sddl += -((fh * ddfh - dfh * dfh) / (fh * fh) * fp);
                }
            }
        }
    }
    // rows g = 0..7 of the warp (lanes with q == 0), butterfly in a fixed shape
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        sl += __shfl_xor_sync(0xffffffffu, sl, off);
        sdl += __shfl_xor_sync(0xffffffffu, sdl, off);
        sddl += __shfl_xor_sync(0xffffffffu, sddl, off);
    }
    if (lane == 0) {
        double* o = p.partial + ((size_t)blockIdx.x * 4 + warp) * 3;
        o[0] = sl; o[1] = sdl; o[2] = sddl;
    }
}

// P, P', P'' of the branch under the Newton iteration written as DMMA B fragments (logical state order): element
// ((ks NTP + pr) 32 + lane) 2 + e  =  M[i][j],  i = 8 (2 pr + e) + lane / 4,  j = 4 ks + lane % 4; zero on the padding.
// Arithmetic exactly as pt3_kernel (cond.cuh), i.e. as lfuntdd forms the three matrices.
struct Pt3FragParams {
    int n, kind, nr, nclass, NT, KS;
    long long set_stride, p3stride;
    const double* sets;
    const int* set_of_class;
    const double* mult;
    const NewtonState* st;    // device-side iteration: t is st->t_eval
    double t;
    double* out;              // [nclass][3][p3stride]
};

__global__ void __launch_bounds__(256) pt3_frag_kernel(const Pt3FragParams p)
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
    const int NTP = (p.NT + 1) / 2;
    double* o = p.out + (size_t)cls * 3 * p.p3stride;
    const int total = p.KS * NTP * 64;
    // grid.y blocks share the matrix: one element (and its n-term sum) per thread -- a single block per class made this kernel
    // the longest of a Newton step on small data (100 us at n = 61 whatever the number of patterns)
    for (int e0 = blockIdx.y * 256 + tid; e0 < total; e0 += gridDim.y * 256) {
        const int e = e0 & 1, lane = (e0 >> 1) & 31, f = e0 >> 6, pr = f % NTP, ks = f / NTP;
        const int i = 8 * (2 * pr + e) + (lane >> 2), j = 4 * ks + (lane & 3);
        double P = 0.0, dP = 0.0, ddP = 0.0;
        if (i < n && j < n && 2 * pr + e < p.NT) {
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
            else {                                              // Cijk form, src/treesub.c:8280-8287
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
        }
        o[e0] = P; o[p.p3stride + e0] = dP; o[2 * p.p3stride + e0] = ddP;
    }
}

}  // namespace pb200
