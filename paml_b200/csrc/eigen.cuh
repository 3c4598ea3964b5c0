// eigen.cuh -- batched spectral decomposition of reversible rate matrices on the device.
//
// Replaces eigenQREV (src/tools.c:5023-5110): Q = S diag(pi) is symmetrised to A = pi^{1/2} Q pi^{-1/2}, A = R diag(Root) R^T
// is found, and U = pi^{-1/2} R, V = R^T pi^{1/2} are returned, states with pi <= 1e-100 being dropped from the
// problem and re-inserted as identity rows/columns exactly as the reference does (src/tools.c:5055-5102).  The
// reference solves the symmetric problem with Householder + implicit QL (eigenRealSym, src/tools.c:5116-5309), a
// strictly sequential algorithm; here ONE CTA per matrix runs a parallel-order cyclic two-sided Jacobi iteration:
// the m/2 disjoint rotations of a round-robin round are applied at once, thread (bi, bj) owning the 2x2 block in
// rows of pair bi and columns of pair bj, double-buffered so that a round costs a single barrier.  Jacobi delivers
// eigenvalues in working precision and orthogonal eigenvectors, so P(t) = U exp(t Root) V agrees with the
// reference's to rounding (the decomposition itself is unique only up to order and sign: eigenvalues are sorted in
// decreasing order like EigenSort, src/tools.c:5136-5158).  Several matrices (site classes, branch types) run as
// several CTAs in the same launch; the sets are written straight into the staging area pt_kernel reads.
#pragma once
#include <cuda_runtime.h>

namespace pb200 {

struct EigenParams {
    int n;                   // states (<= 64)
    const double* Q;         // [njob][n*n] row-major rate matrices
    const double* pi;        // [njob][n]
    const double* scale;     // [njob] factor applied to Root (1/mean rate; 1 = none), src/codeml.c:3314-3321
    const int* job_set;      // [njob] index of the spectral set job j produces
    double* sets;            // [nset][set_stride]: Root[n], U[n*n], V[n*n]
    long long set_stride;
    int* info;               // [njob] sweeps used (negative: not converged)
    int max_sweeps;
};

constexpr int EIG_THREADS = 1024;
constexpr int EIG_MAXM = 64;

__device__ __forceinline__ void jacobi_rotation(double app, double aqq, double apq, double& c, double& s, double& t)
{
    if (apq == 0.0) { c = 1.0; s = 0.0; t = 0.0; return; }
    const double d = aqq - app;
    const double r = sqrt(fma(d, d, 4.0 * apq * apq));
    t = 2.0 * apq / (fabs(d) + r);
    if (d < 0) t = -t;
    c = 1.0 / sqrt(fma(t, t, 1.0));
    s = t * c;
}

__global__ void __launch_bounds__(EIG_THREADS, 1) eigen_qrev_kernel(const EigenParams p)
{
    extern __shared__ __align__(16) double esm[];
    __shared__ int s_idx[EIG_MAXM];
    __shared__ int s_rank[EIG_MAXM];
    __shared__ double s_ps[EIG_MAXM];
    __shared__ double s_red[2 * (EIG_THREADS / 32)];
    __shared__ int s_nnew;

    const int job = blockIdx.x, tid = threadIdx.x, n = p.n;
    const double* Q = p.Q + (size_t)job * n * n;
    const double* pi = p.pi + (size_t)job * n;

    if (tid == 0) {
        int nn = 0;
        for (int j = 0; j < n; j++)
            if (pi[j] > 1e-100) { s_idx[nn] = j; s_ps[nn] = sqrt(pi[j]); nn++; }
        s_nnew = nn;
    }
    __syncthreads();
    const int nnew = s_nnew;
    const int m = nnew + (nnew & 1), half = m / 2, LD = m + 1;
    double* A0 = esm;
    double* A1 = A0 + m * LD;
    double* V = A1 + m * LD;

    // A = pi^{1/2} Q pi^{-1/2}, built from the lower triangle of Q like the reference (src/tools.c:5047-5049, 5056-5066)
    for (int e = tid; e < m * m; e += EIG_THREADS) {
        const int i = e / m, j = e - i * m;
        double a = 0.0;
        if (i < nnew && j < nnew) {
            if (i == j) a = Q[s_idx[i] * n + s_idx[i]];
            else {
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                a = Q[s_idx[hi] * n + s_idx[lo]] * s_ps[hi] / s_ps[lo];
            }
        }
        A0[i * LD + j] = a; A1[i * LD + j] = a;
        V[i * LD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int bi = tid >> 5, bj = tid & 31;
    const bool upper = (bi <= bj) && (bj < half);
    // V work: lower-triangle threads, (half-1)/2 of them per pair, each taking a strided share of the m rows
    const int per_pair = (half - 1) / 2 > 0 ? (half - 1) / 2 : 1;
    int vpair = -1, vsub = 0;
    if (half == 1) { if (tid == 1) { vpair = 0; vsub = 0; } }
    else if (bi > bj && bi < half) {
        const int li = bi * (bi - 1) / 2 + bj;
        if (li < per_pair * half) { vpair = li / per_pair; vsub = li - vpair * per_pair; }
    }
    const int nplayers = m - 1;

    double frob2 = 0.0;
    {
        double loc = 0.0;
        for (int e = tid; e < m * m; e += EIG_THREADS) { const int i = e / m, j = e - i * m; const double a = A0[i * LD + j]; loc += a * a; }
        for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = loc;
        __syncthreads();
        for (int w = 0; w < EIG_THREADS / 32; w++) frob2 += s_red[w];
        __syncthreads();
    }

    int sweeps = 0;
    bool converged = (m < 2) || frob2 == 0.0;
    double* Ac = A0;
    double* An = A1;
    while (!converged && sweeps < p.max_sweeps) {
        for (int r = 0; r < nplayers; r++) {
            // round-robin pairing of round r: pair 0 = (m-1, r), pair k = (r+k, r-k) mod (m-1)
            auto pair_of = [&](int k, int& pp, int& qq) {
                int a, b;
                if (k == 0) { a = m - 1; b = r % nplayers; }
                else { a = (r + k) % nplayers; b = (r - k + 2 * nplayers) % nplayers; }
                pp = a < b ? a : b; qq = a < b ? b : a;
            };
            if (upper) {
                int p1, q1, p2, q2;
                pair_of(bi, p1, q1); pair_of(bj, p2, q2);
                double c1, s1, t1, c2, s2, t2;
                const double app = Ac[p1 * LD + p1], aqq = Ac[q1 * LD + q1], apq = Ac[p1 * LD + q1];
                jacobi_rotation(app, aqq, apq, c1, s1, t1);
                if (bi == bj) {
                    An[p1 * LD + p1] = app - t1 * apq;
                    An[q1 * LD + q1] = aqq + t1 * apq;
                    An[p1 * LD + q1] = 0.0; An[q1 * LD + p1] = 0.0;
                }
                else {
                    jacobi_rotation(Ac[p2 * LD + p2], Ac[q2 * LD + q2], Ac[p2 * LD + q2], c2, s2, t2);
                    const double apu = Ac[p1 * LD + p2], apw = Ac[p1 * LD + q2], aqu = Ac[q1 * LD + p2], aqw = Ac[q1 * LD + q2];
                    const double rpu = c1 * apu - s1 * aqu, rqu = s1 * apu + c1 * aqu;
                    const double rpw = c1 * apw - s1 * aqw, rqw = s1 * apw + c1 * aqw;
                    const double npu = c2 * rpu - s2 * rpw, npw = s2 * rpu + c2 * rpw;
                    const double nqu = c2 * rqu - s2 * rqw, nqw = s2 * rqu + c2 * rqw;
                    An[p1 * LD + p2] = npu; An[p2 * LD + p1] = npu;
                    An[p1 * LD + q2] = npw; An[q2 * LD + p1] = npw;
                    An[q1 * LD + p2] = nqu; An[p2 * LD + q1] = nqu;
                    An[q1 * LD + q2] = nqw; An[q2 * LD + q1] = nqw;
                }
            }
            else if (vpair >= 0) {
                int pp, qq;
                pair_of(vpair, pp, qq);
                double c, s, t;
                jacobi_rotation(Ac[pp * LD + pp], Ac[qq * LD + qq], Ac[pp * LD + qq], c, s, t);
                for (int i = vsub; i < m; i += per_pair) {
                    const double vp = V[i * LD + pp], vq = V[i * LD + qq];
                    V[i * LD + pp] = c * vp - s * vq;
                    V[i * LD + qq] = s * vp + c * vq;
                }
            }
            __syncthreads();
            double* tmp = Ac; Ac = An; An = tmp;
        }
        sweeps++;
        // convergence: off-diagonal Frobenius norm against the whole matrix
        double loc = 0.0;
        for (int e = tid; e < m * m; e += EIG_THREADS) { const int i = e / m, j = e - i * m; if (i != j) { const double a = Ac[i * LD + j]; loc += a * a; } }
        for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = loc;
        __syncthreads();
        double off2 = 0.0;
        for (int w = 0; w < EIG_THREADS / 32; w++) off2 += s_red[w];
        __syncthreads();
        converged = off2 <= 1e-33 * frob2;          // off <= 3e-17 |A|_F
    }

    // eigenvalues in decreasing order (ties: by index), like EigenSort
    if (tid < nnew) {
        const double d = Ac[tid * LD + tid];
        int rk = 0;
        for (int j = 0; j < nnew; j++) { const double dj = Ac[j * LD + j]; if (dj > d || (dj == d && j < tid)) rk++; }
        s_rank[tid] = rk;
    }
    __syncthreads();

    double* set = p.sets + (size_t)p.job_set[job] * p.set_stride;
    double* Root = set;
    double* U = set + n;
    double* Vo = set + n + (size_t)n * n;
    const double sc = p.scale ? p.scale[job] : 1.0;
    // dropped states: Root 0, identity row/column (src/tools.c:5070-5101)
    for (int e = tid; e < n * n; e += EIG_THREADS) { const int i = e / n, j = e - i * n; U[e] = (i == j) ? 1.0 : 0.0; Vo[e] = (i == j) ? 1.0 : 0.0; }
    if (tid < n) Root[tid] = 0.0;
    __syncthreads();
    if (tid < nnew) Root[s_idx[s_rank[tid]]] = Ac[tid * LD + tid] * sc;
    for (int e = tid; e < nnew * nnew; e += EIG_THREADS) {
        const int i = e / nnew, k = e - i * nnew;          // component i of eigenvector k
        const double v = V[i * LD + k];
        const int col = s_idx[s_rank[k]];                   // where eigen-pair k goes
        U[s_idx[i] * n + col] = v / s_ps[i];
        Vo[col * n + s_idx[i]] = v * s_ps[i];
    }
    if (tid == 0 && p.info) p.info[job] = converged ? sweeps : -sweeps - 1;
}

}  // namespace pb200
