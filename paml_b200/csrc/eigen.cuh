// eigen.cuh -- batched spectral decomposition of reversible rate matrices on the device.
//
// Replaces eigenQREV (src/tools.c:5023-5110): Q = S diag(pi) is symmetrised to A = pi^{1/2} Q pi^{-1/2}, A = R diag(Root) R^T
// is found, and U = pi^{-1/2} R, V = R^T pi^{1/2} are returned, states with pi <= 1e-100 being dropped from the
// problem and re-inserted as identity rows/columns exactly as the reference does (src/tools.c:5055-5102).  The
// reference solves the symmetric problem with Householder + implicit QL (eigenRealSym, src/tools.c:5116-5309), a
// strictly sequential algorithm; here ONE CTA per matrix runs a parallel-order cyclic two-sided Jacobi iteration:
// a round-robin round has m/2 disjoint pivot pairs; their rotations are computed by m/2 threads into shared memory,
// then thread (bi, bj), bi <= bj, applies rotation bi from the left and rotation bj from the right to the 2x2 block
// in rows of pair bi and columns of pair bj (and writes its mirror), reading matrix A_cur and writing A_next, while
// the eigenvector columns are rotated in place by all threads: two barriers per round, m-1 rounds per sweep.  The
// rotation is computed with two FP64 reciprocal square roots and no division (it is the round's critical path).
// Jacobi delivers eigenvalues to working precision and orthogonal eigenvectors, so P(t) = U exp(t Root) V agrees
// with the reference's to rounding (the decomposition itself is unique only up to order and sign: eigenvalues are
// sorted in decreasing order like EigenSort, src/tools.c:5136-5158).  Several matrices (site classes, branch types,
// branches) run as several CTAs of one launch; results are written in the UVROOT set layout pt_kernel reads.
#pragma once
#include <cuda_runtime.h>

namespace pb200 {

struct EigenParams {
    int n;                   // states (<= 64)
    const double* Q;         // [njob][n*n] row-major rate matrices
    const double* pi;        // [njob][n]
    double* sets;            // [njob][set_stride]: Root[n], U[n*n], V[n*n]
    long long set_stride;
    int* info;               // [njob] sweeps used (negative: not converged)
    int max_sweeps;
    // Warm start (optional): the eigenvectors of job j's PREVIOUS solution, [njob][EIG_WARM_DOUBLES] = {tag, V column-major m x (m+1)}.
    // use_warm: start the iteration from them (A' = R0^T A R0 is then nearly diagonal when the matrix changed little -- the optimiser
    // perturbs one parameter at a time -- and two sweeps instead of seven or eight suffice); the final vectors are always stored back.
    double* warm;
    int use_warm;
};
constexpr int EIG_WARM_DOUBLES = 8 + 64 * 65;

constexpr int EIG_THREADS = 1024;
constexpr int EIG_MAXM = 64;

__global__ void __launch_bounds__(EIG_THREADS, 1) eigen_qrev_kernel(const EigenParams p)
{
    extern __shared__ __align__(16) double esm[];
    __shared__ int s_idx[EIG_MAXM];
    __shared__ int s_rank[EIG_MAXM];
    __shared__ double s_ps[EIG_MAXM];
    __shared__ double s_c[EIG_MAXM / 2], s_s[EIG_MAXM / 2];
    __shared__ int s_p[EIG_MAXM / 2], s_q[EIG_MAXM / 2];
    __shared__ double s_red[EIG_THREADS / 32];
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

    // A = pi^{1/2} Q pi^{-1/2}, built from the lower triangle of Q like the reference (src/tools.c:5047-5049, 5056-5066);
    // an odd problem is padded with one zero row/column, which no rotation ever mixes with the rest
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
        V[i * LD + j] = (i == j) ? 1.0 : 0.0;                 // eigenvectors, COLUMN-major: V[k * LD + i] = component i of vector k
    }
    __syncthreads();

    double* wbuf = p.warm ? p.warm + (size_t)job * EIG_WARM_DOUBLES : nullptr;
    // the tag of a stored basis: dimension of the reduced problem and of the alphabet (a basis of another shape is ignored)
    const double tag = (double)(nnew * 1024 + n);
    if (wbuf && p.use_warm && wbuf[0] == tag && m >= 2) {
        // V = R0;  T = A R0 (into A1);  A' = R0^T T (into A0, then copied to A1).  Jacobi then works on A' and keeps rotating V = R0 J1 J2 ...
        const double* R0 = wbuf + 8;
        for (int e = tid; e < m * LD; e += EIG_THREADS) V[e] = R0[e];
        __syncthreads();
        for (int e = tid; e < m * m; e += EIG_THREADS) {
            const int i = e / m, k = e - i * m;
            double t = 0.0;
            for (int j = 0; j < m; j++) t = fma(A0[i * LD + j], V[k * LD + j], t);
            A1[i * LD + k] = t;
        }
        __syncthreads();
        for (int e = tid; e < m * m; e += EIG_THREADS) {
            const int k = e / m, l = e - k * m;
            double t = 0.0;
            for (int i = 0; i < m; i++) t = fma(V[k * LD + i], A1[i * LD + l], t);
            A0[k * LD + l] = t;
        }
        __syncthreads();
        for (int e = tid; e < m * m; e += EIG_THREADS) { const int i = e / m, j = e - i * m; A1[i * LD + j] = A0[i * LD + j]; }
        __syncthreads();
    }

    const int bi = tid >> 5, bj = tid & 31;
    const bool upper = (bi <= bj) && (bj < half);
    const int nplayers = m - 1;
    // this thread's eigenvector items (pair index, component): consecutive lanes take consecutive components -> conflict-free
    //: half * m <= 2048 items over 1024 threads
    int vk0 = -1, vi0 = 0, vk1 = -1, vi1 = 0;
    if (tid < half * m) { vk0 = tid / m; vi0 = tid - vk0 * m; }
    if (tid + EIG_THREADS < half * m) { vk1 = (tid + EIG_THREADS) / m; vi1 = tid + EIG_THREADS - vk1 * m; }

    // block-wide sum of squares of the symmetric matrix held in the upper triangle, with or without its diagonal
    auto sumsq = [&](const double* A, bool with_diag) -> double {
        double loc = 0.0;
        for (int e = tid; e < m * m; e += EIG_THREADS) {
            const int i = e / m, j = e - i * m;
            if (i < j) { const double a = A[i * LD + j]; loc += 2.0 * a * a; }
            else if (i == j && with_diag) { const double a = A[i * LD + j]; loc += a * a; }
        }
        for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = loc;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < EIG_THREADS / 32; w++) tot += s_red[w];
        __syncthreads();
        return tot;
    };
    const double frob2 = sumsq(A0, true);

    int sweeps = 0;
    bool converged = (m < 2) || frob2 == 0.0;
    bool last = false;                       // the sweep that starts below the quadratic-convergence threshold is the last
    if (!converged && wbuf && p.use_warm && wbuf[0] == tag) {
        // a warm start may already be within the quadratic-convergence threshold: then ONE sweep reaches the rounding floor
        const double off0 = sumsq(A0, false);
        if (off0 <= 1e-18 * frob2) last = true;
    }
    double* Ac = A0;
    double* An = A1;
    while (!converged && sweeps < p.max_sweeps) {
        for (int r = 0; r < nplayers; r++) {
            // round-robin pairing of round r: pair 0 = (m-1, r), pair k = (r+k, r-k) mod (m-1)
            if (tid < half) {
                int a, b;
                if (tid == 0) { a = m - 1; b = r; }
                else { a = r + tid; if (a >= nplayers) a -= nplayers; b = r - tid; if (b < 0) b += nplayers; }
                const int pp = a < b ? a : b, qq = a < b ? b : a;
                const double app = Ac[pp * LD + pp], aqq = Ac[qq * LD + qq], apq = Ac[pp * LD + qq];
                double c = 1.0, s = 0.0;
                // inner rotation (|theta| <= pi/4) annihilating a_pq: tan 2theta = 2 a_pq / (a_qq - a_pp).  Formulated with two
                // reciprocal square roots and no division (FP64 div / sqrt are long dependent sequences and this
                // thread is the critical path of the round): cos 2theta = |d| / h, sin 2theta = sgn(d) 2 a_pq / h,
                // h = hypot(d, 2 a_pq); cos theta = sqrt((1 + cos 2theta) / 2), sin theta = sin 2theta / (2 cos theta).
                const double d = aqq - app, two_apq = 2.0 * apq;
                const double h2 = fma(d, d, two_apq * two_apq);
                if (apq != 0.0 && h2 > 1e-290) {            // (below: nothing to rotate in working precision)
                    const double ih = rsqrt(h2);
                    const double cos2 = fabs(d) * ih;
                    const double sin2 = (d < 0 ? -two_apq : two_apq) * ih;
                    const double x = 0.5 + 0.5 * cos2;      // in [0.5, 1]
                    const double ix = rsqrt(x);
                    c = x * ix;
                    s = 0.5 * sin2 * ix;
                }
                s_c[tid] = c; s_s[tid] = s; s_p[tid] = pp; s_q[tid] = qq;
            }
            __syncthreads();
            if (upper) {
                const int p1 = s_p[bi], q1 = s_q[bi], p2 = s_p[bj], q2 = s_q[bj];
                const double c1 = s_c[bi], s1 = s_s[bi], c2 = s_c[bj], s2 = s_s[bj];
                // A is symmetric: only [min][max] is kept up to date (half the stores, no strided mirror writes)
                const int ipu = p1 < p2 ? p1 * LD + p2 : p2 * LD + p1, ipw = p1 < q2 ? p1 * LD + q2 : q2 * LD + p1;
                const int iqu = q1 < p2 ? q1 * LD + p2 : p2 * LD + q1, iqw = q1 < q2 ? q1 * LD + q2 : q2 * LD + q1;
                const double apu = Ac[ipu], apw = Ac[ipw], aqu = Ac[iqu], aqw = Ac[iqw];
                // rows of pair 1 rotated (J^T from the left), then columns of pair 2 (J from the right)
                const double rpu = c1 * apu - s1 * aqu, rqu = s1 * apu + c1 * aqu;
                const double rpw = c1 * apw - s1 * aqw, rqw = s1 * apw + c1 * aqw;
                const double npu = c2 * rpu - s2 * rpw, npw = s2 * rpu + c2 * rpw;
                const double nqu = c2 * rqu - s2 * rqw, nqw = s2 * rqu + c2 * rqw;
                if (bi == bj) {
                    An[ipu] = npu; An[iqw] = nqw;
                    An[ipw] = 0.0;                          // a_pq, annihilated by construction (ipw == [p1][q1])
                }
                else {
                    An[ipu] = npu; An[ipw] = npw; An[iqu] = nqu; An[iqw] = nqw;
                }
            }
            // eigenvector columns of every pair, in place (pairs are disjoint): item = (pair, row), at most two per thread
            if (vk0 >= 0) {
                const int pp = s_p[vk0], qq = s_q[vk0];
                const double c = s_c[vk0], s = s_s[vk0];
                const double vp = V[pp * LD + vi0], vq = V[qq * LD + vi0];
                V[pp * LD + vi0] = c * vp - s * vq;
                V[qq * LD + vi0] = s * vp + c * vq;
            }
            if (vk1 >= 0) {
                const int pp = s_p[vk1], qq = s_q[vk1];
                const double c = s_c[vk1], s = s_s[vk1];
                const double vp = V[pp * LD + vi1], vq = V[qq * LD + vi1];
                V[pp * LD + vi1] = c * vp - s * vq;
                V[qq * LD + vi1] = s * vp + c * vq;
            }
            __syncthreads();
            double* tmp = Ac; Ac = An; An = tmp;
        }
        sweeps++;
        // Jacobi converges quadratically: once the off-diagonal norm is below 1e-9 |A|_F one more sweep takes it to the
        // rounding floor.  (A plain "off <= eps |A|" test would never fire: the floor itself is ~1e-15 |A|_F.)
        const double off2 = sumsq(Ac, false);
        if (last || off2 <= 1e-30 * frob2) converged = true;
        else if (off2 <= 1e-18 * frob2) last = true;
    }

    if (wbuf) {                               // this solution's vectors are the next one's starting basis
        for (int e = tid; e < m * LD; e += EIG_THREADS) wbuf[8 + e] = V[e];
        if (tid == 0) wbuf[0] = converged ? tag : 0.0;
    }
    // eigenvalues in decreasing order (ties: by index), like EigenSort
    if (tid < nnew) {
        const double d = Ac[tid * LD + tid];
        int rk = 0;
        for (int j = 0; j < nnew; j++) { const double dj = Ac[j * LD + j]; if (dj > d || (dj == d && j < tid)) rk++; }
        s_rank[tid] = rk;
    }
    __syncthreads();

    double* set = p.sets + (size_t)job * p.set_stride;
    double* Root = set;
    double* U = set + n;
    double* Vo = set + n + (size_t)n * n;
    // dropped states: Root 0, identity row/column (src/tools.c:5070-5101)
    for (int e = tid; e < n * n; e += EIG_THREADS) { const int i = e / n, j = e - i * n; U[e] = (i == j) ? 1.0 : 0.0; Vo[e] = (i == j) ? 1.0 : 0.0; }
    if (tid < n) Root[tid] = 0.0;
    __syncthreads();
    if (tid < nnew) Root[s_idx[s_rank[tid]]] = Ac[tid * LD + tid];
    for (int e = tid; e < nnew * nnew; e += EIG_THREADS) {
        const int i = e / nnew, k = e - i * nnew;          // component i of eigenvector k
        const double v = V[k * LD + i];
        const int col = s_idx[s_rank[k]];                   // where eigen-pair k goes
        U[s_idx[i] * n + col] = v / s_ps[i];
        Vo[col * n + s_idx[i]] = v * s_ps[i];
    }
    if (tid == 0 && p.info) p.info[job] = converged ? sweeps : -sweeps - 1;
}

}  // namespace pb200
