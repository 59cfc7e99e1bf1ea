// standardize=True: per-group whitening of a Gram matrix (reference _lasso.py:249-252 and,
// for the ridged variant, :776-789).
//
// The reference penalises ||X_g b_g||_2 (or ||sqrtm(X_g^T X_g + sqrt(delta_g) I) b_g||_2)
// instead of ||b_g||_2.  With A_g = G_gg (+ sqrt(delta_g) I) = R_g^T R_g (Cholesky, R_g upper
// triangular) the norm is ||R_g b_g||, so in the variables gamma_g = R_g b_g the problem is
// a plain group Lasso on the design X W, W = blockdiag(R_g^{-1}), i.e. on the Gram
// W^T G W.  The ridge 1/2 delta_g ||b_g||^2 = 1/2 delta_g ||W_g gamma_g||^2 is smooth and is
// folded into the Gram (+ n delta_g W_g^T W_g on the diagonal block, the Gram of sqrt(n delta_g)
// W_g appended to the design), so the solver sees a pure group Lasso.  All of it is O(p^2 m)
// (m = group size) HBM/L2-bound work on CUDA cores; the factorisation of a group runs in one
// CTA on its m x m block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slm {

// group of feature j (groups contiguous: gptr[g] <= j < gptr[g+1])
__device__ __forceinline__ int find_group(const int* __restrict__ gptr, int Gn, int j) {
    int lo = 0, hi = Gn - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(gptr + mid) <= j)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// One CTA per group: R = chol(A_g) (upper, A = R^T R) in `scratch`, then W_g = R^{-1} (upper
// triangular, strictly lower part zeroed) in `W`; both m x m row-major at wptr[g].
// info[0] counts groups whose block is not numerically positive definite.
__global__ void __launch_bounds__(256) group_chol_inv_kernel(const double* __restrict__ G, long long pa,
                                                             const int* __restrict__ gptr,
                                                             const long long* __restrict__ wptr,
                                                             const double* __restrict__ shift,
                                                             double* __restrict__ W, double* __restrict__ scratch,
                                                             int* __restrict__ info) {
    const int g = blockIdx.x;
    const int g0 = gptr[g], m = gptr[g + 1] - g0;
    if (m <= 0) return;
    double* R = scratch + wptr[g];
    double* Wi = W + wptr[g];
    const double sh = shift ? shift[g] : 0.0;
    __shared__ double s_piv;
    __shared__ double s_dmax;
    __shared__ int s_bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < m * m; e += nt) {
        const int i = e / m, j = e - i * m;
        R[e] = G[(long long)(g0 + i) * pa + g0 + j] + (i == j ? sh : 0.0);
    }
    if (tid == 0) {
        s_bad = 0;
        s_dmax = 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double d = 0.0;
        for (int i = 0; i < m; ++i) d = fmax(d, R[i * m + i]);
        s_dmax = d;
    }
    __syncthreads();
    const double tol = s_dmax * 1e-13 * (double)m;
    for (int k = 0; k < m; ++k) {
        if (tid == 0) {
            const double piv = R[k * m + k];
            if (!(piv > tol)) s_bad = 1;
            s_piv = piv > tol ? sqrt(piv) : 1.0;
        }
        __syncthreads();
        const double rkk = s_piv;
        for (int j = k + tid; j < m; j += nt) R[k * m + j] = (j == k) ? rkk : R[k * m + j] / rkk;
        __syncthreads();
        // trailing update of the upper triangle: A[i][j] -= R[k][i] R[k][j], k < i <= j
        const int t = m - k - 1;
        for (int e = tid; e < t * t; e += nt) {
            const int i = k + 1 + e / t, j = k + 1 + e % t;
            if (j >= i) R[i * m + j] -= R[k * m + i] * R[k * m + j];
        }
        __syncthreads();
    }
    // W = R^{-1}: column j by back substitution (columns are independent)
    for (int j = tid; j < m; j += nt) {
        for (int i = m - 1; i > j; --i) Wi[i * m + j] = 0.0;
        Wi[j * m + j] = 1.0 / R[j * m + j];
        for (int i = j - 1; i >= 0; --i) {
            double acc = 0.0;
            for (int k = i + 1; k <= j; ++k) acc += R[i * m + k] * Wi[k * m + j];
            Wi[i * m + j] = -acc / R[i * m + i];
        }
    }
    if (tid == 0 && s_bad) atomicAdd(info, 1);
}

// M = blockdiag(W)^T G on rows < p (rows >= p copied): M[r][j] = sum_{k' <= r'} W_g[k'][r'] G[g0+k'][j]
__global__ void __launch_bounds__(256) whiten_left_kernel(const double* __restrict__ G, long long pa, int p,
                                                          const int* __restrict__ gptr,
                                                          const long long* __restrict__ wptr, int Gn,
                                                          const double* __restrict__ W, double* __restrict__ M) {
    const int r = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= pa) return;
    if (r >= p) {
        M[(long long)r * pa + j] = G[(long long)r * pa + j];
        return;
    }
    const int g = find_group(gptr, Gn, r);
    const int g0 = __ldg(gptr + g), m = __ldg(gptr + g + 1) - g0, rl = r - g0;
    const double* Wg = W + __ldg(wptr + g);
    double acc = 0.0;
    for (int k = 0; k <= rl; ++k) acc += __ldg(Wg + (long long)k * m + rl) * G[(long long)(g0 + k) * pa + j];
    M[(long long)r * pa + j] = acc;
}

// Gout = M blockdiag(W) on columns < p (columns >= p copied), plus the folded ridge
// ridge_scale * ridge[g] * (W_g^T W_g) on the diagonal blocks.
__global__ void __launch_bounds__(256) whiten_right_kernel(const double* __restrict__ M, long long pa, int p,
                                                           const int* __restrict__ gptr,
                                                           const long long* __restrict__ wptr, int Gn,
                                                           const double* __restrict__ W,
                                                           const double* __restrict__ ridge, double ridge_scale,
                                                           double* __restrict__ Gout) {
    const int i = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= pa) return;
    const double* Mi = M + (long long)i * pa;
    if (c >= p) {
        Gout[(long long)i * pa + c] = Mi[c];
        return;
    }
    const int g = find_group(gptr, Gn, c);
    const int g0 = __ldg(gptr + g), m = __ldg(gptr + g + 1) - g0, cl = c - g0;
    const double* Wg = W + __ldg(wptr + g);
    double acc = 0.0;
    for (int k = 0; k <= cl; ++k) acc += Mi[g0 + k] * __ldg(Wg + (long long)k * m + cl);
    if (ridge && i >= g0 && i < g0 + m) {
        const double rg = ridge_scale * __ldg(ridge + g);
        if (rg != 0.0) {
            const int il = i - g0;
            const int kmax = il < cl ? il : cl;
            double ww = 0.0;
            for (int k = 0; k <= kmax; ++k) ww += __ldg(Wg + (long long)k * m + il) * __ldg(Wg + (long long)k * m + cl);
            acc += rg * ww;
        }
    }
    Gout[(long long)i * pa + c] = acc;
}

// b_g = W_g gamma_g per grid column: B[r][k] = sum_{c' >= r'} W_g[r'][c'] Bg[g0+c'][k]
__global__ void __launch_bounds__(128) unwhiten_coef_kernel(const double* __restrict__ Bg, int p, long long ldz,
                                                            int K, const int* __restrict__ gptr,
                                                            const long long* __restrict__ wptr, int Gn,
                                                            const double* __restrict__ W, double* __restrict__ B) {
    const int r = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p || k >= K) return;
    const int g = find_group(gptr, Gn, r);
    const int g0 = __ldg(gptr + g), m = __ldg(gptr + g + 1) - g0, rl = r - g0;
    const double* Wg = W + __ldg(wptr + g);
    double acc = 0.0;
    for (int c = rl; c < m; ++c) acc += __ldg(Wg + (long long)rl * m + c) * Bg[(long long)(g0 + c) * ldz + k];
    B[(long long)r * ldz + k] = acc;
}

}  // namespace slm
