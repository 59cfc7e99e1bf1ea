// Second-order phase of the batched solve (ill-conditioned pure-group problems, DESIGN section 2
// item 13): one lock-step damped Newton step on the active groups of k slow columns, entirely in
// this file's kernels -- Hessian assembly straight from the Gram, a batched blocked Cholesky
// factorisation (64-wide panels: diagonal block factored and inverted in shared memory, panel
// through the inverse, trailing update on the FP64 tensor-core GEMM of gemm_f64.cuh), blocked
// triangular solves, and the Armijo line search with objective differences formed without
// cancellation.  (Round 1 did these steps with torch operations and cuSOLVER's potrf.)
//
// On the active manifold of column c (groups with ||x_g|| > 0; identity on the other coordinates)
//     phi(x) = 1/(2n) x'Gx - c'x/n + sum_g w_g ||x_g|| + 1/2 sum_g d_g ||x_g||^2
//     grad   = (Gx - c)/n + w_g x_g/||x_g|| + d_g x_g
//     H      = G_AA/n + blockdiag_g( w_g/||x_g|| (I - u_g u_g') ) + diag(d),   u_g = x_g/||x_g||
// Layouts: X, GX, U, KK, DP, GRAD, DIR are [k][ldv] (one row per column, solver feature order);
// H is [k][ldh][ldh] row-major, lower triangle = the factor after factorisation; group tables W2,
// D2, NRM are [k][Gn].
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace slm {

constexpr int NW_NB = 64;   // panel width of the blocked Cholesky
constexpr int NW_T = 256;   // threads per block of the vector kernels
constexpr size_t NW_TILE_SMEM = 2 * NW_NB * (NW_NB + 1) * sizeof(double);  // two padded 64 x 64 tiles

__device__ __forceinline__ double nw_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = (l < NW_T / 32) ? red[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// per column: group norms, activity, u = x/||x_g||, kk = w_g/||x_g||, dp = d_g (active coordinates)
__global__ void __launch_bounds__(NW_T) newton_prepare_kernel(int p, int Gn, const int* __restrict__ gptr,
                                                              const double* __restrict__ X, long long ldv,
                                                              const double* __restrict__ W2, const double* __restrict__ D2,
                                                              double* __restrict__ NRM, double* __restrict__ U,
                                                              double* __restrict__ KK, double* __restrict__ DP) {
    const int c = blockIdx.x;
    const double* x = X + (long long)c * ldv;
    for (int g = threadIdx.x; g < Gn; g += NW_T) {
        const int ja = gptr[g], jb = gptr[g + 1];
        double ss = 0.0;
        for (int j = ja; j < jb; ++j) ss += x[j] * x[j];
        const double nrm = sqrt(ss);
        NRM[(long long)c * Gn + g] = nrm;
        const double w = W2[(long long)c * Gn + g];
        const double d = D2 ? D2[(long long)c * Gn + g] : 0.0;
        for (int j = ja; j < jb; ++j) {
            const long long e = (long long)c * ldv + j;
            if (nrm > 0.0) {
                U[e] = x[j] / nrm;
                KK[e] = w / nrm;
                DP[e] = d;
            } else {
                U[e] = 0.0;
                KK[e] = -1.0;  // marks an inactive coordinate
                DP[e] = 0.0;
            }
        }
    }
}

// X[c][j] -> Z[fold[c]][j][slot[c]] (the Gram apply's feature-major layout; Z zero-filled by the caller)
__global__ void __launch_bounds__(NW_T) newton_pack_kernel(int p, const int* __restrict__ fold, const int* __restrict__ slot,
                                                           const double* __restrict__ X, long long ldv,
                                                           double* __restrict__ Z, long long ldz) {
    const int c = blockIdx.y;
    const int j = blockIdx.x * NW_T + threadIdx.x;
    if (j >= p) return;
    Z[((long long)fold[c] * p + j) * ldz + slot[c]] = X[(long long)c * ldv + j];
}

// gradient on the active coordinates from GX = G x:  grad = (Gx - c)/n + kk x + dp x  (kk x = w u)
// gs (the quadratic part, stored in GS) is kept for the line search; GZ = G Z in the apply's layout
__global__ void __launch_bounds__(NW_T) newton_grad_kernel(int p, const double* __restrict__ G, long long g_stride,
                                                           long long pa, const int* __restrict__ fold,
                                                           const int* __restrict__ slot, const double* __restrict__ nobs,
                                                           const double* __restrict__ X, const double* __restrict__ GZ,
                                                           long long ldz, long long ldv, const double* __restrict__ KK,
                                                           const double* __restrict__ DP, double* __restrict__ GS,
                                                           double* __restrict__ GRAD) {
    const int c = blockIdx.y;
    const int j = blockIdx.x * NW_T + threadIdx.x;
    if (j >= p) return;
    const long long e = (long long)c * ldv + j;
    const double kk = KK[e];
    if (kk < 0.0) {
        GS[e] = 0.0;
        GRAD[e] = 0.0;
        return;
    }
    const double n = nobs[c];
    const double cj = G[(long long)fold[c] * g_stride + (long long)p * pa + j];
    const double gx = GZ[((long long)fold[c] * p + j) * ldz + slot[c]];  // (G x)_j in the apply's layout
    const double gs = (gx - cj) / n;
    GS[e] = gs;
    GRAD[e] = gs + (kk + DP[e]) * X[e];
}

// H[c][i][j] (all entries; symmetric)
__global__ void __launch_bounds__(NW_T) newton_hessian_kernel(int p, const double* __restrict__ G, long long g_stride,
                                                              long long pa, const int* __restrict__ fold,
                                                              const double* __restrict__ nobs,
                                                              const int* __restrict__ gid, const double* __restrict__ U,
                                                              const double* __restrict__ KK, const double* __restrict__ DP,
                                                              long long ldv, double* __restrict__ H, long long ldh) {
    const int c = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * NW_T + threadIdx.x;
    if (j >= ldh || i >= ldh) return;
    double* Hc = H + (long long)c * ldh * ldh;
    if (i >= p || j >= p) {  // padding rows / columns: identity
        Hc[(long long)i * ldh + j] = (i == j) ? 1.0 : 0.0;
        return;
    }
    const long long ei = (long long)c * ldv + i, ej = (long long)c * ldv + j;
    const double kki = KK[ei], kkj = KK[ej];
    double v = 0.0;
    if (kki >= 0.0 && kkj >= 0.0) {
        v = G[(long long)fold[c] * g_stride + (long long)i * pa + j] / nobs[c];
        if (gid[i] == gid[j]) v -= kki * U[ei] * U[ej];
        if (i == j) v += kki + DP[ei];
    } else if (i == j) {
        v = 1.0;
    }
    Hc[(long long)i * ldh + j] = v;
}

// ---- blocked Cholesky ------------------------------------------------------------------------------
// diagonal block [j0, j0+nb): factor in shared memory, write L (lower) back, invert it, keep the
// inverse in INV[c][panel][NB][NB] (row-major, lower triangular).  info[c] != 0: not positive definite.
__global__ void __launch_bounds__(NW_T) chol_diag_kernel(double* __restrict__ H, long long ldh, int j0, int nb,
                                                         double* __restrict__ INV, int npanels, int panel,
                                                         int* __restrict__ info) {
    extern __shared__ double nw_sh[];
    double (*S)[NW_NB + 1] = reinterpret_cast<double (*)[NW_NB + 1]>(nw_sh);
    double (*V)[NW_NB + 1] = reinterpret_cast<double (*)[NW_NB + 1]>(nw_sh + NW_NB * (NW_NB + 1));
    __shared__ int bad;
    const int c = blockIdx.x, tid = threadIdx.x;
    double* Hc = H + (long long)c * ldh * ldh + (long long)j0 * ldh + j0;
    if (tid == 0) bad = 0;
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) {
        const int i = e / NW_NB, j = e % NW_NB;
        S[i][j] = (i < nb && j < nb && j <= i) ? Hc[(long long)i * ldh + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        const double piv = S[k][k];
        if (!(piv > 0.0)) {
            if (tid == 0) bad = 1;
        }
        __syncthreads();
        const double sq = piv > 0.0 ? sqrt(piv) : 1.0;
        if (tid == 0) S[k][k] = sq;
        for (int i = k + 1 + tid; i < nb; i += NW_T) S[i][k] /= sq;
        __syncthreads();
        // trailing update of the lower triangle: S[i][j] -= S[i][k] S[j][k], k < j <= i < nb
        // (threads as a 16 x 16 patch stepping over the triangle: no integer division in the loop)
        for (int i = k + 1 + (tid >> 4); i < nb; i += 16) {
            const double sik = S[i][k];
            for (int j = k + 1 + (tid & 15); j <= i; j += 16) S[i][j] -= sik * S[j][k];
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += NW_T) {
        const int i = e / nb, j = e % nb;
        if (j <= i) Hc[(long long)i * ldh + j] = S[i][j];
    }
    // V = L^{-1}: column q by forward substitution (thread q), rows sequential
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) V[e / NW_NB][e % NW_NB] = 0.0;
    __syncthreads();
    if (tid < nb) {
        const int q = tid;
        for (int i = q; i < nb; ++i) {
            double s = (i == q) ? 1.0 : 0.0;
            for (int m = q; m < i; ++m) s -= S[i][m] * V[m][q];
            V[i][q] = s / S[i][i];
        }
    }
    __syncthreads();
    double* out = INV + ((long long)c * npanels + panel) * NW_NB * NW_NB;
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) out[e] = V[e / NW_NB][e % NW_NB];
    if (tid == 0 && bad) info[c] = 1;
}

// panel below the diagonal block: L21 = A21 L11^{-T}, i.e. out[r][q] = sum_{m <= q} A[r][m] inv[q][m].
// Written in place and, transposed, into the GEMM operands of the trailing update:
// PT[c][q][r] = out, NPT[c][q][r] = -out (row r counted from the top of the matrix).
__global__ void __launch_bounds__(NW_T) chol_panel_kernel(double* __restrict__ H, long long ldh, int p, int j0, int nb,
                                                          const double* __restrict__ INV, int npanels, int panel,
                                                          double* __restrict__ PT, double* __restrict__ NPT,
                                                          long long ldpt) {
    extern __shared__ double nw_sh[];
    double (*A)[NW_NB + 1] = reinterpret_cast<double (*)[NW_NB + 1]>(nw_sh);
    double (*Vi)[NW_NB + 1] = reinterpret_cast<double (*)[NW_NB + 1]>(nw_sh + NW_NB * (NW_NB + 1));
    const int c = blockIdx.y, tid = threadIdx.x;
    const int r0 = j0 + nb + blockIdx.x * NW_NB;  // first row of this tile
    if (r0 >= p) return;
    const int nr = min(NW_NB, p - r0);
    double* Hc = H + (long long)c * ldh * ldh;
    const double* inv = INV + ((long long)c * npanels + panel) * NW_NB * NW_NB;
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) {
        const int i = e / NW_NB, j = e % NW_NB;
        A[i][j] = (i < nr && j < nb) ? Hc[(long long)(r0 + i) * ldh + j0 + j] : 0.0;
        Vi[i][j] = inv[e];
    }
    __syncthreads();
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) {
        const int r = e % NW_NB, q = e / NW_NB;  // consecutive threads: consecutive rows (coalesced PT writes)
        if (r >= nr || q >= nb) continue;
        double s = 0.0;
        for (int m = 0; m <= q; ++m) s += A[r][m] * Vi[q][m];
        Hc[(long long)(r0 + r) * ldh + j0 + q] = s;
        const long long t = ((long long)c * NW_NB + q) * ldpt + r0 + r;
        PT[t] = s;
        NPT[t] = -s;
    }
}

// solve L L' d = rhs for every column with the blocked factor: forward (y = L^{-1} rhs) and backward
// (d = L^{-T} y) substitution by panels; the diagonal blocks through their stored inverses.
// rhs = -GRAD; DIR receives d (inactive coordinates: rhs = 0 -> d = 0).  A warp reads contiguous
// row segments of the factor (lanes over columns), partial sums meet in shared memory.
__global__ void __launch_bounds__(NW_T) chol_solve_kernel(const double* __restrict__ H, long long ldh, int p,
                                                          const double* __restrict__ INV, int npanels,
                                                          const double* __restrict__ GRAD, double* __restrict__ DIR,
                                                          long long ldv) {
    extern __shared__ double sh[];  // y[ldh] + t[NB] + part[8][NB]
    double* y = sh;
    double* t = sh + ldh;
    double* part = t + NW_NB;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = NW_T / 32;
    const double* Hc = H + (long long)c * ldh * ldh;
    for (int j = tid; j < p; j += NW_T) y[j] = -GRAD[(long long)c * ldv + j];
    __syncthreads();
    // forward: t = rhs_j - L[j-rows, 0:j0] y[0:j0]; y_j = inv_jj t
    for (int pn = 0; pn < npanels; ++pn) {
        const int j0 = pn * NW_NB, nb = min(NW_NB, p - j0);
        for (int r = warp; r < nb; r += NWARP) {
            const double* row = Hc + (long long)(j0 + r) * ldh;
            double acc = 0.0;
            for (int m = lane; m < j0; m += 32) acc += row[m] * y[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) t[r] = y[j0 + r] - acc;
        }
        __syncthreads();
        const double* inv = INV + ((long long)c * npanels + pn) * NW_NB * NW_NB;
        double s = 0.0;
        if (tid < nb)
            for (int m = 0; m <= tid; ++m) s += inv[tid * NW_NB + m] * t[m];
        __syncthreads();
        if (tid < nb) y[j0 + tid] = s;
        __syncthreads();
    }
    // backward: d_j = inv_jj' (y_j - sum_{i > j} L[i][j-cols]' d_i)
    for (int pn = npanels - 1; pn >= 0; --pn) {
        const int j0 = pn * NW_NB, nb = min(NW_NB, p - j0);
        double a0 = 0.0, a1 = 0.0;  // columns j0 + lane, j0 + lane + 32
        for (int i = j0 + nb + warp; i < p; i += NWARP) {
            const double* row = Hc + (long long)i * ldh + j0;
            const double yi = y[i];
            if (lane < nb) a0 += row[lane] * yi;
            if (lane + 32 < nb) a1 += row[lane + 32] * yi;
        }
        part[warp * NW_NB + lane] = a0;
        part[warp * NW_NB + lane + 32] = a1;
        __syncthreads();
        if (tid < NW_NB) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) acc += part[w * NW_NB + tid];
            t[tid] = (tid < nb) ? y[j0 + tid] - acc : 0.0;
        }
        __syncthreads();
        const double* inv = INV + ((long long)c * npanels + pn) * NW_NB * NW_NB;
        double s = 0.0;
        if (tid < nb)
            for (int m = tid; m < nb; ++m) s += inv[m * NW_NB + tid] * t[m];
        __syncthreads();
        if (tid < nb) y[j0 + tid] = s;
        __syncthreads();
    }
    for (int j = tid; j < p; j += NW_T) DIR[(long long)c * ldv + j] = y[j];
}

// Armijo backtracking on phi(x + t d) - phi(x) (formed without cancellation), one block per column.
// q2 = d'(G_AA/n)d comes from the Newton system itself: d'Hd = -grad'd, H = G_AA/n + curvature terms.
// out[c] = {accepted (0/1), t, decrement = -grad'd, 0}.  X is updated in place when accepted.
__global__ void __launch_bounds__(NW_T) newton_linesearch_kernel(int p, int Gn, const int* __restrict__ gptr,
                                                                 double* __restrict__ X, const double* __restrict__ DIR,
                                                                 const double* __restrict__ GS, const double* __restrict__ GRAD,
                                                                 const double* __restrict__ U, const double* __restrict__ KK,
                                                                 const double* __restrict__ DP, long long ldv,
                                                                 const double* __restrict__ W2, const double* __restrict__ NRM,
                                                                 const int* __restrict__ info, double* __restrict__ out) {
    extern __shared__ double sh[];  // g1[Gn], g2[Gn]
    __shared__ double red[NW_T / 32];
    __shared__ double tsel;
    double* g1 = sh;
    double* g2 = sh + Gn;
    const int c = blockIdx.x, tid = threadIdx.x;
    double* x = X + (long long)c * ldv;
    const double* d = DIR + (long long)c * ldv;
    double q1 = 0.0, gd = 0.0, r1 = 0.0, r2 = 0.0, cdd = 0.0;
    for (int g = tid; g < Gn; g += NW_T) {
        const int ja = gptr[g], jb = gptr[g + 1];
        double a1 = 0.0, a2 = 0.0, ud = 0.0, kkg = 0.0;
        for (int j = ja; j < jb; ++j) {
            const long long e = (long long)c * ldv + j;
            const double dj = d[j], xj = x[j];
            a1 += xj * dj;
            a2 += dj * dj;
            q1 += GS[e] * dj;
            gd += GRAD[e] * dj;
            const double kk = KK[e];
            if (kk >= 0.0) {
                r1 += DP[e] * xj * dj;
                r2 += DP[e] * dj * dj;
                cdd += (kk + DP[e]) * dj * dj;
                ud += U[e] * dj;
                kkg = kk;
            }
        }
        cdd -= kkg * ud * ud;
        g1[g] = a1;
        g2[g] = a2;
    }
    q1 = nw_block_sum(q1, red);
    gd = nw_block_sum(gd, red);
    r1 = nw_block_sum(r1, red);
    r2 = nw_block_sum(r2, red);
    cdd = nw_block_sum(cdd, red);
    const double dec = -gd;
    const double q2 = dec - cdd;  // d'(G_AA/n)d
    const bool usable = info[c] == 0 && dec > 0.0 && isfinite(dec);
    double t = 1.0;
    bool accepted = false;
    if (usable) {
        for (int ls = 0; ls < 24 && !accepted; ++ls) {
            double dpen = 0.0;
            for (int g = tid; g < Gn; g += NW_T) {
                const double nrm = NRM[(long long)c * Gn + g];
                const double dsq = 2.0 * t * g1[g] + t * t * g2[g];  // ||x_g + t d_g||^2 - ||x_g||^2
                const double nn = sqrt(fmax(nrm * nrm + dsq, 0.0));
                const double safe = nrm > 0.0 ? nrm : 1.0;
                dpen += W2[(long long)c * Gn + g] * dsq / (nn + safe);
            }
            dpen = nw_block_sum(dpen, red);
            const double dphi = t * q1 + 0.5 * t * t * q2 + t * r1 + 0.5 * t * t * r2 + dpen;
            if (dphi <= 1e-4 * t * gd)
                accepted = true;
            else
                t *= 0.5;
        }
    }
    if (tid == 0) tsel = accepted ? t : 0.0;
    __syncthreads();
    const double ts = tsel;
    if (ts > 0.0)
        for (int j = tid; j < p; j += NW_T) x[j] += ts * d[j];
    if (tid == 0) {
        out[4 * c + 0] = accepted ? 1.0 : 0.0;
        out[4 * c + 1] = ts;
        out[4 * c + 2] = dec;
        out[4 * c + 3] = (double)info[c];
    }
}

}  // namespace slm
