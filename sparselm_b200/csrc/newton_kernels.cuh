// Second-order phase of the batched solve (ill-conditioned pure-group problems, DESIGN section 2
// item 13): one lock-step damped Newton step on the active groups of k slow columns, entirely in
// this file's kernels -- Hessian assembly straight from the Gram, a batched blocked Cholesky
// factorisation H = U'U (64-wide panels: diagonal block factored and inverted in shared memory, the row
// panel through the inverse, trailing update on the FP64 tensor-core GEMM straight from those rows), blocked
// triangular solves, and the Armijo line search with objective differences formed without
// cancellation.  (Round 1 did these steps with torch operations and cuSOLVER's potrf.)
//
// On the active manifold of column c (groups with ||x_g|| > 0; identity on the other coordinates)
//     phi(x) = 1/(2n) x'Gx - c'x/n + sum_g w_g ||x_g|| + 1/2 sum_g d_g ||x_g||^2
//     grad   = (Gx - c)/n + w_g x_g/||x_g|| + d_g x_g
//     H      = G_AA/n + blockdiag_g( w_g/||x_g|| (I - u_g u_g') ) + diag(d),   u_g = x_g/||x_g||
// Layouts: X, GX, U, KK, DP, GRAD, DIR are [k][ldv] (one row per column, solver feature order);
// H is [k][ldh][ldh] row-major, upper triangle = the factor U after factorisation; group tables W2,
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

// ---- blocked Cholesky, upper form H = U'U ------------------------------------------------------
// Working on the UPPER triangle of the row-major matrix keeps every bulk access on contiguous row
// segments: the panel to the right of a diagonal block is 64 full rows, and those rows ARE the
// K-major operands of the trailing update  H22 -= U12' U12  on the tensor-core GEMM (SYM, negate,
// upper tiles only: no transposed copies, no mirrored scattered stores).
//
// diagonal block [j0, j0+nb): factor in shared memory (A = U'U, U upper), write U back, invert it,
// keep the inverse in INV[c][panel][NB][NB] (row-major, upper triangular).  info[c] != 0: not
// positive definite.
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
    // S holds the block TRANSPOSED (S[j][i] = A[i][j], i <= j): the factorisation below is the
    // familiar lower one on S, i.e. S = U'
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) {
        const int i = e / NW_NB, j = e % NW_NB;  // consecutive threads: consecutive columns of a row (coalesced)
        S[j][i] = (i < nb && j < nb && i <= j) ? Hc[(long long)i * ldh + j] : (i == j ? 1.0 : 0.0);
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
        // trailing update of the lower triangle of S (threads as a 16 x 16 patch: no integer division)
        for (int i = k + 1 + (tid >> 4); i < nb; i += 16) {
            const double sik = S[i][k];
            for (int j = k + 1 + (tid & 15); j <= i; j += 16) S[i][j] -= sik * S[j][k];
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += NW_T) {
        const int i = e / nb, j = e % nb;
        if (i <= j) Hc[(long long)i * ldh + j] = S[j][i];  // U[i][j] = S[j][i]
    }
    // V = S^{-1} (lower) = U^{-T}: column q by forward substitution (thread q)
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) V[e / NW_NB][e % NW_NB] = 0.0;
    __syncthreads();
    if (tid < nb) {
        const int q = tid;
        for (int i = q; i < nb; ++i) {
            double a = (i == q) ? 1.0 : 0.0;
            for (int m = q; m < i; ++m) a -= S[i][m] * V[m][q];
            V[i][q] = a / S[i][i];
        }
    }
    __syncthreads();
    // stored as U^{-1} (upper): INV[i][j] = V[j][i]
    double* out = INV + ((long long)c * npanels + panel) * NW_NB * NW_NB;
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) out[e] = V[e % NW_NB][e / NW_NB];
    if (tid == 0 && bad) info[c] = 1;
}

// panel to the right of the diagonal block, in place: U12 = U11^{-T} A12, i.e.
// out[q][x] = sum_{m <= q} inv[m][q] A[m][x]  (inv = U11^{-1}, upper).  One block per 64 columns.
__global__ void __launch_bounds__(NW_T) chol_panel_kernel(double* __restrict__ H, long long ldh, int p, int j0, int nb,
                                                          const double* __restrict__ INV, int npanels, int panel) {
    extern __shared__ double nw_sh[];
    double (*A)[NW_NB + 1] = reinterpret_cast<double (*)[NW_NB + 1]>(nw_sh);
    double (*Vi)[NW_NB + 1] = reinterpret_cast<double (*)[NW_NB + 1]>(nw_sh + NW_NB * (NW_NB + 1));
    const int c = blockIdx.y, tid = threadIdx.x;
    const int x0 = j0 + nb + blockIdx.x * NW_NB;  // first column of this tile
    if (x0 >= p) return;
    const int nx = min(NW_NB, p - x0);
    double* Hc = H + (long long)c * ldh * ldh;
    const double* inv = INV + ((long long)c * npanels + panel) * NW_NB * NW_NB;
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) {
        const int m = e / NW_NB, x = e % NW_NB;  // rows of the panel are contiguous: coalesced
        A[m][x] = (m < nb && x < nx) ? Hc[(long long)(j0 + m) * ldh + x0 + x] : 0.0;
        Vi[m][x] = inv[e];
    }
    __syncthreads();
    for (int e = tid; e < NW_NB * NW_NB; e += NW_T) {
        const int q = e / NW_NB, x = e % NW_NB;
        if (q >= nb || x >= nx) continue;
        double a = 0.0;
        for (int m = 0; m <= q; ++m) a += Vi[m][q] * A[m][x];
        Hc[(long long)(j0 + q) * ldh + x0 + x] = a;
    }
}

// solve U'U d = rhs for every column with the blocked factor: forward (U'y = rhs) and backward (U d = y)
// substitution by panels; the diagonal blocks through their stored inverses.  rhs = -GRAD; DIR
// receives d (inactive coordinates: rhs = 0 -> d = 0).  A warp reads contiguous row segments of the
// factor (lanes over columns), partial sums meet in shared memory.
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
    // forward: y_j = U_jj^{-T} (rhs_j - sum_{i < j} U[i-rows][j-cols]' y_i)
    for (int pn = 0; pn < npanels; ++pn) {
        const int j0 = pn * NW_NB, nb = min(NW_NB, p - j0);
        double a0 = 0.0, a1 = 0.0;  // columns j0 + lane, j0 + lane + 32
        for (int i = warp; i < j0; i += NWARP) {
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
        double a = 0.0;
        if (tid < nb)
            for (int m = 0; m <= tid; ++m) a += inv[m * NW_NB + tid] * t[m];  // (U^{-T} t)_q = sum_{m <= q} inv[m][q] t_m
        __syncthreads();
        if (tid < nb) y[j0 + tid] = a;
        __syncthreads();
    }
    // backward: d_j = U_jj^{-1} (y_j - U[j-rows][right cols] d_right)
    for (int pn = npanels - 1; pn >= 0; --pn) {
        const int j0 = pn * NW_NB, nb = min(NW_NB, p - j0);
        for (int r = warp; r < nb; r += NWARP) {
            const double* row = Hc + (long long)(j0 + r) * ldh;
            double acc = 0.0;
            for (int m = j0 + nb + lane; m < p; m += 32) acc += row[m] * y[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) t[r] = y[j0 + r] - acc;
        }
        __syncthreads();
        const double* inv = INV + ((long long)c * npanels + pn) * NW_NB * NW_NB;
        double a = 0.0;
        if (tid < nb)
            for (int m = tid; m < nb; ++m) a += inv[tid * NW_NB + m] * t[m];  // upper: row q, columns m >= q
        __syncthreads();
        if (tid < nb) y[j0 + tid] = a;
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
