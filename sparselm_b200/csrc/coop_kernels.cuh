// Few-column solves on a large design (a plain estimator.fit, the refit of a CV search, an
// adaptive chain): one cooperative grid runs MANY complete iterations per launch.
//
// With one to four grid columns the regular path is pure latency: per iteration a stream-K
// GEMM that moves a few MB, two prox kernels and a support-list kernel, ~100 us of launches
// and drains for ~5 us of work.  Here the SMs of the GPU share ONE column: CTA b owns a
// contiguous range of whole groups (rows r0..r1), keeps their state (z, beta, G beta) in
// registers and computes its rows of G z from the support rows of G (by symmetry
// G[j][k] = G[k][j]: for every non-zero z_k one coalesced segment G[k][r0..r1) out of L2).
// The only exchange per iteration is one grid-wide barrier: before it every CTA publishes
// its beta+ rows and its share of the restart dot sum (z - b+)(b+ - b); after it every CTA
// forms theta and the whole new z = b+ + theta (b+ - b) itself (compacted to the support
// list of the next product on the fly).  The arithmetic is that of gemm apply +
// prox_main_kernel + prox_momentum_kernel (same restart rule, fixed summation orders), so
// the regular kernels -- the convergence checks -- interleave freely, as they do with
// fista_small_kernel.
//
// Buffers: beta of the last three iterations rotate through buf[0..2] ([nprob][p], contiguous
// per problem so that the post-barrier reads are coalesced); a buffer is overwritten two
// barriers after its last reader.  dpart[2][nprob][NB] is double-buffered the same way.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "solver_kernels.cuh"

namespace slm {

namespace cg = cooperative_groups;

constexpr int kCoopMaxProb = 4;   // grid columns a cooperative launch iterates on
constexpr int CO_T = 512;            // threads per CTA
constexpr int CO_W = CO_T / 32;      // warps = splits of the contraction
constexpr int CO_ROWS = 256;         // most rows a CTA may own (one owner thread per row)
constexpr int CO_RT = CO_ROWS / 32;  // row tiles of 32 per CTA (upper bound)

struct CoopArgs {
    double* buf[3];  // [nprob][p] each
    double* dpart;   // [2][nprob][NB]
    int nprob;
    int seg;  // per-warp capacity of the support list in shared memory
    int pf[kCoopMaxProb], pk[kCoopMaxProb];  // grid mode: (fold, column) of every problem
    const int* plist;                        // cluster mode: device list, fold * 65536 + column
};

// (fold, column) of the active grid columns, in fold-major order (one block; the host knows
// their number from the convergence check and sizes the cluster launch with it)
__global__ void coop_list_kernel(const __grid_constant__ SolveDev sp, int* __restrict__ plist) {
    if (threadIdx.x != 0) return;
    int q = 0;
    for (int f = 0; f < sp.F; ++f)
        for (int k = 0; k < sp.K[f]; ++k)
            if (sp.flag[(long long)f * sp.ldz + k] == 0) plist[q++] = f * 65536 + k;
}

__host__ __device__ inline int coop_seg(int p) { return ((p + 31) / 32 + CO_W - 1) / CO_W * 32; }
__host__ __device__ inline size_t coop_smem(int p) {
    // support list (index + value) per warp, partial products [CO_W][CO_ROWS], us[CO_ROWS], 64 scalars
    return (size_t)CO_W * coop_seg(p) * (sizeof(int) + sizeof(double)) +
           sizeof(double) * (CO_W * CO_ROWS + CO_ROWS + 64);
}

// CLUSTER = false: one cooperative grid, blockIdx.y = problem, gridDim.x CTAs per problem, grid
// barrier.  CLUSTER = true: one thread-block cluster per problem (any number of problems, each
// iterating at its own pace), cluster barrier.
template <bool GROUPED, bool CLUSTER>
__global__ void __launch_bounds__(CO_T, 1) fista_coop_kernel(const __grid_constant__ SolveDev sp,
                                                              const __grid_constant__ CoopArgs ca, int par,
                                                              int n_inner) {
    cg::grid_group grid = cg::this_grid();
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smraw[];
    const int p = sp.p, seg = ca.seg;
    double* sval = reinterpret_cast<double*>(smraw);              // [CO_W][seg] z of the support rows
    double* parts = sval + (size_t)CO_W * seg;                    // [CO_W][CO_ROWS]
    double* us = parts + CO_W * CO_ROWS;                          // [CO_ROWS]
    double* red = us + CO_ROWS;                                   // [64]
    int* sidx = reinterpret_cast<int*>(red + 64);                 // [CO_W][seg]
    __shared__ int scnt[CO_W];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NB = CLUSTER ? (int)cluster.num_blocks() : (int)gridDim.x;
    const int b = CLUSTER ? (int)cluster.block_rank() : (int)blockIdx.x;
    const int prob = CLUSTER ? (int)blockIdx.x / NB : (int)blockIdx.y;
    const int f = CLUSTER ? ca.plist[prob] >> 16 : ca.pf[prob];
    const int k = CLUSTER ? (ca.plist[prob] & 0xffff) : ca.pk[prob];
    const long long ldz = sp.ldz;
    const long long colbase = (long long)f * ldz + k;
    const bool active = sp.flag[colbase] == 0;  // a converged / frozen column only keeps the barriers

    // rows of this CTA: whole groups, balanced by rows
    int r0, r1;
    if (GROUPED) {
        auto first_group_at = [&](int target) {  // smallest g with gptr[g] >= target
            int lo = 0, hi = sp.Gn;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sp.gptr[mid] >= target)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            return lo;
        };
        const int ga = b == 0 ? 0 : first_group_at((int)(((long long)b * p + NB - 1) / NB));
        const int gb = b == NB - 1 ? sp.Gn : first_group_at((int)(((long long)(b + 1) * p + NB - 1) / NB));
        r0 = sp.gptr[ga];
        r1 = sp.gptr[gb];
    } else {
        r0 = (int)((long long)b * p / NB);
        r1 = (int)((long long)(b + 1) * p / NB);
    }
    const int nrows = r1 - r0;  // <= CO_ROWS (the host checked p/NB + largest group)
    const int row = r0 + tid;
    const bool owner = active && tid < nrows;

    const double* __restrict__ Gf = sp.G + (long long)f * sp.g_stride;
    const int ko = sp.colmap ? sp.colmap[colbase] : k;
    const double n = sp.n_obs[f], step = sp.lips_dev ? 1.0 / sp.lips_dev[f] : sp.step[f];
    const double son = step / n;
    const long long sbase = (long long)f * p * ldz + k;
    double zi = 0.0, bi = 0.0, gbi = 0.0, cj = 0.0, tw1 = 0.0, tw2 = 0.0, rd = 1.0;
    int ja = tid, jb = tid + 1;  // group of this row, relative to r0
    if (owner) {
        const long long e = sbase + (long long)row * ldz;
        zi = sp.Z[e];
        bi = sp.B[e];
        gbi = sp.GB[e];
        cj = Gf[(long long)p * sp.pa + row];
        const double w1 = sp.W1 ? sp.W1[(long long)f * p * ldz + (long long)row * ldz + ko]
                                : (sp.lam1 ? sp.lam1[(long long)f * ldz + ko] : 0.0);
        tw1 = step * w1;
        int g = row;
        if (GROUPED) {
            int lo = 0, hi = sp.Gn - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (sp.gptr[mid] <= row)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            g = lo;
            ja = sp.gptr[g] - r0;
            jb = sp.gptr[g + 1] - r0;
        }
        const long long gbase = (long long)f * sp.Gn * ldz + (long long)g * ldz + ko;
        tw2 = step * (sp.W2 ? sp.W2[gbase] : 0.0);
        rd = 1.0 / (1.0 + step * (sp.D2 ? sp.D2[gbase] : 0.0));
    }
    double theta = active ? sp.theta[par][colbase] : 0.0;
    double tm = active ? sp.tmom[par][colbase] : 1.0;

    double* bufs[3] = {ca.buf[0] + (size_t)prob * p, ca.buf[1] + (size_t)prob * p, ca.buf[2] + (size_t)prob * p};
    int cur = 0;  // buffer holding beta of the current iteration
    if (owner) bufs[0][row] = bi;

    // support list of z for the first product: straight from the state array
    const int nblk32 = (p + 31) / 32;
    if (active) {
        int cnt = 0;
        for (int blk = warp; blk < nblk32; blk += CO_W) {
            const int kk = blk * 32 + lane;
            const double z = kk < p ? sp.Z[sbase + (long long)kk * ldz] : 0.0;
            const bool nz = z != 0.0;
            const unsigned bal = __ballot_sync(0xffffffffu, nz);
            if (nz) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                sidx[warp * seg + pos] = kk;
                sval[warp * seg + pos] = z;
            }
            cnt += __popc(bal);
        }
        if (lane == 0) scnt[warp] = cnt;
    }
    __syncthreads();

    const int ntile = (nrows + 31) / 32;
#pragma unroll 1
    for (int it = 0; it < n_inner; ++it) {
        double bn = 0.0;
        if (active) {
            // ---- rows r0..r1 of G z: warp w contracts over its share of the support ----------
            double acc[CO_RT];
#pragma unroll
            for (int t = 0; t < CO_RT; ++t) acc[t] = 0.0;
            const int cnt = scnt[warp];
            const int* __restrict__ wi = sidx + warp * seg;
            const double* __restrict__ wv = sval + warp * seg;
            const double* __restrict__ gcol = Gf + r0 + lane;
            int i = 0;
            constexpr int U = 8;  // independent row segments in flight per thread and row tile
            for (; i + U <= cnt; i += U) {
                const double* gp[U];
                double zz[U];
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    gp[q] = gcol + (long long)wi[i + q] * sp.pa;
                    zz[q] = wv[i + q];
                }
#pragma unroll
                for (int t = 0; t < CO_RT; ++t) {
                    if (t < ntile && t * 32 + lane < nrows) {
                        double a[U];
#pragma unroll
                        for (int q = 0; q < U; ++q) a[q] = gp[q][t * 32];
#pragma unroll
                        for (int q = 0; q < U; ++q) acc[t] += a[q] * zz[q];
                    }
                }
            }
            for (; i < cnt; ++i) {
                const double* g0 = gcol + (long long)wi[i] * sp.pa;
                const double z0 = wv[i];
#pragma unroll
                for (int t = 0; t < CO_RT; ++t)
                    if (t < ntile && t * 32 + lane < nrows) acc[t] += g0[t * 32] * z0;
            }
#pragma unroll
            for (int t = 0; t < CO_RT; ++t)
                if (t < ntile) parts[warp * CO_ROWS + t * 32 + lane] = acc[t];
            __syncthreads();

            // ---- prox of the owned rows ------------------------------------------------------
            double u = 0.0;
            if (owner) {
                double gz = parts[tid];
#pragma unroll
                for (int w = 1; w < CO_W; ++w) gz += parts[w * CO_ROWS + tid];
                gbi = (gz + theta * gbi) / (1.0 + theta);
                u = softt(zi - son * (gz - cj), tw1);
                us[tid] = u;
            }
            __syncthreads();
            double d = 0.0;
            if (owner) {
                double ss = 0.0;
                if (GROUPED) {
                    for (int j = ja; j < jb; ++j) ss += us[j] * us[j];
                } else {
                    ss = u * u;
                }
                const double nrm = sqrt(ss);
                const double scale = (nrm > 0.0 ? fmax(0.0, 1.0 - tw2 / nrm) : 0.0) * rd;
                bn = scale * u;
                d = (zi - bn) * (bn - bi);
                bufs[(cur + 1) % 3][row] = bn;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if (lane == 0) red[warp] = d;
            __syncthreads();
            if (tid == 0) {
                double ds = 0.0;
                for (int w = 0; w < CO_W; ++w) ds += red[w];
                ca.dpart[((size_t)(it & 1) * ca.nprob + prob) * NB + b] = ds;
            }
        }
        if (CLUSTER)
            cluster.sync();
        else
            grid.sync();
        if (active) {
            // ---- restart test and momentum coefficient (every CTA, same fixed order) ---------
            if (warp == 0) {
                const double* dp = ca.dpart + ((size_t)(it & 1) * ca.nprob + prob) * NB;
                double ds = 0.0;
                for (int q = lane; q < NB; q += 32) ds += __ldcg(dp + q);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
                if (lane == 0) {
                    double tn = 0.5 * (1.0 + sqrt(1.0 + 4.0 * tm * tm));
                    double th = (tm - 1.0) / tn;
                    if (ds > 0.0) {  // gradient-scheme adaptive restart
                        th = 0.0;
                        tn = 1.0;
                    }
                    red[32] = th;
                    red[33] = tn;
                }
            }
            __syncthreads();
            theta = red[32];
            tm = red[33];
            // ---- new z = b+ + theta (b+ - b), compacted to the next product's support list ----
            const double* __restrict__ bN = bufs[(cur + 1) % 3];
            const double* __restrict__ bO = bufs[cur];
            int cnt = 0;
            for (int blk = warp; blk < nblk32; blk += CO_W) {
                const int kk = blk * 32 + lane;
                double z = 0.0;
                if (kk < p) {
                    const double x1 = __ldcg(bN + kk), x0 = __ldcg(bO + kk);
                    z = x1 + theta * (x1 - x0);
                }
                const bool nz = z != 0.0;
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (nz) {
                    const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                    sidx[warp * seg + pos] = kk;
                    sval[warp * seg + pos] = z;
                }
                cnt += __popc(bal);
            }
            if (lane == 0) scnt[warp] = cnt;
            if (owner) {
                zi = bn + theta * (bn - bi);
                bi = bn;
            }
            cur = (cur + 1) % 3;
            __syncthreads();
        }
    }
    if (owner) {
        const long long e = sbase + (long long)row * ldz;
        sp.Z[e] = zi;
        sp.B[e] = bi;
        sp.GB[e] = gbi;
    }
    if (active && b == 0 && tid == 0) {
        const int po = par ^ (n_inner & 1);
        sp.theta[po][colbase] = theta;
        sp.tmom[po][colbase] = tm;
    }
}

}  // namespace slm
