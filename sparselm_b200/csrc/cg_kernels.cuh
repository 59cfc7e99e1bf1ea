// Conjugate gradients on a device-resident Gram: the unpenalised member of the family
// (OrdinaryLeastSquares, reference model/_ols.py:57-65) solves G b = c with the same
// tensor-core Gram apply as the penalised estimators; the vector part of an iteration
// (two dot products, three axpy's) is p elements and runs in ONE block so that the
// scalars never leave the device and the sums have a fixed order.
//
// Layout: X8, D8, GD8 are [p][8] (the apply's narrowest right-hand side: column 0 is the
// vector, columns 1..7 stay zero), R is [p], sc[0] = r.r, sc[1] = c.c, sc[2] = d.Gd of the
// last step.  c is row p of the Gram (X^T y), as everywhere else in the engine.
#pragma once
#include <cuda_runtime.h>

namespace slm {

constexpr int CG_T = 1024;

// sum over the block in a fixed order; every thread gets the result
__device__ __forceinline__ double cg_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // red may still be read from the previous call
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = (l < CG_T / 32) ? red[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// r = c - (G + diag(shift)) x (GX8 == NULL: x = 0, r = c), d = r.  shift (may be NULL): the ridge
// n * delta_g(j) of an unpenalised ridged problem (alpha = 0, delta > 0).
__global__ void __launch_bounds__(CG_T) cg_start_kernel(const double* __restrict__ G, long long pa, int p,
                                                        double* __restrict__ X8, const double* __restrict__ GX8,
                                                        const double* __restrict__ shift, double* __restrict__ R,
                                                        double* __restrict__ D8, double* __restrict__ sc) {
    __shared__ double red[CG_T / 32];
    const double* __restrict__ c = G + (long long)p * pa;
    double rs = 0.0, cc = 0.0;
    for (int j = threadIdx.x; j < p; j += CG_T) {
        const double cj = c[j];
        double r = cj;
        if (GX8)
            r -= GX8[(long long)j * 8] + (shift ? shift[j] * X8[(long long)j * 8] : 0.0);
        else
            X8[(long long)j * 8] = 0.0;
        R[j] = r;
        D8[(long long)j * 8] = r;
        rs += r * r;
        cc += cj * cj;
    }
    rs = cg_block_sum(rs, red);
    cc = cg_block_sum(cc, red);
    if (threadIdx.x == 0) {
        sc[0] = rs;
        sc[1] = cc;
        sc[2] = 0.0;
    }
}

// one CG step given GD8 = G d (the operator is G + diag(shift))
__global__ void __launch_bounds__(CG_T) cg_step_kernel(int p, double* __restrict__ X8, double* __restrict__ R,
                                                       double* __restrict__ D8, const double* __restrict__ GD8,
                                                       const double* __restrict__ shift, double* __restrict__ sc) {
    __shared__ double red[CG_T / 32];
    const double rs = sc[0];
    double dgd = 0.0;
    for (int j = threadIdx.x; j < p; j += CG_T) {
        const double d = D8[(long long)j * 8];
        dgd += d * (GD8[(long long)j * 8] + (shift ? shift[j] * d : 0.0));
    }
    dgd = cg_block_sum(dgd, red);
    // d in the null space of G (or r = 0): nothing left to gain along d
    const double a = (dgd > 0.0 && rs > 0.0) ? rs / dgd : 0.0;
    double rsn = 0.0;
    for (int j = threadIdx.x; j < p; j += CG_T) {
        const long long e = (long long)j * 8;
        X8[e] += a * D8[e];
        const double r = R[j] - a * (GD8[e] + (shift ? shift[j] * D8[e] : 0.0));
        R[j] = r;
        rsn += r * r;
    }
    rsn = cg_block_sum(rsn, red);
    const double b = rs > 0.0 ? rsn / rs : 0.0;
    for (int j = threadIdx.x; j < p; j += CG_T) {
        const long long e = (long long)j * 8;
        D8[e] = R[j] + b * D8[e];
    }
    if (threadIdx.x == 0) {
        sc[0] = rsn;
        sc[2] = dgd;
    }
}

}  // namespace slm
