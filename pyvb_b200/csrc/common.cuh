// Shared device helpers for the pyvb_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pyvb_b200.h"

namespace pyvb {

// ---- stat buffer layout (must match pyvb_b200/_layout.py) -------------------
struct StatLayout {
    int D, q, P;
    size_t t1, bst, ast, cnt, colx, S, zsum, scal, len;
    __host__ __device__ StatLayout(int D_, int q_) : D(D_), q(q_) {
        P = q * (q + 1) / 2;
        t1 = 0;
        bst = t1 + (size_t)D * P;
        ast = bst + (size_t)D * q;
        cnt = ast + (size_t)D * q;
        colx = cnt + D;
        S = colx + D;
        zsum = S + P;
        scal = zsum + q;
        len = scal + PYVB_NSCAL;
    }
};

__host__ __device__ inline int tri(int i) { return i * (i + 1) / 2; }
// first <w_d> column of a Gw row: P rounded up to a multiple of 8 (one DMMA column group)
__host__ __device__ inline int gw_woff(int q) { return (tri(q) + 7) & ~7; }

// packed index p -> (i, j), i >= j
__device__ inline void unpack_p(int p, int &i, int &j) {
    int r = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tri(r) > p) --r;
    while (tri(r + 1) <= p) ++r;
    i = r;
    j = p - tri(r);
}

__device__ inline double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; every thread gets the total.  `sh` holds >= 33 doubles.
// Fixed summation order => deterministic.
__device__ inline double block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nwarp) ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

}  // namespace pyvb
