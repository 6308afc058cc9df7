// DFMA issue rate as a function of the operand pattern (B200): does a stream of  a[j] = fma(t, B[j], a[j])  -- two distinct 64-bit
// register operands per instruction, the inner loop of the Gauss-Jordan K2 kernel -- run at the rate of  c = fma(c, k1, k2) ?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_operands tools/dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NA, int NB, int MODE>
__global__ void __launch_bounds__(256, 1) k(double *out, const double *in, int iters) {
    double a[NA], B[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = in[threadIdx.x + i];
#pragma unroll
    for (int i = 0; i < NB; ++i) B[i] = in[threadIdx.x + 100 + i];
    double t0 = in[threadIdx.x + 300], t1 = in[threadIdx.x + 301];
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = fma(a[i], t0, t1);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NA; ++i) a[i] = fma(t0, B[i % NB], a[i]);
        } else if (MODE == 3 || MODE == 4) {      // the kernel's order with the B values refilled from shared memory (broadcast LDS.128)
            extern __shared__ double2 sm[];
            const int half = (threadIdx.x >> 4) & 1;
#pragma unroll
            for (int i = 0; i < NA / 2; i += 2) {
                a[i] = fma(t0, B[i % NB], a[i]);
                a[i + NA / 2] = fma(t1, B[i % NB], a[i + NA / 2]);
                a[i + 1] = fma(t0, B[(i + 1) % NB], a[i + 1]);
                a[i + 1 + NA / 2] = fma(t1, B[(i + 1) % NB], a[i + 1 + NA / 2]);
                const double2 v = sm[(MODE == 3 ? half : (threadIdx.x & 31)) * 16 + (i / 2) + (it & 1) * 512];
                B[i % NB] = v.x;
                B[(i + 1) % NB] = v.y;
            }
        } else {                                  // two rows share every B value (the kernel's order)
#pragma unroll
            for (int i = 0; i < NA / 2; ++i) {
                a[i] = fma(t0, B[i % NB], a[i]);
                a[i + NA / 2] = fma(t1, B[i % NB], a[i + NA / 2]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NA, int NB, int MODE>
void run(const char *name, int warps, double *out, double *in) {
    int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NA, NB, MODE><<<148, 32 * warps, 32768>>>(out, in, 10);
    cudaEventRecord(e0);
    k<NA, NB, MODE><<<148, 32 * warps, 32768>>>(out, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_per_clk_sm = (double)NA * iters * 32 * warps / (ms * 1e-3 * 1.965e9);
    printf("%-44s warps/SM %2d: %.3f ms  %.1f FMA/clk/SM  (%.2f clk per DFMA and SMSP)\n", name, warps, ms, fma_per_clk_sm,
           128.0 / fma_per_clk_sm);
}

int main() {
    double *out, *in;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    for (int w : {4, 8}) {
        if (w == 4) {
            run<64, 32, 0>("c = fma(c, k1, k2), 64 chains", 4, out, in);
            run<64, 32, 1>("a[j] = fma(t, B[j], a[j]), 64 acc, 32 B", 4, out, in);
            run<64, 32, 2>("two rows per B value, 64 acc, 16 B used", 4, out, in);
            run<64, 32, 3>("+ broadcast LDS.128 per 4 DFMA", 4, out, in);
            run<64, 32, 4>("+ per-lane LDS.128 per 4 DFMA", 4, out, in);
        } else {
            run<64, 32, 0>("c = fma(c, k1, k2), 64 chains", 8, out, in);
            run<64, 32, 1>("a[j] = fma(t, B[j], a[j]), 64 acc, 32 B", 8, out, in);
            run<64, 32, 2>("two rows per B value, 64 acc, 16 B used", 8, out, in);
            run<64, 32, 3>("+ broadcast LDS.128 per 4 DFMA", 8, out, in);
            run<64, 32, 4>("+ per-lane LDS.128 per 4 DFMA", 8, out, in);
        }
    }
    return 0;
}
