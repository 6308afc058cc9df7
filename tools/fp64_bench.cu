// FP64 pipe microbenchmark for B200: DMMA.8x8x4 rate, DFMA rate, and both together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_bench tools/fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double *out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// even warps DMMA, odd warps DFMA
template <int ILP>
__global__ void k_mix(double *out, int iters, double a, double b) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double c[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) s += c[i];
    } else {
        double c[ILP][2];
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32;
        for (int bps : {1, 2}) {
            if (warps * bps > 64) continue;
            const int blocks = sms * bps;
            float ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
            double fl = (double)blocks * warps * iters * 8 * 512.0;
            float ms2 = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl2 = (double)blocks * threads * iters * 8 * 2.0;
            float ms3 = timeit([&] { k_mix<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl3 = (double)blocks * (warps / 2) * iters * 8 * (512.0 + 64.0);
            printf("warps/blk %2d blk/SM %d : DMMA %7.2f TF  (%.3f ms) | DFMA %7.2f TF | mix %7.2f TF (dmma part %.2f, dfma part %.2f; ms %.3f)\n",
                   warps, bps, fl / ms * 1e-9, ms, fl2 / ms2 * 1e-9, fl3 / ms3 * 1e-9,
                   (double)blocks * (warps / 2) * iters * 8 * 512.0 / ms3 * 1e-9,
                   (double)blocks * (warps / 2) * iters * 8 * 64.0 / ms3 * 1e-9, ms3);
        }
    }
    // ILP sweep at 8 warps, 1 block/SM
    {
        const int blocks = sms, threads = 256;
        float a = timeit([&] { k_dmma<1><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
        float b = timeit([&] { k_dmma<2><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
        float c = timeit([&] { k_dmma<4><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
        float d = timeit([&] { k_dmma<16><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
        double base = (double)blocks * 8 * iters * 512.0 * 1e-9;
        printf("8 warps ILP1 %.2f TF ILP2 %.2f ILP4 %.2f ILP16 %.2f  (latency/ILP1: %.1f ns per dmma)\n",
               base / a, 2 * base / b, 4 * base / c, 16 * base / d, a * 1e6 / iters);
    }
    return 0;
}
