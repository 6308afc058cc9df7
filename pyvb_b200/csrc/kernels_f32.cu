// FP32 variant of the Z step contraction (K1) on the 5th-generation tensor cores: tcgen05.mma with TMEM
// accumulators, operands staged by TMA, warp-specialised persistent kernel.
//
// Reference arithmetic: Multiplication.pass_up_m1_m2, nodes/node.py:203-227 (m1 = tr(<w_i w_j^T> Lambda_n),
// m2 = <W>^T sum_m2) with Lambda_n = tau diag(mask_n) -- the same contraction as the FP64 DMMA kernel
// (kernels_dmma.cu), here as a bf16 x 3 split:
//     [qprec_n | eta_n] = [P0 | h0] + tau * ( mask_n @ (G_h + G_m + G_l)  +  (x_h + x_m)_n @ (W_h + W_m + W_l) )
// * the mask is exactly representable in bf16 and G = G_h + G_m + G_l to 24 bits, so mask @ G is an FP32-exact
//   product accumulated in FP32 (TMEM);  the eta columns of G hold -mu_d w_d (the "- m1 <Mu>" of node.py:105-107)
// * x is kept as two bf16 planes (16-bit mantissa), W as three; products x_h W_{h,m,l} + x_m W_{h,m}
// Data (static over the sweeps, prepared once):  planes bf16 [3][N][D] = mask | x_h | x_m  (0 where missing)
// Per sweep (tiny):                               GT bf16 [3][NCP][D] (K-major B operand), WT bf16 [3][q][D]
// Output: MZ32 float [N][NCP] rows [qprec packed (P) | pad | eta (q) | pad], NCP = (PP + q) rounded up to 64.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include "umma.cuh"

namespace pyvb {

namespace {

__host__ __device__ constexpr int f_tri(int i) { return i * (i + 1) / 2; }
__host__ __device__ constexpr int f_pp(int q) { return (f_tri(q) + 15) & ~15; }   // first eta / zbar column
__host__ __device__ constexpr int f_ncp(int q) { return (f_pp(q) + q + 63) & ~63; }

// ------------------------------------------------------------------ operand preparation
__device__ __forceinline__ void split3(double v, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l) {
    h = __double2bfloat16(v);
    const double r1 = v - (double)__bfloat162float(h);
    m = __double2bfloat16(r1);
    const double r2 = r1 - (double)__bfloat162float(m);
    l = __double2bfloat16(r2);
}

// planes[0] = mask, planes[1] = x_h, planes[2] = x_m  (each [N][D]); missing entries are zero in all three
__global__ void __launch_bounds__(256)
prepare_x_kernel(long long N, int D, const double *__restrict__ X, long long ldx, __nv_bfloat16 *__restrict__ planes) {
    const long long total = N * (long long)D;
    const size_t plane = (size_t)total;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / D;
        const int d = (int)(e - n * D);
        const double x = X[n * ldx + d];
        const bool ob = (x == x);
        __nv_bfloat16 h, m, l;
        split3(ob ? x : 0.0, h, m, l);
        planes[e] = __float2bfloat16(ob ? 1.0f : 0.0f);
        planes[plane + e] = h;
        planes[2 * plane + e] = m;
    }
}

// GT[p][c][d]: column c of the accumulator row ([G_d packed | pad | -mu_d w_d | pad]) for data dimension d, plane p;
// WT[p][i][d] = plane p of <w_di>.  One thread per (c, d), d fastest (coalesced writes).
__global__ void __launch_bounds__(256)
pack_gw_f32_kernel(int D, int q, const double *__restrict__ Wbar, const double *__restrict__ Wvar,
                   const double *__restrict__ mu, __nv_bfloat16 *__restrict__ GT, __nv_bfloat16 *__restrict__ WT) {
    const int P = f_tri(q), PP = f_pp(q), NCP = f_ncp(q);
    const long long total = (long long)(NCP + q) * D;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e / D), d = (int)(e % D);
        double v = 0.0;
        __nv_bfloat16 *dst;
        size_t plane;
        if (c < NCP) {
            if (c < P) {
                int i, j;
                unpack_p(c, i, j);
                v = Wbar[(size_t)d * q + i] * Wbar[(size_t)d * q + j];
                if (i == j) v += Wvar[(size_t)d * q + i];
            } else if (c >= PP && c < PP + q) {
                v = -mu[d] * Wbar[(size_t)d * q + (c - PP)];
            }
            dst = GT + (size_t)c * D + d;
            plane = (size_t)NCP * D;
        } else {
            v = Wbar[(size_t)d * q + (c - NCP)];
            dst = WT + (size_t)(c - NCP) * D + d;
            plane = (size_t)q * D;
        }
        __nv_bfloat16 h, m, l;
        split3(v, h, m, l);
        dst[0] = h;
        dst[plane] = m;
        dst[2 * plane] = l;
    }
}

// ------------------------------------------------------------------ tensor maps (bf16, 3-D: [plane][row][col])
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_f32() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}
cudaError_t make_map_bf16_3d(CUtensorMap *m, const void *base, uint64_t cols, uint64_t rows, uint64_t planes,
                             uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode_f32();
    if (!enc) return cudaErrorNotSupported;
    cuuint64_t dims[3] = {cols, rows, planes};
    cuuint64_t strides[2] = {cols * 2, cols * rows * 2};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

__device__ __forceinline__ void tma_load_3d(void *dst_smem, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ K1, FP32 variant
template <int Q> struct FC;      // NT: accumulator columns per CTA tile (multiple of 64, <= 256); ST: pipeline stages
template <> struct FC<16> { static constexpr int NT = 192, ST = 3; };
template <> struct FC<32> { static constexpr int NT = 192, ST = 3; };
template <> struct FC<64> { static constexpr int NT = 256, ST = 2; };

template <int Q> struct FT {
    static constexpr int P = f_tri(Q), PP = f_pp(Q), NCP = f_ncp(Q);
    static constexpr int NT = FC<Q>::NT, ST = FC<Q>::ST;
    static constexpr int NCT = (NCP + NT - 1) / NT;          // column tiles
    static constexpr int ECT = PP / NT;                      // the column tile that holds the eta columns
    static constexpr int BM = 128, BK = 32;                  // rows per tile, K elements per stage (64-byte rows)
    static constexpr int A_B = BM * BK * 2;                  // one A plane tile (bytes)
    static constexpr int G_B = NT * BK * 2;                  // one G plane tile
    static constexpr int W_B = Q * BK * 2;                   // one W plane tile
    static constexpr int STAGE_B = 3 * A_B + 3 * G_B + 3 * W_B;
    static constexpr int NTHR = 6 * 32;                      // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue
    static constexpr size_t SMEM = 1024 + (size_t)ST * STAGE_B + (size_t)(P + Q) * 4 + (2 * ST + 4) * 8 + 16;
    static_assert((PP + Q - 1) / NT == ECT, "the eta columns must not straddle two column tiles");
    static_assert(A_B % 512 == 0 && G_B % 512 == 0 && W_B % 512 == 0, "SWIZZLE_64B tiles must stay 512-byte aligned");
};

template <int Q>
__global__ void __launch_bounds__(FT<Q>::NTHR, 1)
zstep_f32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                 const __grid_constant__ CUtensorMap tmW, long long N, int D, const double *__restrict__ P0,
                 const double *__restrict__ h0, const double *__restrict__ gl, float *__restrict__ MZ, int ntiles) {
    using T = FT<Q>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *stage0 = smem;
    float *p0v = reinterpret_cast<float *>(smem + T::ST * T::STAGE_B);
    float *h0s = p0v + T::P;
    uint64_t *full = reinterpret_cast<uint64_t *>(h0s + Q);
    uint64_t *empty = full + T::ST;
    uint64_t *tfull = empty + T::ST;        // [2] accumulator buffer complete
    uint64_t *tempty = tfull + 2;           // [2] accumulator buffer drained by the epilogue
    uint32_t *tbase = reinterpret_cast<uint32_t *>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = D / T::BK;

    for (int p = tid; p < T::P; p += T::NTHR) {
        int i, j;
        unpack_p(p, i, j);
        p0v[p] = (float)P0[i * Q + j];
    }
    if (tid < Q) h0s[tid] = (float)h0[tid];
    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < T::ST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) umma::tmem_alloc(tbase, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tbase;

    if (warp == 0) {
        // ===================== TMA producer (one lane) =====================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int ct = tile % T::NCT;
                const int row0 = (tile / T::NCT) * T::BM;
                const bool eta = (ct == T::ECT);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % T::ST;
                    umma::mbar_wait_bounded(&empty[s], (uint32_t)(((it / T::ST) & 1) ^ 1));
                    unsigned char *st = stage0 + s * T::STAGE_B;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(T::A_B + 3 * T::G_B + (eta ? 2 * T::A_B + 3 * T::W_B : 0)));
                    tma_load_3d(st, &tmA, kb * T::BK, row0, 0, &full[s]);                 // mask
                    for (int p = 0; p < 3; ++p)
                        tma_load_3d(st + 3 * T::A_B + p * T::G_B, &tmG, kb * T::BK, ct * T::NT, p, &full[s]);
                    if (eta) {
                        tma_load_3d(st + T::A_B, &tmA, kb * T::BK, row0, 1, &full[s]);    // x_h
                        tma_load_3d(st + 2 * T::A_B, &tmA, kb * T::BK, row0, 2, &full[s]);
                        for (int p = 0; p < 3; ++p)
                            tma_load_3d(st + 3 * T::A_B + 3 * T::G_B + p * T::W_B, &tmW, kb * T::BK, 0, p, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one lane) =====================
        if (lane == 0) {
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                const int ct = tile % T::NCT;
                const bool eta = (ct == T::ECT);
                const int nt = (T::NCP - ct * T::NT < T::NT) ? (T::NCP - ct * T::NT) : T::NT;
                const int buf = tl & 1;
                umma::mbar_wait_bounded(&tempty[buf], (uint32_t)(((tl >> 1) & 1) ^ 1));
                umma::fence_after_sync();
                const uint32_t dacc = tmem + (uint32_t)(buf * 256);
                const uint32_t deta = dacc + (uint32_t)(T::PP - T::ECT * T::NT);
                const uint32_t id_g = umma::idesc_bf16_f32(T::BM, nt, 0, 0);
                const uint32_t id_w = umma::idesc_bf16_f32(T::BM, Q, 0, 0);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % T::ST;
                    umma::mbar_wait_bounded(&full[s], (uint32_t)((it / T::ST) & 1));
                    umma::fence_after_sync();
                    const uint32_t a0 = smem_u32(stage0 + s * T::STAGE_B);
                    const uint32_t g0 = a0 + 3 * T::A_B, w0 = g0 + 3 * T::G_B;
#pragma unroll
                    for (int ks = 0; ks < T::BK / 16; ++ks) {
#pragma unroll
                        for (int p = 0; p < 3; ++p)
                            umma::mma_bf16(dacc, umma::desc_kmajor_sw64(a0, ks), umma::desc_kmajor_sw64(g0 + p * T::G_B, ks),
                                           id_g, (kb | ks | p) ? 1u : 0u);
                        if (eta) {
                            const uint64_t xh = umma::desc_kmajor_sw64(a0 + T::A_B, ks);
                            const uint64_t xm = umma::desc_kmajor_sw64(a0 + 2 * T::A_B, ks);
                            umma::mma_bf16(deta, xh, umma::desc_kmajor_sw64(w0, ks), id_w, 1u);
                            umma::mma_bf16(deta, xh, umma::desc_kmajor_sw64(w0 + T::W_B, ks), id_w, 1u);
                            umma::mma_bf16(deta, xm, umma::desc_kmajor_sw64(w0, ks), id_w, 1u);
                            umma::mma_bf16(deta, xh, umma::desc_kmajor_sw64(w0 + 2 * T::W_B, ks), id_w, 1u);
                            umma::mma_bf16(deta, xm, umma::desc_kmajor_sw64(w0 + T::W_B, ks), id_w, 1u);
                        }
                    }
                    umma::mma_commit(&empty[s]);               // the stage is free once these MMAs have read it
                }
                umma::mma_commit(&tfull[buf]);                 // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> [P0 | h0] + tau * acc -> global rows =====================
        const int wq = warp & 3;                               // TMEM lane quarter this warp may access
        const float tau = (float)gl[PYVB_GL_TAU];
        int tl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const int ct = tile % T::NCT;
            const long long row = (long long)(tile / T::NCT) * T::BM + wq * 32 + lane;
            const int c0 = ct * T::NT;
            const int nt = (T::NCP - c0 < T::NT) ? (T::NCP - c0) : T::NT;
            const int buf = tl & 1;
            umma::mbar_wait_bounded(&tfull[buf], (uint32_t)((tl >> 1) & 1));
            umma::fence_after_sync();
            const uint32_t taddr = tmem + (uint32_t)(buf * 256) + ((uint32_t)(wq * 32) << 16);
            float *orow = MZ + row * T::NCP + c0;
            for (int cc = 0; cc < nt; cc += 16) {
                uint32_t v[16];
                umma::tmem_ld16(taddr + (uint32_t)cc, v);
                umma::tmem_ld_wait();
                if (row < N) {
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + cc + j;
                        const float a = __uint_as_float(v[j]);
                        o[j] = (c < T::P) ? fmaf(tau, a, p0v[c]) : (c >= T::PP && c < T::PP + Q) ? fmaf(tau, a, h0s[c - T::PP]) : 0.0f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4 *>(orow + cc + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                }
            }
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    // ---- teardown
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

template <int Q>
cudaError_t launch_zstep_f32_q(long long N, int D, const void *planes, const void *GT, const void *WT, const double *P0,
                               const double *h0, const double *gl, float *MZ, cudaStream_t st) {
    using T = FT<Q>;
    CUtensorMap tmA, tmG, tmW;
    cudaError_t e = make_map_bf16_3d(&tmA, planes, (uint64_t)D, (uint64_t)N, 3, T::BK, T::BM, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
    e = make_map_bf16_3d(&tmG, GT, (uint64_t)D, (uint64_t)T::NCP, 3, T::BK, T::NT, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
    e = make_map_bf16_3d(&tmW, WT, (uint64_t)D, (uint64_t)Q, 3, T::BK, Q, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zstep_f32_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e != cudaSuccess) return e;
    const long long nt = ((N + T::BM - 1) / T::BM) * T::NCT;
    int grid = (int)(nt < 148 ? nt : 148);
    zstep_f32_kernel<Q><<<grid, T::NTHR, T::SMEM, st>>>(tmA, tmG, tmW, N, D, P0, h0, gl, MZ, (int)nt);
    return cudaGetLastError();
}

}  // namespace

int f32_ncp(int q) { return f_ncp(q); }
int f32_zoff(int q) { return f_pp(q); }
bool f32_supported(int D, int q) { return (q == 16 || q == 32 || q == 64) && D >= 32 && (D % 32) == 0; }

cudaError_t launch_prepare_x_f32(long long N, int D, const double *X, long long ldx, void *planes, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    long long b = (N * (long long)D + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    prepare_x_kernel<<<(unsigned)b, 256, 0, st>>>(N, D, X, ldx, static_cast<__nv_bfloat16 *>(planes));
    return cudaGetLastError();
}

cudaError_t launch_pack_gw_f32(int D, int q, const double *Wbar, const double *Wvar, const double *mu, void *GT, void *WT,
                               cudaStream_t st) {
    long long b = ((long long)(f_ncp(q) + q) * D + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    pack_gw_f32_kernel<<<(unsigned)b, 256, 0, st>>>(D, q, Wbar, Wvar, mu, static_cast<__nv_bfloat16 *>(GT),
                                                    static_cast<__nv_bfloat16 *>(WT));
    return cudaGetLastError();
}

cudaError_t launch_zstep_f32(long long N, int D, int q, const void *planes, const void *GT, const void *WT,
                             const double *P0, const double *h0, const double *gl, float *MZ, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    switch (q) {
        case 16: return launch_zstep_f32_q<16>(N, D, planes, GT, WT, P0, h0, gl, MZ, st);
        case 32: return launch_zstep_f32_q<32>(N, D, planes, GT, WT, P0, h0, gl, MZ, st);
        case 64: return launch_zstep_f32_q<64>(N, D, planes, GT, WT, P0, h0, gl, MZ, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace pyvb
