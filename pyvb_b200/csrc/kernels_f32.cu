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
// Output: MZ32 float [N][NCP] rows [eta (q) | qprec packed (P) | pad], NCP = (q + P) rounded up to 64.  (eta / zbar
// come FIRST in the FP32 row: the statistics kernel then finds zbar in the first 64-column atom of its B tiles.)
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include "umma.cuh"

namespace pyvb {

namespace {

__host__ __device__ constexpr int f_tri(int i) { return i * (i + 1) / 2; }
__host__ __device__ constexpr int f_ncp(int q) { return (q + f_tri(q) + 63) & ~63; }   // floats per FP32 row

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

// GT[p][c][d]: column c of the accumulator row ([-mu_d w_d (q) | G_d packed (P) | pad]) for data dimension d, plane p;
// WT[p][i][d] = plane p of <w_di>.  One thread per (c, d), d fastest (coalesced writes).
__global__ void __launch_bounds__(256)
pack_gw_f32_kernel(int D, int q, const double *__restrict__ Wbar, const double *__restrict__ Wvar,
                   const double *__restrict__ mu, __nv_bfloat16 *__restrict__ GT, __nv_bfloat16 *__restrict__ WT) {
    const int P = f_tri(q), NCP = f_ncp(q);
    const long long total = (long long)(NCP + q) * D;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e / D), d = (int)(e % D);
        double v = 0.0;
        __nv_bfloat16 *dst;
        size_t plane;
        if (c < NCP) {
            if (c < q) {
                v = -mu[d] * Wbar[(size_t)d * q + c];
            } else if (c < q + P) {
                int i, j;
                unpack_p(c - q, i, j);
                v = Wbar[(size_t)d * q + i] * Wbar[(size_t)d * q + j];
                if (i == j) v += Wvar[(size_t)d * q + i];
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
// rows_alloc: rows per plane in the allocation (>= rows: a row sub-range of a bigger array keeps its plane stride)
cudaError_t make_map_bf16_3d(CUtensorMap *m, const void *base, uint64_t cols, uint64_t rows, uint64_t rows_alloc,
                             uint64_t planes, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode_f32();
    if (!enc) return cudaErrorNotSupported;
    cuuint64_t dims[3] = {cols, rows, planes};
    cuuint64_t strides[2] = {cols * 2, cols * rows_alloc * 2};
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
    static constexpr int P = f_tri(Q), NCP = f_ncp(Q);
    static constexpr int NT = FC<Q>::NT, ST = FC<Q>::ST;
    static constexpr int NCT = (NCP + NT - 1) / NT;          // column tiles
    static constexpr int ECT = 0;                            // the column tile that holds the eta columns [0, q)
    static constexpr int BM = 128, BK = 32;                  // rows per tile, K elements per stage (64-byte rows)
    static constexpr int A_B = BM * BK * 2;                  // one A plane tile (bytes)
    static constexpr int G_B = NT * BK * 2;                  // one G plane tile
    static constexpr int W_B = Q * BK * 2;                   // one W plane tile
    static constexpr int STAGE_B = 3 * A_B + 3 * G_B + 3 * W_B;
    static constexpr int NTHR = 6 * 32;                      // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue
    static constexpr size_t SMEM = 1024 + (size_t)ST * STAGE_B + (size_t)(P + Q) * 4 + (2 * ST + 4) * 8 + 16;
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
        // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues: a loop
        // entered by lane 0 alone makes the compiler wrap every TMA / tcgen05 instruction in an election loop) ==========
        const bool leader = elect_one();
        {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int ct = tile % T::NCT;
                const int row0 = (tile / T::NCT) * T::BM;
                const bool eta = (ct == T::ECT);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % T::ST;
                    umma::mbar_wait_bounded(&empty[s], (uint32_t)(((it / T::ST) & 1) ^ 1));
                    unsigned char *st = stage0 + s * T::STAGE_B;
                    if (leader) {
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
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane issues) =====================
        const bool leader = elect_one();
        {
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                const int ct = tile % T::NCT;
                const bool eta = (ct == T::ECT);
                const int nt = (T::NCP - ct * T::NT < T::NT) ? (T::NCP - ct * T::NT) : T::NT;
                const int buf = tl & 1;
                umma::mbar_wait_bounded(&tempty[buf], (uint32_t)(((tl >> 1) & 1) ^ 1));
                umma::fence_after_sync();
                const uint32_t dacc = tmem + (uint32_t)(buf * 256);
                const uint32_t deta = dacc;                     // eta = the first q accumulator columns
                const uint32_t id_g = umma::idesc_bf16_f32(T::BM, nt, 0, 0);
                const uint32_t id_w = umma::idesc_bf16_f32(T::BM, Q, 0, 0);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % T::ST;
                    umma::mbar_wait_bounded(&full[s], (uint32_t)((it / T::ST) & 1));
                    umma::fence_after_sync();
                    const uint32_t a0 = smem_u32(stage0 + s * T::STAGE_B);
                    const uint32_t g0 = a0 + 3 * T::A_B, w0 = g0 + 3 * T::G_B;
                    if (leader) {
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
                    __syncwarp();
                }
                if (leader) umma::mma_commit(&tfull[buf]);     // accumulator complete
                __syncwarp();
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
                        o[j] = (c < Q) ? fmaf(tau, a, h0s[c]) : (c < Q + T::P) ? fmaf(tau, a, p0v[c - Q]) : 0.0f;
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
cudaError_t launch_zstep_f32_q(long long N, long long nalloc, int D, const void *planes, const void *GT, const void *WT,
                               const double *P0, const double *h0, const double *gl, float *MZ, cudaStream_t st) {
    using T = FT<Q>;
    CUtensorMap tmA, tmG, tmW;
    cudaError_t e = make_map_bf16_3d(&tmA, planes, (uint64_t)D, (uint64_t)N, (uint64_t)nalloc, 3, T::BK, T::BM,
                                     CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
    e = make_map_bf16_3d(&tmG, GT, (uint64_t)D, (uint64_t)T::NCP, (uint64_t)T::NCP, 3, T::BK, T::NT, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
    e = make_map_bf16_3d(&tmW, WT, (uint64_t)D, (uint64_t)Q, (uint64_t)Q, 3, T::BK, Q, CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zstep_f32_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e != cudaSuccess) return e;
    const long long nt = ((N + T::BM - 1) / T::BM) * T::NCT;
    int grid = (int)(nt < 148 ? nt : 148);
    zstep_f32_kernel<Q><<<grid, T::NTHR, T::SMEM, st>>>(tmA, tmG, tmW, N, D, P0, h0, gl, MZ, (int)nt);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ K3, FP32 variant: statistics GEMM
// T[d][c] = sum_n A[n][d] B[n][c] over a chunk of rows: hstack.pass_up_m1_m2's sums over the rows
// (nodes/nodes_todo.py:50-61).  A = mask (O-type: T1 = O^T vec<zz^T>, Bst = O^T zbar) or x_h + x_m (X-type:
// Ast = (O.X)^T zbar); B = the three-way bf16 split of the finished FP32 rows [zbar | <zz^T> packed] (planes MP
// written by the batched solve).  Both operands sit in memory with the contraction index (the row n) OUTERMOST, so
// they are fed to tcgen05.mma as MN-major tiles: TMA boxes of [rows][64 elements] with SWIZZLE_128B are exactly
// the canonical MN-major atoms, no transposition anywhere.  Accumulators: 128 d (TMEM lanes) x NT columns + 64
// columns for the X-type product (zbar = the first q columns of the first 64-column atom).  Partial sums per row
// chunk leave as FP64 into the same workspace layout as the FP64 path; the fixed-order second stage is shared.
template <int Q> struct SFC;
template <> struct SFC<16> { static constexpr int NT = 192, ST = 3; };
template <> struct SFC<32> { static constexpr int NT = 192, ST = 3; };
template <> struct SFC<64> { static constexpr int NT = 256, ST = 3; };

template <int Q> struct SFT {
    static constexpr int P = f_tri(Q), NCP = f_ncp(Q);
    static constexpr int NT = SFC<Q>::NT, ST = SFC<Q>::ST;
    static constexpr int NCT = (NCP + NT - 1) / NT;
    static constexpr int BD = 128, BKN = 32;                  // data dimensions per tile, rows per stage
    static constexpr int BLK_B = BKN * 128;                   // one [BKN rows][64 elements] box (bytes)
    static constexpr int A_B = 2 * BLK_B;                     // one A plane tile: 128 d
    static constexpr int B_B = (NT / 64) * BLK_B;             // one B plane tile: NT columns
    static constexpr int STAGE_B = 3 * A_B + 3 * B_B;
    static constexpr int NTHR = 6 * 32;
    static constexpr int TCOLS = (NT + 64 <= 256) ? 256 : 512;
    static constexpr size_t SMEM = 1024 + (size_t)ST * STAGE_B + (2 * ST + 2) * 8 + 16;
    static_assert(NT % 64 == 0 && BLK_B % 1024 == 0, "MN-major SWIZZLE_128B atoms");
};

template <int Q>
__global__ void __launch_bounds__(SFT<Q>::NTHR, 1)
stats_f32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, long long N, int D,
                 double *__restrict__ ws, long long rows_per_chunk) {
    using T = SFT<Q>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *stage0 = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + T::ST * T::STAGE_B);
    uint64_t *empty = full + T::ST;
    uint64_t *tfull = empty + T::ST;
    uint32_t *tbase = reinterpret_cast<uint32_t *>(tfull + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ct = (int)(blockIdx.x % T::NCT);
    const int d0 = (int)(blockIdx.x / T::NCT) * T::BD;
    const int c0 = ct * T::NT;
    const int nt = (T::NCP - c0 < T::NT) ? (T::NCP - c0) : T::NT;
    const bool xt = (ct == 0);                                // this tile also carries the X-type product
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > N) r1 = N;
    const int nsteps = (r1 > r0) ? (int)((r1 - r0 + T::BKN - 1) / T::BKN) : 0;

    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < T::ST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tfull[0], 1);
        mbar_fence_init();
    }
    if (warp == 1) umma::tmem_alloc(tbase, T::TCOLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tbase;

    if (warp == 0) {
        const bool leader = elect_one();                   // whole warp, one elected lane issues (see zstep_f32_kernel)
        {
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % T::ST;
                umma::mbar_wait_bounded(&empty[s], (uint32_t)(((it / T::ST) & 1) ^ 1));
                unsigned char *st = stage0 + s * T::STAGE_B;
                const int nb = (int)(r0 + (long long)it * T::BKN);       // rows >= N are zero-filled by the copies
                if (leader) {
                mbar_arrive_expect_tx(&full[s], (uint32_t)((xt ? 3 : 1) * T::A_B + 3 * (nt / 64) * T::BLK_B));
                for (int p = 0; p < (xt ? 3 : 1); ++p)
                    for (int j = 0; j < 2; ++j)
                        tma_load_3d(st + p * T::A_B + j * T::BLK_B, &tmA, d0 + j * 64, nb, p, &full[s]);
                for (int p = 0; p < 3; ++p)
                    for (int j = 0; j < nt / 64; ++j)
                        tma_load_3d(st + 3 * T::A_B + p * T::B_B + j * T::BLK_B, &tmB, c0 + j * 64, nb, p, &full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        {
            const uint32_t id_t = umma::idesc_bf16_f32(T::BD, nt, 1, 1);
            const uint32_t id_x = umma::idesc_bf16_f32(T::BD, 64, 1, 1);
            const uint32_t dT = tmem, dX = tmem + (uint32_t)T::NT;
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % T::ST;
                umma::mbar_wait_bounded(&full[s], (uint32_t)((it / T::ST) & 1));
                umma::fence_after_sync();
                const uint32_t a0 = smem_u32(stage0 + s * T::STAGE_B);
                const uint32_t b0 = a0 + 3 * T::A_B;
                if (leader) {
#pragma unroll
                for (int ks = 0; ks < T::BKN / 16; ++ks) {
                    const uint64_t am = umma::desc_mnmajor_sw128(a0, ks, T::BLK_B);
#pragma unroll
                    for (int p = 0; p < 3; ++p)
                        umma::mma_bf16(dT, am, umma::desc_mnmajor_sw128(b0 + p * T::B_B, ks, T::BLK_B), id_t,
                                       (it | ks | p) ? 1u : 0u);
                    if (xt) {
                        const uint64_t xh = umma::desc_mnmajor_sw128(a0 + T::A_B, ks, T::BLK_B);
                        const uint64_t xm = umma::desc_mnmajor_sw128(a0 + 2 * T::A_B, ks, T::BLK_B);
                        const uint64_t zh = umma::desc_mnmajor_sw128(b0, ks, T::BLK_B);
                        const uint64_t zm = umma::desc_mnmajor_sw128(b0 + T::B_B, ks, T::BLK_B);
                        const uint64_t zl = umma::desc_mnmajor_sw128(b0 + 2 * T::B_B, ks, T::BLK_B);
                        umma::mma_bf16(dX, xh, zh, id_x, (it | ks) ? 1u : 0u);
                        umma::mma_bf16(dX, xh, zm, id_x, 1u);
                        umma::mma_bf16(dX, xm, zh, id_x, 1u);
                        umma::mma_bf16(dX, xh, zl, id_x, 1u);
                        umma::mma_bf16(dX, xm, zm, id_x, 1u);
                    }
                }
                umma::mma_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (leader) umma::mma_commit(&tfull[0]);
            __syncwarp();
        }
    } else {
        // ===================== epilogue: lane = data dimension d, columns -> T1 / Bst / Ast partials (FP64) ==========
        const int wq = warp & 3;
        const int d = d0 + wq * 32 + lane;
        const StatLayout L(D, Q);
        double *out = ws + (size_t)blockIdx.y * L.len;
        const bool live = nsteps > 0;                          // an empty chunk contributes zeros
        if (live) umma::mbar_wait_bounded(&tfull[0], 0u);
        umma::fence_after_sync();
        const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16);
        for (int cc = 0; cc < nt; cc += 16) {
            uint32_t v[16];
            umma::tmem_ld16(taddr + (uint32_t)cc, v);
            umma::tmem_ld_wait();
            if (d < D) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = c0 + cc + j;
                    const double a = live ? (double)__uint_as_float(v[j]) : 0.0;
                    if (c < Q) out[L.bst + (size_t)d * Q + c] = a;
                    else if (c < Q + T::P) out[L.t1 + (size_t)d * T::P + (c - Q)] = a;
                }
            }
        }
        if (xt) {
            for (int cc = 0; cc < Q; cc += 16) {
                uint32_t v[16];
                umma::tmem_ld16(taddr + (uint32_t)(T::NT + cc), v);
                umma::tmem_ld_wait();
                if (d < D) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        out[L.ast + (size_t)d * Q + cc + j] = live ? (double)__uint_as_float(v[j]) : 0.0;
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, T::TCOLS);
}

template <int Q>
cudaError_t launch_stats_f32_q(long long N, long long nalloc, int D, const void *planes, const void *MP, double *ws,
                               int nchunks, cudaStream_t st) {
    using T = SFT<Q>;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map_bf16_3d(&tmA, planes, (uint64_t)D, (uint64_t)N, (uint64_t)nalloc, 3, 64, T::BKN,
                                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (e != cudaSuccess) return e;
    e = make_map_bf16_3d(&tmB, MP, (uint64_t)T::NCP, (uint64_t)N, (uint64_t)nalloc, 3, 64, T::BKN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(stats_f32_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e != cudaSuccess) return e;
    long long rpc = (N + nchunks - 1) / nchunks;
    rpc = ((rpc + T::BKN - 1) / T::BKN) * T::BKN;
    if (rpc < T::BKN) rpc = T::BKN;
    dim3 grid((unsigned)(((D + T::BD - 1) / T::BD) * T::NCT), (unsigned)nchunks);
    stats_f32_kernel<Q><<<grid, T::NTHR, T::SMEM, st>>>(tmA, tmB, N, D, ws, rpc);
    return cudaGetLastError();
}

}  // namespace

// row chunks of the FP32 statistics kernel: ~2 waves of CTAs, at least 64 pipeline steps (2048 rows) per chunk
int stats_f32_nchunks(long long N, int D, int q) {
    const int nct = (f_ncp(q) + (q == 64 ? 256 : 192) - 1) / (q == 64 ? 256 : 192);
    const long long per = (long long)((D + 127) / 128) * nct;
    long long c = (2 * 148 + per - 1) / per;
    const long long by_rows = (N + 2047) / 2048;
    if (c > by_rows) c = by_rows;
    if (c < 1) c = 1;
    if (c > 1024) c = 1024;
    return (int)c;
}

cudaError_t launch_stats_f32(long long N, long long nalloc, int D, int q, const void *planes, const void *MP, double *ws,
                             int nchunks, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    switch (q) {
        case 16: return launch_stats_f32_q<16>(N, nalloc, D, planes, MP, ws, nchunks, st);
        case 32: return launch_stats_f32_q<32>(N, nalloc, D, planes, MP, ws, nchunks, st);
        case 64: return launch_stats_f32_q<64>(N, nalloc, D, planes, MP, ws, nchunks, st);
    }
    return cudaErrorNotSupported;
}

int f32_ncp(int q) { return f_ncp(q); }
int f32_zoff(int q) { (void)q; return 0; }
int f32_poff(int q) { return q; }
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

cudaError_t launch_zstep_f32(long long N, long long nalloc, int D, int q, const void *planes, const void *GT,
                             const void *WT, const double *P0, const double *h0, const double *gl, float *MZ,
                             cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    switch (q) {
        case 16: return launch_zstep_f32_q<16>(N, nalloc, D, planes, GT, WT, P0, h0, gl, MZ, st);
        case 32: return launch_zstep_f32_q<32>(N, nalloc, D, planes, GT, WT, P0, h0, gl, MZ, st);
        case 64: return launch_zstep_f32_q<64>(N, nalloc, D, planes, GT, WT, P0, h0, gl, MZ, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace pyvb
