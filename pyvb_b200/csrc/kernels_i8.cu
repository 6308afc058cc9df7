// Exact FP64 mask contractions on the INT8 tensor cores (tcgen05.mma kind::i8, INT32 accumulators in TMEM).
//
// The dominant part of the Z step, qprec_n = P0 + tau * sum_d O_nd G_d (Multiplication.pass_up_m1_m2's
// m1 = tr(<w_i w_j^T> Lambda_n), nodes/node.py:213-227, with a masked precision), multiplies a 0/1 matrix with a small
// real one.  The mask is exact in int8; every column of G is written in fixed point with its own scale,
//     G[d][c] = scale_c 2^-54 sum_{t<7} digit_t[d][c] 256^t,     digit_t in [-128, 127]   (balanced base 256),
// |round(G / scale_c 2^54)| <= 2^54 < 2^55 = the range of seven balanced digits, i.e. every entry is kept to
// scale_c 2^-55 absolute -- finer than the FP64 spacing of the column's large entries.  Then
//     mask @ G = scale_c 2^-54 sum_t 256^t (mask @ digit_t)
// where every mask @ digit_t is an EXACT integer GEMM (|sum| <= 128 D << 2^31).  The seven INT32 results per output
// are recombined in the epilogue (pairs in INT32, then three FP64 FMAs): no accumulation error at all, one rounding
// at the end -- at least as accurate as an FP64 accumulation, at the INT8 tensor rate instead of the FP64 (DMMA) one.
// The eta columns (both operands real) stay on the DMMA kernel (ZT<Q, true> in kernels_dmma.cu).
//
// Tiling: CTA tile = 128 rows x 32 output columns x 7 digit planes = one tcgen05.mma with N = 224 per 32-byte K step.
// The 128 x D mask block of a row tile stays RESIDENT in shared memory (D / 64 chunks of 128 x 64 bytes, each with its
// own full / empty mbarrier) while the CTA walks over the column tiles; only the 14 KB digit tiles stream through a
// 4-stage ring (they come from L2: the whole digit array is a few MB).  TMEM holds two accumulators (2 x 256 columns)
// so that the epilogue of one tile overlaps the MMAs of the next.  Warp 0: TMA producer, warp 1: MMA issuer, warps
// 2-9: epilogue (two warps per TMEM lane quarter, 16 columns each).  Operands are K-major TMA tiles with 64-byte rows
// (SWIZZLE_64B), the same byte geometry as the bf16 kernels of kernels_f32.cu.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include "umma.cuh"

namespace pyvb {

namespace {

__host__ __device__ constexpr int i_tri(int i) { return i * (i + 1) / 2; }
__host__ __device__ constexpr int i_nc8(int q) { return (i_tri(q) + 31) & ~31; }          // packed columns rounded to 32
constexpr int NPL = 7;                                                                    // digit planes
constexpr int BM = 128, BKB = 64, ST = 4, CT = 32;       // rows per tile, K bytes per chunk, digit stages, columns per tile
constexpr int A_B = BM * BKB, B_B = NPL * CT * BKB;      // 8192, 14336
constexpr int NTHR = 10 * 32;

// ------------------------------------------------------------------ operand preparation
__global__ void __launch_bounds__(256)
prepare_mask_i8_kernel(long long N, int D, const double *__restrict__ X, long long ldx, signed char *__restrict__ mask) {
    const long long total = N * (long long)(D / 4);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / (D / 4);
        const int d = (int)(e - n * (D / 4)) * 4;
        const double2 a = *reinterpret_cast<const double2 *>(X + n * ldx + d);
        const double2 b = *reinterpret_cast<const double2 *>(X + n * ldx + d + 2);
        char4 m;
        m.x = (a.x == a.x) ? 1 : 0;
        m.y = (a.y == a.y) ? 1 : 0;
        m.z = (b.x == b.x) ? 1 : 0;
        m.w = (b.y == b.y) ? 1 : 0;
        *reinterpret_cast<char4 *>(mask + n * D + d) = m;
    }
}

// one CTA per output column c: scale_c = max_d |G[d][c]|, then the seven balanced base-256 digits of
// round(G / scale_c * 2^54).  GI[ct][plane][c % 32][d] (int8), ct = c / 32: the B tile of column tile ct is one
// [224 rows][D] block.
__global__ void __launch_bounds__(256)
pack_g_i8_kernel(int D, int q, const double *__restrict__ Wbar, const double *__restrict__ Wvar,
                 signed char *__restrict__ GI, double *__restrict__ gscale) {
    __shared__ double sh[33];
    const int P = i_tri(q);
    const int c = blockIdx.x;
    int i = 0, j = 0;
    if (c < P) unpack_p(c, i, j);
    double mx = 0.0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double g = 0.0;
        if (c < P) {
            g = Wbar[(size_t)d * q + i] * Wbar[(size_t)d * q + j];
            if (i == j) g += Wvar[(size_t)d * q + i];
        }
        mx = fmax(mx, fabs(g));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (threadIdx.x == 0) sh[32] = t;
    }
    __syncthreads();
    const double scale = (sh[32] > 0.0 && sh[32] < 1e300) ? sh[32] : 1.0;
    if (threadIdx.x == 0) gscale[c] = scale;
    const double inv = 18014398509481984.0 / scale;                     // 2^54 / scale
    signed char *base = GI + ((size_t)(c >> 5) * NPL * CT + (c & 31)) * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double g = 0.0;
        if (c < P) {
            g = Wbar[(size_t)d * q + i] * Wbar[(size_t)d * q + j];
            if (i == j) g += Wvar[(size_t)d * q + i];
        }
        long long v = __double2ll_rn(g * inv);
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const long long dg = ((v + 128) & 255) - 128;               // balanced digit in [-128, 127]
            base[(size_t)t * CT * D + d] = (signed char)dg;
            v = (v - dg) >> 8;
        }
    }
}

// ------------------------------------------------------------------ tensor maps (uint8)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_i8() {
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
// [planes][rows][cols] bytes, row pitch = cols
cudaError_t make_map_u8_3d(CUtensorMap *m, const void *base, uint64_t cols, uint64_t rows, uint64_t planes,
                           uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_i8();
    if (!enc) return cudaErrorNotSupported;
    cuuint64_t dims[3] = {cols, rows, planes};
    cuuint64_t strides[2] = {cols, cols * rows};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
__device__ __forceinline__ void tma_load_3d_i8(void *dst_smem, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// exact int32 -> double without the conversion unit: 2^52 + 2^31 + a has the integer in its low mantissa word
__device__ __forceinline__ double i2d(int a) {
    return __hiloint2double(0x43300000, (int)((unsigned)a ^ 0x80000000u)) - 4503601774854144.0;
}
// sum_t 256^t a_t for seven INT32 accumulators (|a_t| <= 2^22: pairs fit INT32); one rounding (the last FMA)
__device__ __forceinline__ double combine7(const int (&a)[NPL]) {
    const int t01 = a[0] + (a[1] * 256), t23 = a[2] + (a[3] * 256), t45 = a[4] + (a[5] * 256);
    double v = fma(i2d(a[6]), 65536.0, i2d(t45));
    v = fma(v, 65536.0, i2d(t23));
    return fma(v, 65536.0, i2d(t01));
}

struct I8Geom {
    int P, PP, NC8, NCT;
    __host__ __device__ explicit I8Geom(int q) {
        P = i_tri(q);
        PP = (P + 7) & ~7;
        NC8 = i_nc8(q);
        NCT = NC8 / CT;
    }
};
size_t i8_smem_bytes(int D, int q) {
    const I8Geom g(q);
    const int nk = D / BKB;
    return 1024 + (size_t)nk * A_B + (size_t)ST * B_B + (size_t)(g.PP + g.NC8) * 8 + (size_t)(2 * nk + 2 * ST + 4) * 8 + 16;
}

// ------------------------------------------------------------------ K1-i8: qprec columns of the MZ rows
__global__ void __launch_bounds__(NTHR, 1)
zstep_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, long long N, int D, int q,
                const double *__restrict__ P0, const double *__restrict__ gscale, const double *__restrict__ gl,
                double *__restrict__ MZ, int ldmz, int nrb) {
    const I8Geom G(q);
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int nk = D / BKB;
    unsigned char *a_base = smem;                                              // [nk][128 x 64 B] resident mask block
    unsigned char *b_base = smem + (size_t)nk * A_B;                           // [ST][224 x 64 B] digit tiles
    double *p0v = reinterpret_cast<double *>(b_base + ST * B_B);               // [PP]: packed P0, zero pad
    double *fcol = p0v + G.PP;                                                 // [NC8]: tau * scale_c * 2^-54
    uint64_t *afull = reinterpret_cast<uint64_t *>(fcol + G.NC8);              // [nk]
    uint64_t *aempty = afull + nk;                                             // [nk]
    uint64_t *full = aempty + nk;                                              // [ST]
    uint64_t *empty = full + ST;                                               // [ST]
    uint64_t *tfull = empty + ST;                                              // [2]
    uint64_t *tempty = tfull + 2;                                              // [2]
    uint32_t *tbase = reinterpret_cast<uint32_t *>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double tau = gl[PYVB_GL_TAU];

    for (int p = tid; p < G.PP; p += NTHR) {
        double v = 0.0;
        if (p < G.P) {
            int i, j;
            unpack_p(p, i, j);
            v = P0[i * q + j];
        }
        p0v[p] = v;
    }
    for (int c = tid; c < G.NC8; c += NTHR) fcol[c] = tau * gscale[c] * 5.5511151231257827e-17;   // 2^-54
    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int k = 0; k < nk; ++k) {
            mbar_init(&afull[k], 1);
            mbar_init(&aempty[k], 1);
        }
        for (int s = 0; s < ST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 8);
        }
        mbar_fence_init();
    }
    if (warp == 1) umma::tmem_alloc(tbase, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tbase;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0, rl = 0;
            for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x, ++rl) {
                const int row0 = rb * BM;
                for (int ct = 0; ct < G.NCT; ++ct) {
                    for (int kb = 0; kb < nk; ++kb, ++it) {
                        if (ct == 0) {      // the mask chunk of the previous row block has been consumed by its last column tile
                            umma::mbar_wait_bounded(&aempty[kb], (uint32_t)((rl & 1) ^ 1));
                            mbar_arrive_expect_tx(&afull[kb], (uint32_t)A_B);
                            tma_load_3d_i8(a_base + (size_t)kb * A_B, &tmA, kb * BKB, row0, 0, &afull[kb]);   // rows past N: zero fill
                        }
                        const int s = it % ST;
                        umma::mbar_wait_bounded(&empty[s], (uint32_t)(((it / ST) & 1) ^ 1));
                        mbar_arrive_expect_tx(&full[s], (uint32_t)B_B);
                        tma_load_3d_i8(b_base + s * B_B, &tmB, kb * BKB, 0, ct, &full[s]);                    // 7 planes x 32 columns
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma::idesc_s8_s32(BM, NPL * CT);
            int it = 0, tl = 0, rl = 0;
            for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x, ++rl) {
                for (int ct = 0; ct < G.NCT; ++ct, ++tl) {
                    const int buf = tl & 1;
                    umma::mbar_wait_bounded(&tempty[buf], (uint32_t)(((tl >> 1) & 1) ^ 1));
                    umma::fence_after_sync();
                    const uint32_t dacc = tmem + (uint32_t)(buf * 256);
                    for (int kb = 0; kb < nk; ++kb, ++it) {
                        const int s = it % ST;
                        if (ct == 0) umma::mbar_wait_bounded(&afull[kb], (uint32_t)(rl & 1));
                        umma::mbar_wait_bounded(&full[s], (uint32_t)((it / ST) & 1));
                        umma::fence_after_sync();
                        const uint32_t a0 = smem_u32(a_base + (size_t)kb * A_B), b0 = smem_u32(b_base + s * B_B);
#pragma unroll
                        for (int ks = 0; ks < BKB / 32; ++ks)
                            umma::mma_i8(dacc, umma::desc_kmajor_sw64(a0, ks), umma::desc_kmajor_sw64(b0, ks), idesc,
                                         (kb | ks) ? 1u : 0u);
                        umma::mma_commit(&empty[s]);
                        if (ct == G.NCT - 1) umma::mma_commit(&aempty[kb]);
                    }
                    umma::mma_commit(&tfull[buf]);
                }
            }
        }
    } else {
        // ===================== epilogue: 7 INT32 planes -> FP64 -> P0 + f_c * v -> MZ row ========
        const int wq = warp & 3;                               // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                      // which 16 of the tile's 32 columns
        int tl = 0;
        for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            const long long row = (long long)rb * BM + wq * 32 + lane;
            for (int ct = 0; ct < G.NCT; ++ct, ++tl) {
                const int c0 = ct * CT + half * 16;
                const int buf = tl & 1;
                umma::mbar_wait_bounded(&tfull[buf], (uint32_t)((tl >> 1) & 1));
                umma::fence_after_sync();
                const uint32_t taddr = tmem + (uint32_t)(buf * 256 + half * 16) + ((uint32_t)(wq * 32) << 16);
                uint32_t a[2][NPL][8];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int p = 0; p < NPL; ++p) tmem_ld8(taddr + (uint32_t)(p * CT + ch * 8), a[ch][p]);
                umma::tmem_ld_wait();
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);       // the accumulator buffer is free again
                if (row < N) {
                    double *orow = MZ + row * ldmz + c0;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        if (c0 + ch * 8 < G.PP) {               // PP is a multiple of 8: a chunk is entirely in or out
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                int d0[NPL], d1[NPL];
#pragma unroll
                                for (int p = 0; p < NPL; ++p) {
                                    d0[p] = (int)a[ch][p][k];
                                    d1[p] = (int)a[ch][p][k + 1];
                                }
                                const int c = c0 + ch * 8 + k;
                                double2 o;
                                o.x = fma(fcol[c], combine7(d0), p0v[c]);
                                o.y = fma(fcol[c + 1], combine7(d1), p0v[c + 1]);
                                *reinterpret_cast<double2 *>(orow + ch * 8 + k) = o;
                            }
                        }
                    }
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

bool i8_supported(int D, int q) {
    return (q == 16 || q == 32 || q == 64) && D >= 64 && (D % 64) == 0 && i8_smem_bytes(D, q) <= 227 * 1024;
}
size_t i8_digits_bytes(int D, int q) { return (size_t)(i_nc8(q) / CT) * NPL * CT * D; }
int i8_ncols(int q) { return i_nc8(q); }

cudaError_t launch_prepare_mask_i8(long long N, int D, const double *X, long long ldx, void *mask, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    long long b = (N * (long long)(D / 4) + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    prepare_mask_i8_kernel<<<(unsigned)b, 256, 0, st>>>(N, D, X, ldx, static_cast<signed char *>(mask));
    return cudaGetLastError();
}

cudaError_t launch_pack_g_i8(int D, int q, const double *Wbar, const double *Wvar, void *GI, double *gscale,
                             cudaStream_t st) {
    pack_g_i8_kernel<<<i_nc8(q), 256, 0, st>>>(D, q, Wbar, Wvar, static_cast<signed char *>(GI), gscale);
    return cudaGetLastError();
}

cudaError_t launch_zstep_i8(long long N, int D, int q, const void *mask, const void *GI, const double *P0,
                            const double *gscale, const double *gl, double *MZ, int ldmz, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    if (!i8_supported(D, q)) return cudaErrorNotSupported;
    const I8Geom g(q);
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map_u8_3d(&tmA, mask, (uint64_t)D, (uint64_t)N, 1, BKB, BM);
    if (e != cudaSuccess) return e;
    e = make_map_u8_3d(&tmB, GI, (uint64_t)D, (uint64_t)(NPL * CT), (uint64_t)g.NCT, BKB, NPL * CT);
    if (e != cudaSuccess) return e;
    const size_t smem = i8_smem_bytes(D, q);
    e = cudaFuncSetAttribute(zstep_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long nrb = (N + BM - 1) / BM;
    const int grid = (int)(nrb < 148 ? nrb : 148);
    zstep_i8_kernel<<<grid, NTHR, smem, st>>>(tmA, tmB, N, D, q, P0, gscale, gl, MZ, ldmz, (int)nrb);
    return cudaGetLastError();
}

}  // namespace pyvb
