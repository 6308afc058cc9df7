// K2, blocked: batched q x q SPD inverse / solve with the FP64 tensor cores (DMMA.8x8x4), one warp per matrix.
//
// Replaces, per row n, the reference's  cho_factor(qprec) / cho_solve(., I) / dot(qcov, .)  and
// q_ln_det = .5/log(prod(diag(chol)))  (nodes/gaussian.py:117-123) for q in {8, 16, 32, 64}.
//
// The matrix lives in shared memory as NB x NB lower-triangular storage of 8 x 8 blocks (NB = q/8, 64 doubles per
// block, 32-byte chunks XOR-swizzled so that every DMMA fragment pattern -- row-wise, transposed, accumulator --
// is bank-conflict free without padding).  Three in-place block sweeps, LAPACK potrf / trtri / lauum style:
//   1. left-looking Cholesky:   C_ij = A_ij - sum_k L_ik L_jk^T (DMMA);  diagonal block factored AND inverted in
//      registers with warp shuffles (it already sits in the accumulator layout);  L_ij = C_ij X_jj^T (DMMA)
//   2. X = L^-1:                X_ik = -X_ii sum_j L_ij X_jk   (DMMA, a block row at a time)
//   3. Sigma = X^T X:           S_ij = sum_k X_ki^T X_kj       (DMMA)
// then zbar = Sigma eta, <zz^T> = Sigma + zbar zbar^T, ln prod diag chol from the pivots.  Only the 8 x 8 diagonal
// blocks run on the FP64 FMA pipe (8 sequential pivots each); everything else is tensor-core work.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace pyvb {

namespace {

__host__ __device__ constexpr int kb_tri(int i) { return i * (i + 1) / 2; }
__host__ __device__ constexpr int kb_pitch(int q) {          // == pyvb_gw_pitch(q)
    int p = ((kb_tri(q) + 7) & ~7) + q + 1;
    while ((p % 8) != 4) ++p;
    return p;
}
// element (r, c) of a swizzled 8 x 8 block
__host__ __device__ constexpr int kb_sw(int r, int c) { return r * 8 + ((((c >> 2) ^ ((r >> 1) & 1))) << 2) + (c & 3); }
// block (i, j), i >= j
__host__ __device__ constexpr int kb_boff(int i, int j) { return (kb_tri(i) + j) * 64; }

template <int Q> struct KBC;
template <> struct KBC<8>  { static constexpr int WARPS = 8,  OCC = 3; static constexpr bool ZS = true; };
template <> struct KBC<16> { static constexpr int WARPS = 8,  OCC = 2; static constexpr bool ZS = true; };
template <> struct KBC<32> { static constexpr int WARPS = 10, OCC = 2; static constexpr bool ZS = true; };
template <> struct KBC<64> { static constexpr int WARPS = 11, OCC = 1; static constexpr bool ZS = false; };

template <int Q> struct KB {
    static constexpr int NB = Q / 8, NBLK = kb_tri(NB);
    static constexpr int P = kb_tri(Q), PP = (P + 7) & ~7, OROW = PP + Q, LDG = kb_pitch(Q);
    static constexpr int WARPS = KBC<Q>::WARPS, OCC = KBC<Q>::OCC;
    static constexpr bool ZS = KBC<Q>::ZS;
    static constexpr int KW = OROW + PYVB_ZS_EXTRA;
    static constexpr int WARP_D = NBLK * 64 + 2 * Q + (ZS ? OROW : 0) + 4;     // blocks | eta | z | column sums | scalars
    static constexpr int TAB_B = ((P * 4) + 15) & ~15;                         // one uint32 table
    static constexpr int UNR = (P / 32 >= 16) ? 16 : (P + 31) / 32;            // global loads in flight per lane
    static constexpr size_t SMEM = (size_t)TAB_B + (size_t)WARPS * WARP_D * 8;
};

// Row format of the batched solve.  FP64: the interleaved MZ rows [packed (P) | pad | eta/zbar (q)], pitch
// pyvb_mz_pitch(q).  FP32 variant: float rows [eta/zbar (q) | packed (P) | pad] of pitch pyvb_f32_pitch(q); the
// arithmetic stays FP64 (the rows are converted on load / store), and the finished row is also written as a
// three-way bf16 split (planes [3][N][pitch]) -- the B operand of the FP32 statistics kernel.
template <int Q, bool F32> struct KIO;
template <int Q> struct KIO<Q, false> {
    using type = double;
    static constexpr int PITCH = kb_pitch(Q), POFF = 0, ZOFF = (kb_tri(Q) + 7) & ~7, USED = ZOFF + Q;
};
template <int Q> struct KIO<Q, true> {
    using type = float;
    static constexpr int PITCH = (Q + kb_tri(Q) + 63) & ~63, POFF = Q, ZOFF = 0, USED = Q + kb_tri(Q);
};

__device__ __forceinline__ void store_split3(__nv_bfloat16 *dst, size_t plane, float v) {
    const __nv_bfloat16 h = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(m);
    dst[0] = h;
    dst[plane] = m;
    dst[2 * plane] = __float2bfloat16(r2);
}

// 8 x 8 diagonal block, accumulator layout (lane l: row l/4, columns 2(l%4), 2(l%4)+1).  In: the SPD block A
// (lower triangle valid).  Out: X = chol(A)^-1 (lower triangular, exact zeros above the diagonal); lprod is
// multiplied by prod_k 1/l_kk.  Right-looking factorisation and right-looking inversion fused in one sweep:
// per pivot one broadcast, one rsqrt, and independent multiply-adds.
__device__ __forceinline__ void diag_chol_inv(double c0, double c1, double &x0, double &x1, double &lprod) {
    const int lane = threadIdx.x & 31;
    const int gid = lane >> 2, qd = lane & 3;
    x0 = (gid == 2 * qd) ? 1.0 : 0.0;
    x1 = (gid == 2 * qd + 1) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int kq = k >> 1;
        const double ck = (k & 1) ? c1 : c0;
        const double d = __shfl_sync(0xffffffffu, ck, k * 4 + kq);                 // pivot a_kk
        const double rinv = rsqrt(d);
        lprod *= rinv;
        const double lik = __shfl_sync(0xffffffffu, ck, (lane & ~3) | kq) * rinv;   // l_(gid, k)   (gid >= k)
        const double lj0 = __shfl_sync(0xffffffffu, ck, (2 * qd) * 4 + kq) * rinv;  // l_(2qd, k)
        const double lj1 = __shfl_sync(0xffffffffu, ck, (2 * qd + 1) * 4 + kq) * rinv;
        if (2 * qd > k) c0 = fma(-lik, lj0, c0);
        if (2 * qd + 1 > k) c1 = fma(-lik, lj1, c1);
        // row k of X is final once scaled by 1/l_kk; eliminate it from the rows below
        if (gid == k) {
            x0 *= rinv;
            x1 *= rinv;
        }
        const double xk0 = __shfl_sync(0xffffffffu, x0, k * 4 + qd);
        const double xk1 = __shfl_sync(0xffffffffu, x1, k * 4 + qd);
        if (gid > k) {
            x0 = fma(-lik, xk0, x0);
            x1 = fma(-lik, xk1, x1);
        }
    }
    if (gid < 2 * qd) x0 = 0.0;
    if (gid < 2 * qd + 1) x1 = 0.0;
}

template <int Q, bool F32>
__global__ void __launch_bounds__(32 * KB<Q>::WARPS, KB<Q>::OCC)
zsolve_blocked_kernel(long long N, typename KIO<Q, F32>::type *__restrict__ MZ, double *__restrict__ Sig,
                      double *__restrict__ logdet, double *gl, double *__restrict__ zsums,
                      __nv_bfloat16 *__restrict__ MP) {
    using T = KB<Q>;
    using IO = KIO<Q, F32>;
    using io_t = typename IO::type;
    constexpr int NB = T::NB;
    extern __shared__ __align__(16) unsigned char smem_kb[];
    // packed index p -> offset in the block storage (bits 0-11) | i (bits 12-17) | j (bits 18-23)
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_kb);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *blk = reinterpret_cast<double *>(smem_kb + T::TAB_B) + (size_t)warp * T::WARP_D;
    double *eta = blk + T::NBLK * 64;
    double *zv = eta + Q;
    double *csum = zv + Q;                                       // [OROW] when ZS
    double *wsc = csum + (T::ZS ? T::OROW : 0);                  // [4]

    for (int p = tid; p < T::P; p += 32 * T::WARPS) {
        int i, j;
        unpack_p(p, i, j);
        tab[p] = (uint32_t)(kb_boff(i >> 3, j >> 3) + kb_sw(i & 7, j & 7)) | ((uint32_t)i << 12) | ((uint32_t)j << 18);
    }
    if (T::ZS)
        for (int c = lane; c < T::OROW; c += 32) csum[c] = 0.0;
    __syncthreads();

    const int gid = lane >> 2, qd = lane & 3;
    const int oA0 = kb_sw(gid, qd), oA1 = kb_sw(gid, qd + 4);    // row-wise fragment   M[gid][qd + 4h]
    const int oT0 = kb_sw(qd, gid), oT1 = kb_sw(qd + 4, gid);    // transposed fragment M[qd + 4h][gid]
    const int oC = kb_sw(gid, 2 * qd);                           // accumulator pair    M[gid][2qd, 2qd+1]
    double s_qld = 0.0, s_ld = 0.0, s_n = 0.0;

    const long long nwarps = (long long)gridDim.x * T::WARPS;
    for (long long n = (long long)blockIdx.x * T::WARPS + warp; n < N; n += nwarps) {
        io_t *row = MZ + n * IO::PITCH;
        // the row after this one: pull it into L2 now, it is read ~50k cycles from now
        if (lane == 0 && n + nwarps < N)
            prefetch_l2(row + nwarps * IO::PITCH, (uint32_t)((IO::USED * sizeof(io_t) + 15) & ~15u));
        // ---- unpack [qprec packed | eta] into the block storage (coalesced global reads, UNR in flight)
        {
            const double e0 = (lane < Q) ? (double)row[IO::ZOFF + lane] : 0.0;
            const double e1 = (Q > 32) ? (double)row[IO::ZOFF + 32 + (lane & 31)] : 0.0;
#pragma unroll 1
            for (int base = 0; base < T::P; base += 32 * T::UNR) {
                double v[T::UNR];
#pragma unroll
                for (int u = 0; u < T::UNR; ++u) {
                    const int p = base + 32 * u + lane;
                    v[u] = (p < T::P) ? (double)row[IO::POFF + p] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < T::UNR; ++u) {
                    const int p = base + 32 * u + lane;
                    if (p < T::P) blk[tab[p] & 0xfff] = v[u];
                }
            }
            if (lane < Q) eta[lane] = e0;
            if (Q > 32) eta[32 + lane] = e1;
        }
        __syncwarp();

        // ---- 1. left-looking block Cholesky; diagonal blocks are replaced by their inverses X_jj
        double mant = 1.0;            // prod_k 1/l_kk = mant * 2^esum, renormalised after every diagonal block
        int esum = 0;
        bool ok = true;
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            double acc[NB][2];
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) acc[ib][0] = acc[ib][1] = 0.0;
#pragma unroll
            for (int kb = 0; kb < jb; ++kb) {
                const double *Bj = blk + kb_boff(jb, kb);
                const double b0 = Bj[oA0], b1 = Bj[oA1];
#pragma unroll
                for (int ib = jb; ib < NB; ++ib) {
                    const double *Ai = blk + kb_boff(ib, kb);
                    dmma884(acc[ib][0], acc[ib][1], Ai[oA0], b0);
                    dmma884(acc[ib][0], acc[ib][1], Ai[oA1], b1);
                }
            }
#pragma unroll
            for (int ib = jb; ib < NB; ++ib) {
                const double2 a = *reinterpret_cast<const double2 *>(blk + kb_boff(ib, jb) + oC);
                acc[ib][0] = a.x - acc[ib][0];
                acc[ib][1] = a.y - acc[ib][1];
            }
            double x0, x1;
            diag_chol_inv(acc[jb][0], acc[jb][1], x0, x1, mant);
            {
                ok = ok && (mant - mant == 0.0);                 // NaN / inf <=> a pivot was <= 0
                const long long bits = __double_as_longlong(mant);
                esum += (int)((bits >> 52) & 0x7ff) - 1023;
                mant = __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3ff0000000000000LL);
            }
            *reinterpret_cast<double2 *>(blk + kb_boff(jb, jb) + oC) = make_double2(x0, x1);
#pragma unroll
            for (int ib = jb + 1; ib < NB; ++ib)
                *reinterpret_cast<double2 *>(blk + kb_boff(ib, jb) + oC) = make_double2(acc[ib][0], acc[ib][1]);
            __syncwarp();
            if (jb + 1 < NB) {
                // L_ij = C_ij X_jj^T
                const double *Xj = blk + kb_boff(jb, jb);
                const double xb0 = Xj[oA0], xb1 = Xj[oA1];
#pragma unroll
                for (int ib = jb + 1; ib < NB; ++ib) {
                    const double *Ci = blk + kb_boff(ib, jb);
                    acc[ib][0] = acc[ib][1] = 0.0;
                    dmma884(acc[ib][0], acc[ib][1], Ci[oA0], xb0);
                    dmma884(acc[ib][0], acc[ib][1], Ci[oA1], xb1);
                }
                __syncwarp();
#pragma unroll
                for (int ib = jb + 1; ib < NB; ++ib)
                    *reinterpret_cast<double2 *>(blk + kb_boff(ib, jb) + oC) = make_double2(acc[ib][0], acc[ib][1]);
                __syncwarp();
            }
        }

        // ---- 2. X = L^-1, one block row at a time:  X_ik = -X_ii * sum_{k <= j < i} L_ij X_jk
#pragma unroll
        for (int ib = 1; ib < NB; ++ib) {
            double acc[NB][2];
#pragma unroll
            for (int kb = 0; kb < NB; ++kb) acc[kb][0] = acc[kb][1] = 0.0;
#pragma unroll
            for (int jb = 0; jb < ib; ++jb) {
                const double *Lij = blk + kb_boff(ib, jb);
                const double a0 = Lij[oA0], a1 = Lij[oA1];
#pragma unroll
                for (int kb = 0; kb <= jb; ++kb) {
                    const double *Xjk = blk + kb_boff(jb, kb);
                    dmma884(acc[kb][0], acc[kb][1], a0, Xjk[oT0]);
                    dmma884(acc[kb][0], acc[kb][1], a1, Xjk[oT1]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int kb = 0; kb < ib; ++kb)
                *reinterpret_cast<double2 *>(blk + kb_boff(ib, kb) + oC) = make_double2(acc[kb][0], acc[kb][1]);
            __syncwarp();
            const double *Xii = blk + kb_boff(ib, ib);
            const double xa0 = Xii[oA0], xa1 = Xii[oA1];
#pragma unroll
            for (int kb = 0; kb < ib; ++kb) {
                const double *S = blk + kb_boff(ib, kb);
                acc[kb][0] = acc[kb][1] = 0.0;
                dmma884(acc[kb][0], acc[kb][1], xa0, S[oT0]);
                dmma884(acc[kb][0], acc[kb][1], xa1, S[oT1]);
            }
            __syncwarp();
#pragma unroll
            for (int kb = 0; kb < ib; ++kb)
                *reinterpret_cast<double2 *>(blk + kb_boff(ib, kb) + oC) = make_double2(-acc[kb][0], -acc[kb][1]);
            __syncwarp();
        }

        // ---- 3. Sigma = X^T X (lower blocks, diagonal blocks come out full):  S_ij = sum_{k >= i} X_ki^T X_kj
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            double acc[NB][2];
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll
            for (int k = i; k < NB; ++k) {
                const double *Xki = blk + kb_boff(k, i);
                const double a0 = Xki[oT0], a1 = Xki[oT1];
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    const double *Xkj = blk + kb_boff(k, j);
                    dmma884(acc[j][0], acc[j][1], a0, Xkj[oT0]);
                    dmma884(acc[j][0], acc[j][1], a1, Xkj[oT1]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j <= i; ++j)
                *reinterpret_cast<double2 *>(blk + kb_boff(i, j) + oC) = make_double2(acc[j][0], acc[j][1]);
        }
        __syncwarp();

        // ---- zbar = Sigma eta on the tensor cores: B = eta_j broadcast over the 8 columns, so every accumulator
        //      column holds the block row's part of z
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            double z0 = 0.0, z1 = 0.0;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const double *B = blk + (j <= i ? kb_boff(i, j) : kb_boff(j, i));
                const double a0 = B[j <= i ? oA0 : oT0], a1 = B[j <= i ? oA1 : oT1];
                dmma884(z0, z1, a0, eta[j * 8 + qd]);
                dmma884(z0, z1, a1, eta[j * 8 + qd + 4]);
            }
            if (qd == 0) zv[i * 8 + gid] = z0;
        }
        __syncwarp();
        const double ldsum = ok ? -fma((double)esum, 0.69314718055994530942, log(mant)) : __longlong_as_double(0x7ff8000000000000LL);

        // ---- outputs: [<zz^T> packed | pad (left as it is: zeros) | zbar], optional Sigma, log-det
        double *sg = (Sig != nullptr) ? (Sig + n * T::P) : nullptr;
#pragma unroll 4
        for (int p = lane; p < T::P; p += 32) {
            const uint32_t t = tab[p];
            const double s = blk[t & 0xfff];
            const double m = fma(zv[(t >> 12) & 63], zv[t >> 18], s);
            row[IO::POFF + p] = (io_t)m;
            if (F32) store_split3(MP + n * IO::PITCH + IO::POFF + p, (size_t)N * IO::PITCH, (float)m);
            if (sg) sg[p] = s;
            if (T::ZS) csum[p] += m;
        }
        for (int c = lane; c < Q; c += 32) {
            const double z = zv[c];
            row[IO::ZOFF + c] = (io_t)z;
            if (F32) store_split3(MP + n * IO::PITCH + IO::ZOFF + c, (size_t)N * IO::PITCH, (float)z);
            if (T::ZS) csum[T::PP + c] += z;
        }
        if (lane == 0) {
            logdet[n] = ldsum;
            s_qld += 0.5 / ldsum;
            s_ld += ldsum;
            s_n += 1.0;
            if (!(ldsum - ldsum == 0.0)) atomicAdd(&gl[PYVB_GL_NONPD], 1.0);   // NaN / inf <=> a pivot was <= 0
        }
        __syncwarp();
    }

    if (!T::ZS || zsums == nullptr) return;
    if (lane == 0) {
        wsc[0] = s_qld;
        wsc[1] = s_ld;
        wsc[2] = s_n;
        wsc[3] = 0.0;
    }
    __syncthreads();
    double *out = zsums + (size_t)blockIdx.x * T::KW;
    const double *w0 = reinterpret_cast<const double *>(smem_kb + T::TAB_B) + T::NBLK * 64 + 2 * Q;   // csum of warp 0
    for (int c = tid; c < T::KW; c += 32 * T::WARPS) {
        double a = 0.0;
        for (int w = 0; w < T::WARPS; ++w) a += w0[(size_t)w * T::WARP_D + c];   // [csum OROW | scalars 4] is contiguous
        out[c] = a;
    }
}

template <int Q, bool F32>
cudaError_t launch_blocked_q(long long N, void *MZ, double *Sig, double *logdet, double *gl, double *zsums, void *MP,
                             cudaStream_t st) {
    using T = KB<Q>;
    cudaError_t e = cudaFuncSetAttribute(zsolve_blocked_kernel<Q, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)T::SMEM);
    if (e != cudaSuccess) return e;
    const int blocks = zsolve_blocked_blocks(N, Q);
    zsolve_blocked_kernel<Q, F32><<<blocks, 32 * T::WARPS, T::SMEM, st>>>(
        N, static_cast<typename KIO<Q, F32>::type *>(MZ), Sig, logdet, gl, T::ZS ? zsums : nullptr,
        static_cast<__nv_bfloat16 *>(MP));
    return cudaGetLastError();
}

}  // namespace

int zsolve_blocked_blocks(long long N, int q) {
    int warps = 8, occ = 2;
    switch (q) {
        case 8: warps = KB<8>::WARPS; occ = KB<8>::OCC; break;
        case 16: warps = KB<16>::WARPS; occ = KB<16>::OCC; break;
        case 32: warps = KB<32>::WARPS; occ = KB<32>::OCC; break;
        case 64: warps = KB<64>::WARPS; occ = KB<64>::OCC; break;
        default: return 0;
    }
    long long b = (N + warps - 1) / warps;
    if (b > 148LL * occ) b = 148LL * occ;
    if (b < 1) b = 1;
    return (int)b;
}

int zsolve_blocked_kw(int q) {
    switch (q) {
        case 8: return KB<8>::KW;
        case 16: return KB<16>::KW;
        case 32: return KB<32>::KW;
    }
    return 0;   // q = 64: no column-sum partials (the statistics pass sums the MZ rows itself)
}

cudaError_t launch_zsolve_blocked(long long N, int q, double *MZ, double *Sig, double *logdet, double *gl,
                                  double *zsums, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    switch (q) {
        case 8: return launch_blocked_q<8, false>(N, MZ, Sig, logdet, gl, zsums, nullptr, st);
        case 16: return launch_blocked_q<16, false>(N, MZ, Sig, logdet, gl, zsums, nullptr, st);
        case 32: return launch_blocked_q<32, false>(N, MZ, Sig, logdet, gl, zsums, nullptr, st);
        case 64: return launch_blocked_q<64, false>(N, MZ, Sig, logdet, gl, zsums, nullptr, st);
    }
    return cudaErrorNotSupported;
}

void zsolve_partials_f32(long long N, int q, int &nblk, int &kw) {
    kw = (q == 16 || q == 32) ? zsolve_blocked_kw(q) : 0;
    nblk = (kw > 0 && N > 0) ? zsolve_blocked_blocks(N, q) : 0;
}

// FP32 rows (see KIO): in place on MZ32, plus the bf16 x 3 split of the finished rows into MP [3][N][pitch]
cudaError_t launch_zsolve_f32(long long N, int q, float *MZ32, void *MP, double *Sig, double *logdet, double *gl,
                              double *zsums, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    switch (q) {
        case 16: return launch_blocked_q<16, true>(N, MZ32, Sig, logdet, gl, zsums, MP, st);
        case 32: return launch_blocked_q<32, true>(N, MZ32, Sig, logdet, gl, zsums, MP, st);
        case 64: return launch_blocked_q<64, true>(N, MZ32, Sig, logdet, gl, zsums, MP, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace pyvb
