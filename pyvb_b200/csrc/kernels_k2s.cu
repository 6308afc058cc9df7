// K2 as a BLOCKED symmetric sweep: scalar panel elimination + FP64 tensor-core trailing update, IN PLACE in the packed rows as
// they lie in HBM.
//
// Replaces cho_factor / cho_solve(., I) / dot(qcov, .) / q_ln_det of Gaussian.update (nodes/gaussian.py:117-123).
//
// The Gauss-Jordan kernel (kernels_k2g.cu) is FP64-ISSUE bound: q^3 DFMA per matrix, one issue slot per 32 of them, and half of
// them redundant (it cannot use the symmetry).  Here the matrix is swept 8 pivots (one tile column K) at a time with the
// symmetric sweep operator
//
//   M_KK <- -inv(M_KK),    M_IK <- M_IK inv(M_KK),    M_IJ <- M_IJ - M_IK inv(M_KK) M_KJ     (I, J != K),
//
// which keeps the state symmetric, so that only the lower triangle exists, every intermediate overwrites the slots of the block
// it replaces (element (i, j), i >= j, at i (i + 1) / 2 + j: the HBM layout; a reader of the upper half swaps the indices) and
// after q / 8 sweeps the row holds -Sigma:
//
//   panel   q x 8 column block K in ROW layout (lane = row): the scalar sweep of kernels_k2g.cu over its 8 pivots -- the tile
//           rows publish one entry each per pivot, every row adds t_i b_j over the 8 columns: 8 q^2 FMA per matrix instead of q^3.
//           The multipliers come from 1 x 1 pivots exactly as in the unblocked elimination: NO explicit inverse of a block is
//           ever applied (a recursive Schur-complement version with explicit 8 x 8 / 16 x 16 inverses was 10 - 400 x less
//           accurate than the Cholesky route at cond 1e4 - 1e6 in the lane-level restatement and was dropped);
//   update  M_IJ += (-T_I) old_J^T for the lower tiles outside row / column K: DMMA.8x8x4, 256 FMA per issue slot.  The A
//           fragment of T_I and the B fragment of old_J^T (= the A fragment of the old panel tile, read before the panel is
//           overwritten) are plain LDS.64 from the packed rows, the accumulators go back with STS.64.
//
// ln prod diag chol = 1/2 sum ln(pivots): the pivots are those of the unblocked elimination.  zbar = Sigma eta: lane = row,
// symmetric reads.  <zz^T> = Sigma + zbar zbar^T: tile by tile in the accumulator layout.  Column sums / maxima bounds /
// log-det scalars per CTA as in the other K2 kernels (partial layout of the blocked kernel).  oracle/sweep_oracle.py restates
// the kernel lane by lane on the packed row.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace pyvb {

namespace {

__host__ __device__ constexpr int s_tri(int i) { return i * (i + 1) / 2; }
__host__ __device__ constexpr int s_pitch(int q) {          // == pyvb_mz_pitch(q)
    int p = ((s_tri(q) + 7) & ~7) + q + 1;
    while ((p % 8) != 4) ++p;
    return p;
}

// MPW matrices per warp and stage (one bulk copy each way; 2 or 4), UM of them unrolled together in the tensor-core phases
template <int Q, int MPW_, int WARPS_, int STAGES_, int UM_> struct SB {
    static constexpr int MPW = MPW_, WARPS = WARPS_, STAGES = STAGES_, UM = UM_;
    static constexpr int NBT = Q / 8;                          // tiles per dimension
    static constexpr int P = s_tri(Q), PP = (P + 7) & ~7, OROW = PP + Q, PITCH = s_pitch(Q);
    static constexpr int KW = 2 * OROW + PYVB_ZS_EXTRA;       // [column sums OROW | 4 scalars | bounds on the column maxima OROW]
    static constexpr int STAGE_D = MPW * PITCH;               // the rows of a group as they lie in HBM
    static constexpr int BC_D = 2 * MPW * 8;                  // published pivot columns of the diagonal sweep, two parities
    static constexpr int NTR = (Q > 32) ? Q / 32 : 1;         // rows of the maxima a lane tracks
    static constexpr int WARP_D = STAGES * STAGE_D + BC_D + OROW + 4 + 2 * Q + 2;   // stages | .. | csum | scalars | maxima | mbarriers
    static constexpr size_t SMEM = (size_t)WARPS * WARP_D * 8;
    static_assert(MPW == 1 || MPW == 2 || MPW == 4, "the diagonal sweep spreads MPW x 8 tile rows over the lanes");
    static_assert((STAGE_D % 2) == 0 && (OROW % 2) == 0 && (WARP_D % 2) == 0, "16-byte alignment of the per-warp arrays");
    static_assert(STAGES == 1 || STAGES == 2, "one or two stages");
};

// 1 / d to the last bit or so: MUFU.RCP64H seed + two Newton steps (as kernels_k2g.cu)
__device__ __forceinline__ double s_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// ---- 8 x 8 pivot tiles K of the warp's MPW matrices: M_KK <- -inv(M_KK) in place, lane (m, r) = row r of matrix m's tile (with
// MPW < 4 the other lanes shadow a lane of the same row: same arithmetic, same stores).  The sweep of kernels_k2g.cu at q = 8:
// every lane publishes  b_i = a[i][k] s_i  (s_i = 1 before row i's pivot, -1 / d_i after it), a[i][j] += t_i b_j  with
// t_i = -a[i][k] / d_k;  the pivot row keeps t = 0 and gets a 1 in column k, its scaling by 1 / d_k is deferred to the end.
template <int K, typename T>
__device__ __forceinline__ void s_pivot_tile(double *stg, double *bc, int lane, double &pr, bool &pos) {
    constexpr int C0 = 8 * K;
    const int m = (lane >> 3) % T::MPW, r = lane & 7;
    double *st = stg + m * T::PITCH;
    const int i = C0 + r;
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (j <= r) ? st[s_tri(i) + C0 + j] : st[s_tri(C0 + j) + i];
    double sinv = 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double *b = bc + (k & 1) * (T::MPW * 8) + m * 8;
        b[r] = a[k] * sinv;
        __syncwarp();
        double B[8];
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
            const double2 v = *reinterpret_cast<const double2 *>(b + 2 * j2);
            B[2 * j2] = v.x;
            B[2 * j2 + 1] = v.y;
        }
        const double rc = s_rcp(B[k]);
        const bool piv = (r == k);
        double t = -a[k] * rc;
        t = piv ? 0.0 : t;
        sinv = piv ? -rc : sinv;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j != k) a[j] = fma(t, B[j], a[j]);
        a[k] = piv ? 1.0 : t;
    }
    if ((lane >> 3) < T::MPW) {
        pr *= -sinv;                                            // 1 / d_r
        pos = pos && (-sinv > 0.0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j <= r) st[s_tri(i) + C0 + j] = a[j] * sinv;        // row r of inv(M_KK) is a[] (-sinv): the state keeps MINUS the inverse
    __syncwarp();
}

// ---- sweep of tile column K of the warp's MPW matrices, everything in place.  In: the symmetric sweep state M (lower triangle).
// Out:   M_KK <- -inv(M_KK),   M_IK <- M_IK inv(M_KK),   M_IJ <- M_IJ - M_IK inv(M_KK) M_KJ      (I, J != K).
template <int K, typename T>
__device__ __forceinline__ void s_sweep_tile(double *stg, double *bc, int lane, double &pr, bool &pos) {
    constexpr int NBT = T::NBT, C0 = 8 * K;
    const int gid = lane >> 2, qd = lane & 3;
    s_pivot_tile<K, T>(stg, bc, lane, pr, pos);
#pragma unroll(T::UM)
    for (int m = 0; m < T::MPW; ++m) {
        double *st = stg + m * T::PITCH;
        // A fragments of the OLD panel tiles (J, K), J != K (== B fragments of their transposes)
        double of[NBT][2];
#pragma unroll
        for (int J = 0; J < NBT; ++J)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (J == K) continue;
                of[J][h] = (J > K) ? st[s_tri(8 * J + gid) + C0 + 4 * h + qd] : st[s_tri(C0 + 4 * h + qd) + 8 * J + gid];
            }
        // B fragments of -inv(M_KK) (symmetric: the stored lower half serves both)
        double pb[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int rr = C0 + 4 * h + qd, cc = C0 + gid;
            pb[h] = (rr >= cc) ? st[s_tri(rr) + cc] : st[s_tri(cc) + rr];
        }
        // -T_J = old_J (-inv(M_KK))
        double nt[NBT][2];
#pragma unroll
        for (int J = 0; J < NBT; ++J) {
            if (J == K) continue;
            nt[J][0] = nt[J][1] = 0.0;
            dmma884(nt[J][0], nt[J][1], of[J][0], pb[0]);
            dmma884(nt[J][0], nt[J][1], of[J][1], pb[1]);
        }
        __syncwarp();                                           // every lane has read the old panel
#pragma unroll
        for (int J = 0; J < NBT; ++J) {
            if (J == K) continue;
            if (J > K) {
                double *o = st + s_tri(8 * J + gid) + C0 + 2 * qd;
                o[0] = -nt[J][0];
                o[1] = -nt[J][1];
            } else {                                            // above the pivot tile: the slot of the transposed element
                st[s_tri(C0 + 2 * qd) + 8 * J + gid] = -nt[J][0];
                st[s_tri(C0 + 2 * qd + 1) + 8 * J + gid] = -nt[J][1];
            }
        }
        __syncwarp();
        // trailing update on the tensor cores: M_IJ += (-T_I) old_J^T for the lower tiles I >= J, I, J != K
        double ta[NBT][2];
#pragma unroll
        for (int I = 0; I < NBT; ++I)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (I == K) continue;
                ta[I][h] = -((I > K) ? st[s_tri(8 * I + gid) + C0 + 4 * h + qd] : st[s_tri(C0 + 4 * h + qd) + 8 * I + gid]);
            }
#pragma unroll
        for (int I = 0; I < NBT; ++I) {
            if (I == K) continue;
#pragma unroll
            for (int J = 0; J <= I; ++J) {
                if (J == K) continue;
                double *o = st + s_tri(8 * I + gid) + 8 * J + 2 * qd;
                double c0 = o[0], c1 = o[1];                    // (a diagonal tile: the upper half is junk that is never stored)
                dmma884(c0, c1, ta[I][0], of[J][0]);
                dmma884(c0, c1, ta[I][1], of[J][1]);
                if (I > J || 2 * qd <= gid) o[0] = c0;
                if (I > J || 2 * qd + 1 <= gid) o[1] = c1;
            }
        }
    }
    __syncwarp();
}

template <int K, typename T>
__device__ __forceinline__ void s_sweep_all(double *stg, double *bc, int lane, double &pr, bool &pos) {
    if constexpr (K < T::NBT) {
        s_sweep_tile<K, T>(stg, bc, lane, pr, pos);
        s_sweep_all<K + 1, T>(stg, bc, lane, pr, pos);
    }
}

template <int Q, int MPW_, int WARPS_, int STAGES_, int UM_>
__global__ void __launch_bounds__(32 * WARPS_, 1)
zsolve_sweep_kernel(long long N, double *__restrict__ MZ, double *__restrict__ Sig, double *__restrict__ logdet, double *gl,
                    double *__restrict__ zsums, const double *__restrict__ cond, const I8Check chk) {
    using T = SB<Q, MPW_, WARPS_, STAGES_, UM_>;
    constexpr int MPW = T::MPW, P = T::P, PP = T::PP, PITCH = T::PITCH, NBT = T::NBT, NTR = T::NTR;
    constexpr int NPASS = MPW * Q / 32;                          // passes of the lane = row loops over the warp's 4 Q rows
    if (cond != nullptr && !(*cond > 0.0)) return;               // conditional (fall-back) launch: nothing to redo
    __shared__ double s_chk[T::WARPS + 1];
    extern __shared__ __align__(16) double smem_sb[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, qd = lane & 3;
    double *stage = smem_sb + (size_t)warp * T::WARP_D;
    double *bc = stage + T::STAGES * T::STAGE_D;
    double *csum = bc + T::BC_D;                                  // [OROW]
    double *wsc = csum + T::OROW;                                 // [4]
    double *wmx = wsc + 4;                                        // [2 Q]: max_n <z_i z_i>, max_n |<z_i>| of this warp's rows
    uint64_t *bar = reinterpret_cast<uint64_t *>(wmx + 2 * Q);

    for (int c = lane; c < T::OROW; c += 32) csum[c] = 0.0;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    // INT8 guard (kernels.h: I8Check): a row whose largest diagonal entry is below `thr` carries too much fixed-point rounding
    if (chk.gscale != nullptr) {                                 // kernel-uniform
        double mx = 0.0;
        for (int c = tid; c < chk.ncols; c += 32 * T::WARPS) mx = fmax(mx, chk.gscale[c]);
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_chk[warp] = mx;
    }
    mbar_fence_init();
    __syncthreads();
    double thr = -1.0;
    if (chk.gscale != nullptr) {
        double mx = s_chk[0];
        for (int w = 1; w < T::WARPS; ++w) mx = fmax(mx, s_chk[w]);
        thr = gl[PYVB_GL_TAU] * chk.fac * mx;
    }
    if (cond != nullptr && blockIdx.x == 0 && tid == 0) gl[PYVB_GL_I8FALL] += 1.0;

    double dmx[NTR], zmx[NTR], s_qld = 0.0, s_ld = 0.0, s_n = 0.0;
#pragma unroll
    for (int s = 0; s < NTR; ++s) dmx[s] = zmx[s] = 0.0;

    const long long nwarps = (long long)gridDim.x * T::WARPS;
    const long long ngroups = (N + MPW - 1) / MPW;
    long long g = (long long)blockIdx.x * T::WARPS + warp;
    if (g < ngroups && lane == 0) {                               // the first group of this warp
        const long long left = N - g * MPW;
        const uint32_t bytes = (uint32_t)((left < MPW ? left : MPW) * PITCH * 8);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(stage, MZ + g * MPW * PITCH, bytes, bar);
    }
    for (int it = 0; g < ngroups; g += nwarps, ++it) {
        const int sidx = (T::STAGES == 2) ? (it & 1) : 0;
        double *stg = stage + sidx * T::STAGE_D;                  // this group's rows
        const long long n0 = g * MPW;
        const int nval = (N - n0 < MPW) ? (int)(N - n0) : MPW;
        const long long gn = g + nwarps;                         // the warp's next group
        if (T::STAGES == 2) {                                    // into the other stage now (its store has been read by then)
            if (lane == 0 && gn < ngroups) {
                bulk_wait_read_all();
                const long long left = N - gn * MPW;
                const uint32_t bytes = (uint32_t)((left < MPW ? left : MPW) * PITCH * 8);
                mbar_arrive_expect_tx(bar + (sidx ^ 1), bytes);
                bulk_g2s(stage + (sidx ^ 1) * T::STAGE_D, MZ + gn * MPW * PITCH, bytes, bar + (sidx ^ 1));
            }
            mbar_wait(bar + sidx, (uint32_t)((it >> 1) & 1));
        } else {
            mbar_wait(bar, (uint32_t)(it & 1));
        }

        {   // ---- the sweep; it leaves -Sigma in place.  Lane (m8, r8) = pivot-tile row r8 of matrix m8 (with MPW < 4: shadows)
            const int m8 = lane >> 3, r8 = lane & 7;
            if (thr >= 0.0) {                                    // INT8 guard: largest diagonal entry of qprec (kernel-uniform branch)
                const double *st = stg + (m8 % MPW) * PITCH;
                double dm = 0.0;
#pragma unroll
                for (int t = 0; t < NBT; ++t) dm = fmax(dm, st[s_tri(r8 + 8 * t) + r8 + 8 * t]);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
                if (r8 == 0 && m8 < nval && m8 < MPW && thr > dm) atomicAdd(&gl[PYVB_GL_I8BAD], 1.0);
            }
            // ln prod diag chol = -1/2 sum ln(1 / d_i); a pivot <= 0 (or NaN) poisons the logarithm
            double pr = 1.0;
            bool pos = true;
            s_sweep_all<0, T>(stg, bc, lane, pr, pos);
            double lg = pos ? log(pr) : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
            const double ldsum = -0.5 * lg;
            if (r8 == 0 && m8 < nval && m8 < MPW) {
                logdet[n0 + m8] = ldsum;
                s_qld += 0.5 / ldsum;
                s_ld += ldsum;
                s_n += 1.0;
                if (!(ldsum - ldsum == 0.0)) atomicAdd(&gl[PYVB_GL_NONPD], 1.0);   // NaN / inf <=> a pivot was <= 0
            }
        }
        if (Sig != nullptr) {                                    // kernel-uniform (the Sigma output is optional)
            for (int m = 0; m < nval; ++m)
                for (int c = lane; c < P; c += 32) Sig[(n0 + m) * P + c] = -stg[m * PITCH + c];
        }

        // ---- zbar = Sigma eta = -(state) eta: lane = row (pass p takes rows 32 p .. 32 p + 31 of the warp's MPW Q rows)
        double z[NPASS];
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
            const int idx = p * 32 + lane, m = idx / Q, i = idx % Q;
            const double *st = stg + m * PITCH;
            const double *rowp = st + s_tri(i), *colp = st + i;
            double acc = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < Q / 2; ++j2) {
                const double2 E = *reinterpret_cast<const double2 *>(st + PP + 2 * j2);
                const double v0 = (2 * j2 <= i) ? rowp[2 * j2] : colp[s_tri(2 * j2)];
                const double v1 = (2 * j2 + 1 <= i) ? rowp[2 * j2 + 1] : colp[s_tri(2 * j2 + 1)];
                acc = fma(v1, E.y, fma(v0, E.x, acc));
            }
            acc = -acc;
            z[p] = acc;
            if (m < nval) {
                const int s = (NTR > 1) ? (p % NTR) : 0;
                dmx[s] = fmax(dmx[s], fma(acc, acc, -rowp[i]));  // <z_i z_i> exactly as it is stored below
                zmx[s] = fmax(zmx[s], fabs(acc));
            }
        }
        __syncwarp();                                            // every lane has read eta
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
            const int idx = p * 32 + lane, m = idx / Q, i = idx % Q;
            stg[m * PITCH + PP + i] = z[p];                      // zbar takes eta's place
        }
        __syncwarp();
        // ---- <zz^T> = Sigma + zbar zbar^T, tile by tile in the accumulator layout, packed in place
#pragma unroll(T::UM)
        for (int m = 0; m < MPW; ++m) {
            double *st = stg + m * PITCH;
            double zi[NBT];
            double2 zj[NBT];
#pragma unroll
            for (int I = 0; I < NBT; ++I) {
                zi[I] = st[PP + 8 * I + gid];
                zj[I] = *reinterpret_cast<const double2 *>(st + PP + 8 * I + 2 * qd);
            }
#pragma unroll
            for (int I = 0; I < NBT; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J) {
                    double *o = st + s_tri(8 * I + gid) + 8 * J + 2 * qd;
                    if (I > J || 2 * qd <= gid) o[0] = fma(zi[I], zj[J].x, -o[0]);
                    if (I > J || 2 * qd + 1 <= gid) o[1] = fma(zi[I], zj[J].y, -o[1]);
                }
        }
        fence_async_smem();
        __syncwarp();
        // ---- rows back to HBM (one bulk store), column sums of the finished rows, next group in
        if (lane == 0) {
            bulk_s2g(MZ + n0 * PITCH, stg, (uint32_t)(nval * PITCH * 8));
            bulk_commit();
        }
        if (zsums != nullptr) {                                  // kernel-uniform
            for (int c = lane; c < T::OROW; c += 32) {
                double v = 0.0;
#pragma unroll
                for (int mm = 0; mm < MPW; ++mm)
                    if (mm < nval) v += stg[mm * PITCH + c];
                csum[c] += v;
            }
        }
        __syncwarp();
        if (T::STAGES == 1 && lane == 0 && gn < ngroups) {       // one stage: the store has to have read it first
            bulk_wait_read_all();
            const long long left = N - gn * MPW;
            const uint32_t bytes = (uint32_t)((left < MPW ? left : MPW) * PITCH * 8);
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(stage, MZ + gn * MPW * PITCH, bytes, bar);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

    if (zsums == nullptr) return;                                // kernel-uniform
    // ---- CTA partial: [column sums | sum 0.5/logdet | sum logdet | rows | 0 | bounds on the column maxima]
    s_qld = warp_sum(s_qld);
    s_ld = warp_sum(s_ld);
    s_n = warp_sum(s_n);
    if (lane == 0) {
        wsc[0] = s_qld;
        wsc[1] = s_ld;
        wsc[2] = s_n;
        wsc[3] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < NTR; ++s) {
#pragma unroll
        for (int o = Q; o < 32; o <<= 1) {                        // q < 32: several lanes track the same row
            dmx[s] = fmax(dmx[s], __shfl_xor_sync(0xffffffffu, dmx[s], o));
            zmx[s] = fmax(zmx[s], __shfl_xor_sync(0xffffffffu, zmx[s], o));
        }
        if (lane < Q) {
            wmx[s * 32 + lane] = dmx[s];
            wmx[Q + s * 32 + lane] = zmx[s];
        }
    }
    __syncthreads();
    double *out = zsums + (size_t)blockIdx.x * T::KW;
    const double *w0 = smem_sb + T::STAGES * T::STAGE_D + T::BC_D;    // csum of warp 0
    for (int c = tid; c < T::OROW + 4; c += 32 * T::WARPS) {
        double v = 0.0;
        for (int w = 0; w < T::WARPS; ++w) v += w0[(size_t)w * T::WARP_D + c];   // [csum OROW | scalars 4] is contiguous
        out[c] = v;
    }
    // CTA maxima of the diagonal second moments and of |z| (folded into warp 0's stage), then the column bounds:
    // |<z_i z_j>| <= sqrt(<z_i z_i> <z_j z_j>) because <zz^T> is PSD
    const double *m0 = w0 + T::OROW + 4;
    double *fm = smem_sb;                                        // warp 0's stage is free now: [2 Q] folded maxima
    for (int c = tid; c < 2 * Q; c += 32 * T::WARPS) {
        double v = 0.0;
        for (int w = 0; w < T::WARPS; ++w) v = fmax(v, m0[(size_t)w * T::WARP_D + c]);
        fm[c] = v;
    }
    __syncthreads();
    for (int c = tid; c < T::OROW; c += 32 * T::WARPS) {
        double v = 0.0;
        if (c < P) {
            int i, j;
            unpack_p(c, i, j);
            v = sqrt(fm[i] * fm[j]);
        } else if (c >= PP) {
            v = fm[Q + c - PP];
        }
        out[T::OROW + 4 + c] = v;
    }
}

template <int Q, int MPW, int WARPS, int STAGES, int UM>
cudaError_t launch_sweep_cfg(long long N, double *MZ, double *Sig, double *logdet, double *gl, double *zsums, cudaStream_t st,
                             const double *cond, I8Check chk) {
    using T = SB<Q, MPW, WARPS, STAGES, UM>;
    static_assert(T::SMEM <= 232448 - 1024, "shared memory of one CTA");
    cudaError_t e = cudaFuncSetAttribute(zsolve_sweep_kernel<Q, MPW, WARPS, STAGES, UM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e != cudaSuccess) return e;
    zsolve_sweep_kernel<Q, MPW, WARPS, STAGES, UM><<<zsolve_sweep_blocks(N, Q), 32 * WARPS, T::SMEM, st>>>(N, MZ, Sig, logdet, gl,
                                                                                                            zsums, cond, chk);
    return cudaGetLastError();
}

#define SB_CONFIGS(X) \
    X(16, 4, 16, 2, 4) X(16, 4, 16, 2, 1) X(16, 2, 16, 2, 2) \
    X(32, 4, 9, 1, 4) X(32, 4, 9, 1, 2) X(32, 4, 9, 1, 1) X(32, 2, 15, 1, 2) X(32, 2, 15, 1, 1) X(32, 2, 9, 2, 2) X(32, 4, 5, 2, 4) \
    X(32, 2, 12, 1, 2) X(32, 4, 8, 1, 4) \
    X(64, 1, 6, 1, 1) X(64, 1, 4, 2, 1) X(64, 2, 4, 1, 2) X(64, 2, 4, 1, 1)

// built configurations (matrices per warp and stage, warps per CTA, stages, matrices unrolled together); PYVB_SWEEP =
// "mpw,warps,stages,unroll" picks another built one (measurements)
void sweep_config(int q, int &mpw, int &warps, int &stages, int &um) {
    if (q == 16) mpw = 4, warps = 16, stages = 2, um = 4;
    else if (q == 32) mpw = 4, warps = 8, stages = 1, um = 4;
    else mpw = 1, warps = 6, stages = 1, um = 1;
    const char *e = getenv("PYVB_SWEEP");
    int a = 0, b = 0, c = 0, d = 0;
    if (e && sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) {
#define SB_HAVE(Q_, M_, W_, S_, U_) if (q == Q_ && a == M_ && b == W_ && c == S_ && d == U_) mpw = a, warps = b, stages = c, um = d;
        SB_CONFIGS(SB_HAVE)
#undef SB_HAVE
    }
}

}  // namespace

int zsolve_sweep_blocks(long long N, int q) {
    if (q != 16 && q != 32 && q != 64) return 0;
    int mpw, warps, stages, um;
    sweep_config(q, mpw, warps, stages, um);
    long long b = (N + (long long)mpw * warps - 1) / ((long long)mpw * warps);
    if (b > 148) b = 148;
    if (b < 1) b = 1;
    return (int)b;
}

int zsolve_sweep_kw(int q) {
    return q == 16 ? SB<16, 4, 16, 2, 4>::KW : q == 32 ? SB<32, 4, 8, 1, 4>::KW : q == 64 ? SB<64, 1, 6, 1, 1>::KW : 0;
}

cudaError_t launch_zsolve_sweep(long long N, int q, double *MZ, double *Sig, double *logdet, double *gl, double *zsums,
                                cudaStream_t st, const double *cond, I8Check chk) {
    if (N <= 0) return cudaSuccess;
    int mpw, warps, stages, um;
    sweep_config(q, mpw, warps, stages, um);
#define SB_CASE(Q_, M_, W_, S_, U_) \
    if (q == Q_ && mpw == M_ && warps == W_ && stages == S_ && um == U_) \
        return launch_sweep_cfg<Q_, M_, W_, S_, U_>(N, MZ, Sig, logdet, gl, zsums, st, cond, chk);
    SB_CONFIGS(SB_CASE)
#undef SB_CASE
    return cudaErrorNotSupported;
}

}  // namespace pyvb
