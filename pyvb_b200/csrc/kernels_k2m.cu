// K2, blocked with lane-parallel diagonal blocks (q = 32, 64): batched q x q SPD inverse / solve, MPW matrices per warp.
//
// Replaces, per row n, the reference's  cho_factor(qprec) / cho_solve(., I) / dot(qcov, .)  and
// q_ln_det = .5/log(prod(diag(chol)))  (nodes/gaussian.py:117-123).
//
// Same block algorithm and shared-memory storage as the one-matrix-per-warp kernel (kernels_k2.cu: NB x NB lower
// triangle of swizzled 8 x 8 blocks, potrf / trtri / lauum sweeps on DMMA.8x8x4), with ONE difference that removes most
// of its instructions: the 8 x 8 DIAGONAL blocks.  Factoring and inverting an 8 x 8 block is a chain of 8 dependent
// pivots; done warp-cooperatively (shuffles, kernels_k2.cu) it costs ~320 warp instructions per block -- 2/3 of all the
// instructions of a q = 32 matrix -- with every lane holding two entries.  Here a warp owns MPW matrices at a time and
// LANE m factors and inverts the diagonal block of matrix m on its own, in registers (the straight-line code of the
// thread-per-matrix kernel, chol8.cuh: chol_inv8): the same ~320 instructions now serve MPW matrices.  The off-diagonal
// work stays on the tensor cores, one matrix after the other (two interleaved: independent accumulator chains).
#include <stdlib.h>

#include "chol8.cuh"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace pyvb {

namespace {

__host__ __device__ constexpr int km_tri(int i) { return i * (i + 1) / 2; }
__host__ __device__ constexpr int km_pitch(int q) {          // == pyvb_gw_pitch(q)
    int p = ((km_tri(q) + 7) & ~7) + q + 1;
    while ((p % 8) != 4) ++p;
    return p;
}
// element (r, c) of a swizzled 8 x 8 block; block (i, j), i >= j  (the layout of kernels_k2.cu)
__host__ __device__ constexpr int km_sw(int r, int c) { return r * 8 + ((((c >> 2) ^ ((r >> 1) & 1))) << 2) + (c & 3); }
__host__ __device__ constexpr int km_boff(int i, int j) { return (km_tri(i) + j) * 64; }

template <int Q> struct KMC;
// MI: matrices whose tensor-core instruction streams are interleaved (independent accumulator chains hide the DMMA and
// shared-memory latencies: with 8 warps per SM the kernel is latency-bound, not issue-bound)
template <> struct KMC<16> { static constexpr int WARPS = 8, MPW = 8, MI = 4; static constexpr bool ZS = true; };
template <> struct KMC<32> { static constexpr int WARPS = 8, MPW = 4, MI = 4; static constexpr bool ZS = true; };
template <> struct KMC<64> { static constexpr int WARPS = 5, MPW = 2, MI = 2; static constexpr bool ZS = false; };

template <int Q> struct KM {
    static constexpr int NB = Q / 8, NBLK = km_tri(NB);
    static constexpr int P = km_tri(Q), PP = (P + 7) & ~7, OROW = PP + Q, PITCH = km_pitch(Q);
    static constexpr int WARPS = KMC<Q>::WARPS, MPW = KMC<Q>::MPW;
    static constexpr bool ZS = KMC<Q>::ZS;
    static constexpr int KW = 2 * OROW + PYVB_ZS_EXTRA;   // [column sums OROW | 4 scalars | bounds on the column maxima OROW]
    // per matrix: blocks | eta | z, padded so that the MPW matrices of a warp start 2 doubles (4 banks) apart modulo the
    // 32 banks: the per-lane accesses of the diagonal phase (lane m -> matrix m, same element) are conflict-free
    static constexpr int MAT_D = NBLK * 64 + 2 * Q + 2;
    static constexpr int WARP_D = MPW * MAT_D + (ZS ? OROW : 0) + 4 + (ZS ? 2 * Q : 0) + 2;
    static constexpr int TAB_B = ((P * 4) + 15) & ~15;
    static constexpr int UNR = (P / 32 >= 16) ? 16 : (P + 31) / 32;
    static constexpr size_t SMEM = (size_t)TAB_B + (size_t)WARPS * WARP_D * 8;
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int Q>
__global__ void __launch_bounds__(32 * KM<Q>::WARPS, 1)
zsolve_lanediag_kernel(long long N, double *__restrict__ MZ, double *__restrict__ Sig, double *__restrict__ logdet,
                       double *gl, double *__restrict__ zsums, const double *__restrict__ cond, const I8Check chk) {
    using T = KM<Q>;
    if (cond != nullptr && !(*cond > 0.0)) return;               // conditional (fall-back) launch: nothing to redo
    __shared__ double s_chk[T::WARPS + 1];
    constexpr int NB = T::NB, MPW = T::MPW, MI = KMC<Q>::MI;      // MI matrices interleaved in the tensor-core phases
    static_assert(MPW % MI == 0, "MPW");
    extern __shared__ __align__(16) unsigned char smem_km[];
    // packed index p -> offset in the block storage (bits 0-11) | i (bits 12-17) | j (bits 18-23)
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_km);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *wbase = reinterpret_cast<double *>(smem_km + T::TAB_B) + (size_t)warp * T::WARP_D;
    double *csum = wbase + MPW * T::MAT_D;                       // [OROW] when ZS
    double *wsc = csum + (T::ZS ? T::OROW : 0);                  // [4]
    double *wmx = wsc + 4;                                       // [2 Q] when ZS: max_n <z_i z_i>, max_n |<z_i>| of this warp's rows
    double dmx[(Q + 31) / 32], zmx[(Q + 31) / 32];
#pragma unroll
    for (int k = 0; k < (Q + 31) / 32; ++k) dmx[k] = zmx[k] = 0.0;

    for (int p = tid; p < T::P; p += 32 * T::WARPS) {
        int i, j;
        unpack_p(p, i, j);
        tab[p] = (uint32_t)(km_boff(i >> 3, j >> 3) + km_sw(i & 7, j & 7)) | ((uint32_t)i << 12) | ((uint32_t)j << 18);
    }
    if (T::ZS)
        for (int c = lane; c < T::OROW; c += 32) csum[c] = 0.0;
    // INT8 guard (kernels.h: I8Check): a row whose largest diagonal entry is below `thr` carries too much fixed-point rounding
    if (chk.gscale != nullptr) {                                 // kernel-uniform
        double m = 0.0;
        for (int c = tid; c < chk.ncols; c += 32 * T::WARPS) m = fmax(m, chk.gscale[c]);
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_chk[warp] = m;
    }
    __syncthreads();
    double thr = -1.0;
    if (chk.gscale != nullptr) {
        double m = s_chk[0];
        for (int w = 1; w < T::WARPS; ++w) m = fmax(m, s_chk[w]);
        thr = gl[PYVB_GL_TAU] * chk.fac * m;
    }
    if (cond != nullptr && blockIdx.x == 0 && tid == 0) gl[PYVB_GL_I8FALL] += 1.0;

    const int gid = lane >> 2, qd = lane & 3;
    const int oA0 = km_sw(gid, qd), oA1 = km_sw(gid, qd + 4);    // row-wise fragment   M[gid][qd + 4h]
    const int oT0 = km_sw(qd, gid), oT1 = km_sw(qd + 4, gid);    // transposed fragment M[qd + 4h][gid]
    const int oC = km_sw(gid, 2 * qd);                           // accumulator pair    M[gid][2qd, 2qd+1]
    double s_qld = 0.0, s_ld = 0.0, s_n = 0.0;
    double *myblk = wbase + (lane < MPW ? lane : 0) * T::MAT_D;  // the matrix whose diagonal blocks this lane factors

    const long long nwarps = (long long)gridDim.x * T::WARPS;
    const long long ngroups = (N + MPW - 1) / MPW;
    for (long long g = (long long)blockIdx.x * T::WARPS + warp; g < ngroups; g += nwarps) {
        const long long n0 = g * MPW;
        if (lane == 0 && (g + nwarps) * MPW + MPW <= N)
            prefetch_l2(MZ + (g + nwarps) * MPW * T::PITCH, (uint32_t)(MPW * T::PITCH * sizeof(double)));
        // ---- unpack [qprec packed | eta] of the MPW rows into the block storage (coalesced global reads, UNR in flight)
#pragma unroll 1
        for (int m = 0; m < MPW; ++m) {
            const long long n = (n0 + m < N) ? n0 + m : N - 1;   // tail: redo the last row, store nothing
            const double *row = MZ + n * T::PITCH;
            double *blk = wbase + m * T::MAT_D, *eta = blk + T::NBLK * 64;
            const double e0 = (lane < Q) ? row[T::PP + lane] : 0.0;
            const double e1 = (Q > 32) ? row[T::PP + 32 + lane] : 0.0;
#pragma unroll 1
            for (int base = 0; base < T::P; base += 32 * T::UNR) {
                double v[T::UNR];
#pragma unroll
                for (int u = 0; u < T::UNR; ++u) {
                    const int p = base + 32 * u + lane;
                    v[u] = (p < T::P) ? row[p] : 0.0;
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
        if (thr >= 0.0) {                                        // diagonal of qprec, before the factorisation overwrites it
#pragma unroll 1
            for (int m = 0; m < MPW; ++m) {
                const double *blk = wbase + m * T::MAT_D;
                double dm = 0.0;
#pragma unroll
                for (int i = lane; i < Q; i += 32) dm = fmax(dm, blk[km_boff(i >> 3, i >> 3) + km_sw(i & 7, i & 7)]);
                for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
                if (lane == 0 && n0 + m < N && thr > dm) atomicAdd(&gl[PYVB_GL_I8BAD], 1.0);
            }
        }

        // ---- 1. left-looking block Cholesky; diagonal blocks are replaced by their inverses X_jj
        double ldsum = 0.0;                                      // lane m < MPW: ln prod diag chol of matrix m
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            // C_ij = A_ij - sum_k L_ik L_jk^T for the block column jb (tensor cores), MI matrices at a time
#pragma unroll 1
            for (int m0 = 0; m0 < MPW; m0 += MI) {
                double acc[MI][NB][2];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                    for (int ib = 0; ib < NB; ++ib) acc[mi][ib][0] = acc[mi][ib][1] = 0.0;
#pragma unroll
                    for (int kb = 0; kb < jb; ++kb) {
                        const double *Bj = blk + km_boff(jb, kb);
                        const double b0 = Bj[oA0], b1 = Bj[oA1];
#pragma unroll
                        for (int ib = jb; ib < NB; ++ib) {
                            const double *Ai = blk + km_boff(ib, kb);
                            dmma884(acc[mi][ib][0], acc[mi][ib][1], Ai[oA0], b0);
                            dmma884(acc[mi][ib][0], acc[mi][ib][1], Ai[oA1], b1);
                        }
                    }
                }
                if (jb > 0) {
                    __syncwarp();
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) {
                        double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                        for (int ib = jb; ib < NB; ++ib) {
                            double2 *dst = reinterpret_cast<double2 *>(blk + km_boff(ib, jb) + oC);
                            const double2 a = *dst;
                            *dst = make_double2(a.x - acc[mi][ib][0], a.y - acc[mi][ib][1]);
                        }
                    }
                }
            }
            __syncwarp();
            // the diagonal block of matrix m, factored and inverted by lane m in registers
            if (lane < MPW) {
                double *D = myblk + km_boff(jb, jb);
                double A[36];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) A[c8_idx(i, j)] = D[km_sw(i, j)];
                double lp = 1.0;
                chol_inv8(A, lp);
                ldsum -= log(lp);                                // NaN / inf <=> a pivot was <= 0
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) D[km_sw(i, j)] = (j <= i) ? A[c8_idx(i, j)] : 0.0;
            }
            __syncwarp();
            if (jb + 1 < NB) {
                // L_ij = C_ij X_jj^T
#pragma unroll 1
                for (int m0 = 0; m0 < MPW; m0 += MI) {
                    double acc[MI][NB][2];
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) {
                        const double *blk = wbase + (m0 + mi) * T::MAT_D;
                        const double *Xj = blk + km_boff(jb, jb);
                        const double xb0 = Xj[oA0], xb1 = Xj[oA1];
#pragma unroll
                        for (int ib = jb + 1; ib < NB; ++ib) {
                            const double *Ci = blk + km_boff(ib, jb);
                            acc[mi][ib][0] = acc[mi][ib][1] = 0.0;
                            dmma884(acc[mi][ib][0], acc[mi][ib][1], Ci[oA0], xb0);
                            dmma884(acc[mi][ib][0], acc[mi][ib][1], Ci[oA1], xb1);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) {
                        double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                        for (int ib = jb + 1; ib < NB; ++ib)
                            *reinterpret_cast<double2 *>(blk + km_boff(ib, jb) + oC) = make_double2(acc[mi][ib][0], acc[mi][ib][1]);
                    }
                }
                __syncwarp();
            }
        }

        // ---- 2. X = L^-1, one block row at a time:  X_ik = -X_ii * sum_{k <= j < i} L_ij X_jk
#pragma unroll 1
        for (int m0 = 0; m0 < MPW; m0 += MI) {
#pragma unroll
            for (int ib = 1; ib < NB; ++ib) {
                double acc[MI][NB][2];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                    for (int kb = 0; kb < NB; ++kb) acc[mi][kb][0] = acc[mi][kb][1] = 0.0;
#pragma unroll
                    for (int jb = 0; jb < ib; ++jb) {
                        const double *Lij = blk + km_boff(ib, jb);
                        const double a0 = Lij[oA0], a1 = Lij[oA1];
#pragma unroll
                        for (int kb = 0; kb <= jb; ++kb) {
                            const double *Xjk = blk + km_boff(jb, kb);
                            dmma884(acc[mi][kb][0], acc[mi][kb][1], a0, Xjk[oT0]);
                            dmma884(acc[mi][kb][0], acc[mi][kb][1], a1, Xjk[oT1]);
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                    for (int kb = 0; kb < ib; ++kb)
                        *reinterpret_cast<double2 *>(blk + km_boff(ib, kb) + oC) = make_double2(acc[mi][kb][0], acc[mi][kb][1]);
                }
                __syncwarp();
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const double *blk = wbase + (m0 + mi) * T::MAT_D;
                    const double *Xii = blk + km_boff(ib, ib);
                    const double xa0 = Xii[oA0], xa1 = Xii[oA1];
#pragma unroll
                    for (int kb = 0; kb < ib; ++kb) {
                        const double *S = blk + km_boff(ib, kb);
                        acc[mi][kb][0] = acc[mi][kb][1] = 0.0;
                        dmma884(acc[mi][kb][0], acc[mi][kb][1], xa0, S[oT0]);
                        dmma884(acc[mi][kb][0], acc[mi][kb][1], xa1, S[oT1]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                    for (int kb = 0; kb < ib; ++kb)
                        *reinterpret_cast<double2 *>(blk + km_boff(ib, kb) + oC) = make_double2(-acc[mi][kb][0], -acc[mi][kb][1]);
                }
                __syncwarp();
            }

            // ---- 3. Sigma = X^T X (lower blocks, diagonal blocks come out full):  S_ij = sum_{k >= i} X_ki^T X_kj
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                double acc[MI][NB][2];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                    for (int j = 0; j < NB; ++j) acc[mi][j][0] = acc[mi][j][1] = 0.0;
#pragma unroll
                    for (int k = i; k < NB; ++k) {
                        const double *Xki = blk + km_boff(k, i);
                        const double a0 = Xki[oT0], a1 = Xki[oT1];
#pragma unroll
                        for (int j = 0; j <= i; ++j) {
                            const double *Xkj = blk + km_boff(k, j);
                            dmma884(acc[mi][j][0], acc[mi][j][1], a0, Xkj[oT0]);
                            dmma884(acc[mi][j][0], acc[mi][j][1], a1, Xkj[oT1]);
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    double *blk = wbase + (m0 + mi) * T::MAT_D;
#pragma unroll
                    for (int j = 0; j <= i; ++j)
                        *reinterpret_cast<double2 *>(blk + km_boff(i, j) + oC) = make_double2(acc[mi][j][0], acc[mi][j][1]);
                }
            }
            __syncwarp();

            // ---- zbar = Sigma eta on the tensor cores: B = eta_j broadcast over the 8 columns, so every accumulator
            //      column holds the block row's part of z
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                double *blk = wbase + (m0 + mi) * T::MAT_D;
                const double *eta = blk + T::NBLK * 64;
                double *zv = blk + T::NBLK * 64 + Q;
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    double z0 = 0.0, z1 = 0.0;
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const double *B = blk + (j <= i ? km_boff(i, j) : km_boff(j, i));
                        const double a0 = B[j <= i ? oA0 : oT0], a1 = B[j <= i ? oA1 : oT1];
                        dmma884(z0, z1, a0, eta[j * 8 + qd]);
                        dmma884(z0, z1, a1, eta[j * 8 + qd + 4]);
                    }
                    if (qd == 0) zv[i * 8 + gid] = z0;
                }
            }
        }
        __syncwarp();

        // ---- outputs: [<zz^T> packed | pad (left as it is: zeros) | zbar], optional Sigma, log-det
#pragma unroll 1
        for (int m = 0; m < MPW; ++m) {
            const double ld = __shfl_sync(0xffffffffu, ldsum, m);
            const long long n = n0 + m;
            if (n >= N) break;                                   // warp-uniform
            double *blk = wbase + m * T::MAT_D;
            const double *zv = blk + T::NBLK * 64 + Q;
            double *row = MZ + n * T::PITCH;
            double *sg = (Sig != nullptr) ? (Sig + n * T::P) : nullptr;
#pragma unroll 4
            for (int p = lane; p < T::P; p += 32) {
                const uint32_t t = tab[p];
                const double s = blk[t & 0xfff];
                const double mm = fma(zv[(t >> 12) & 63], zv[t >> 18], s);
                row[p] = mm;
                if (sg) sg[p] = s;
                if (T::ZS) csum[p] += mm;
            }
#pragma unroll
            for (int k = 0; k < (Q + 31) / 32; ++k) {
                const int c = lane + 32 * k;
                if (c >= Q) continue;
                const double z = zv[c];
                row[T::PP + c] = z;
                if (T::ZS) {
                    csum[T::PP + c] += z;
                    // <z_c z_c> = Sigma_cc + z_c^2: with the diagonal maxima, |<z_i z_j>| <= sqrt(<z_i z_i> <z_j z_j>) bounds
                    // every column of the packed rows (the fixed-point scales of the INT8 statistics)
                    dmx[k] = fmax(dmx[k], fma(z, z, blk[km_boff(c >> 3, c >> 3) + km_sw(c & 7, c & 7)]));
                    zmx[k] = fmax(zmx[k], fabs(z));
                }
            }
            if (lane == 0) {
                logdet[n] = ld;
                s_qld += 0.5 / ld;
                s_ld += ld;
                s_n += 1.0;
                if (!(ld - ld == 0.0)) atomicAdd(&gl[PYVB_GL_NONPD], 1.0);   // NaN / inf <=> a pivot was <= 0
            }
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
#pragma unroll
    for (int k = 0; k < (Q + 31) / 32; ++k)
        if (lane + 32 * k < Q) {
            wmx[lane + 32 * k] = dmx[k];
            wmx[Q + lane + 32 * k] = zmx[k];
        }
    __syncthreads();
    double *out = zsums + (size_t)blockIdx.x * T::KW;
    const double *w0 = reinterpret_cast<const double *>(smem_km + T::TAB_B) + MPW * T::MAT_D;   // csum of warp 0
    for (int c = tid; c < T::OROW + 4; c += 32 * T::WARPS) {
        double a = 0.0;
        for (int w = 0; w < T::WARPS; ++w) a += w0[(size_t)w * T::WARP_D + c];   // [csum OROW | scalars 4] is contiguous
        out[c] = a;
    }
    // CTA maxima of the diagonal second moments and of |z| (fold the warps into warp 0's slots), then the column bounds
    double *m0 = const_cast<double *>(w0) + T::OROW + 4;
    __syncthreads();
    for (int c = tid; c < 2 * Q; c += 32 * T::WARPS) {
        double a = 0.0;
        for (int w = 0; w < T::WARPS; ++w) a = fmax(a, m0[(size_t)w * T::WARP_D + c]);
        m0[c] = a;
    }
    __syncthreads();
    for (int c = tid; c < T::OROW; c += 32 * T::WARPS) {
        double a = 0.0;
        if (c < T::P) {
            const uint32_t t = tab[c];
            a = sqrt(m0[(t >> 12) & 63] * m0[t >> 18]);
        } else if (c >= T::PP) {
            a = m0[Q + c - T::PP];
        }
        out[T::OROW + 4 + c] = a;
    }
}

template <int Q>
cudaError_t launch_lanediag_q(long long N, double *MZ, double *Sig, double *logdet, double *gl, double *zsums,
                              cudaStream_t st, const double *cond, I8Check chk) {
    using T = KM<Q>;
    cudaError_t e = cudaFuncSetAttribute(zsolve_lanediag_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e != cudaSuccess) return e;
    const int blocks = zsolve_lanediag_blocks(N, Q);
    zsolve_lanediag_kernel<Q><<<blocks, 32 * T::WARPS, T::SMEM, st>>>(N, MZ, Sig, logdet, gl, T::ZS ? zsums : nullptr, cond, chk);
    return cudaGetLastError();
}

}  // namespace

int zsolve_lanediag_blocks(long long N, int q) {
    int warps, mpw;
    switch (q) {
        case 16: warps = KM<16>::WARPS; mpw = KM<16>::MPW; break;
        case 32: warps = KM<32>::WARPS; mpw = KM<32>::MPW; break;
        case 64: warps = KM<64>::WARPS; mpw = KM<64>::MPW; break;
        default: return 0;
    }
    long long b = (N + (long long)mpw * warps - 1) / ((long long)mpw * warps);
    if (b > 148) b = 148;
    if (b < 1) b = 1;
    return (int)b;
}

int zsolve_lanediag_kw(int q) {
    switch (q) {
        case 16: return KM<16>::KW;
        case 32: return KM<32>::KW;
    }
    return 0;   // q = 64: no column-sum partials (the statistics pass sums the MZ rows itself)
}

cudaError_t launch_zsolve_lanediag(long long N, int q, double *MZ, double *Sig, double *logdet, double *gl, double *zsums,
                                   cudaStream_t st, const double *cond, I8Check chk) {
    if (N <= 0) return cudaSuccess;
    switch (q) {
        case 16: return launch_lanediag_q<16>(N, MZ, Sig, logdet, gl, zsums, st, cond, chk);
        case 32: return launch_lanediag_q<32>(N, MZ, Sig, logdet, gl, zsums, st, cond, chk);
        case 64: return launch_lanediag_q<64>(N, MZ, Sig, logdet, gl, zsums, st, cond, chk);
    }
    return cudaErrorNotSupported;
}

}  // namespace pyvb
