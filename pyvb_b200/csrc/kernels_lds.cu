// VB smoother of a linear dynamic system, batched over independent sequences (BASELINE config 5): one WARP per
// sequence, the whole sequence resident in shared memory, whole VB iterations inside one launch.
//
// Reference: examples/Linear_Dynamic_System.py:47-76 -- per iteration all X_t forwards, all X_t backwards, the A
// columns, the C columns, Q, R -- i.e. per node
//   X_t   Gaussian.update (nodes/gaussian.py:102-123) with Multiplication.pass_up_m1_m2 (nodes/node.py:203-227):
//         prec_t = Qbar (I at t = 0) + <A^T Qbar A> [t < T-1] + <C^T Rbar C>,
//         mean_t = prec_t^-1 (Qbar A x_{t-1} + A^T Qbar x_{t+1} + C^T Rbar y_t)
//   A_i, C_i   hstack.pass_up_m1_m2 (nodes/nodes_todo.py:43-62), Gauss-Seidel over the columns
//   Q, R  DiagonalGamma.update (nodes/nodes_todo.py:187-190)
// The precisions do not depend on the states: three distinct q x q matrices per sequence and iteration (t = 0,
// interior, t = T-1).  They are factored and inverted with the warp-cooperative 8 x 8 Cholesky of the batched solve
// (chol8.cuh), the smoother gain K_s = Sigma_s [Qbar A | A^T Qbar | C^T Rbar] (8 x 24) is spread over the lanes
// (lane = row i, quarter p: 6 coefficients), and one Gauss-Seidel step is 6 FMAs + two shuffle reductions per lane.
// State and observation dimensions are padded to 8 inside the kernel (q, d <= 8).
//
// The sweep over t (SCAN = true, the default).  A forward sweep is the recurrence  x_t = K1 x_{t-1} + u_t  with
// u_t = K2 x_{t+1} + K3 y_t  made of values the sweep does not change, and K1 the same matrix for every interior t (the
// backward sweep: x_t = K2 x_{t+1} + u'_t,  u'_t = K1 x_{t-1} + K3 y_t).  Done one step after the other -- SCAN = false, the
// first version, kept as the cross-check -- it is a chain of 2T dependent steps of ~180 clocks each (6 % of the HBM roof).  Here:
//   A. u_t for all t in parallel, four time steps per instruction (lane = (t mod 4, row)), written over x_t in place;
//   B. the interior range is cut into <= 32 chunks of L steps, lane = chunk, K1 in the lane's registers:
//      pass 1 runs every chunk from a zero start (its end value e_c), the chunk starts follow from
//      E_c = e_c + K1^L E_{c-1}  (K1^L by binary powering, a short serial loop over the chunks), pass 2 reruns every chunk
//      from its true start and stores x_t.
// Same arithmetic as the step-by-step sweep up to the rounding of the chunk starts (1e-16 relative; the recurrence is a
// contraction), 2 x 8 matrix-vector steps deep instead of 2T.
#include <stdlib.h>

#include "chol8.cuh"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace pyvb {

namespace {

template <bool SCAN> struct LWC { static constexpr int LW = SCAN ? 1 : 4; };    // warps (sequences) per CTA (SCAN: as many one-warp CTAs as fit)

// per-warp shared memory (doubles): xs [(T + 2)][8] (one zero row before and after), ys [T][d] (+ 8: the padded
// columns of the last row are read with zero coefficients), then the small arrays; SCAN: + the interior gain [8][24] and the
// chunk ends [32][8]
__host__ __device__ inline size_t lds_warp_doubles(int T, int d, bool scan) {
    return (size_t)(T + 2) * 8 + (((size_t)T * d + 8 + 1) & ~(size_t)1) + 64 * 14 + (scan ? 192 + 256 : 0);
}

// dst = A B (8 x 8, row-major, shared memory; dst distinct from A and B); lane (gid, qd) owns dst[gid][2qd], [2qd + 1]
__device__ __forceinline__ void mat8_mul(double *dst, const double *A, const double *B, int gid, int qd) {
    double r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double a = A[gid * 8 + k];
        const double2 b = *reinterpret_cast<const double2 *>(B + k * 8 + 2 * qd);
        r0 = fma(a, b.x, r0);
        r1 = fma(a, b.y, r1);
    }
    __syncwarp();
    *reinterpret_cast<double2 *>(dst + gid * 8 + 2 * qd) = make_double2(r0, r1);
    __syncwarp();
}

template <bool SCAN>
__global__ void __launch_bounds__(32 * LWC<SCAN>::LW)
lds_iterate_kernel(int B, int T, int q, int d, const double *__restrict__ Y, double *__restrict__ X,
                   double *__restrict__ Xcov3, double *__restrict__ A, double *__restrict__ Avar, double *__restrict__ C,
                   double *__restrict__ Cvar, double *__restrict__ Qa, double *__restrict__ Qb, double *__restrict__ Ra,
                   double *__restrict__ Rb, double alpha0, double a0, double b0, int niters, double *status,
                   const double *__restrict__ Aknown) {
    extern __shared__ __align__(16) double smem_l[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, qd = lane & 3;
    constexpr int LW = LWC<SCAN>::LW;
    double *base = smem_l + (size_t)warp * lds_warp_doubles(T, d, SCAN);
    double *xs = base + 8;                                  // xs[t * 8 + i], t = -1 .. T valid (zero rows at the ends)
    double *ys = base + (size_t)(T + 2) * 8;
    double *pA = ys + (((size_t)T * d + 8 + 1) & ~(size_t)1);   // [k][i]
    double *pAv = pA + 64, *pC = pAv + 64, *pCv = pC + 64;
    double *Sg = pCv + 64;                                  // [3][64]: Sigma_0, Sigma_interior, Sigma_{T-1}
    double *tmp = Sg + 192;                                 // [64]
    double *sSA = tmp + 64, *sSC = sSA + 64, *sXX1 = sSC + 64, *sYX = sXX1 + 64;
    double *qbar = sYX + 64, *rbar = qbar + 8, *yy = rbar + 8, *xx0 = yy + 8;   // [8] each
    double *pK = xx0 + 8;                                   // [k][i]: known entries of A (NaN = free), examples/LDS_knowns_in_A.py:72-74
    double *Kf = pK + 64;                                   // SCAN: [8][24] the interior smoother gain
    double *ce = Kf + 192;                                  // SCAN: [32][8] chunk ends

    for (int b = blockIdx.x * LW + warp; b < B; b += gridDim.x * LW) {
        // ---- load the sequence and its parameters (padded to 8 x 8)
        for (int i = lane; i < (T + 2) * 8; i += 32) base[i] = 0.0;
        for (int i = lane; i < T * d + 8; i += 32) ys[i] = 0.0;
        for (int i = lane; i < 64 * 4; i += 32) pA[i] = 0.0;
        __syncwarp();
        const double *Yb = Y + (size_t)b * T * d;
        double *Xb = X + (size_t)b * T * q;
        // (asynchronous copies, all in flight at once: with ~7 warps per SM a register-staged loop is a chain of HBM latencies)
        for (int i = lane; i < T * d; i += 32) ldgsts8(ys + i, Yb + i);
        for (int t0 = 0; t0 < T; t0 += 4) {                     // four rows per instruction: lane = (row mod 4, column)
            const int t = t0 + (lane >> 3), i = lane & 7;
            if (t < T && i < q) ldgsts8(xs + t * 8 + i, Xb + t * q + i);
        }
        for (int i = lane; i < q * q; i += 32) {
            ldgsts8(pA + (i / q) * 8 + (i % q), A + (size_t)b * q * q + i);
            ldgsts8(pAv + (i / q) * 8 + (i % q), Avar + (size_t)b * q * q + i);
        }
        for (int i = lane; i < d * q; i += 32) {
            ldgsts8(pC + (i / q) * 8 + (i % q), C + (size_t)b * d * q + i);
            ldgsts8(pCv + (i / q) * 8 + (i % q), Cvar + (size_t)b * d * q + i);
        }
        ldgsts_wait_all();
        for (int i = lane; i < 64; i += 32) pK[i] = __longlong_as_double(0x7ff8000000000000LL);
        __syncwarp();
        if (Aknown != nullptr)
            for (int i = lane; i < q * q; i += 32) pK[(i / q) * 8 + (i % q)] = Aknown[(size_t)b * q * q + i];
        double qb_l = (lane < q) ? Qb[(size_t)b * q + lane] : 1.0;            // lane k: Q row k; lane 8 + k: R row k
        double qa_l = (lane < q) ? Qa[(size_t)b * q + lane] : 1.0;
        if (lane >= 8 && lane < 8 + d) {
            qb_l = Rb[(size_t)b * d + lane - 8];
            qa_l = Ra[(size_t)b * d + lane - 8];
        }
        bool ok = true;
        __syncwarp();

        for (int iter = 0; iter < niters; ++iter) {
            // ---- expected precisions; padded state dimensions get unit precision (decoupled)
            if (lane < 8) qbar[lane] = (lane < q) ? qa_l / qb_l : 1.0;
            if (lane >= 8 && lane < 16) rbar[lane - 8] = (lane - 8 < d) ? qa_l / qb_l : 0.0;
            __syncwarp();
            // ---- the three posterior precisions in the accumulator layout (lane: row gid, columns 2qd, 2qd + 1)
            double c0[3], c1[3], x0[3], x1[3], lp[3] = {1.0, 1.0, 1.0};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = gid, j = 2 * qd + e;
                double aqa = 0.0, crc = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    aqa = fma(qbar[k] * pA[k * 8 + i], pA[k * 8 + j], aqa);
                    crc = fma(rbar[k] * pC[k * 8 + i], pC[k * 8 + j], crc);
                    if (i == j) {
                        aqa = fma(qbar[k], pAv[k * 8 + i], aqa);
                        crc = fma(rbar[k], pCv[k * 8 + i], crc);
                    }
                }
                const double eye = (i == j) ? 1.0 : 0.0, qd_ = (i == j) ? qbar[i] : 0.0;
                const double p0 = eye + aqa + crc, pi = qd_ + aqa + crc, pT = qd_ + crc;
                if (e == 0) { c0[0] = p0; c0[1] = pi; c0[2] = pT; } else { c1[0] = p0; c1[1] = pi; c1[2] = pT; }
            }
            diag_chol_inv<3>(c0, c1, x0, x1, lp);
            ok = ok && (lp[0] - lp[0] == 0.0) && (lp[1] - lp[1] == 0.0) && (lp[2] - lp[2] == 0.0);
            // Sigma_s = X_s^T X_s
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                tmp[gid * 8 + 2 * qd] = x0[s];
                tmp[gid * 8 + 2 * qd + 1] = x1[s];
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = gid, j = 2 * qd + e;
                    double sg = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) sg = fma(tmp[k * 8 + i], tmp[k * 8 + j], sg);
                    Sg[s * 64 + i * 8 + j] = sg;
                }
                __syncwarp();
            }
            // ---- smoother gains: lane (row i = gid, quarter p = qd) holds K_s[i][6p .. 6p+5],
            //      inputs v = [x_{t-1} (8) | x_{t+1} (8) | y_t (8)]
            double kc[3][6];
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int c = qd * 6 + m;
                double k0 = 0.0, k1 = 0.0, k2 = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    double mjc;
                    if (c < 8) mjc = qbar[j] * pA[j * 8 + c];                        // (Qbar A)[j][c]
                    else if (c < 16) mjc = pA[(c - 8) * 8 + j] * qbar[c - 8];        // (A^T Qbar)[j][c-8]
                    else mjc = pC[(c - 16) * 8 + j] * rbar[c - 16];                  // (C^T Rbar)[j][c-16]
                    k0 = fma(Sg[gid * 8 + j], mjc, k0);
                    k1 = fma(Sg[64 + gid * 8 + j], mjc, k1);
                    k2 = fma(Sg[128 + gid * 8 + j], mjc, k2);
                }
                kc[0][m] = (c < 8) ? 0.0 : k0;                                      // t = 0: no x_{t-1} term
                kc[1][m] = k1;
                kc[2][m] = (c >= 8 && c < 16) ? 0.0 : k2;                           // t = T-1: no x_{t+1} term
            }
            // offsets of this lane's six inputs relative to row t: x rows t-1 / t+1 live in xs, y in ys
            const double *src[6];
            int str[6];                                             // row pitch of the input: 8 (states) or d (observations)
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int c = qd * 6 + m;
                src[m] = (c < 8) ? (xs + c - 8) : (c < 16) ? (xs + c) : (ys + c - 16);
                str[m] = (c < 16) ? 8 : d;
            }
            auto step = [&](int t, const double (&kk)[6]) {
                double a = 0.0, bsum = 0.0;
#pragma unroll
                for (int m = 0; m < 6; m += 2) {
                    a = fma(kk[m], src[m][t * str[m]], a);           // (padded observation columns: zero gain, finite data)
                    bsum = fma(kk[m + 1], src[m + 1][t * str[m + 1]], bsum);
                }
                a += bsum;
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                if (qd == 0) xs[t * 8 + gid] = a;                   // row t is not an input of step t
                __syncwarp();
            };
            if (!SCAN) {
                step(0, kc[0]);
                for (int t = 1; t < T - 1; ++t) step(t, kc[1]);
                step(T - 1, kc[2]);
                step(T - 1, kc[2]);
                for (int t = T - 2; t >= 1; --t) step(t, kc[1]);
                step(0, kc[0]);
            } else {
#pragma unroll
                for (int m = 0; m < 6; ++m) Kf[gid * 24 + qd * 6 + m] = kc[1][m];
                __syncwarp();
                const int n = T - 2;                                // interior steps t = 1 .. T-2 (T >= 3)
                const int L = (n + 31) / 32, nch = (n + L - 1) / L;
                // one sweep over the interior: fwd: t = 1 .. T-2 from x_0;  !fwd: t = T-2 .. 1 from x_{T-1}
                auto sweep = [&](const bool fwd) {
                    {   // A. u_t = [K2 | K3] [x_{t+1} | y_t]  (fwd)  /  [K1 | K3] [x_{t-1} | y_t]: lane = (t mod 4, row)
                        const int tt = lane >> 3, i = lane & 7;
                        double F[16];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            F[c] = Kf[i * 24 + (fwd ? 8 : 0) + c];
                            F[8 + c] = Kf[i * 24 + 16 + c];
                        }
                        for (int g0 = 0; g0 < n; g0 += 8) {             // two groups of four time steps in flight
                            double acc[2];
                            int tw[2];
                            bool act[2];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                int idx = g0 + 4 * h + tt;
                                act[h] = idx < n;
                                if (!act[h]) idx = n - 1;
                                const int t = fwd ? 1 + idx : T - 2 - idx;
                                tw[h] = t;
                                const double *xr = xs + (fwd ? t + 1 : t - 1) * 8, *yr = ys + t * d;
                                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                                for (int c = 0; c < 8; c += 4) {
                                    const double2 xv = *reinterpret_cast<const double2 *>(xr + c);
                                    const double2 xw = *reinterpret_cast<const double2 *>(xr + c + 2);
                                    a0 = fma(F[c], xv.x, a0);
                                    a1 = fma(F[c + 1], xv.y, a1);
                                    a2 = fma(F[c + 2], xw.x, a2);
                                    a3 = fma(F[c + 3], xw.y, a3);
                                }
#pragma unroll
                                for (int c = 0; c < 8; c += 4) {        // (padded observation columns: zero gain, finite data)
                                    a0 = fma(F[8 + c], yr[c], a0);
                                    a1 = fma(F[9 + c], yr[c + 1], a1);
                                    a2 = fma(F[10 + c], yr[c + 2], a2);
                                    a3 = fma(F[11 + c], yr[c + 3], a3);
                                }
                                acc[h] = (a0 + a1) + (a2 + a3);
                            }
                            __syncwarp();                               // a row written here is an input of a neighbouring lane group
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                if (act[h]) xs[tw[h] * 8 + i] = acc[h];
                        }
                        __syncwarp();
                    }
                    // B. the recurrence matrix, row-major in `tmp`, and its L-th power in Mr (binary powering; the statistics
                    //    arrays are free until the sweeps are over)
                    for (int e = lane; e < 64; e += 32) tmp[e] = Kf[(e >> 3) * 24 + (fwd ? 0 : 8) + (e & 7)];
                    __syncwarp();
                    double *Mr = sSA, *Mt = sSC, *Mb = sXX1, *Mt2 = sYX;
                    if (nch > 1) {
                        bool have = false;
                        const double *bs = tmp;
                        for (int e = L; e; e >>= 1) {
                            if (e & 1) {
                                if (!have) {
                                    for (int k = lane; k < 64; k += 32) Mr[k] = bs[k];
                                    __syncwarp();
                                    have = true;
                                } else {
                                    mat8_mul(Mt, Mr, bs, gid, qd);
                                    double *sw = Mr; Mr = Mt; Mt = sw;
                                }
                            }
                            if (e >> 1) {
                                double *dst = (bs == Mb) ? Mt2 : Mb;
                                mat8_mul(dst, bs, bs, gid, qd);
                                bs = dst;
                            }
                        }
                    }
                    double Kr[8][8];
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            const double2 v = *reinterpret_cast<const double2 *>(tmp + r * 8 + j);
                            Kr[r][j] = v.x;
                            Kr[r][j + 1] = v.y;
                        }
                    const double *start0 = fwd ? xs : xs + (T - 1) * 8;   // the row the first chunk starts from
                    // one chunk, from the start value x[]; STORE: x_t replaces u_t
                    auto chunk = [&](double (&x)[8], const bool store) {
                        for (int sidx = 0; sidx < L; ++sidx) {
                            const int idx = lane * L + sidx;
                            if (idx < n) {
                                double *row = xs + (fwd ? 1 + idx : T - 2 - idx) * 8;
                                double xn[8];
#pragma unroll
                                for (int r = 0; r < 8; r += 2) {
                                    const double2 u = *reinterpret_cast<const double2 *>(row + r);
                                    xn[r] = u.x;
                                    xn[r + 1] = u.y;
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j)
#pragma unroll
                                    for (int r = 0; r < 8; ++r) xn[r] = fma(Kr[r][j], x[j], xn[r]);
#pragma unroll
                                for (int r = 0; r < 8; ++r) x[r] = xn[r];
                                if (store) {
#pragma unroll
                                    for (int r = 0; r < 8; r += 2)
                                        *reinterpret_cast<double2 *>(row + r) = make_double2(xn[r], xn[r + 1]);
                                }
                            }
                        }
                    };
                    double x[8];
                    if (nch > 1) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) x[r] = 0.0;
                        chunk(x, false);                                 // pass 1: e_c
#pragma unroll
                        for (int r = 0; r < 8; r += 2) *reinterpret_cast<double2 *>(ce + lane * 8 + r) = make_double2(x[r], x[r + 1]);
                        __syncwarp();
                        double Pr[8];                                    // lane i < 8: row i of K^L
#pragma unroll
                        for (int j = 0; j < 8; ++j) Pr[j] = Mr[(lane & 7) * 8 + j];
                        for (int cc = 0; cc + 1 < nch; ++cc) {           // E_c = e_c + K^L E_{c-1}
                            const double *Ep = cc ? ce + (cc - 1) * 8 : start0;
                            double v = ce[cc * 8 + (lane & 7)], w = 0.0;
#pragma unroll
                            for (int j = 0; j < 8; j += 2) {
                                v = fma(Pr[j], Ep[j], v);
                                w = fma(Pr[j + 1], Ep[j + 1], w);
                            }
                            __syncwarp();
                            if (lane < 8) ce[cc * 8 + lane] = v + w;
                            __syncwarp();
                        }
                    }
                    {   // pass 2 from the true chunk starts
                        const double *sp = (lane == 0 || nch == 1) ? start0 : ce + (lane - 1) * 8;
#pragma unroll
                        for (int r = 0; r < 8; ++r) x[r] = sp[r];
                        chunk(x, true);
                    }
                    __syncwarp();
                };
                step(0, kc[0]);
                sweep(true);
                step(T - 1, kc[2]);
                step(T - 1, kc[2]);
                sweep(false);
                step(0, kc[0]);
            }

            // ---- sufficient statistics of the sequence; lane owns entries (gid, 2qd), (gid, 2qd + 1)
            double sxx0 = 0.0, sxx1 = 0.0, sx10 = 0.0, sx11 = 0.0, syx0 = 0.0, syx1 = 0.0, syy = 0.0;
            double xp0 = 0.0, xp1 = 0.0;                            // x_{t-1}[2qd], x_{t-1}[2qd+1]
            for (int t = 0; t < T; ++t) {
                const double xg = xs[t * 8 + gid], yg = (gid < d) ? ys[t * d + gid] : 0.0;
                const double2 xj = *reinterpret_cast<const double2 *>(xs + t * 8 + 2 * qd);
                sxx0 = fma(xg, xj.x, sxx0);
                sxx1 = fma(xg, xj.y, sxx1);
                sx10 = fma(xg, xp0, sx10);                          // t = 0: x_{-1} = 0
                sx11 = fma(xg, xp1, sx11);
                syx0 = fma(yg, xj.x, syx0);
                syx1 = fma(yg, xj.y, syx1);
                syy = fma(yg, yg, syy);
                xp0 = xj.x;
                xp1 = xj.y;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = gid, j = 2 * qd + e;
                const double sxx = e ? sxx1 : sxx0;
                const double sc = sxx + Sg[i * 8 + j] + (double)(T - 2) * Sg[64 + i * 8 + j] + Sg[128 + i * 8 + j];
                sSC[i * 8 + j] = sc;
                sSA[i * 8 + j] = sc - (xs[(T - 1) * 8 + i] * xs[(T - 1) * 8 + j] + Sg[128 + i * 8 + j]);
                sXX1[i * 8 + j] = e ? sx11 : sx10;
                sYX[i * 8 + j] = e ? syx1 : syx0;
            }
            if (qd == 0) {
                yy[gid] = syy;
                xx0[gid] = xs[gid] * xs[gid] + Sg[gid * 8 + gid];   // <x_0 x_0^T>_kk
            }
            __syncwarp();

            // ---- parameters: lane k < 8 owns row k of A (and Q_k), lane 8 + k row k of C (and R_k)
            if (lane < 16) {
                const bool isA = lane < 8;
                const int k = isA ? lane : lane - 8;
                const bool live = isA ? (k < q) : (k < d);
                double *pm = isA ? pA : pC, *pv = isA ? pAv : pCv;
                const double *S = isA ? sSA : sSC, *cross = isA ? sXX1 : sYX;
                const double lam = isA ? qbar[k] : rbar[k];
                if (live) {
                    double a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = pm[k * 8 + j];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < q) {
                            const double prec = alpha0 + lam * S[i * 8 + i];
                            double m2 = cross[k * 8 + i];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (j != i) m2 = fma(-S[i * 8 + j], a[j], m2);
                            a[i] = lam * m2 / prec;
                            pv[k * 8 + i] = 1.0 / prec;
                            // a partially observed column (Gaussian.update, gaussian.py:125-134) with a diagonal posterior
                            // covariance: the known entries are clamped (value, zero variance), the others are untouched
                            const double kv = isA ? pK[k * 8 + i] : __longlong_as_double(0x7ff8000000000000LL);
                            if (kv == kv) {
                                a[i] = kv;
                                pv[k * 8 + i] = 0.0;
                            }
                        }
                    }
                    double quad = 0.0, lin = 0.0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        pm[k * 8 + i] = a[i];
                        lin = fma(a[i], cross[k * 8 + i], lin);
                        quad = fma(pv[k * 8 + i], S[i * 8 + i], quad);
#pragma unroll
                        for (int j = 0; j < 8; ++j) quad = fma(a[i] * a[j], S[i * 8 + j], quad);
                    }
                    if (isA) {
                        qb_l = b0 + 0.5 * (sSC[k * 8 + k] - xx0[k]) + 0.5 * quad - lin;
                        qa_l = a0 + 0.5 * (double)(T - 1);
                    } else {
                        qb_l = b0 + 0.5 * yy[k] + 0.5 * quad - lin;
                        qa_l = a0 + 0.5 * (double)T;
                    }
                }
            }
            __syncwarp();
        }

        // ---- write back
        for (int t0 = 0; t0 < T; t0 += 4) {
            const int t = t0 + (lane >> 3), i = lane & 7;
            if (t < T && i < q) Xb[t * q + i] = xs[t * 8 + i];
        }
        for (int i = lane; i < q * q; i += 32) {
            A[(size_t)b * q * q + i] = pA[(i / q) * 8 + (i % q)];
            Avar[(size_t)b * q * q + i] = pAv[(i / q) * 8 + (i % q)];
        }
        for (int i = lane; i < d * q; i += 32) {
            C[(size_t)b * d * q + i] = pC[(i / q) * 8 + (i % q)];
            Cvar[(size_t)b * d * q + i] = pCv[(i / q) * 8 + (i % q)];
        }
        for (int i = lane; i < 3 * q * q; i += 32) {
            const int s = i / (q * q), r = i % (q * q);
            Xcov3[(size_t)b * 3 * q * q + i] = Sg[s * 64 + (r / q) * 8 + (r % q)];
        }
        if (lane < q) {
            Qa[(size_t)b * q + lane] = qa_l;
            Qb[(size_t)b * q + lane] = qb_l;
        }
        if (lane >= 8 && lane < 8 + d) {
            Ra[(size_t)b * d + lane - 8] = qa_l;
            Rb[(size_t)b * d + lane - 8] = qb_l;
        }
        if (lane == 0 && !ok) atomicAdd(status, 1.0);              // a posterior precision was not positive definite
        __syncwarp();
    }
}

}  // namespace

// PYVB_LDS = serial: the step-by-step sweep (cross-check)
static bool lds_scan() {
    const char *e = getenv("PYVB_LDS");
    return !(e && e[0] == 's' && e[1] == 'e');
}

size_t lds_smem_bytes(int T, int d) {
    const bool scan = lds_scan();
    return (size_t)(scan ? LWC<true>::LW : LWC<false>::LW) * lds_warp_doubles(T, d, scan) * sizeof(double);
}

template <bool SCAN>
static cudaError_t launch_lds_t(int B, int T, int q, int d, const double *Y, double *X, double *Xcov3, double *A, double *Avar,
                                double *C, double *Cvar, double *Qa, double *Qb, double *Ra, double *Rb, double alpha0,
                                double a0, double b0, int niters, double *status, cudaStream_t st, const double *Aknown) {
    constexpr int LW = LWC<SCAN>::LW;
    const size_t smem = (size_t)LW * lds_warp_doubles(T, d, SCAN) * sizeof(double);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(lds_iterate_kernel<SCAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    long long blocks = ((long long)B + LW - 1) / LW;
    if (blocks > 148LL * per_sm) blocks = 148LL * per_sm;
    lds_iterate_kernel<SCAN><<<(unsigned)blocks, 32 * LW, smem, st>>>(B, T, q, d, Y, X, Xcov3, A, Avar, C, Cvar, Qa, Qb, Ra,
                                                                     Rb, alpha0, a0, b0, niters, status, Aknown);
    return cudaGetLastError();
}

cudaError_t launch_lds_iterate(int B, int T, int q, int d, const double *Y, double *X, double *Xcov3, double *A,
                               double *Avar, double *C, double *Cvar, double *Qa, double *Qb, double *Ra, double *Rb,
                               double alpha0, double a0, double b0, int niters, double *status, cudaStream_t st,
                               const double *Aknown) {
    if (B <= 0 || niters <= 0) return cudaSuccess;
    if (lds_scan())
        return launch_lds_t<true>(B, T, q, d, Y, X, Xcov3, A, Avar, C, Cvar, Qa, Qb, Ra, Rb, alpha0, a0, b0, niters, status, st,
                                  Aknown);
    return launch_lds_t<false>(B, T, q, d, Y, X, Xcov3, A, Avar, C, Cvar, Qa, Qb, Ra, Rb, alpha0, a0, b0, niters, status, st,
                               Aknown);
}

}  // namespace pyvb
