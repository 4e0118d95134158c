// qp_solver.h -- batched dense friction-pyramid QP, one thread team per problem, float64.
//
// Replaces the qpOASES call of SolveMPC (/root/reference/quadruped/src/controllers/mpc/
// qr_mpc_interface.cpp:414-438; constraint rows from ResizeQPMats :230-240):
//     min 1/2 x'Hx + g'x   s.t.  0 <= mu_*fx + fz, 0 <= -mu_*fx + fz, 0 <= mu_*fy + fz,
//                                0 <= -mu_*fy + fz, 0 <= fz <= ub        per foot-step (mu_ = 1/mu)
// Foot-steps with ub == 0 (swing) are pinned to f = 0 by these rows, so their variables are
// eliminated up front; the row 0 <= fz is implied by the pyramid rows and is dropped.  What is left
// are nf "stance" foot-steps, n = 3*nf variables and 5 inequalities on each 3-vector.
//
// Method (not qpOASES' online active set -- one working-set change per iteration is inherently
// sequential and needs 30-130 of them):
//   1. Block active-set iteration from a cold start.  Every round guesses, per foot-step, which of
//      its rows are active.  The active rows of a foot-step define f = Z y + p with Z (3 x d, d <= 3);
//      all free directions of all foot-steps are packed into one dense reduced system
//      (Z'HZ) y = -Z'(Hp + g), solved by a blocked Cholesky.  Then the multipliers of the active rows
//      and the values of the inactive rows are checked for every foot-step at once and the guesses
//      corrected (violated rows added, rows with a negative multiplier dropped).  When a round
//      changes nothing the KKT conditions hold to feas_tol / mult_tol, i.e. the point is THE optimum
//      of the strictly convex QP (H >= 2*alpha*I), not an approximation of it.  Typically 4-10 rounds.
//   2. Fallback (about 0.1-1 % of instances, where the block updates cycle): Mehrotra
//      predictor-corrector interior point to identify the active set (A'DA is block-diagonal 3x3 so
//      an iteration is one Cholesky of H + blkdiag and two solves), then the same verification.
//
// Storage: symmetric matrices are lower block-triangular with full 3x3 blocks, block (S,T), S >= T,
// at 9*(S*(S+1)/2 + T).  H and the matrix under factorisation both live in shared memory; the
// vectors only the fallback needs live in a per-CTA global scratch.
#pragma once

#include "qr_team.h"
#include "../../include/qr_gpu.h"

#ifndef QR_TRACE_ROUND
#define QR_TRACE_ROUND(round, nred, W) ((void)0)   // scratch analysis hook (host emulation only)
#endif

#ifndef QR_CYCLE_CHANGES
#define QR_CYCLE_CHANGES 9
#endif
#ifndef QR_COARSE_MAX_ROUNDS
#define QR_COARSE_MAX_ROUNDS 4   // rounds spent on the coarse problem (its last guess is used either way); measured on B200, A1 trot 65536: cap 3 / 4 / 5 / 7 / 16 -> 3.89 / 3.84 / 3.78 / 3.69 / 3.56 M QP/s, mixed gaits best at 4-5
#endif
#ifndef QR_CYCLE_CHANGES_DIV
#define QR_CYCLE_CHANGES_DIV 4
#endif

struct QrQpWork {
    int nf;             // stance foot-steps
    int k8;             // device: K is in shared memory with room for the 8x8-tile layout of chol8.h (set by the carve)
    double mu_;         // 1/mu as the reference rounds it (float32 value)
    double* Hs;         // [9*ntri(cap)] shared: symmetric block-packed Hessian
    double* K;          // [9*ntri(cap)] shared: matrix under factorisation
    double* Dinv;       // [9*cap]  inverses of the 3x3 pivot blocks of the LDL' factorisation
    double* zv;         // [9*cap]  basis vector (3 doubles) of every reduced variable
    double* ps;         // [3*cap]  offsets p of f = Z y + p
    double* g;          // [3*cap]
    double* xn;         // [3*cap]  current / final iterate
    double* q;          // [3*cap]  H p + g, later H x + g
    double* wv;         // [3*cap]  right-hand side / solve work vector
    double* dx;         // [3*cap]  solve result
    double* ubz;        // [cap]
    int* act;           // [cap] active-row bit masks (bit c = row c; bit 4 = cap)
    int* vert;          // [cap] foot-step pinned to the apex f = 0
    int* flag;          // [cap]
    int* foff;          // [cap+1] first reduced variable of every foot-step
    int* rfoot;         // [3*cap] foot-step of every reduced variable (-1: padding)
    unsigned short* tri;// [ntri(cap)] lower-triangular index decode table, (I << 8) | J
    int* hist;          // [24 + cap] or null: cycle detection -- hashes of the last 16 guesses, two hash accumulators,
                        // [18] trigger flag, [24 + f] how often foot-step f has changed its guess
    // fallback-only vectors (global scratch)
    double *x, *dxa, *rd, *yv;          // [3*cap]
    double *s, *lam, *dsa, *dla, *rc, *dl;  // [5*cap]
    double* red;                        // [4*cap]
};

QR_DEV int qr_blk(int S, int T) { return 9 * ((S * (S + 1)) / 2 + T); }
// The matrix under LDL' factorisation (the reduced system) is packed by COLUMNS instead: block (I,J), I >= J, of an
// nb x nb block matrix at 9*(J*nb - J(J-1)/2 + I - J).  The trailing sub-matrix of every elimination step is then
// one contiguous run of blocks, so the 32 lanes of a warp update 32 consecutive tiles: their 64-bit accesses fall
// on distinct banks (with the row-packed layout the gaps between rows cost 27-41 % extra shared-memory
// wavefronts in this loop -- the shared-memory pipe is the unit that limits the kernel's throughput).
QR_DEV int qr_kcol(int nb, int J) { return J * nb - (J * (J - 1)) / 2; }
QR_DEV int qr_kblk(int nb, int I, int J) { return 9 * (qr_kcol(nb, J) + (I - J)); }

// Decode a lower-triangular linear index idx -> (I, J), I >= J (used once per launch to fill W.tri).
QR_DEV void qr_tri_decode(int idx, int& I, int& J) {
    int i = (int)((sqrtf(8.f * (float)idx + 1.f) - 1.f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= idx) ++i;
    while (i * (i + 1) / 2 > idx) --i;
    I = i;
    J = idx - i * (i + 1) / 2;
}

QR_DEV double qr_rsqrt(double v) {
#ifdef QR_ON_DEVICE
    return rsqrt(v);
#else
    return 1.0 / sqrt(v);
#endif
}

// Row i of (Hs * v), Hs symmetric block-packed.  Three independent accumulators keep the FMA
// dependency chain a third as long.
// SMALL (the WBC force QP: at most four blocks, Hessian in shared memory): loops kept rolled -- the default fourfold
// unrolling is only code there, and that kernel is bound by instruction fetch.
template <bool SMALL>
QR_DEV double qr_sym_matvec_row_t(const double* Hs, const double* v, int nf, int i) {
    const int S = i / 3, a = i - 3 * S;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const double* row = Hs + qr_blk(S, 0) + 3 * a;
#pragma unroll(SMALL ? 1 : 4)
    for (int T = 0; T <= S; ++T, row += 9) {
        a0 += row[0] * v[3 * T];
        a1 += row[1] * v[3 * T + 1];
        a2 += row[2] * v[3 * T + 2];
    }
    const double* col = Hs + qr_blk(S + 1, S) + a;
#pragma unroll(SMALL ? 1 : 4)
    for (int T = S + 1; T < nf; ++T) {
        a0 += col[0] * v[3 * T];
        a1 += col[3] * v[3 * T + 1];
        a2 += col[6] * v[3 * T + 2];
        col += 9 * (T + 1);
    }
    return (a0 + a1) + a2;
}
QR_DEV double qr_sym_matvec_row(const double* Hs, const double* v, int nf, int i) { return qr_sym_matvec_row_t<false>(Hs, v, nf, i); }

#if defined(QR_ON_DEVICE)
// q = Hs v + g on the rows of the foot-steps with act != 0 (all rows when act is null), four lanes per row: lane `part`
// of a quad takes the blocks T = part (mod 4) and the partial sums meet by shuffle.  For the long-horizon classes, whose
// Hessian streams from L2: with one thread per row a row is a serial chain of up to 72 dependent-latency L2 loads on the
// 40 % of the threads whose foot-step has active rows (17 k cycles per round at h = 30).  Different summation order than
// qr_sym_matvec_row (1e-16 relative), which is why the h <= 16 classes do not use it.
template <int NT>
__device__ __forceinline__ void qr_sym_matvec_quad(const double* Hs, const double* v, const double* g, const int* act,
                                                   int nf, double* q) {
    const int n = 3 * nf;
    for (int item = threadIdx.x; item < 4 * n; item += NT) {   // NT and 4n are multiples of 4: a quad stays together
        const int i = item >> 2, part = item & 3;
        const int S = i / 3, a = i - 3 * S;
        if (act && !act[S]) continue;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const double* row = Hs + qr_blk(S, part) + 3 * a;
#pragma unroll 4
        for (int T = part; T <= S; T += 4, row += 36) {
            a0 += row[0] * v[3 * T];
            a1 += row[1] * v[3 * T + 1];
            a2 += row[2] * v[3 * T + 2];
        }
#pragma unroll 4
        for (int T = S + 1 + ((part - (S + 1)) & 3); T < nf; T += 4) {
            const double* col = Hs + qr_blk(T, S) + a;
            a0 += col[0] * v[3 * T];
            a1 += col[3] * v[3 * T + 1];
            a2 += col[6] * v[3 * T + 2];
        }
        double s = (a0 + a1) + a2;
        const unsigned m = 0xFu << (threadIdx.x & 28);
        s += __shfl_xor_sync(m, s, 1);
        s += __shfl_xor_sync(m, s, 2);
        if (part == 0) q[i] = s + g[i];
    }
}
#endif

// Row i of (Hs * p) where p is non-zero only on foot-steps whose cap row is active (act bit 4).
QR_DEV double qr_sym_matvec_row_capped(const double* Hs, const double* p, const int* act, int nf, int i) {
    const int S = i / 3, a = i - 3 * S;
    double acc = 0.0;
    for (int T = 0; T < nf; ++T) {
        if (!(act[T] & 16)) continue;
        if (T <= S) {
            const double* row = Hs + qr_blk(S, T) + 3 * a;
            acc += row[0] * p[3 * T] + row[1] * p[3 * T + 1] + row[2] * p[3 * T + 2];
        } else {
            const double* col = Hs + qr_blk(T, S) + a;
            acc += col[0] * p[3 * T] + col[3] * p[3 * T + 1] + col[6] * p[3 * T + 2];
        }
    }
    return acc;
}

// 1/d for a positive normal d.  Device: hardware seed (rcp.approx, about 20 bits) refined by two Newton steps
// to double precision -- a 60-cycle dependent chain instead of the ~130 of the IEEE division subroutine; it
// sits on the critical path of every factorisation step.
QR_DEV double qr_rcp_pos(double d) {
#ifdef QR_ON_DEVICE
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / d;
#endif
}

// Inverse of a symmetric positive definite 3x3 block [a b c; b d e; c e f] by its adjugate: one
// reciprocal and no square root, so the dependent chain is short (cofactors -> determinant -> 1/det).
// Written as a full symmetric 3x3 (row-major) to o[0..8].  A non-positive determinant is clamped
// (the caller's verification and the final finiteness check catch a breakdown).
QR_DEV void qr_inv3_sym(double a, double b, double c, double d, double e, double f, double* o) {
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
    double det = a * c00 + b * c01 + c * c02;
    if (!(det > 1e-300)) det = 1e-300;
    const double r = qr_rcp_pos(det);
    o[0] = c00 * r; o[1] = c01 * r; o[2] = c02 * r;
    o[3] = o[1];    o[4] = c11 * r; o[5] = c12 * r;
    o[6] = o[2];    o[7] = o[5];    o[8] = c22 * r;
}

#include "chol8.h"   // blocked Cholesky on the FP64 tensor cores for the large reduced systems (device only)

// Block LDL' factorisation K = L D L' of the leading nb x nb blocks (3x3 pivot blocks, no pivoting:
// K is symmetric positive definite), fused with the forward substitution of the right-hand side
// W.wv when with_rhs != 0.
//
// One barrier per block column.  In step Kc every thread owning a trailing tile (I,J), I >= J > Kc,
// reads the UNSCALED column tiles W_IK, W_JK and the pivot inverse Dinv_K and applies
//     A_IJ -= (W_IK Dinv_K) W_JK'
// (the scaling by Dinv_K is redone per tile: redundant flops are cheaper than a second barrier).
// The thread that produces the next pivot block (Kc+1,Kc+1) inverts it on the spot and publishes
// Dinv_{K+1}; that block is never written back.  The right-hand side is treated as one more block
// row: y_J -= W_JK (Dinv_K y_K).  On exit: off-diagonal tiles hold W = L D (unscaled columns),
// W.Dinv the pivot inverses, and W.wv = L^{-1} rhs.
// Device: teams of at least this many threads factorise with one lane per tile ROW instead of one per tile (see
// qr_ldl_factor).  Measured on B200 (A1 trot, 65536 instances): the row form issues 1.6x the instructions of the
// tile form in the factorisation (21 loads per 18 DFMA instead of 33 per 54), which costs the 96-thread throughput
// kernels 10 % (4.12 -> 3.70 M QP/s: at six CTAs per SM they are bound by issued instructions, not by the latency of a
// step), while the kernels that run ONE 256-thread CTA per SM -- the latency kernel and the long-horizon size
// classes -- are bound by the latency of a step and gain from it.
#ifndef QR_LDL_ROWSPLIT_MIN_NT
#define QR_LDL_ROWSPLIT_MIN_NT 256
#endif
#ifndef QR_LDL_ROWSPLIT_MAX_NB
#define QR_LDL_ROWSPLIT_MAX_NB 24   // larger systems (long horizons): back to the tile form (h = 30: 165 k QP/s against 138 k with rows)
#endif

template <int NT>
QR_DEV void qr_ldl_factor(QrQpWork& W, int nb, int with_rhs QR_PROF_ARG) {
    double* K = W.K;
    double* y = W.wv;
    QR_THREADS(t) {
        if (t == 0) {
            const double* D = K;
            qr_inv3_sym(D[0], D[3], D[6], D[4], D[7], D[8], W.Dinv);
        }
    }
    QR_SYNC();
#if defined(QR_ON_DEVICE)
    if (NT >= QR_LDL_ROWSPLIT_MIN_NT && nb <= QR_LDL_ROWSPLIT_MAX_NB) {
    // Device: one lane per tile ROW.  With the coarse prediction in front, most factorisations have 9..17 block columns
    // and most of their steps fewer tiles than the team has lanes, so a step costs the latency of one item, not its
    // flops: a row (18 DFMA, 21 loads) is a third of a tile (54 DFMA, 33 loads), three times as many lanes take part,
    // and the warp instructions of a step cover three times fewer tiles' worth of work each.  The three rows of the
    // next pivot block sit on lanes 0..2 of warp 0; lane 0 gathers them by shuffle and inverts the block one step ahead
    // as before.  Every entry is the same expression as in the tile form below (host emulation), evaluated once.
    for (int Kc = 0; Kc < nb; ++Kc) {
        const int nrem = nb - Kc - 1;
        const int ntile = (nrem * (nrem + 1)) / 2;
        const int kc0 = qr_kcol(nb, Kc), kc1 = qr_kcol(nb, Kc + 1);
        const double* Di = W.Dinv + 9 * Kc;
        const int nitem = 3 * ntile + (with_rhs ? 3 * nrem : 0);
        const double d0 = Di[0], d1 = Di[1], d2 = Di[2], d4 = Di[4], d5 = Di[5], d8 = Di[8];
        for (int base = 0; base < nitem; base += NT) {
            const int item = base + (int)threadIdx.x;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            double* dst = nullptr;
            if (item < 3 * ntile) {
                const int tile = item / 3, q = item - 3 * tile;
                const int code = W.tri[ntile - 1 - tile];
                const int I = nb - 1 - (code & 255), J = nb - 1 - (code >> 8);
                const double* wi = K + 9 * (kc0 + (I - Kc)) + 3 * q;
                const double* wj = K + 9 * (kc0 + (J - Kc));
                double* a = K + 9 * (kc1 + tile) + 3 * q;
                const double w0 = wi[0], w1 = wi[1], w2 = wi[2];
                const double t0 = w0 * d0 + w1 * d1 + w2 * d2;
                const double t1 = w0 * d1 + w1 * d4 + w2 * d5;
                const double t2 = w0 * d2 + w1 * d5 + w2 * d8;
                r0 = a[0] - (t0 * wj[0] + t1 * wj[1] + t2 * wj[2]);
                r1 = a[1] - (t0 * wj[3] + t1 * wj[4] + t2 * wj[5]);
                r2 = a[2] - (t0 * wj[6] + t1 * wj[7] + t2 * wj[8]);
                if (tile != 0) dst = a;   // tile 0 = the next pivot block (Kc+1, Kc+1): consumed below, never written back
            } else if (item < nitem) {
                const int j = item - 3 * ntile, Jb = j / 3, q = j - 3 * Jb;
                const int J = Kc + 1 + Jb;
                const double y0 = y[3 * Kc], y1 = y[3 * Kc + 1], y2 = y[3 * Kc + 2];
                const double t0 = d0 * y0 + d1 * y1 + d2 * y2;
                const double t1 = d1 * y0 + d4 * y1 + d5 * y2;
                const double t2 = d2 * y0 + d5 * y1 + d8 * y2;
                const double* wj = K + 9 * (kc0 + (J - Kc)) + 3 * q;
                y[3 * J + q] -= wj[0] * t0 + wj[1] * t1 + wj[2] * t2;
            }
            if (base == 0 && threadIdx.x < 32 && ntile > 0) {   // warp 0, first pass: lanes 0..2 hold the pivot rows
                const double p10 = __shfl_sync(0xffffffffu, r0, 1), p11 = __shfl_sync(0xffffffffu, r1, 1);
                const double p20 = __shfl_sync(0xffffffffu, r0, 2), p21 = __shfl_sync(0xffffffffu, r1, 2);
                const double p22 = __shfl_sync(0xffffffffu, r2, 2);
                if (threadIdx.x == 0) qr_inv3_sym(r0, p10, p20, p11, p21, p22, W.Dinv + 9 * (Kc + 1));
            }
            if (dst) { dst[0] = r0; dst[1] = r1; dst[2] = r2; }
        }
        QR_SYNC();
        QR_PROF(11);
    }
    return;
    }
#endif
    for (int Kc = 0; Kc < nb; ++Kc) {
        const int nrem = nb - Kc - 1;
        const int ntile = (nrem * (nrem + 1)) / 2;
        const int kc0 = qr_kcol(nb, Kc), kc1 = qr_kcol(nb, Kc + 1);
        const double* Di = W.Dinv + 9 * Kc;
        QR_FOR(idx, ntile + (with_rhs ? nrem : 0)) {
            const double d0 = Di[0], d1 = Di[1], d2 = Di[2], d4 = Di[4], d5 = Di[5], d8 = Di[8];
            if (idx < ntile) {
                // tiles in memory order: idx-th block of the (contiguous) trailing columns Kc+1 .. nb-1
                const int code = W.tri[ntile - 1 - idx];
                const int I = nb - 1 - (code & 255), J = nb - 1 - (code >> 8);
                const double* wi = K + 9 * (kc0 + (I - Kc));
                const double* wj = K + 9 * (kc0 + (J - Kc));
                double* a = K + 9 * (kc1 + idx);
                double m[9], t[9], r[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) { m[e] = wj[e]; r[e] = a[e]; }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double w0 = wi[3 * q], w1 = wi[3 * q + 1], w2 = wi[3 * q + 2];
                    t[3 * q] = w0 * d0 + w1 * d1 + w2 * d2;
                    t[3 * q + 1] = w0 * d1 + w1 * d4 + w2 * d5;
                    t[3 * q + 2] = w0 * d2 + w1 * d5 + w2 * d8;
                }
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        r[3 * q + c] -= t[3 * q] * m[3 * c] + t[3 * q + 1] * m[3 * c + 1] + t[3 * q + 2] * m[3 * c + 2];
                if (I == Kc + 1 && J == Kc + 1) {
                    qr_inv3_sym(r[0], r[3], r[6], r[4], r[7], r[8], W.Dinv + 9 * (Kc + 1));
                } else {
#pragma unroll
                    for (int e = 0; e < 9; ++e) a[e] = r[e];
                }
            } else {
                const int J = Kc + 1 + (idx - ntile);
                const double y0 = y[3 * Kc], y1 = y[3 * Kc + 1], y2 = y[3 * Kc + 2];
                const double t0 = d0 * y0 + d1 * y1 + d2 * y2;
                const double t1 = d1 * y0 + d4 * y1 + d5 * y2;
                const double t2 = d2 * y0 + d5 * y1 + d8 * y2;
                const double* wj = K + 9 * (kc0 + (J - Kc));
                y[3 * J] -= wj[0] * t0 + wj[1] * t1 + wj[2] * t2;
                y[3 * J + 1] -= wj[3] * t0 + wj[4] * t1 + wj[5] * t2;
                y[3 * J + 2] -= wj[6] * t0 + wj[7] * t1 + wj[8] * t2;
            }
        }
        QR_SYNC();
        QR_PROF(11);
    }
}

// Stand-alone forward substitution W.wv <- L^{-1} W.wv for a further right-hand side (the
// interior-point fallback solves twice per factorisation).
template <int NT>
QR_DEV void qr_ldl_forward(QrQpWork& W, int nb QR_PROF_ARG) {
    const double* K = W.K;
    double* y = W.wv;
    for (int Kc = 0; Kc + 1 < nb; ++Kc) {
        const double* Di = W.Dinv + 9 * Kc;
        QR_FOR(idx, 3 * (nb - Kc - 1)) {
            const int I = Kc + 1 + idx / 3, a = idx % 3;
            const double y0 = y[3 * Kc], y1 = y[3 * Kc + 1], y2 = y[3 * Kc + 2];
            const double t0 = Di[0] * y0 + Di[1] * y1 + Di[2] * y2;
            const double t1 = Di[3] * y0 + Di[4] * y1 + Di[5] * y2;
            const double t2 = Di[6] * y0 + Di[7] * y1 + Di[8] * y2;
            const double* r = K + qr_kblk(nb, I, Kc) + 3 * a;
            y[3 * I + a] -= r[0] * t0 + r[1] * t1 + r[2] * t2;
        }
        QR_SYNC();
    }
    QR_PROF(12);
}

// Backward substitution: out = L'^{-1} D^{-1} W.wv, i.e. x_K = Dinv_K (y_K - sum_{I>K} W_IK' x_I).
// In place on W.wv (a step reads block Kc and updates blocks J < Kc).
//
// On the device, systems of up to 32 blocks are solved by ONE warp without any CTA barrier: lane J keeps
// block row J of the right-hand side in registers, the pivot row's solution travels by warp shuffle and
// the next step's tile W_KJ is fetched while the current one is applied (one step is a dependent chain of
// about 75 cycles instead of a shared-memory round trip plus a CTA barrier).
template <int NT>
QR_DEV void qr_ldl_backward(QrQpWork& W, int nb, double* out QR_PROF_ARG) {
    const double* K = W.K;
    double* y = W.wv;
#ifdef QR_ON_DEVICE
    if (NT >= 32 && nb <= 32) {
        if (qr_tid<NT>() < 32) {
            const int lane = qr_tid<NT>();
            const int me = lane < nb ? lane : 0;
            double y0 = y[3 * me], y1 = y[3 * me + 1], y2 = y[3 * me + 2];
            const double* Di = W.Dinv + 9 * me;
            const double d0 = Di[0], d1 = Di[1], d2 = Di[2], d4 = Di[4], d5 = Di[5], d8 = Di[8];
            double b[9];
            {
                const double* blk = K + qr_kblk(nb, nb - 1, lane < nb - 1 ? lane : 0);
#pragma unroll
                for (int e = 0; e < 9; ++e) b[e] = blk[e];
            }
            for (int Kc = nb - 1; Kc >= 0; --Kc) {
                // every lane forms Dinv*y of its own row; only lane Kc's is final and gets broadcast
                const double x0 = d0 * y0 + d1 * y1 + d2 * y2;
                const double x1 = d1 * y0 + d4 * y1 + d5 * y2;
                const double x2 = d2 * y0 + d5 * y1 + d8 * y2;
                double n[9];
                if (Kc > 0) {
                    const double* blk = K + qr_kblk(nb, Kc - 1, lane < Kc - 1 ? lane : 0);
#pragma unroll
                    for (int e = 0; e < 9; ++e) n[e] = blk[e];
                }
                const double bx0 = __shfl_sync(0xffffffffu, x0, Kc);
                const double bx1 = __shfl_sync(0xffffffffu, x1, Kc);
                const double bx2 = __shfl_sync(0xffffffffu, x2, Kc);
                if (lane == Kc) { out[3 * Kc] = bx0; out[3 * Kc + 1] = bx1; out[3 * Kc + 2] = bx2; }
                if (lane < Kc) {
                    y0 -= b[0] * bx0 + b[3] * bx1 + b[6] * bx2;
                    y1 -= b[1] * bx0 + b[4] * bx1 + b[7] * bx2;
                    y2 -= b[2] * bx0 + b[5] * bx1 + b[8] * bx2;
                }
#pragma unroll
                for (int e = 0; e < 9; ++e) b[e] = n[e];
            }
        }
        QR_SYNC();
        QR_PROF(13);
        return;
    }
#endif
    for (int Kc = nb - 1; Kc >= 0; --Kc) {
        const double* Di = W.Dinv + 9 * Kc;
        QR_FOR(idx, 3 * (Kc + 1)) {
            const int J = idx / 3, a = idx % 3;
            const double b0 = y[3 * Kc], b1 = y[3 * Kc + 1], b2 = y[3 * Kc + 2];
            const double x0 = Di[0] * b0 + Di[1] * b1 + Di[2] * b2;
            const double x1 = Di[3] * b0 + Di[4] * b1 + Di[5] * b2;
            const double x2 = Di[6] * b0 + Di[7] * b1 + Di[8] * b2;
            if (J == Kc) {
                out[3 * Kc + a] = (a == 0 ? x0 : (a == 1 ? x1 : x2));
            } else {
                const double* blk = K + qr_kblk(nb, Kc, J);   // W_KJ, column a of it
                y[3 * J + a] -= blk[a] * x0 + blk[3 + a] * x1 + blk[6 + a] * x2;
            }
        }
        QR_SYNC();
    }
    QR_PROF(13);
}

// Constraint values of one foot-step: c0..c3 pyramid faces, c4 = ub - fz.
QR_DEV void qr_foot_constraints(double mu_, double ub, const double* f, double* c) {
    c[0] = mu_ * f[0] + f[2];
    c[1] = -mu_ * f[0] + f[2];
    c[2] = mu_ * f[1] + f[2];
    c[3] = -mu_ * f[1] + f[2];
    c[4] = ub - f[2];
}
// A_f' t for one foot-step (t has 5 entries).
QR_DEV void qr_foot_At(double mu_, const double* t, double* o) {
    o[0] = mu_ * (t[0] - t[1]);
    o[1] = mu_ * (t[2] - t[3]);
    o[2] = t[0] + t[1] + t[2] + t[3] - t[4];
}

// Basis of one foot-step's active face: f = Z y + p with the d free directions in the first d
// columns of Z (3x3 row-major).  act bits 0..3 = pyramid faces, bit 4 = cap.
// Returns d, or -1 when the active rows pin f = 0 (apex of the pyramid).
QR_DEV int qr_foot_basis(int act, double mu_, double ub, double* Z, double* p) {
    const int a0 = act & 1, a1 = (act >> 1) & 1, a2 = (act >> 2) & 1, a3 = (act >> 3) & 1, cap = (act >> 4) & 1;
    for (int e = 0; e < 9; ++e) Z[e] = 0.0;
    p[0] = p[1] = p[2] = 0.0;
    if (a0 + a1 == 2 || a2 + a3 == 2 || a0 + a1 + a2 + a3 >= 3) return -1;
    const double im = 1.0 / mu_;
    // active face mu_*fx + fz = 0 -> fx = -fz/mu_ ; face -mu_*fx + fz = 0 -> fx = +fz/mu_
    const double kx = a0 ? -im : (a1 ? im : 0.0);
    const double ky = a2 ? -im : (a3 ? im : 0.0);
    int d = 0;
    if (!(a0 | a1)) { Z[d] = 1.0; ++d; }                 // e_x
    if (!(a2 | a3)) { Z[3 + d] = 1.0; ++d; }             // e_y
    if (cap) {
        p[0] = kx * ub; p[1] = ky * ub; p[2] = ub;
    } else {
        Z[d] = kx; Z[3 + d] = ky; Z[6 + d] = 1.0; ++d;   // (kx, ky, 1) * fz
    }
    return d;
}

// Block active-set iteration (see the header comment).  cold = 1: start with no row active;
// cold = 0: start from the active set the interior-point iterate (W.s, W.lam) suggests;
// cold = 2: start from the guess the caller left in W.act (the coarse prediction of mpc_problem.h).
// Result in W.xn.  Returns rounds used; *ok = 1 when the KKT conditions were verified.
template <int NT>
QR_DEV int qr_active_set(QrQpWork& W, const qr_qp_options& opt, int* ok, int cold, int max_rounds QR_PROF_ARG) {
    const int nf = W.nf, n = 3 * nf;
    const double mu_ = W.mu_;
    const double im = 1.0 / mu_;
    if (cold != 2) {
        QR_FOR(f, nf) {
            int a = 0;
            if (!cold)
                for (int c = 0; c < 5; ++c)
                    if (W.s[5 * f + c] < opt.act_kappa * W.lam[5 * f + c]) a |= (1 << c);
            W.act[f] = a;
        }
    }
    // Cycle detection.  The block updates are a primal-dual active-set iteration and can enter a cycle (period 2-8 in
    // practice, 0.5 % of the Lite3 trot instances).  Every round's new guess is hashed; when a hash of the last 16
    // rounds comes back the iteration continues in RESTRICTED mode: per round only one foot-step -- the one with the
    // most negative multiplier -- may drop rows and only one -- the one with the most violated row -- may add rows.
    // That ends the cycles seen so far within 4-8 more rounds instead of running into the interior-point fallback.
    // Quasi-cycles (a few foot-steps rotate through the same guesses while others wobble, so that no hash repeats)
    // are caught by counting how often each foot-step has changed its guess: in converging instances -- including the
    // long "waves" of 15-20 rounds -- a foot-step changes at most 8 times.
    int restricted = 0;
    const int kchg = QR_CYCLE_CHANGES + (nf > 24 ? (nf - 24) / QR_CYCLE_CHANGES_DIV : 0);   // longer horizons: longer waves
    if (W.hist) { QR_FOR(i, 24 + nf) W.hist[i] = 0; }
    QR_SYNC();
    *ok = 0;
    int round = 0;
    for (; round < max_rounds; ++round) {
        // ---- bases of the current guess; K is free here, its head is scratch for the 3x3 bases
        double* Zfull = W.K;
        QR_FOR(f, nf) {
            const int d = qr_foot_basis(W.act[f], mu_, W.ubz[f], Zfull + 9 * f, W.ps + 3 * f);
            W.vert[f] = d < 0;
            W.flag[f] = d < 0 ? 0 : d;
        }
        QR_SYNC();
        QR_PROF(1);
        // ---- pack the free directions: every foot-step finds its offset (sum of d_g, g < f) itself
        QR_FOR(f, nf) {
            int off = 0;
            for (int gidx = 0; gidx < f; ++gidx) off += W.flag[gidx];
            const int d = W.flag[f];
            W.foff[f] = off;
            for (int c = 0; c < d; ++c) {
                W.rfoot[off + c] = f;
                W.zv[3 * (off + c)] = Zfull[9 * f + c];
                W.zv[3 * (off + c) + 1] = Zfull[9 * f + 3 + c];
                W.zv[3 * (off + c) + 2] = Zfull[9 * f + 6 + c];
            }
            if (f == nf - 1) {
                const int tot = off + d;
                W.foff[nf] = tot;
                for (int r = tot; r < 3 * ((tot + 2) / 3); ++r) W.rfoot[r] = -1;
            }
        }
        // hp = H p + g  -> q   (p is non-zero only where the cap row is active)
        QR_FOR(i, n) W.q[i] = qr_sym_matvec_row_capped(W.Hs, W.ps, W.act, nf, i) + W.g[i];
        QR_SYNC();
        QR_PROF(3);
        const int nred = W.foff[nf];
        const int nbr = (nred + 2) / 3;
        QR_TRACE_ROUND(round, nred, W);
#if defined(QR_ON_DEVICE)
        // large systems of the long-horizon classes: 8x8 tiles, blocked Cholesky on the FP64 tensor cores (chol8.h)
        const bool big = NT >= QR_CHOL8_MIN_NT && W.k8 && nbr >= QR_CHOL8_MIN_NB && !(opt.flags & QR_QP_SCALAR_FACTOR);
        const int nt8 = qr_k8_nt(nred);
#else
        const bool big = false;
        const int nt8 = 0;
#endif
        // ---- reduced matrix Z'HZ (column-packed, see qr_kblk), padded to a multiple of 3 with identity.
        // One thread per PAIR of foot-steps: the 3x3 block H_{f1 f2} is loaded once and contributes its d1 x d2
        // entries Z_f1' (H_{f1 f2} Z_f2) -- the same products in the same order as an entry-wise evaluation, with a
        // ninth of its loads and a quarter of its instructions (this phase was 17 % of the kernel's instructions).
#if defined(QR_ON_DEVICE)
        if (big) {
            // the same entries written to the 8x8-tile layout.  These classes run without a register cap, so the two
            // loops are unrolled under predicates (nine independent entries in flight instead of one) and the tile
            // address of a column is formed once.
            QR_FOR(pidx, (nf * (nf + 1)) / 2) {
                const int code = W.tri[pidx];
                const int f1 = code >> 8, f2 = code & 255;   // f1 >= f2
                const int d1 = W.flag[f1], d2 = W.flag[f2];
                if (d1 != 0 && d2 != 0) {
                    const double* Hb = W.Hs + qr_blk(f1, f2);
                    double hb[9];
#pragma unroll
                    for (int e = 0; e < 9; ++e) hb[e] = Hb[e];
                    const int o1 = W.foff[f1], o2 = W.foff[f2];
                    double y[9];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double* yp = W.zv + 3 * (o1 + (a < d1 ? a : 0));
                        y[3 * a] = yp[0]; y[3 * a + 1] = yp[1]; y[3 * a + 2] = yp[2];
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (c < d2) {
                            const double* z2 = W.zv + 3 * (o2 + c);
                            const double z0 = z2[0], z1 = z2[1], zz = z2[2];
                            const double t0 = hb[0] * z0 + hb[1] * z1 + hb[2] * zz;
                            const double t1 = hb[3] * z0 + hb[4] * z1 + hb[5] * zz;
                            const double t2 = hb[6] * z0 + hb[7] * z1 + hb[8] * zz;
                            const int r2 = o2 + c, Jt = r2 >> 3, cj = r2 & 7;
                            double* colp = W.K + 64 * (Jt * nt8 - ((Jt * (Jt - 1)) >> 1) - Jt);
#pragma unroll
                            for (int a = 0; a < 3; ++a) {
                                const int r1 = o1 + a;
                                if (a < d1 && r2 <= r1) {
                                    const double val = y[3 * a] * t0 + y[3 * a + 1] * t1 + y[3 * a + 2] * t2;
                                    const int It = r1 >> 3, ri = r1 & 7;
                                    colp[64 * It + qr_k8_swz(ri, cj)] = val;
                                    if (It == Jt) colp[64 * It + qr_k8_swz(cj, ri)] = val;
                                }
                            }
                        }
                    }
                }
            }
        } else
#endif
        QR_FOR(pidx, (nf * (nf + 1)) / 2) {
            const int code = W.tri[pidx];
            const int f1 = code >> 8, f2 = code & 255;   // f1 >= f2
            const int d1 = W.flag[f1], d2 = W.flag[f2];
            if (d1 != 0 && d2 != 0) {
                const double* Hb = W.Hs + qr_blk(f1, f2);
                double hb[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) hb[e] = Hb[e];
                const int o1 = W.foff[f1], o2 = W.foff[f2];
#pragma unroll 1
                for (int c = 0; c < d2; ++c) {   // one column of the block at a time keeps the register footprint small
                    const double* z2 = W.zv + 3 * (o2 + c);
                    const double z0 = z2[0], z1 = z2[1], zz = z2[2];
                    const double t0 = hb[0] * z0 + hb[1] * z1 + hb[2] * zz;
                    const double t1 = hb[3] * z0 + hb[4] * z1 + hb[5] * zz;
                    const double t2 = hb[6] * z0 + hb[7] * z1 + hb[8] * zz;
                    const int r2 = o2 + c, J = r2 / 3, rj = r2 - 3 * J;
#pragma unroll 1
                    for (int a = 0; a < d1; ++a) {
                        const int r1 = o1 + a;
                        if (r2 <= r1) {
                            const double* y = W.zv + 3 * r1;
                            const double val = y[0] * t0 + y[1] * t1 + y[2] * t2;
                            const int I = r1 / 3, ri = r1 - 3 * I;
                            double* blk = W.K + qr_kblk(nbr, I, J);
                            blk[3 * ri + rj] = val;
                            if (I == J) blk[3 * rj + ri] = val;   // diagonal blocks are stored in full
                        }
                    }
                }
            }
        }
#if defined(QR_ON_DEVICE)
        if (big) {   // identity padding rows nred .. 8*nt8-1, right-hand side padded with zeros
            QR_FOR(idx, (8 * nt8 - nred) * 8 * nt8) {
                const int r = nred + idx / (8 * nt8), c = idx - (r - nred) * (8 * nt8);
                if (c <= r) {
                    W.K[qr_k8_idx(nt8, r, c)] = (r == c) ? 1.0 : 0.0;
                    if ((c >> 3) == (r >> 3)) W.K[qr_k8_idx(nt8, c, r)] = (r == c) ? 1.0 : 0.0;
                }
            }
            QR_FOR(r, 8 * nt8) {
                const int f = r < nred ? W.rfoot[r] : -1;
                W.wv[r] = f < 0 ? 0.0
                                : -(W.zv[3 * r] * W.q[3 * f] + W.zv[3 * r + 1] * W.q[3 * f + 1] + W.zv[3 * r + 2] * W.q[3 * f + 2]);
            }
        } else {
#endif
        QR_FOR(idx, (3 * nbr - nred) * 3 * nbr) {   // identity padding rows nred .. 3*nbr-1
            const int r = nred + idx / (3 * nbr), c = idx - (r - nred) * (3 * nbr);
            if (c <= r) {
                double* blk = W.K + qr_kblk(nbr, nbr - 1, c / 3);
                blk[3 * (r % 3) + c % 3] = (r == c) ? 1.0 : 0.0;
                if (c / 3 == nbr - 1) blk[3 * (c % 3) + r % 3] = (r == c) ? 1.0 : 0.0;
            }
        }
        QR_FOR(r, 3 * nbr) {
            const int f = W.rfoot[r];
            W.wv[r] = f < 0 ? 0.0
                            : -(W.zv[3 * r] * W.q[3 * f] + W.zv[3 * r + 1] * W.q[3 * f + 1] + W.zv[3 * r + 2] * W.q[3 * f + 2]);
        }
#if defined(QR_ON_DEVICE)
        }
#endif
        QR_SYNC();
        QR_PROF(4);
#if defined(QR_ON_DEVICE)
        if (big) {
            if constexpr (NT >= QR_CHOL8_MIN_NT) {
                qr_chol8_factor<NT>(W.K, W.wv, W.tri, nt8, 1);
                QR_PROF(11);
                qr_chol8_backward<NT>(W.K, W.wv, W.Dinv, nt8, W.dx, nred);   // Dinv is free on this path: 16 doubles of scratch
                QR_PROF(13);
            }
        } else
#endif
        if (nbr > 0) {   // nbr == 0: every foot-step is pinned to the apex, x = p and only the verification is left
            qr_ldl_factor<NT>(W, nbr, 1 QR_PROF_PASS);
            qr_ldl_backward<NT>(W, nbr, W.dx QR_PROF_PASS);
        }
        QR_FOR(f, nf) {
            double x0 = W.ps[3 * f], x1 = W.ps[3 * f + 1], x2 = W.ps[3 * f + 2];
            const int off = W.foff[f], d = W.flag[f];
            for (int c = 0; c < d; ++c) {
                const double y = W.dx[off + c];
                x0 += y * W.zv[3 * (off + c)];
                x1 += y * W.zv[3 * (off + c) + 1];
                x2 += y * W.zv[3 * (off + c) + 2];
            }
            W.xn[3 * f] = x0; W.xn[3 * f + 1] = x1; W.xn[3 * f + 2] = x2;
        }
        QR_SYNC();
        QR_PROF(5);
        // gradient rows of the foot-steps with active rows only: the verification reads nothing else (a free foot-step's
        // gradient is zero by construction), and with the Hessian in L2 this phase is bound by its L2 -> SM traffic
#if defined(QR_ON_DEVICE)
        if (NT >= QR_CHOL8_MIN_NT && W.k8) {
            if constexpr (NT >= QR_CHOL8_MIN_NT) qr_sym_matvec_quad<NT>(W.Hs, W.xn, W.g, W.act, nf, W.q);
        } else
#endif
        QR_FOR(i, n) {
            if (W.act[i / 3]) W.q[i] = qr_sym_matvec_row_t<(NT <= 64)>(W.Hs, W.xn, nf, i) + W.g[i];
        }
        QR_SYNC();
        QR_PROF(6);
        // ---- verify / correct the active sets
        int changed = 0;
        QR_FOR(f, nf) {
            const double* r = W.q + 3 * f;
            const int act = W.act[f];
            int nact = act;
            double score = 0.0;   // most negative multiplier of a foot-step that wants to drop rows
            double ascore = 0.0;  // most negative value of a violated row of a foot-step that wants to add rows
            if (W.vert[f]) {
                // apex: the gradient must lie in the cone spanned by the four face normals
                if (r[2] < (fabs(r[0]) + fabs(r[1])) * im - opt.mult_tol) {
                    score = r[2] - (fabs(r[0]) + fabs(r[1])) * im;
                    const double l0 = r[0] > 0.0 ? r[0] * im : 0.0, l1 = r[0] < 0.0 ? -r[0] * im : 0.0;
                    const double l2 = r[1] > 0.0 ? r[1] * im : 0.0, l3 = r[1] < 0.0 ? -r[1] * im : 0.0;
                    nact = (l0 > 0.0 ? 1 : 0) | (l1 > 0.0 ? 2 : 0) | (l2 > 0.0 ? 4 : 0) | (l3 > 0.0 ? 8 : 0);
                    const int hasx = nact & 3, hasy = nact & 12;
                    if (hasx && hasy) {
                        // least-squares multipliers on the edge; drop the face that would pull inward
                        const double sx = (nact & 1) ? 1.0 : -1.0, sy = (nact & 4) ? 1.0 : -1.0;
                        const double b0 = sx * mu_ * r[0] + r[2], b1 = sy * mu_ * r[1] + r[2];
                        const double dd = mu_ * mu_ + 1.0, det = dd * dd - 1.0;
                        const double lx = (dd * b0 - b1) / det, ly = (dd * b1 - b0) / det;
                        if (lx < 0.0 || ly < 0.0) {
                            if (lx < ly) nact &= ~3; else nact &= ~12;
                        }
                    }
                }
            } else {
                double c[5];
                qr_foot_constraints(mu_, W.ubz[f], W.xn + 3 * f, c);
                int viol = 0;
                for (int k = 0; k < 5; ++k)
                    if (!((act >> k) & 1) && c[k] < -opt.feas_tol) { viol |= (1 << k); ascore = qr_min(ascore, c[k]); }
                if (viol) {
                    nact = act | viol;
                } else if (act) {
                    // multipliers of the (independent) active rows: r = sum lambda_c a_c
                    const double lx = (act & 1) ? r[0] * im : ((act & 2) ? -r[0] * im : 0.0);
                    const double ly = (act & 4) ? r[1] * im : ((act & 8) ? -r[1] * im : 0.0);
                    const double lc = (act & 16) ? (lx + ly - r[2]) : 0.0;
                    double worst = -opt.mult_tol;
                    int drop = 0;
                    if ((act & 3) && lx < worst) { worst = lx; drop = act & 3; }
                    if ((act & 12) && ly < worst) { worst = ly; drop = act & 12; }
                    if ((act & 16) && lc < worst) { worst = lc; drop = 16; }
                    if (drop) { nact = act & ~drop; score = worst; }
                }
            }
            W.act[f] = nact;
            changed |= (nact != act);
            if (W.hist) {
                if (restricted) {
                    W.flag[f] = act;                                   // (flag, wv, dx are dead until the next round)
                    const int grows = (nact | act) == nact;
                    W.wv[f] = (nact != act && !grows) ? score : 1.0;   // a drop has a negative score,
                    W.dx[f] = (nact != act && grows) ? ascore : 1.0;   // and so has an addition
                } else {
                    unsigned hsh = ((unsigned)f * 37u + (unsigned)nact + 1u) * 2654435761u;
                    hsh ^= hsh >> 15;
                    QR_ATOMIC_ADD(&W.hist[16 + (round & 1)], (int)(hsh * 2246822519u));
                    if (nact != act && ++W.hist[24 + f] >= kchg) W.hist[18] = 1;
                }
            }
        }
        const int any_changed = QR_ANY(changed);
        QR_PROF(7);
        if (!any_changed) { *ok = 1; ++round; break; }
        if (W.hist) {
            if (restricted) {
                QR_FOR(f, nf) {
                    const double* sv = W.wv[f] < 0.0 ? W.wv : W.dx;
                    const double sc = sv[f];
                    if (sc < 0.0) {
                        int keep = 1;
                        for (int gidx = 0; gidx < nf; ++gidx) {
                            const double o = sv[gidx];
                            if (o < sc || (o == sc && gidx < f)) keep = 0;
                        }
                        if (!keep) W.act[f] = W.flag[f];
                    }
                }
                QR_SYNC();
            } else {
                const int cur = W.hist[16 + (round & 1)];
                int hit = 0;
                for (int j = 0; j < 16; ++j) hit |= (j != (round & 15)) & (W.hist[j] == cur) & (j < round || round >= 16);
                QR_THREADS(t) {
                    if (t == 0) { W.hist[round & 15] = cur; W.hist[16 + ((round + 1) & 1)] = 0; }
                }
                restricted = hit | W.hist[18];
            }
        }
    }
    return round;
}

// ------------------------------------------------------------------------------------------
// Fallback linear algebra: blocked Cholesky with triangular solves against the 3x3 diagonal factors
// (backward stable; the explicit pivot inverses of qr_ldl_factor lose too many digits on the
// interior-point matrices, whose barrier terms span ten orders of magnitude).
// ------------------------------------------------------------------------------------------
// Cholesky factor of a 3x3 SPD block given by its lower part, as the reciprocal pivots and the
// off-diagonal entries (all a thread needs for its own forward substitution).
struct QrChol3 {
    double i00, i11, i22, l10, l20, l21;
};
QR_DEV QrChol3 qr_chol3(const double* D) {
    QrChol3 c;
    c.i00 = qr_rsqrt(qr_max(D[0], 1e-300));
    c.l10 = D[3] * c.i00;
    c.l20 = D[6] * c.i00;
    c.i11 = qr_rsqrt(qr_max(D[4] - c.l10 * c.l10, 1e-300));
    c.l21 = (D[7] - c.l20 * c.l10) * c.i11;
    c.i22 = qr_rsqrt(qr_max(D[8] - c.l20 * c.l20 - c.l21 * c.l21, 1e-300));
    return c;
}

// In-place blocked Cholesky K = L L' of the leading nb x nb blocks.  Off-diagonal blocks of K are
// overwritten with L; the diagonal blocks are left untouched and their inverse factors go to Dinv
// (that is all the solves need).  Non-positive pivots are clamped (the caller's verification or the
// final finiteness check catches a breakdown).
template <int NT>
QR_DEV void qr_blk_cholesky(QrQpWork& W, int nb) {  // interior-point fallback only
    double* K = W.K;
    for (int Kc = 0; Kc < nb; ++Kc) {
        // panel: every row of the block column solves against the (redundantly factorised) diagonal block
        QR_FOR(idx, 3 * (nb - Kc)) {
            const int I = Kc + idx / 3, a = idx % 3;
            const QrChol3 c = qr_chol3(K + qr_blk(Kc, Kc));
            if (I == Kc) {
                double* Di = W.Dinv + 9 * Kc;
                if (a == 0) {
                    Di[0] = c.i00; Di[1] = 0.0; Di[2] = 0.0;
                } else if (a == 1) {
                    Di[3] = -c.l10 * c.i00 * c.i11; Di[4] = c.i11; Di[5] = 0.0;
                } else {
                    const double inv10 = -c.l10 * c.i00 * c.i11;
                    Di[6] = -(c.l20 * c.i00 + c.l21 * inv10) * c.i22;
                    Di[7] = -c.l21 * c.i11 * c.i22;
                    Di[8] = c.i22;
                }
            } else {
                double* r = K + qr_blk(I, Kc) + 3 * a;
                const double x0 = r[0] * c.i00;
                const double x1 = (r[1] - x0 * c.l10) * c.i11;
                const double x2 = (r[2] - x0 * c.l20 - x1 * c.l21) * c.i22;
                r[0] = x0; r[1] = x1; r[2] = x2;
            }
        }
        QR_SYNC();
        // trailing update A_IJ -= L_IK L_JK'
        const int nrem = nb - Kc - 1;
        QR_FOR(idx, (nrem * (nrem + 1)) / 2) {
            const int code = W.tri[idx];
            const int I = (code >> 8) + Kc + 1, J = (code & 255) + Kc + 1;
            const int rowI = (I * (I + 1)) / 2;
            const double* li = K + 9 * (rowI + Kc);
            const double* lj = K + qr_blk(J, Kc);
            double* a = K + 9 * (rowI + J);
            double l[9], m[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) { l[e] = li[e]; m[e] = lj[e]; }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c2 = 0; c2 < 3; ++c2)
                    a[3 * r + c2] -= l[3 * r] * m[3 * c2] + l[3 * r + 1] * m[3 * c2 + 1] + l[3 * r + 2] * m[3 * c2 + 2];
        }
        QR_SYNC();
    }
}

// Solve (L L') out = W.wv on the leading nb blocks.  W.wv and W.yv are destroyed.
template <int NT>
QR_DEV void qr_blk_solve(QrQpWork& W, int nb, double* out) {
    const double* K = W.K;
    double* v = W.wv;
    double* y = W.yv;
    // forward: L y = v
    for (int Kc = 0; Kc < nb; ++Kc) {
        QR_FOR(idx, 3 * (nb - Kc)) {
            const int I = Kc + idx / 3, a = idx % 3;
            const double* Di = W.Dinv + 9 * Kc;
            const double b0 = v[3 * Kc], b1 = v[3 * Kc + 1], b2 = v[3 * Kc + 2];
            const double y0 = Di[0] * b0;
            const double y1 = Di[3] * b0 + Di[4] * b1;
            const double y2 = Di[6] * b0 + Di[7] * b1 + Di[8] * b2;
            if (I == Kc) {
                y[3 * Kc + a] = (a == 0 ? y0 : (a == 1 ? y1 : y2));
            } else {
                const double* r = K + qr_blk(I, Kc) + 3 * a;
                v[3 * I + a] -= r[0] * y0 + r[1] * y1 + r[2] * y2;
            }
        }
        QR_SYNC();
    }
    // backward: L' out = y, in place on y (a step reads block Kc of y and updates blocks J < Kc)
    for (int Kc = nb - 1; Kc >= 0; --Kc) {
        QR_FOR(idx, 3 * (Kc + 1)) {
            const int J = idx / 3, a = idx % 3;
            const double* Di = W.Dinv + 9 * Kc;
            const double b0 = y[3 * Kc], b1 = y[3 * Kc + 1], b2 = y[3 * Kc + 2];
            // x_K = Linv' b
            const double x0 = Di[0] * b0 + Di[3] * b1 + Di[6] * b2;
            const double x1 = Di[4] * b1 + Di[7] * b2;
            const double x2 = Di[8] * b2;
            if (J == Kc) {
                out[3 * Kc + a] = (a == 0 ? x0 : (a == 1 ? x1 : x2));
            } else {
                const double* blk = K + qr_blk(Kc, J);   // L_KJ, column a of it
                y[3 * J + a] -= blk[a] * x0 + blk[3 + a] * x1 + blk[6 + a] * x2;
            }
        }
        QR_SYNC();
    }
}

// ------------------------------------------------------------------------------------------
// Fallback: interior point.  Rare, so it is written for clarity; its vectors are in global scratch.
// ------------------------------------------------------------------------------------------
template <int NT>
QR_DEV double qr_red_max(const double* r, int cnt) {
    double m = 0.0;
    for (int i = 0; i < cnt; ++i) m = qr_max(m, r[i]);
    return m;
}
template <int NT>
QR_DEV double qr_red_sum(const double* r, int cnt) {
    double m = 0.0;
    for (int i = 0; i < cnt; ++i) m += r[i];
    return m;
}
template <int NT>
QR_DEV double qr_red_min1(const double* r, int cnt) {
    double m = 1.0;
    for (int i = 0; i < cnt; ++i) m = qr_min(m, r[i]);
    return m;
}

// Copy Hs -> K and add the barrier blocks A'DA (D = lam/s) on the diagonal.
template <int NT>
QR_DEV void qr_build_kkt(QrQpWork& W) {
    const int nf = W.nf, ntri = (nf * (nf + 1)) / 2;
    QR_FOR(idx, 9 * ntri) W.K[idx] = W.Hs[idx];
    QR_SYNC();
    QR_FOR(f, nf) {
        const double* s = W.s + 5 * f;
        const double* l = W.lam + 5 * f;
        const double d0 = l[0] / s[0], d1 = l[1] / s[1], d2 = l[2] / s[2], d3 = l[3] / s[3], d4 = l[4] / s[4];
        const double m = W.mu_;
        double* D = W.K + qr_blk(f, f);
        const double xx = m * m * (d0 + d1), yy = m * m * (d2 + d3);
        const double xz = m * (d0 - d1), yz = m * (d2 - d3);
        D[0] += xx; D[4] += yy; D[8] += d0 + d1 + d2 + d3 + d4;
        D[2] += xz; D[6] += xz; D[5] += yz; D[7] += yz;
    }
    QR_SYNC();
}

// Interior-point phase.  On exit W.x, W.s, W.lam hold the final iterate.  Returns iterations used;
// *converged tells whether the tolerance was met.
template <int NT>
QR_DEV int qr_ipm(QrQpWork& W, const qr_qp_options& opt, double tol, int* converged QR_PROF_ARG) {
    const int nf = W.nf, n = 3 * nf, m = 5 * nf;
    const double mu_ = W.mu_;
    double* red = W.red;
    // ---- starting point: strictly interior, fz = ub/4
    QR_FOR(f, nf) {
        W.x[3 * f] = 0.0; W.x[3 * f + 1] = 0.0; W.x[3 * f + 2] = 0.25 * W.ubz[f];
    }
    QR_SYNC();
    QR_FOR(i, n) W.q[i] = qr_sym_matvec_row(W.Hs, W.x, nf, i) + W.g[i];
    QR_FOR(f, nf) qr_foot_constraints(mu_, W.ubz[f], W.x + 3 * f, W.s + 5 * f);
    QR_FOR(i, n) red[nf + i] = fabs(W.g[i]);
    QR_SYNC();
    QR_FOR(f, nf) {
        red[f] = qr_max(fabs(W.q[3 * f]), qr_max(fabs(W.q[3 * f + 1]), fabs(W.q[3 * f + 2])));
    }
    QR_SYNC();
    const double gscale = qr_max(1.0, qr_red_max<NT>(red + nf, n));
    {
        const double lam0 = qr_red_max<NT>(red, nf) + 1e-3;
        QR_SYNC();
        QR_FOR(f, nf) {
            const double* s = W.s + 5 * f;
            red[f] = s[0] + s[1] + s[2] + s[3] + s[4];
        }
        QR_SYNC();
        const double mu0 = lam0 * qr_red_sum<NT>(red, nf) / (double)m;
        QR_SYNC();
        QR_FOR(c, m) W.lam[c] = mu0 / W.s[c];
        QR_SYNC();
    }
    *converged = 0;
    int it = 0;
    for (; it < opt.max_ipm_iter; ++it) {
        // residuals (q, s are current)
        QR_FOR(f, nf) {
            double at[3];
            qr_foot_At(mu_, W.lam + 5 * f, at);
            double rmax = 0.0, gap = 0.0;
            for (int a = 0; a < 3; ++a) {
                const double r = W.q[3 * f + a] - at[a];
                W.rd[3 * f + a] = r;
                rmax = qr_max(rmax, fabs(r));
            }
            for (int c = 0; c < 5; ++c) gap += W.s[5 * f + c] * W.lam[5 * f + c];
            red[f] = rmax;
            red[nf + f] = gap;
        }
        QR_SYNC();
        const double rdmax = qr_red_max<NT>(red, nf);
        const double gap = qr_red_sum<NT>(red + nf, nf) / (double)m;
        QR_SYNC();
        if (rdmax < tol * gscale && gap < tol) { *converged = 1; break; }

        qr_build_kkt<NT>(W);
        qr_blk_cholesky<NT>(W, nf);
        // predictor: K dxa = -(Hx + g)
        QR_FOR(i, n) W.wv[i] = -W.q[i];
        QR_SYNC();
        qr_blk_solve<NT>(W, nf, W.dxa);
        QR_FOR(f, nf) {
            const double* dxf = W.dxa + 3 * f;
            double ds[5];
            ds[0] = mu_ * dxf[0] + dxf[2]; ds[1] = -mu_ * dxf[0] + dxf[2];
            ds[2] = mu_ * dxf[1] + dxf[2]; ds[3] = -mu_ * dxf[1] + dxf[2];
            ds[4] = -dxf[2];
            double ap = 1e30, ad = 1e30;
            for (int c = 0; c < 5; ++c) {
                const double s = W.s[5 * f + c], l = W.lam[5 * f + c];
                const double dl = -l - (l / s) * ds[c];
                W.dsa[5 * f + c] = ds[c];
                W.dla[5 * f + c] = dl;
                if (ds[c] < 0.0) ap = qr_min(ap, -s / ds[c]);
                if (dl < 0.0) ad = qr_min(ad, -l / dl);
            }
            red[f] = ap;
            red[nf + f] = ad;
        }
        QR_SYNC();
        const double apa = qr_red_min1<NT>(red, nf);
        const double ada = qr_red_min1<NT>(red + nf, nf);
        QR_SYNC();
        QR_FOR(f, nf) {
            double acc = 0.0;
            for (int c = 0; c < 5; ++c)
                acc += (W.s[5 * f + c] + apa * W.dsa[5 * f + c]) * (W.lam[5 * f + c] + ada * W.dla[5 * f + c]);
            red[f] = acc;
        }
        QR_SYNC();
        const double mu_aff = qr_red_sum<NT>(red, nf) / (double)m;
        QR_SYNC();
        double sigma = mu_aff / gap;
        sigma = sigma * sigma * sigma;
        // corrector right-hand side
        QR_FOR(f, nf) {
            double t[5];
            for (int c = 0; c < 5; ++c) {
                const double rc = W.s[5 * f + c] * W.lam[5 * f + c] + W.dsa[5 * f + c] * W.dla[5 * f + c] - sigma * gap;
                W.rc[5 * f + c] = rc;
                t[c] = rc / W.s[5 * f + c];
            }
            double at[3];
            qr_foot_At(mu_, t, at);
            for (int a = 0; a < 3; ++a) W.wv[3 * f + a] = -W.rd[3 * f + a] - at[a];
        }
        QR_SYNC();
        qr_blk_solve<NT>(W, nf, W.dx);
        QR_FOR(f, nf) {
            const double* dxf = W.dx + 3 * f;
            double ds[5];
            ds[0] = mu_ * dxf[0] + dxf[2]; ds[1] = -mu_ * dxf[0] + dxf[2];
            ds[2] = mu_ * dxf[1] + dxf[2]; ds[3] = -mu_ * dxf[1] + dxf[2];
            ds[4] = -dxf[2];
            double ap = 1e30, ad = 1e30;
            for (int c = 0; c < 5; ++c) {
                const double s = W.s[5 * f + c], l = W.lam[5 * f + c];
                const double dl = -(W.rc[5 * f + c] + l * ds[c]) / s;
                W.dl[5 * f + c] = dl;
                if (ds[c] < 0.0) ap = qr_min(ap, -s / ds[c]);
                if (dl < 0.0) ad = qr_min(ad, -l / dl);
            }
            red[f] = ap;
            red[nf + f] = ad;
        }
        QR_SYNC();
        double ap = 1e30, ad = 1e30;
        for (int f = 0; f < nf; ++f) { ap = qr_min(ap, red[f]); ad = qr_min(ad, red[nf + f]); }
        QR_SYNC();
        ap = ap < 1.0 ? 0.995 * ap : 1.0;
        ad = ad < 1.0 ? 0.995 * ad : 1.0;
        QR_FOR(i, n) W.x[i] += ap * W.dx[i];
        QR_FOR(c, m) W.lam[c] += ad * W.dl[c];
        QR_SYNC();
        QR_FOR(i, n) W.q[i] = qr_sym_matvec_row(W.Hs, W.x, nf, i) + W.g[i];
        QR_FOR(f, nf) {
            qr_foot_constraints(mu_, W.ubz[f], W.x + 3 * f, W.s + 5 * f);
            for (int c = 0; c < 5; ++c) W.s[5 * f + c] = qr_max(W.s[5 * f + c], 1e-30);
        }
        QR_SYNC();
    }
    return it;
}

// Optional coarse problem whose solution predicts the active set of the full-size one (built by the caller, see
// qr_mpc_build_coarse in mpc_problem.h): ng tied foot-steps, grp[f] = tied foot-step of foot-step f.
struct QrCoarse {
    int ng;
    int max_rounds;   // rounds spent on the coarse problem at most
    double *Hs, *g, *ubz;
    const int* grp;
};

// Full solve on a prepared workspace (Hs, g, ubz, mu_ set).  Result in W.xn (verified) or W.x.
// Returns the per-instance status code of qr_gpu.h.
//
// Stages: [-1, 0: the iteration on the coarse problems, whose active rows every foot-step of the next level inherits] -> 1: the iteration on
// the full-size problem (cold, or from the inherited guess) -> on failure 2: interior point to identify the active
// set (tightening the tolerance once if the verification still does not settle) and 3: the same verification from
// its guess.  Written as one loop around a single qr_active_set call site so that the (force-inlined) iteration is
// instantiated once.
// L2 (compile time): the workspace class carries a second coarse level (C points at two levels, else at one).  Kept out
// of the instantiations that never use it: two levels make C an indexed array, which lives on the thread's stack, and
// the h = 10 classes lost 4 % to that alone (B200, A1 trot 65536: 4.12 -> 3.96 M QP/s).
template <int NT, bool L2>
QR_DEV int qr_qp_solve(QrQpWork& W, const qr_qp_options& opt, int* ipm_iters, int* as_rounds,
                       const double** result, const QrCoarse* C QR_PROF_ARG) {
    int conv = 0, ok = 0;
    *ipm_iters = 0; *as_rounds = 0;
    *result = W.xn;
    if (W.nf == 0) return 0;
    const int nf_full = W.nf;
    double* const hs_full = W.Hs;
    double* const g_full = W.g;
    double* const ubz_full = W.ubz;
    // stages -1 / 0: the coarse levels C[1] (coarsest, optional) and C[0]; every level starts from the active rows
    // inherited from the level before it
    int stage = 1;
    if (C && C[0].ng > 0) stage = (L2 && C[1].ng > 0) ? -1 : 0;
    int mode = 1, attempt = 0;
    double tol = opt.ipm_tol;
    for (;;) {
        int maxr = opt.max_as_rounds;
        if (stage <= 0) {
            const QrCoarse& L = C[L2 ? -stage : 0];
            W.nf = L.ng; W.Hs = L.Hs; W.g = L.g; W.ubz = L.ubz; maxr = L.max_rounds;
        } else if (stage == 3) { mode = 0; maxr = opt.max_polish_rounds; }
        const int rounds = qr_active_set<NT>(W, opt, &ok, mode, maxr QR_PROF_PASS);
        if (stage <= 0) {
            // every foot-step of the next finer level inherits its tied foot-step's active rows (through W.flag:
            // rewritten in place)
            const int* grp = C[L2 ? -stage : 0].grp;
            const int n_fine = (L2 && stage == -1) ? C[0].ng : nf_full;
            W.nf = nf_full; W.Hs = hs_full; W.g = g_full; W.ubz = ubz_full;
            QR_FOR(f, n_fine) W.flag[f] = W.act[grp[f]];
            QR_SYNC();
            QR_FOR(f, n_fine) W.act[f] = W.flag[f];
            QR_SYNC();
            ++stage; mode = 2;
            continue;
        }
        *as_rounds += rounds;
        if (ok) return 0;
        if (attempt == 2) break;
        *ipm_iters += qr_ipm<NT>(W, opt, tol, &conv QR_PROF_PASS);
        tol *= 1e-2;
        ++attempt;
        stage = 3;
    }
    *result = W.x;
    return 1;
}
