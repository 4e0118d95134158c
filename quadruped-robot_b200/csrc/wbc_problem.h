// wbc_problem.h -- one whole-body-control tick of one robot on one thread team (a warp), float64.
//
// Replaces, for a batch of independent robots, what qrWbcLocomotionController<float>::Run computes on a
// recomputing tick (/root/reference/quadruped/src/controllers/wbc/qr_wbc_locomotion_controller.cpp:108-219):
//   UpdateModel        -> FloatingBaseModel::forwardKinematics / biasAccelerations / contactJacobians /
//                         compositeInertias / massMatrix / generalizedGravityForce / generalizedCoriolisForce
//                         (src/dynamics/floating_base_model.cpp:469-806) and GetModelRes (A^-1)
//   ContactTaskUpdate  -> task_set/*.cpp, qr_single_contact.cpp
//   FindConfiguration  -> qr_multitask_projection.cpp:38-106 (SVD pseudo-inverses, threshold 1e-3)
//   MakeTorque         -> qr_wholebody_impulse_ctrl.cpp:62-299 (weighted pseudo-inverses, threshold 1e-4,
//                         QP, inverse dynamics)
// Differences in HOW, not in WHAT:
//   * everything is evaluated in float64 (the reference: float32 + a float64 QP); the float32 rounding noise of
//     the reference (~1e-6 relative on torques) is larger than the difference to this evaluation;
//   * pseudo-inverses use a one-sided Jacobi SVD whose disjoint column pairs rotate in parallel across lanes;
//   * the WBIC QP  min 0.1|da|^2 + |df|^2  s.t. A6 da - (Jc')6 df = -r6,  Uf (f_des + df) >= b  is reduced to the
//     contact forces f = f_des + df (da is an affine function of df through the 6 equality rows), which leaves
//     a dense friction-pyramid QP in 3*nc <= 12 variables with exactly the constraint structure of the MPC QP --
//     it is solved by the same block active-set solver (qp_solver.h) instead of Goldfarb-Idnani (QuadProg++);
//     the optimum of this strictly convex QP is unique, so both give the same forces.
#pragma once

#include "qp_solver.h"
#include "wbc_model.h"

struct QrWbcArgs {
    const QrWbcModelDev* model;   // device constants
    qr_qp_options opt;
    int batch;
    const float* state;           // [B][37]
    const float* cmd;             // [B][66]
    const int32_t* contact;       // [B][4]
    float *tau32, *fr32, *qdes32, *qddes32;
    double *tau64, *fr64, *qdes64, *qddes64, *dbg;
    int32_t* status;
};

struct QrWbcWork {
    double *st, *cmd;
    double *Xup, *Xur, *Xa, *IC, *T1, *T2;       // 13 / 12 / 13 / 13 / 4 / 4 blocks of 36 (T1, T2: one per leg, reused level by level)
    double *v, *vr, *cj, *cr, *avp, *avr, *ag, *agr, *fvp, *fvr;   // 6-vectors per body
    double *sq, *cq;                              // sin / cos of the joint angles
    double *H, *Ainv, *G, *Cq, *rowbuf, *colbuf;
    double *Jc, *Jcd, *pF, *vF;                   // feet: 4 x (3x18), 4x3, 4x3, 4x3
    double *Jt, *xdd, *jdq, *perr, *dvel;         // tasks (up to 6)
    double *JC, *JCd, *fdes;                      // stacked contacts
    double *N, *M1;                            // 18x18 temporaries
    double *Jpre, *Jbar, *Lam, *LamInv, *tmpA;    // weighted-inverse temporaries
    double *svdB, *svdV, *svdS;                   // Jacobi workspace
    double *qdd, *dq, *qdot, *vec, *tot, *P6, *a0, *A6;
    int* ints;                                    // [0] ntask, [1] nc, [2..5] stance legs, [6..11] task leg (-1 body)
    QrQpWork Q;
};

// Shared-memory plan (doubles).  The spatial-algebra temporaries of the dynamics phase (transforms, composite
// inertias, body velocities...) are dead once H, C, G and the foot Jacobians exist; the task / projector / QP
// workspace of the later phases is laid over them.  This is what sets the number of robots in flight per SM.
QR_HD size_t qr_wbc_dyn_doubles() {
    return 36 * (13 + 12 + 13 + 13 + 4 + 4) + 6 * (13 + 12 + 12 + 12 + 13 + 12 + 13 + 12 + 13 + 12) + 24;
}
QR_HD size_t qr_wbc_late_doubles() {
    size_t d = 6 * 54 + 6 * 12;                       // tasks
    d += 12 * 18 + 12 + 12;                           // stacked contacts
    d += 324 * 2;                                     // N, M1
    d += 54 + 216 + 144 + 144;                        // JtPre (3 x 18), weighted-inverse temporaries (tmpA lives in M1)
    d += 216 + 144 + 12;                              // Jacobi workspace
    d += 18 * 5 + 72 + 6 + 36;
    // QP workspace for up to 4 contact blocks (its interior-point fallback vectors reuse the Jacobi workspace)
    d += 90 + 90 + 36 + 36 + 12 * 6 + 4;
    return d;
}
QR_HD size_t qr_wbc_smem_doubles() {
    size_t d = 37 + 66;
    d += 324 * 2 + 18 * 2 + 36 + 18;                  // H, Ainv, G, Cq, rowbuf, colbuf
    d += 4 * 54 + 12 * 3;                             // foot Jacobians, Jdot qdot, positions, velocities
    const size_t a = qr_wbc_dyn_doubles(), b = qr_wbc_late_doubles();
    return d + (a > b ? a : b) + 8;
}
QR_HD size_t qr_wbc_smem_bytes() {   // per robot, a multiple of 16 (several robots' workspaces sit back to back in one CTA)
    return (qr_wbc_smem_doubles() * sizeof(double) + (16 + 4 * 3 + 5 + 12 + 16) * sizeof(int) + 32 + 15) & ~(size_t)15;
}

QR_DEV void qr_wbc_carve(QrWbcWork& W, unsigned char* base) {
    double* d = reinterpret_cast<double*>(base);
    auto take = [&](size_t n) { double* p = d; d += n; return p; };
    W.st = take(37); W.cmd = take(66);
    W.H = take(324); W.Ainv = take(324); W.G = take(18); W.Cq = take(18); W.rowbuf = take(36); W.colbuf = take(18);
    W.Jc = take(4 * 54); W.Jcd = take(12); W.pF = take(12); W.vF = take(12);
    double* const overlay = d;
    // ---- dynamics phase
    W.Xup = take(36 * 13); W.Xur = take(36 * 12); W.Xa = take(36 * 13); W.IC = take(36 * 13); W.T1 = take(36 * 4); W.T2 = take(36 * 4);
    W.v = take(6 * 13); W.vr = take(6 * 12); W.cj = take(6 * 12); W.cr = take(6 * 12); W.avp = take(6 * 13); W.avr = take(6 * 12);
    W.ag = take(6 * 13); W.agr = take(6 * 12); W.fvp = take(6 * 13); W.fvr = take(6 * 12);
    W.sq = take(12); W.cq = take(12);
    // ---- later phases, over the same memory
    d = overlay;
    W.Jt = take(6 * 54); W.xdd = take(18); W.jdq = take(18); W.perr = take(18); W.dvel = take(18);
    W.JC = take(216); W.JCd = take(12); W.fdes = take(12);
    W.N = take(324); W.M1 = take(324);
    W.Jpre = take(54); W.Jbar = take(216); W.Lam = take(144); W.LamInv = take(144);
    W.svdB = take(216); W.svdV = take(144); W.svdS = take(12);
    W.tmpA = W.M1;   // tm_weighted_inverse holds A^-1 J' across its tm_pinv call (which owns the Jacobi workspace); M1 is
                     // only the scratch of tm_project_out / of the mass-matrix inverse, before and after it
    W.qdd = take(18); W.dq = take(18); W.qdot = take(18); W.vec = take(18); W.tot = take(18); W.P6 = take(72); W.a0 = take(6); W.A6 = take(36);
    QrQpWork& Q = W.Q;
    Q.k8 = 0;
    Q.Hs = take(90); Q.K = take(90); Q.Dinv = take(36); Q.zv = take(36);
    Q.ps = take(12); Q.g = take(12); Q.xn = take(12); Q.q = take(12); Q.wv = take(12); Q.dx = take(12);
    Q.ubz = take(4);
    {   // the interior-point fallback vectors of the QP reuse the Jacobi workspace (all pseudo-inverses are done by then)
        double* f = W.svdB;
        Q.x = f; f += 12; Q.dxa = f; f += 12; Q.rd = f; f += 12; Q.yv = f; f += 12;
        Q.s = f; f += 20; Q.lam = f; f += 20; Q.dsa = f; f += 20; Q.dla = f; f += 20; Q.rc = f; f += 20; Q.dl = f; f += 20;
        Q.red = f;   // 16 <= 216 - 184
    }
    {
        const size_t a = qr_wbc_dyn_doubles(), b = qr_wbc_late_doubles();
        d = overlay + (a > b ? a : b);
    }
    d += 8;
    int* ip = reinterpret_cast<int*>(d);
    W.ints = ip; ip += 16;
    Q.act = ip; ip += 4; Q.flag = ip; ip += 4; Q.vert = ip; ip += 4;
    Q.foff = ip; ip += 5; Q.rfoot = ip; ip += 12;
    Q.tri = reinterpret_cast<unsigned short*>(ip);
    Q.hist = nullptr;   // at most four contact blocks: no cycle detection
}

// ---- small team-parallel dense kernels (row-major, contiguous).  Every helper ends with a barrier. ----
// Inner products are kept rolled: the helpers are inlined at a dozen call sites with constant sizes, and the kernel is
// bound by instruction fetch (fully unrolled, the compiler's choice: 7.6 M robots/s; unroll 4: 8.4 M; rolled: 8.5 M).  The six-term loops
// of the spatial algebra are the opposite case: rolled 7.0 M, unroll 2 7.4 M against 8.5 M fully unrolled -- short bodies
// whose loop overhead is more code than the unrolled products.
#ifndef QR_WBC_UNROLL
#define QR_WBC_UNROLL _Pragma("unroll 1")
#endif
template <int NT> QR_DEV void tm_mul(double* C, const double* A, const double* B, int m, int k, int n) {
    QR_FOR_2D(idx, i, j, m, n) { double s = 0.0; QR_WBC_UNROLL for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j]; C[idx] = s; }
    QR_SYNC();
}
template <int NT> QR_DEV void tm_mul_nt(double* C, const double* A, const double* B, int m, int k, int n) {   // C = A B', B is n x k
    QR_FOR_2D(idx, i, j, m, n) { double s = 0.0; QR_WBC_UNROLL for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l]; C[idx] = s; }
    QR_SYNC();
}
// C = I - A B  (A: m x k, B: k x m)
template <int NT> QR_DEV void tm_eye_minus_mul(double* C, const double* A, const double* B, int m, int k) {
    QR_FOR_2D(idx, i, j, m, m) { double s = (i == j) ? 1.0 : 0.0; QR_WBC_UNROLL for (int l = 0; l < k; ++l) s -= A[i * k + l] * B[l * m + j]; C[idx] = s; }
    QR_SYNC();
}
template <int NT> QR_DEV void tm_copy(double* D, const double* S, int n) { QR_FOR(i, n) D[i] = S[i]; QR_SYNC(); }
// N <- N (I - Jbar Jpre) for an 18 x 18 projector and a three-row task (Jbar 18 x 3, Jpre 3 x 18), evaluated as the
// rank-3 update N - (N Jbar) Jpre: 1.9 k multiply-adds and two phases instead of the 6.8 k and three of forming
// I - Jbar Jpre and multiplying by it (the same matrix up to float64 rounding).  T: 54 doubles of scratch.
template <int NT> QR_DEV void tm_project_out(double* N, const double* Jbar, const double* Jpre, double* T) {
    QR_FOR_2D(idx, i, c, 18, 3) { double s = 0.0; QR_WBC_UNROLL for (int l = 0; l < 18; ++l) s += N[18 * i + l] * Jbar[3 * l + c]; T[idx] = s; }
    QR_SYNC();
    QR_FOR_2D(idx, i, j, 18, 18) N[idx] -= T[3 * i] * Jpre[j] + T[3 * i + 1] * Jpre[18 + j] + T[3 * i + 2] * Jpre[36 + j];
    QR_SYNC();
}

// Pseudo-inverse of J (m x n, m <= 12, n <= 18) with singular values <= thr dropped
// (pseudoInverse, include/quadruped/utils/qr_algebra.h:119-140).  One-sided Jacobi on the columns of
// B = J' (n x m): the m/2 disjoint column pairs of a round-robin round rotate on different lanes.
// out = pinv(J), n x m.
//
// Fast path.  When every singular value is safely above the threshold nothing is dropped and
// pinv(J) = J' (J J')^-1 (or J^-1 for the symmetric positive definite J of WeightedInverse): the Gram matrix is
// inverted by Gauss-Jordan and the bound sigma_min^2 >= 1 / ||(J J')^-1||_F certifies the rank decision with a
// 10x margin.  Anything closer to the threshold (or singular, or non-finite) takes the Jacobi SVD below, so the
// thresholding semantics of the reference are kept exactly where they matter.  The SVD was 63 % of the kernel.
template <int NT> QR_DEV void tm_inverse_spd(QrWbcWork& W, const double* src0, double* A, double* tmp, int n);
template <int NT>
QR_DEV void tm_pinv(QrWbcWork& W, double* out, const double* J, int m, int n, double thr, int sym_psd = 0) {
    double* B = W.svdB;   // n x m
    double* V = W.svdV;   // m x m
    if (m == 1 && n == 1) {   // the reference's scalar special case
        QR_THREADS(t) { if (t == 0) out[0] = J[0] > thr ? 1.0 / J[0] : 0.0; }
        QR_SYNC();
        return;
    }
    {
        double* Gi = V;   // m x m
        if (sym_psd) {
            QR_FOR(idx, m * m) Gi[idx] = J[idx];
        } else {
            QR_FOR_2D(idx, i, j, m, m) {
                double a = 0.0;
                for (int l = 0; l < n; ++l) a += J[i * n + l] * J[j * n + l];
                Gi[idx] = a;
            }
        }
        QR_SYNC();
        tm_inverse_spd<NT>(W, Gi, Gi, B, m);   // svdB is free until the Jacobi fallback below
        int bad = 0;
        double fro = 0.0;
        QR_THREADS(t) {
            if (t == 0) {
                for (int idx = 0; idx < m * m; ++idx) fro += Gi[idx] * Gi[idx];
                const double lam_min = 1.0 / sqrt(fro);           // <= smallest eigenvalue of the inverted matrix
                const double need = sym_psd ? 10.0 * thr : 100.0 * thr * thr;
                bad = !(lam_min > need) || !(fro < 1e300) || !(fro > 0.0);
            }
        }
        if (!QR_ANY(bad)) {
            if (sym_psd) {
                QR_FOR(idx, m * m) out[idx] = Gi[idx];
            } else {
                QR_FOR_2D(idx, i, l, n, m) {
                    double a = 0.0;
                    for (int j = 0; j < m; ++j) a += J[j * n + i] * Gi[j * m + l];
                    out[idx] = a;
                }
            }
            QR_SYNC();
            return;
        }
    }
    QR_FOR(idx, n * m) { const int i = idx / m, j = idx - i * m; B[idx] = J[j * n + i]; }
    QR_FOR(idx, m * m) V[idx] = (idx / m == idx % m) ? 1.0 : 0.0;
    QR_SYNC();
    const int mp = m + (m & 1);           // players of the tournament (a bye when m is odd)
    for (int sweep = 0; sweep < 30; ++sweep) {
        int rotated = 0;
        for (int round = 0; round < mp - 1; ++round) {
            QR_FOR(k, mp / 2) {
                int p, q;
                if (k == 0) { p = round; q = mp - 1; }
                else { p = (round + k) % (mp - 1); q = (round - k + (mp - 1)) % (mp - 1); }
                if (p > q) { const int tmp = p; p = q; q = tmp; }
                if (q < m) {
                    double alpha = 0.0, beta = 0.0, gamma = 0.0;
                    for (int i = 0; i < n; ++i) { const double bp = B[i * m + p], bq = B[i * m + q]; alpha += bp * bp; beta += bq * bq; gamma += bp * bq; }
                    if (fabs(gamma) > 1e-15 * sqrt(alpha * beta) && gamma != 0.0) {
                        rotated = 1;
                        const double zeta = (beta - alpha) / (2.0 * gamma);
                        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
                        for (int i = 0; i < n; ++i) { const double bp = B[i * m + p], bq = B[i * m + q]; B[i * m + p] = c * bp - s * bq; B[i * m + q] = s * bp + c * bq; }
                        for (int i = 0; i < m; ++i) { const double vp = V[i * m + p], vq = V[i * m + q]; V[i * m + p] = c * vp - s * vq; V[i * m + q] = s * vp + c * vq; }
                    }
                }
            }
            QR_SYNC();
        }
        if (!QR_ANY(rotated)) break;
    }
    QR_FOR(j, m) {
        double s2 = 0.0;
        for (int i = 0; i < n; ++i) s2 += B[i * m + j] * B[i * m + j];
        W.svdS[j] = (sqrt(s2) > thr) ? 1.0 / s2 : 0.0;
    }
    QR_SYNC();
    QR_FOR(idx, n * m) {
        const int i = idx / m, l = idx - i * m;
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += B[i * m + j] * V[l * m + j] * W.svdS[j];
        out[idx] = s;
    }
    QR_SYNC();
}

// In-place inverse of a symmetric positive definite n x n matrix by Gauss-Jordan without pivoting
// (GetModelRes: Ainv = A.inverse(), qr_wholebody_impulse_ctrl.cpp:50-58).
template <int NT>
QR_DEV void tm_inverse_spd(QrWbcWork& W, const double* src0, double* A, double* tmp, int n) {
    if (n == 3) {   // task-sized blocks: adjugate, one phase
        double o[9];
        QR_THREADS(t) {
            if (t == 0) {
                qr_inv3_sym(src0[0], src0[3], src0[6], src0[4], src0[7], src0[8], o);
                for (int e = 0; e < 9; ++e) A[e] = o[e];
            }
        }
        QR_SYNC();
        return;
    }
    // Gauss-Jordan without pivoting (symmetric positive definite input), ONE barrier per pivot: step k reads the matrix
    // of step k-1 from one buffer and writes its own into the other, so the pivot row and column need no staging
    // phase of their own (the in-place form needed two barriers per pivot; the 18 + 6..12 + 6 pivots of a tick were
    // 19 % of the kernel).  Every entry is the same expression as before: results are bit-identical.
    const double* s = src0;
    double* d = (n & 1) ? A : tmp;   // after n steps the result sits in A ...
    if (s == d) d = tmp;              // ... unless an in-place call with odd n starts on A: copied back below
    for (int k = 0; k < n; ++k) {
        const double akk = s[k * n + k];
        const double piv = akk > 1e-300 ? qr_rcp_pos(akk) : 1.0 / akk;   // positive pivots: Newton reciprocal (once per thread and step)
        QR_FOR_2D(idx, i, c, n, n) {
            const double rowc = (c == k ? 1.0 : s[k * n + c]) * piv;
            if (i == k) d[idx] = rowc;
            else d[idx] = (c == k ? 0.0 : s[idx]) - s[i * n + k] * rowc;
        }
        QR_SYNC();
        s = d;
        d = (d == A) ? tmp : A;
    }
    if (s != A) {
        QR_FOR(idx, n * n) A[idx] = s[idx];
        QR_SYNC();
    }
}

// WeightedInverse (qr_wholebody_impulse_ctrl.cpp:291-299): Jbar = Ainv J' pinv(J Ainv J', 1e-4), n = 18.
template <int NT>
QR_DEV void tm_weighted_inverse(QrWbcWork& W, double* Jbar, const double* J, int m) {
    tm_mul_nt<NT>(W.tmpA, W.Ainv, J, 18, 18, m);          // 18 x m
    tm_mul<NT>(W.Lam, J, W.tmpA, m, 18, m);               // m x m
    tm_pinv<NT>(W, W.LamInv, W.Lam, m, m, 0.0001, 1);
    tm_mul<NT>(Jbar, W.tmpA, W.LamInv, 18, m, m);
}

// spatial helpers on 6-vectors (include/quadruped/dynamics/spatial.hpp)
QR_DEV void qr_motion_cross_axis(const double* a, int axis, double qd, double* o) {   // a x_m (e_axis qd)
    double b[6] = {0, 0, 0, 0, 0, 0};
    b[axis] = qd;
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    o[3] = a[1] * b[5] - a[2] * b[4] + a[4] * b[2] - a[5] * b[1];
    o[4] = a[2] * b[3] - a[0] * b[5] - a[3] * b[2] + a[5] * b[0];
    o[5] = a[0] * b[4] - a[1] * b[3] + a[3] * b[1] - a[4] * b[0];
}
QR_DEV void qr_force_cross(const double* a, const double* b, double* o) {
    o[0] = b[2] * a[1] - b[1] * a[2] - b[4] * a[5] + b[5] * a[4];
    o[1] = b[0] * a[2] - b[2] * a[0] + b[3] * a[5] - b[5] * a[3];
    o[2] = b[1] * a[0] - b[0] * a[1] - b[3] * a[4] + b[4] * a[3];
    o[3] = b[5] * a[1] - b[4] * a[2]; o[4] = b[3] * a[2] - b[5] * a[0]; o[5] = b[4] * a[0] - b[3] * a[1];
}
QR_DEV void qr_quat_to_rot(const double* q, double* R) {   // quaternionToRotationMatrix: world -> body (utils/qr_se3.h:186-203)
    const double e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3];
    R[0] = 1 - 2 * (e2 * e2 + e3 * e3); R[3] = 2 * (e1 * e2 - e0 * e3); R[6] = 2 * (e1 * e3 + e0 * e2);
    R[1] = 2 * (e1 * e2 + e0 * e3); R[4] = 1 - 2 * (e1 * e1 + e3 * e3); R[7] = 2 * (e2 * e3 - e0 * e1);
    R[2] = 2 * (e1 * e3 - e0 * e2); R[5] = 2 * (e2 * e3 + e0 * e1); R[8] = 1 - 2 * (e1 * e1 + e2 * e2);
}
QR_DEV void qr_sxform(const double* R, const double* r, double* X) {   // [R 0; -R[r]x R]
    for (int i = 0; i < 36; ++i) X[i] = 0.0;
    const double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            X[6 * i + j] = R[3 * i + j];
            X[6 * (3 + i) + 3 + j] = R[3 * i + j];
            X[6 * (3 + i) + j] = -(R[3 * i] * S[j] + R[3 * i + 1] * S[3 + j] + R[3 * i + 2] * S[6 + j]);
        }
}
QR_DEV int qr_axis_of(int joint) { return (joint % 3 == 0) ? 0 : 1; }   // abad: X, hip / knee: Y

// ------------------------------------------------------------------------------------------------
template <int NT>
QR_DEV void qr_wbc_dynamics(const QrWbcModelDev& M, QrWbcWork& W) {
    const double* quat = W.st; const double* pos = W.st + 4; const double* bv = W.st + 7;
    const double* q = W.st + 13; const double* qd = W.st + 25;
    // ---- forwardKinematics (floating_base_model.cpp:469-524)
#if defined(QR_ON_DEVICE)
    QR_FOR(j, 12) sincos(q[j], &W.sq[j], &W.cq[j]);   // one range reduction for both (and half the code of sin + cos)
#else
    QR_FOR(j, 12) { W.sq[j] = sin(q[j]); W.cq[j] = cos(q[j]); }
#endif
    QR_THREADS(t) {
        if (t == 0) {
            double R[9];
            qr_quat_to_rot(quat, R);
            qr_sxform(R, pos, W.Xup);
            for (int i = 0; i < 6; ++i) { W.v[i] = bv[i]; W.avp[i] = 0.0; }
        }
    }
    QR_SYNC();
    QR_FOR(idx, 24 * 36) {   // Xup = XJ * Xtree, Xuprot = XJ * Xrot with XJ = blkdiag(Rq, Rq)
        const int body = idx / 36, e = idx - 36 * body, r = e / 6, c = e - 6 * r;
        const int j = body % 12, rr = r % 3, r0 = r - rr;
        const double* X = body < 12 ? M.Xtree[j] : M.Xrot[j];
        const double s = W.sq[j], cs = W.cq[j];
        double R0, R1, R2;   // row rr of the passive rotation about the joint axis
        if (qr_axis_of(j) == 0) { R0 = rr == 0 ? 1.0 : 0.0; R1 = rr == 1 ? cs : (rr == 2 ? -s : 0.0); R2 = rr == 1 ? s : (rr == 2 ? cs : 0.0); }
        else { R0 = rr == 0 ? cs : (rr == 2 ? s : 0.0); R1 = rr == 1 ? 1.0 : 0.0; R2 = rr == 0 ? -s : (rr == 2 ? cs : 0.0); }
        const double val = R0 * X[6 * r0 + c] + R1 * X[6 * (r0 + 1) + c] + R2 * X[6 * (r0 + 2) + c];
        if (body < 12) W.Xup[36 * (1 + j) + e] = val; else W.Xur[36 * j + e] = val;
    }
    QR_SYNC();
    for (int lvl = 0; lvl < 3; ++lvl) {   // velocities down the four chains
        QR_FOR(idx, 48) {
            const int leg = idx / 12, which = (idx / 6) % 2, r = idx % 6;
            const int j = 3 * leg + lvl, par = lvl == 0 ? 0 : j;   // parent body index (body = 1 + joint)
            const double* X = which == 0 ? W.Xup + 36 * (1 + j) : W.Xur + 36 * j;
            double s = 0.0;
            for (int c = 0; c < 6; ++c) s += X[6 * r + c] * W.v[6 * par + c];
            if (r == qr_axis_of(j)) s += qd[j];
            if (which == 0) W.v[6 * (1 + j) + r] = s; else W.vr[6 * j + r] = s;
        }
        QR_SYNC();
    }
    QR_FOR(idx, 24) {   // c = v x_m vJ
        const int j = idx % 12;
        if (idx < 12) qr_motion_cross_axis(W.v + 6 * (1 + j), qr_axis_of(j), qd[j], W.cj + 6 * j);
        else qr_motion_cross_axis(W.vr + 6 * j, qr_axis_of(j), qd[j], W.cr + 6 * j);
    }
    QR_SYNC();
    for (int lvl = 0; lvl < 3; ++lvl) {   // biasAccelerations (:587-600) and absolute transforms
        QR_FOR(idx, 48) {
            const int leg = idx / 12, which = (idx / 6) % 2, r = idx % 6;
            const int j = 3 * leg + lvl, par = lvl == 0 ? 0 : j;
            const double* X = which == 0 ? W.Xup + 36 * (1 + j) : W.Xur + 36 * j;
            double s = which == 0 ? W.cj[6 * j + r] : W.cr[6 * j + r];
            for (int c = 0; c < 6; ++c) s += X[6 * r + c] * W.avp[6 * par + c];
            if (which == 0) W.avp[6 * (1 + j) + r] = s; else W.avr[6 * j + r] = s;
        }
        QR_FOR(idx, 4 * 36) {
            const int leg = idx / 36, e = idx - 36 * leg, r = e / 6, c = e - 6 * r;
            const int j = 3 * leg + lvl, par = lvl == 0 ? 0 : j;
            const double* Xp = lvl == 0 ? W.Xup : W.Xa + 36 * par;   // Xa[base] = Xup[base]
            double s = 0.0;
            for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + j) + 6 * r + l] * Xp[6 * l + c];
            W.Xa[36 * (1 + j) + e] = s;
        }
        QR_SYNC();
    }
    // ---- feet: position, velocity, contact Jacobian, Jdot*qdot (:503-523, :541-580)
    QR_FOR(leg, 4) {
        const int knee = 3 + 3 * leg;   // body index of the knee link
        const double* Xa = W.Xa + 36 * knee;
        const double* loc = M.foot[leg];
        double E[9], Et[9];             // E = rotation part of Xa (world -> link)
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { E[3 * i + j] = Xa[6 * i + j]; Et[3 * j + i] = Xa[6 * i + j]; }
        // r = -unskew(E' * Xa_bottomleft): origin of the link frame in world coordinates
        double Mx[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Mx[3 * i + j] = Et[3 * i] * Xa[6 * 3 + j] + Et[3 * i + 1] * Xa[6 * 4 + j] + Et[3 * i + 2] * Xa[6 * 5 + j];
        const double r[3] = {-0.5 * (Mx[7] - Mx[5]), -0.5 * (Mx[2] - Mx[6]), -0.5 * (Mx[3] - Mx[1])};
        // Xai = invertSXform(Xa) = sxform(E', -E r); its translation is t = -E r (in link coordinates)
        double tl[3], Xai[36];
        for (int i = 0; i < 3; ++i) tl[i] = -(E[3 * i] * r[0] + E[3 * i + 1] * r[1] + E[3 * i + 2] * r[2]);
        qr_sxform(Et, tl, Xai);
        double vs[6];
        for (int i = 0; i < 6; ++i) { double s = 0.0; for (int c = 0; c < 6; ++c) s += Xai[6 * i + c] * W.v[6 * knee + c]; vs[i] = s; }
        // sXFormPoint(Xai, loc) = R (loc - r) with R, r re-extracted from Xai exactly as the reference does
        // (translationFromSXform): with a quaternion that is only normalised to float precision E is not
        // exactly orthonormal and the shortcut E'(loc - t) differs at the 1e-7 level.
        double p[3], t2[3], M2[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M2[3 * i + j] = Xai[6 * 0 + i] * Xai[6 * 3 + j] + Xai[6 * 1 + i] * Xai[6 * 4 + j] + Xai[6 * 2 + i] * Xai[6 * 5 + j];
        t2[0] = -0.5 * (M2[7] - M2[5]); t2[1] = -0.5 * (M2[2] - M2[6]); t2[2] = -0.5 * (M2[3] - M2[1]);
        for (int i = 0; i < 3; ++i) p[i] = Xai[6 * i] * (loc[0] - t2[0]) + Xai[6 * i + 1] * (loc[1] - t2[1]) + Xai[6 * i + 2] * (loc[2] - t2[2]);
        for (int i = 0; i < 3; ++i) W.pF[3 * leg + i] = p[i];
        W.vF[3 * leg + 0] = vs[3] + (vs[1] * p[2] - vs[2] * p[1]);
        W.vF[3 * leg + 1] = vs[4] + (vs[2] * p[0] - vs[0] * p[2]);
        W.vF[3 * leg + 2] = vs[5] + (vs[0] * p[1] - vs[1] * p[0]);
        // contact Jacobian: Xc = sxform(E', loc); Xout = bottom three rows, pulled back through the chain
        double Xc[36], ac[6], vc[6];
        qr_sxform(Et, loc, Xc);
        for (int i = 0; i < 6; ++i) {
            double sa = 0.0, sv = 0.0;
            for (int c = 0; c < 6; ++c) { sa += Xc[6 * i + c] * W.avp[6 * knee + c]; sv += Xc[6 * i + c] * W.v[6 * knee + c]; }
            ac[i] = sa; vc[i] = sv;
        }
        W.Jcd[3 * leg + 0] = ac[3] + (vc[1] * vc[5] - vc[2] * vc[4]);
        W.Jcd[3 * leg + 1] = ac[4] + (vc[2] * vc[3] - vc[0] * vc[5]);
        W.Jcd[3 * leg + 2] = ac[5] + (vc[0] * vc[4] - vc[1] * vc[3]);
        double Xo[18], Xn[18];
        for (int i = 0; i < 3; ++i) for (int c = 0; c < 6; ++c) Xo[6 * i + c] = Xc[6 * (3 + i) + c];
        double* Jc = W.Jc + 54 * leg;
        for (int i = 0; i < 54; ++i) Jc[i] = 0.0;
        for (int lvl = 2; lvl >= 0; --lvl) {
            const int j = 3 * leg + lvl, ax = qr_axis_of(j);
            for (int i = 0; i < 3; ++i) Jc[18 * i + 6 + j] = Xo[6 * i + ax];
            const double* X = W.Xup + 36 * (1 + j);
            for (int i = 0; i < 3; ++i) for (int c = 0; c < 6; ++c) { double s = 0.0; for (int l = 0; l < 6; ++l) s += Xo[6 * i + l] * X[6 * l + c]; Xn[6 * i + c] = s; }
            for (int i = 0; i < 18; ++i) Xo[i] = Xn[i];
        }
        for (int i = 0; i < 3; ++i) for (int c = 0; c < 6; ++c) Jc[18 * i + c] = Xo[6 * i + c];
    }
    // ---- compositeInertias (:750-767)
    QR_FOR(idx, 13 * 36) W.IC[idx] = M.Ibody[idx / 36][idx % 36];
    QR_SYNC();
    for (int lvl = 2; lvl >= 0; --lvl) {
        QR_FOR(idx, 8 * 36) {   // T1 = IC * Xup, T2 = Irot * Xuprot
            const int leg = (idx / 36) % 4, which = idx / 144, e = idx % 36, r = e / 6, c = e - 6 * r;
            const int j = 3 * leg + lvl;
            const double* I = which == 0 ? W.IC + 36 * (1 + j) : M.Irot[j];
            const double* X = which == 0 ? W.Xup + 36 * (1 + j) : W.Xur + 36 * j;
            double s = 0.0;
            for (int l = 0; l < 6; ++l) s += I[6 * r + l] * X[6 * l + c];
            (which == 0 ? W.T1 : W.T2)[36 * leg + e] = s;   // per level: one 6x6 per leg
        }
        QR_SYNC();
        if (lvl > 0) {
            QR_FOR(idx, 4 * 36) {
                const int leg = idx / 36, e = idx - 36 * leg, r = e / 6, c = e - 6 * r;
                const int j = 3 * leg + lvl;
                double s = 0.0;
                for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + j) + 6 * l + r] * W.T1[36 * leg + 6 * l + c] + W.Xur[36 * j + 6 * l + r] * W.T2[36 * leg + 6 * l + c];
                W.IC[36 * j + e] += s;   // parent body of joint j is body j (= 1 + (j - 1))
            }
        } else {
            QR_FOR(e, 36) {
                const int r = e / 6, c = e - 6 * r;
                double s = 0.0;
                for (int leg = 0; leg < 4; ++leg) {
                    const int j = 3 * leg;
                    for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + j) + 6 * l + r] * W.T1[36 * leg + 6 * l + c] + W.Xur[36 * j + 6 * l + r] * W.T2[36 * leg + 6 * l + c];
                }
                W.IC[e] += s;
            }
        }
        QR_SYNC();
    }
    // ---- massMatrix (:774-806)
    QR_FOR(idx, 324) W.H[idx] = 0.0;
    QR_SYNC();
    QR_FOR(e, 36) W.H[18 * (e / 6) + e % 6] = W.IC[e];
    QR_FOR(j, 12) {
        const int ax = qr_axis_of(j);
        double f[6], fr[6], fn[6];
        for (int r = 0; r < 6; ++r) { f[r] = W.IC[36 * (1 + j) + 6 * r + ax]; fr[r] = M.Irot[j][6 * r + ax]; }
        W.H[18 * (6 + j) + 6 + j] = f[ax] + fr[ax];
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
            for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + j) + 6 * l + r] * f[l] + W.Xur[36 * j + 6 * l + r] * fr[l];
            fn[r] = s;
        }
        for (int r = 0; r < 6; ++r) f[r] = fn[r];
        for (int i = j - 1; i >= 3 * (j / 3); --i) {   // ancestors within the leg
            W.H[18 * (6 + i) + 6 + j] = f[qr_axis_of(i)];
            W.H[18 * (6 + j) + 6 + i] = f[qr_axis_of(i)];
            for (int r = 0; r < 6; ++r) { double s = 0.0; for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + i) + 6 * l + r] * f[l]; fn[r] = s; }
            for (int r = 0; r < 6; ++r) f[r] = fn[r];
        }
        for (int r = 0; r < 6; ++r) { W.H[18 * r + 6 + j] = f[r]; W.H[18 * (6 + j) + r] = f[r]; }
    }
    // ---- generalizedGravityForce (:607-626)
    QR_THREADS(t) {
        if (t == 0) {
            for (int r = 0; r < 6; ++r) W.ag[r] = W.Xup[6 * r + 3] * M.gravity[0] + W.Xup[6 * r + 4] * M.gravity[1] + W.Xup[6 * r + 5] * M.gravity[2];
        }
    }
    QR_SYNC();
    for (int lvl = 0; lvl < 3; ++lvl) {
        QR_FOR(idx, 48) {
            const int leg = idx / 12, which = (idx / 6) % 2, r = idx % 6;
            const int j = 3 * leg + lvl, par = lvl == 0 ? 0 : j;
            const double* X = which == 0 ? W.Xup + 36 * (1 + j) : W.Xur + 36 * j;
            double s = 0.0;
            for (int c = 0; c < 6; ++c) s += X[6 * r + c] * W.ag[6 * par + c];
            if (which == 0) W.ag[6 * (1 + j) + r] = s; else W.agr[6 * j + r] = s;
        }
        QR_SYNC();
    }
    QR_FOR(i, 18) {
        double s = 0.0;
        if (i < 6) { for (int c = 0; c < 6; ++c) s -= W.IC[6 * i + c] * W.ag[c]; }
        else {
            const int j = i - 6, ax = qr_axis_of(j);
            for (int c = 0; c < 6; ++c) s -= W.IC[36 * (1 + j) + 6 * ax + c] * W.ag[6 * (1 + j) + c] + M.Irot[j][6 * ax + c] * W.agr[6 * j + c];
        }
        W.G[i] = s;
    }
    // ---- generalizedCoriolisForce (:633-665)
    QR_FOR(b, 25) {
        const bool rotor = b >= 13;
        const int j = rotor ? b - 13 : b - 1;
        const double* I = rotor ? M.Irot[j] : M.Ibody[b];
        const double* vv = rotor ? W.vr + 6 * j : W.v + 6 * b;
        const double* aa = rotor ? W.avr + 6 * j : W.avp + 6 * b;
        double h[6], fa[6], fc[6];
        for (int r = 0; r < 6; ++r) { double sh = 0.0, sa = 0.0; for (int c = 0; c < 6; ++c) { sh += I[6 * r + c] * vv[c]; sa += I[6 * r + c] * aa[c]; } h[r] = sh; fa[r] = sa; }
        qr_force_cross(vv, h, fc);
        double* o = rotor ? W.fvr + 6 * j : W.fvp + 6 * b;
        for (int r = 0; r < 6; ++r) o[r] = fa[r] + fc[r];
    }
    QR_SYNC();
    for (int lvl = 2; lvl >= 0; --lvl) {
        QR_FOR(leg, 4) { const int j = 3 * leg + lvl, ax = qr_axis_of(j); W.Cq[6 + j] = W.fvp[6 * (1 + j) + ax] + W.fvr[6 * j + ax]; }
        if (lvl > 0) {
            QR_FOR(idx, 24) {
                const int leg = idx / 6, r = idx % 6, j = 3 * leg + lvl;
                double s = 0.0;
                for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + j) + 6 * l + r] * W.fvp[6 * (1 + j) + l] + W.Xur[36 * j + 6 * l + r] * W.fvr[6 * j + l];
                W.fvp[6 * j + r] += s;
            }
        } else {
            QR_FOR(r, 6) {
                double s = 0.0;
                for (int leg = 0; leg < 4; ++leg) {
                    const int j = 3 * leg;
                    for (int l = 0; l < 6; ++l) s += W.Xup[36 * (1 + j) + 6 * l + r] * W.fvp[6 * (1 + j) + l] + W.Xur[36 * j + 6 * l + r] * W.fvr[6 * j + l];
                }
                W.fvp[r] += s;
            }
        }
        QR_SYNC();
    }
    QR_FOR(r, 6) W.Cq[r] = W.fvp[r];
    QR_SYNC();
}

// rpyToQuat (utils/qr_se3.h:229-235): rotationMatrixToQuaternion(Rx(r) Ry(p) Rz(y)), passive rotations
QR_DEV void qr_rpy_to_quat(const double* rpy, double* q) {
#if defined(QR_ON_DEVICE)
    double sn[3], cs[3];
#pragma unroll 1
    for (int i = 0; i < 3; ++i) sincos(rpy[i], &sn[i], &cs[i]);   // one copy of the routine, not three
    const double cr = cs[0], sr = sn[0], cp = cs[1], sp = sn[1], cy = cs[2], sy = sn[2];
#else
    const double cr = cos(rpy[0]), sr = sin(rpy[0]), cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
#endif
    const double Rx[9] = {1, 0, 0, 0, cr, sr, 0, -sr, cr}, Ry[9] = {cp, 0, -sp, 0, 1, 0, sp, 0, cp}, Rz[9] = {cy, sy, 0, -sy, cy, 0, 0, 0, 1};
    double A[9], R1[9], r[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[3 * i + j] = Rx[3 * i] * Ry[j] + Rx[3 * i + 1] * Ry[3 + j] + Rx[3 * i + 2] * Ry[6 + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R1[3 * i + j] = A[3 * i] * Rz[j] + A[3 * i + 1] * Rz[3 + j] + A[3 * i + 2] * Rz[6 + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[3 * i + j] = R1[3 * j + i];   // r = r1'
    const double tr = r[0] + r[4] + r[8];
    if (tr > 0.0) { const double S = sqrt(tr + 1.0) * 2.0; q[0] = 0.25 * S; q[1] = (r[7] - r[5]) / S; q[2] = (r[2] - r[6]) / S; q[3] = (r[3] - r[1]) / S; }
    else if (r[0] > r[4] && r[0] > r[8]) { const double S = sqrt(1.0 + r[0] - r[4] - r[8]) * 2.0; q[0] = (r[7] - r[5]) / S; q[1] = 0.25 * S; q[2] = (r[1] + r[3]) / S; q[3] = (r[2] + r[6]) / S; }
    else if (r[4] > r[8]) { const double S = sqrt(1.0 + r[4] - r[0] - r[8]) * 2.0; q[0] = (r[2] - r[6]) / S; q[1] = (r[1] + r[3]) / S; q[2] = 0.25 * S; q[3] = (r[5] + r[7]) / S; }
    else { const double S = sqrt(1.0 + r[8] - r[0] - r[4]) * 2.0; q[0] = (r[3] - r[1]) / S; q[1] = (r[2] + r[6]) / S; q[2] = (r[5] + r[7]) / S; q[3] = 0.25 * S; }
}

QR_DEV double qr_clip10(double v) { return v < -10.0 ? -10.0 : (v > 10.0 ? 10.0 : v); }

// Tasks and contacts (ContactTaskUpdate, qr_wbc_locomotion_controller.cpp:172-201).
template <int NT>
QR_DEV void qr_wbc_tasks(const QrWbcModelDev& M, QrWbcWork& W, const int32_t* contact) {
    QR_FOR(idx, 6 * 54) W.Jt[idx] = 0.0;
    QR_FOR(idx, 18) { W.jdq[idx] = 0.0; }
    QR_SYNC();
    QR_THREADS(t) {
        if (t == 0) {
            const double* quat = W.st; const double* pos = W.st + 4; const double* bv = W.st + 7;
            const double* c = W.cmd;
            double Rw[9];                      // world -> body; tasks use its transpose (body -> world)
            qr_quat_to_rot(quat, Rw);
            // task 0: body orientation (qr_task_body_orientation.cpp:41-90), Kp 100 / Kd 10
            double qdes[4], qe[4], so3[3];
            qr_rpy_to_quat(c + 9, qdes);
            const double qi[4] = {quat[0], -quat[1], -quat[2], -quat[3]};
            qe[0] = qdes[0] * qi[0] - (qdes[1] * qi[1] + qdes[2] * qi[2] + qdes[3] * qi[3]);
            qe[1] = qdes[0] * qi[1] + qi[0] * qdes[1] + (qdes[2] * qi[3] - qdes[3] * qi[2]);
            qe[2] = qdes[0] * qi[2] + qi[0] * qdes[2] + (qdes[3] * qi[1] - qdes[1] * qi[3]);
            qe[3] = qdes[0] * qi[3] + qi[0] * qdes[3] + (qdes[1] * qi[2] - qdes[2] * qi[1]);
            if (qe[0] < 0.0) for (int i = 0; i < 4; ++i) qe[i] = -qe[i];
            so3[0] = qe[1]; so3[1] = qe[2]; so3[2] = qe[3];
            const double theta = 2.0 * asin(sqrt(so3[0] * so3[0] + so3[1] * so3[1] + so3[2] * so3[2]));
            if (fabs(theta) < 0.0000001) { so3[0] = so3[1] = so3[2] = 0.0; }
            else { const double sh = sin(theta / 2.0); for (int i = 0; i < 3; ++i) { so3[i] /= sh; so3[i] *= theta; } }
            const double dv[3] = {c[63] - bv[0], c[64] - bv[1], c[65] - bv[2]};   // previous command - omega_body
            for (int i = 0; i < 3; ++i) {
                const double ve = Rw[i] * dv[0] + Rw[3 + i] * dv[1] + Rw[6 + i] * dv[2];   // Rot' * dv
                for (int j = 0; j < 3; ++j) W.Jt[18 * i + j] = Rw[3 * j + i];
                W.perr[i] = so3[i]; W.dvel[i] = c[12 + i];
                W.xdd[i] = qr_clip10(M.kp_ori * so3[i] + M.kd_ori * ve + 0.0);
            }
            // task 1: body position (qr_task_body_position.cpp:42-74), Kp 100 / Kd 10
            for (int i = 0; i < 3; ++i) {
                const double vw = Rw[i] * bv[3] + Rw[3 + i] * bv[4] + Rw[6 + i] * bv[5];
                for (int j = 0; j < 3; ++j) W.Jt[54 + 18 * i + 3 + j] = Rw[3 * j + i];
                W.perr[3 + i] = c[i] - pos[i]; W.dvel[3 + i] = c[3 + i];
                W.xdd[3 + i] = qr_clip10(M.kp_pos * (c[i] - pos[i]) + M.kd_pos * (c[3 + i] - vw) + c[6 + i]);
            }
            int nt = 2, nc = 0;
            W.ints[6] = -1; W.ints[7] = -1;
            for (int leg = 0; leg < 4; ++leg) {
                if (contact[leg]) { W.ints[2 + nc] = leg; ++nc; continue; }
                // swing foot: qr_task_link_position.cpp:43-88, Kp 500 / Kd 10, no clipping
                W.ints[6 + nt] = leg;
                for (int i = 0; i < 3; ++i) {
                    const double pe = c[15 + 3 * leg + i] - W.pF[3 * leg + i];
                    W.perr[3 * nt + i] = pe; W.dvel[3 * nt + i] = c[27 + 3 * leg + i];
                    W.xdd[3 * nt + i] = M.kp_foot * pe + M.kd_foot * (c[27 + 3 * leg + i] - W.vF[3 * leg + i]) + c[39 + 3 * leg + i];
                    W.jdq[3 * nt + i] = W.Jcd[3 * leg + i];
                }
                ++nt;
            }
            W.ints[0] = nt; W.ints[1] = nc;
            for (int k = 0; k < nc; ++k)
                for (int i = 0; i < 3; ++i) { W.fdes[3 * k + i] = c[51 + 3 * W.ints[2 + k] + i]; W.JCd[3 * k + i] = W.Jcd[3 * W.ints[2 + k] + i]; }
        }
    }
    QR_SYNC();
    const int nt = W.ints[0], nc = W.ints[1];
    QR_FOR(idx, (nt - 2) * 54) { const int k = 2 + idx / 54; W.Jt[54 * k + idx % 54] = W.Jc[54 * W.ints[6 + k] + idx % 54]; }
    QR_FOR(idx, nc * 54) { const int k = idx / 54; W.JC[idx] = W.Jc[54 * W.ints[2 + k] + idx % 54]; }
    QR_SYNC();
}

// Kinematic WBC: qrMultitaskProjection::FindConfiguration (qr_multitask_projection.cpp:38-106).
template <int NT>
QR_DEV void qr_wbc_kin(QrWbcWork& W) {
    const int nt = W.ints[0], nc = W.ints[1], m = 3 * nc;
    const double thr = 0.001;
    if (nc > 0) {
        tm_pinv<NT>(W, W.Jbar, W.JC, m, 18, thr);                 // 18 x m
        tm_eye_minus_mul<NT>(W.N, W.Jbar, W.JC, 18, m);           // Nc
    } else {
        QR_FOR(idx, 324) W.N[idx] = (idx / 18 == idx % 18) ? 1.0 : 0.0;
        QR_SYNC();
    }
    QR_FOR(i, 18) { W.dq[i] = 0.0; W.qdot[i] = 0.0; }
    QR_SYNC();
    for (int k = 0; k < nt; ++k) {
        const double* Jt = W.Jt + 54 * k;
        tm_mul<NT>(W.Jpre, Jt, W.N, 3, 18, 18);                   // JtPre = Jt N
        tm_pinv<NT>(W, W.Jbar, W.Jpre, 3, 18, thr);               // 18 x 3
        QR_FOR(i, 6) {                                            // residuals e - Jt dq, v - Jt qdot
            const int r = i % 3;
            const double* vecx = i < 3 ? W.dq : W.qdot;
            double s = i < 3 ? W.perr[3 * k + r] : W.dvel[3 * k + r];
            for (int c = 0; c < 18; ++c) s -= Jt[18 * r + c] * vecx[c];
            W.vec[i] = s;
        }
        QR_SYNC();
        QR_FOR(i, 18) {
            W.dq[i] += W.Jbar[3 * i] * W.vec[0] + W.Jbar[3 * i + 1] * W.vec[1] + W.Jbar[3 * i + 2] * W.vec[2];
            W.qdot[i] += W.Jbar[3 * i] * W.vec[3] + W.Jbar[3 * i + 1] * W.vec[4] + W.Jbar[3 * i + 2] * W.vec[5];
        }
        QR_SYNC();
        if (k < nt - 1) tm_project_out<NT>(W.N, W.Jbar, W.Jpre, W.M1);   // N <- N (I - pinv(JtPre) JtPre)
    }
}

// WBIC: qrWholeBodyImpulseCtrl::MakeTorque (qr_wholebody_impulse_ctrl.cpp:62-126) up to the QP.
template <int NT>
QR_DEV void qr_wbc_wbic_stack(QrWbcWork& W) {
    const int nt = W.ints[0], nc = W.ints[1], m = 3 * nc;
    if (nc > 0) {
        tm_weighted_inverse<NT>(W, W.Jbar, W.JC, m);              // JcBar 18 x m
        QR_FOR(i, 18) { double s = 0.0; for (int c = 0; c < m; ++c) s -= W.Jbar[m * i + c] * W.JCd[c]; W.qdd[i] = s; }
        tm_eye_minus_mul<NT>(W.N, W.Jbar, W.JC, 18, m);
    } else {
        QR_FOR(i, 18) W.qdd[i] = 0.0;
        QR_FOR(idx, 324) W.N[idx] = (idx / 18 == idx % 18) ? 1.0 : 0.0;
        QR_SYNC();
    }
    for (int k = 0; k < nt; ++k) {
        const double* Jt = W.Jt + 54 * k;
        tm_mul<NT>(W.Jpre, Jt, W.N, 3, 18, 18);
        tm_weighted_inverse<NT>(W, W.Jbar, W.Jpre, 3);            // JtBar 18 x 3
        QR_FOR(r, 3) {
            double s = W.xdd[3 * k + r] - W.jdq[3 * k + r];
            for (int c = 0; c < 18; ++c) s -= Jt[18 * r + c] * W.qdd[c];
            W.vec[r] = s;
        }
        QR_SYNC();
        QR_FOR(i, 18) W.qdd[i] += W.Jbar[3 * i] * W.vec[0] + W.Jbar[3 * i + 1] * W.vec[1] + W.Jbar[3 * i + 2] * W.vec[2];
        QR_SYNC();
        if (k < nt - 1) tm_project_out<NT>(W.N, W.Jbar, W.Jpre, W.M1);   // N <- N (I - pinv(JtPre) JtPre)
    }
}

// The QP of MakeTorque (SetCost :232-247, SetEqualityConstraint :129-148, SetInequalityConstraint :152-167)
// reduced to the contact forces, and GetSolution (:210-228).  Returns the per-robot status.
template <int NT>
QR_DEV int qr_wbc_qp_and_torque(const QrWbcModelDev& M, const qr_qp_options& opt, QrWbcWork& W) {
    const int nc = W.ints[1], m = 3 * nc;
    // tot = A qdd + C + G - JC' f_des
    QR_FOR(i, 18) {
        double s = W.Cq[i] + W.G[i];
        for (int c = 0; c < 18; ++c) s += W.H[18 * i + c] * W.qdd[c];
        for (int c = 0; c < m; ++c) s -= W.JC[18 * c + i] * W.fdes[c];
        W.tot[i] = s;
    }
    QR_FOR(e, 36) W.A6[e] = W.H[18 * (e / 6) + e % 6];
    QR_SYNC();
    tm_inverse_spd<NT>(W, W.A6, W.A6, W.rowbuf, 6);
    // da = P df - a0 with P = A6^-1 (JC')_{0:6} (6 x m), a0 = A6^-1 tot_{0:6}
    QR_FOR(idx, 6 * m) { const int i = idx / m, c = idx - i * m; double s = 0.0; for (int l = 0; l < 6; ++l) s += W.A6[6 * i + l] * W.JC[18 * c + l]; W.P6[idx] = s; }
    QR_FOR(i, 6) { double s = 0.0; for (int l = 0; l < 6; ++l) s += W.A6[6 * i + l] * W.tot[l]; W.a0[i] = s; }
    QR_SYNC();
    int status = 0;
    QrQpWork& Q = W.Q;
    const double* f = W.fdes;
    if (nc > 0) {
        // Hf = w_fb P'P + w_fr I (block-packed), gf = -w_fb P' a0 - Hf f_des
        QR_FOR(idx, 9 * ((nc * (nc + 1)) / 2)) {
            const int b = idx / 9, e = idx - 9 * b, code = Q.tri[b];
            const int r = 3 * (code >> 8) + e / 3, c = 3 * (code & 255) + e % 3;
            double s = (r == c) ? M.w_fr : 0.0;
            for (int l = 0; l < 6; ++l) s += M.w_fb * W.P6[m * l + r] * W.P6[m * l + c];
            Q.Hs[idx] = s;
        }
        QR_SYNC();
        QR_FOR(r, m) {
            double s = 0.0;
            for (int l = 0; l < 6; ++l) s -= M.w_fb * W.P6[m * l + r] * W.a0[l];
            s -= qr_sym_matvec_row_t<true>(Q.Hs, W.fdes, nc, r);
            Q.g[r] = s;
        }
        QR_FOR(k, nc) Q.ubz[k] = M.max_fz;
        QR_SYNC();
        Q.nf = nc;
        Q.mu_ = 1.0 / M.mu;
        int it = 0, rounds = 0;
        const double* x = Q.xn;
        QR_PROF_DECL;
        status = qr_qp_solve<NT, false>(Q, opt, &it, &rounds, &x, nullptr QR_PROF_PASS);
        f = x;
    }
    // qdd[0:6] += da ; tau = (A qdd + C + G - JC' f)[6:18]
    QR_FOR(i, 6) {
        double s = -W.a0[i];
        for (int c = 0; c < m; ++c) s += W.P6[m * i + c] * (f[c] - W.fdes[c]);
        W.vec[i] = s;
    }
    QR_SYNC();
    QR_FOR(i, 6) W.qdd[i] += W.vec[i];
    QR_SYNC();
    QR_FOR(i, 18) {
        double s = W.Cq[i] + W.G[i];
        for (int c = 0; c < 18; ++c) s += W.H[18 * i + c] * W.qdd[c];
        for (int c = 0; c < m; ++c) s -= W.JC[18 * c + i] * f[c];
        W.tot[i] = s;
    }
    QR_FOR(c, m) W.vec[6 + c] = f[c];   // keep the forces (vec[6..17]) past the QP workspace
    QR_SYNC();
    return status;
}

// Once per launch: the index table of the (at most 4 x 4 block) force QP.
template <int NT>
QR_DEV void qr_wbc_init_tables(QrWbcWork& W) {
    QR_FOR(idx, 10) {
        int I, J;
        qr_tri_decode(idx, I, J);
        W.Q.tri[idx] = (unsigned short)((I << 8) | J);
    }
    QR_SYNC();
}

template <int NT>
QR_DEV void qr_wbc_problem(const QrWbcArgs& A, int prob, QrWbcWork& W) {
    const QrWbcModelDev& M = *A.model;
    QR_FOR(i, 37) W.st[i] = (double)A.state[(size_t)prob * 37 + i];
    QR_FOR(i, 66) W.cmd[i] = (double)A.cmd[(size_t)prob * 66 + i];
    QR_SYNC();
    qr_wbc_dynamics<NT>(M, W);
    tm_inverse_spd<NT>(W, W.H, W.Ainv, W.M1, 18);   // the dynamics temporaries are dead: M1 of the overlay is free
    qr_wbc_tasks<NT>(M, W, A.contact + (size_t)prob * 4);
    qr_wbc_kin<NT>(W);
    qr_wbc_wbic_stack<NT>(W);
    int status = qr_wbc_qp_and_torque<NT>(M, A.opt, W);
    int bad = 0;
    QR_FOR(i, 12) bad |= !(fabs(W.tot[6 + i]) < 1e300);
    if (QR_ANY(bad)) status = 3;
    const int nc = W.ints[1];
    QR_FOR(i, 12) {
        const double tau = status == 3 ? 0.0 : W.tot[6 + i];
        const double qd_ = W.st[13 + i] + W.dq[6 + i], qdd_ = W.qdot[6 + i];
        double fr = 0.0;
        for (int k = 0; k < nc; ++k) if (W.ints[2 + k] == i / 3) fr = W.vec[6 + 3 * k + i % 3];
        const size_t o = (size_t)prob * 12 + i;
        if (A.tau32) A.tau32[o] = (float)tau;
        if (A.tau64) A.tau64[o] = tau;
        if (A.fr32) A.fr32[o] = (float)fr;
        if (A.fr64) A.fr64[o] = fr;
        if (A.qdes32) A.qdes32[o] = (float)qd_;
        if (A.qdes64) A.qdes64[o] = qd_;
        if (A.qddes32) A.qddes32[o] = (float)qdd_;
        if (A.qddes64) A.qddes64[o] = qdd_;
    }
    if (A.dbg) {
        double* o = A.dbg + (size_t)prob * 630;
        QR_FOR(i, 324) o[i] = W.H[i];
        QR_FOR(i, 18) { o[324 + i] = W.G[i]; o[342 + i] = W.Cq[i]; o[612 + i] = W.qdd[i]; }
        QR_FOR(i, 216) o[360 + i] = W.Jc[i];
        QR_FOR(i, 12) { o[576 + i] = W.Jcd[i]; o[588 + i] = W.pF[i]; o[600 + i] = W.vF[i]; }
    }
    QR_THREADS(t) { if (t == 0 && A.status) A.status[prob] = status; }
    QR_SYNC();
}

// Swing foot, MPC mode: SwingFootTrajectory::GenerateTrajectoryPoint -> qrFootParabolaPatternGenerator ->
// qrQuadraticSpline::getPoint (qr_foot_trajectory_generator.cpp:188-215, 322-343; utils/qr_geometry.cpp:157-190),
// with the reference's mixed float / double evaluation.  One thread per foot.
QR_DEV int qr_swing_parabola(const float* start, const float* end, float height, float t_in, int phase_module, float* pos) {
    float phase;
    if (phase_module) {
        if (t_in <= 0.5) phase = (float)QR_DMUL(0.8, sin(QR_DMUL((double)t_in, 3.14159265358979323846)));
        else phase = (float)QR_DADD(0.8, QR_DMUL(QR_DADD((double)t_in, -0.5), 0.4));
    } else {
        phase = t_in;
    }
    if ((double)phase < 0.0 - 1e-3) return 0;
    if ((double)phase >= 0.0 + 1.0 + 1e-3) return 0;
    const float x = QR_FADD(QR_FMUL(QR_FSUB(1.f, phase), start[0]), QR_FMUL(phase, end[0]));
    const float y = QR_FADD(QR_FMUL(QR_FSUB(1.f, phase), start[1]), QR_FMUL(phase, end[1]));
    const float mid = QR_FADD(end[2] > start[2] ? end[2] : start[2], height);
    const float d1 = QR_FSUB(mid, start[2]), d2 = QR_FSUB(end[2], start[2]);
    const float d3 = (float)(0.25 - 0.5);
    const float ca = QR_FDIV(QR_FSUB(d1, QR_FMUL(d2, 0.5f)), d3);
    const float cb = (float)(QR_DADD(QR_DMUL((double)d2, 0.25), -(double)d1) / (double)d3);
    const float cbt = QR_FMUL(cb, phase);
    const double z = QR_DADD(QR_DADD(QR_DMUL((double)ca, QR_DMUL((double)phase, (double)phase)), (double)cbt), (double)start[2]);
    // -1e-3 <= phase < 0: the generator accepts the sample (qr_foot_trajectory_generator.cpp:194) but getPoint refuses
    // it (dt < 0, qr_geometry.cpp:171-173) and leaves its output Point at the default x = 0, which the generator
    // then returns as the foot height.  Found by pinning against the reference-compiled generator.
    pos[0] = x; pos[1] = y; pos[2] = (phase < 0.f) ? 0.f : (float)z;
    return 1;
}
