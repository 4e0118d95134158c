// fb_problem.h -- force-balance stance QP of one robot on one thread.
//
// Replaces Quadruped::ComputeContactForce and its helpers ComputeMassMatrix / ComputeObjectiveMatrix /
// ComputeWeightMatrix / ComputeConstraintMatrix
// (/root/reference/quadruped/src/controllers/balance_controller/qr_qp_torque_optimizer.cpp:31-64, 67-151, 154-182,
//  185-189, 192-301 [control-frame overload], 304-400 [world-frame overload], 403-427) for a batch of robots:
//     min 1/2 x'Gx - a'x,   G = M'QM + regWeight*ones(12,12) + 1e-4*I,   a = ((g + ddq_des)'Q M)'
//     s.t.  fmin_l <= n.x_l <= fmax_l (stance legs),  n.x_l >= 1e-7 and -n.x_l >= 1e-7 (swing legs, as written in the
//           reference),  (mu n +- t1).x_l >= 0,  (mu n +- t2).x_l >= 0
// where x is the reference's QP variable (x = -force; the function returns X = -x as a 4x3 leg-major matrix).
// The QP data are built in float32 with the operation order of the reference's Eigen expressions (coefficient-wise
// products, k ascending, no FMA contraction) and widened to double exactly as the reference does before it calls
// QuadProg++ (GG[i][j] = G(j,i): the solver reads that matrix's upper triangle, i.e. the LOWER triangle of G).
// The frame changes around the QP (Rcb, RigidTransform) stay with the caller: `inertia` is the 3x3 matrix the
// reference inverts, `foot` the 4x3 matrix it hands to ComputeMassMatrix.
#pragma once

#include "small_qp.h"
#include "../../include/qr_gpu.h"

struct QrFbArgs {
    qr_fb_params P;
    int batch;
    const float* inertia;   // [B][9] or null (then P.inertia)
    const float* foot;      // [B][12]  footPositions.row(leg)
    const float* acc;       // [B][6]   desiredAcc
    const int32_t* contact; // [B][4]
    const float* gravity;   // [B][3] or null: g.head(3) (0,0,9.8 unless the control frame is tilted)
    const float* frame;     // [B][9] or null: normal, tangent1, tangent2 rows (default e_z, e_x, e_y)
    float* force_out;       // [B][12]  X(leg, axis) = -x[3 leg + axis]
    int32_t* status_out;    // [B] or null
    int32_t* iters_out;     // [B] or null
};

// Float32 QP data of one robot: Gf (12x12 row-major), af (12), Cf (24 rows of 12), lbf (24).
QR_DEV void qr_fb_build(const qr_fb_params& P, const float* inertia, const float* foot, const float* acc,
                        const int32_t* contact, const float* gravity, const float* frame, float* Gf, float* af,
                        float* Cf, float* lbf) {
    // ---- ComputeMassMatrix (:31-64): invMass = I / m, invInertia = inertia.inverse() (3x3: cofactors / determinant)
    const float* Iw = inertia;
    const float minv = QR_FDIV(1.f, P.mass);
#define QR_COF(a, b, c, d) QR_FSUB(QR_FMUL(Iw[a], Iw[b]), QR_FMUL(Iw[c], Iw[d]))
    const float c00 = QR_COF(4, 8, 5, 7), c01 = QR_COF(5, 6, 3, 8), c02 = QR_COF(3, 7, 4, 6);
    const float det = QR_FADD(QR_FADD(QR_FMUL(Iw[0], c00), QR_FMUL(Iw[1], c01)), QR_FMUL(Iw[2], c02));
    const float id = QR_FDIV(1.f, det);
    float Ii[9];
    Ii[0] = QR_FMUL(c00, id); Ii[1] = QR_FMUL(QR_COF(2, 7, 1, 8), id); Ii[2] = QR_FMUL(QR_COF(1, 5, 2, 4), id);
    Ii[3] = QR_FMUL(c01, id); Ii[4] = QR_FMUL(QR_COF(0, 8, 2, 6), id); Ii[5] = QR_FMUL(QR_COF(2, 3, 0, 5), id);
    Ii[6] = QR_FMUL(c02, id); Ii[7] = QR_FMUL(QR_COF(1, 6, 0, 7), id); Ii[8] = QR_FMUL(QR_COF(0, 4, 1, 3), id);
#undef QR_COF
    float M[6][12];
    for (int leg = 0; leg < 4; ++leg) {
        const float x0 = foot[3 * leg], x1 = foot[3 * leg + 1], x2 = foot[3 * leg + 2];
        const float S[9] = {0.f, -x2, x1, x2, 0.f, -x0, -x1, x0, 0.f};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                M[i][3 * leg + j] = (i == j) ? minv : 0.f;
                M[3 + i][3 * leg + j] = QR_FADD(QR_FADD(QR_FMUL(Ii[3 * i], S[j]), QR_FMUL(Ii[3 * i + 1], S[3 + j])),
                                                QR_FMUL(Ii[3 * i + 2], S[6 + j]));
            }
    }
    // ---- ComputeObjectiveMatrix (:154-182): quadTerm = (M'Q) M + R,  linearTerm = ((g + acc)'Q) M
    float T1[12][6];
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 6; ++j) T1[i][j] = QR_FMUL(M[j][i], P.acc_weight[j]);
    float v[6];
    for (int k = 0; k < 6; ++k) {
        const float gk = k < 3 ? (gravity ? gravity[k] : (k == 2 ? 9.8f : 0.f)) : 0.f;
        v[k] = QR_FMUL(QR_FADD(gk, acc[k]), P.acc_weight[k]);
    }
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) {
            float s = QR_FMUL(T1[i][0], M[0][j]);
            for (int k = 1; k < 6; ++k) s = QR_FADD(s, QR_FMUL(T1[i][k], M[k][j]));
            s = QR_FADD(s, P.reg_weight);                    // + R = ones * regWeight
            if (i == j) s = QR_FADD(s, 1e-4f);               // + W = 1e-4 * I (ComputeWeightMatrix :185-189)
            Gf[12 * i + j] = s;
        }
        float s = QR_FMUL(v[0], M[0][i]);
        for (int k = 1; k < 6; ++k) s = QR_FADD(s, QR_FMUL(v[k], M[k][i]));
        af[i] = s;
    }
    // ---- ComputeConstraintMatrix (:67-151)
    float nrm[3] = {0.f, 0.f, 1.f}, t1[3] = {1.f, 0.f, 0.f}, t2[3] = {0.f, 1.f, 0.f};
    if (frame)
        for (int k = 0; k < 3; ++k) { nrm[k] = frame[k]; t1[k] = frame[3 + k]; t2[k] = frame[6 + k]; }
    for (int i = 0; i < 24 * 12; ++i) Cf[i] = 0.f;
    for (int leg = 0; leg < 4; ++leg) {
        for (int k = 0; k < 3; ++k) {
            Cf[12 * (2 * leg) + 3 * leg + k] = nrm[k];
            Cf[12 * (2 * leg + 1) + 3 * leg + k] = -nrm[k];
            const float mn = QR_FMUL(P.mu, nrm[k]);
            Cf[12 * (8 + 4 * leg) + 3 * leg + k] = QR_FADD(mn, t1[k]);
            Cf[12 * (9 + 4 * leg) + 3 * leg + k] = QR_FSUB(mn, t1[k]);
            Cf[12 * (10 + 4 * leg) + 3 * leg + k] = QR_FADD(mn, t2[k]);
            Cf[12 * (11 + 4 * leg) + 3 * leg + k] = QR_FSUB(mn, t2[k]);
        }
        if (contact[leg] > 0) {
            if (P.world_frame) {   // world-frame overload: float * float, then * 9.8 in double, narrowed (:128-129)
                lbf[2 * leg] = (float)((double)QR_FMUL(P.fmin_ratio[leg], P.mass) * 9.8);
                lbf[2 * leg + 1] = (float)((double)QR_FMUL(-P.fmax_ratio[leg], P.mass) * 9.8);
            } else {               // control-frame overload: fMinRatio * bodyMass * 9.8f in float (:77-78, :90-91)
                lbf[2 * leg] = QR_FMUL(QR_FMUL(P.fmin_ratio[leg], P.mass), 9.8f);
                lbf[2 * leg + 1] = -QR_FMUL(QR_FMUL(P.fmax_ratio[leg], P.mass), 9.8f);
            }
        } else {
            lbf[2 * leg] = 1e-7f;
            lbf[2 * leg + 1] = 1e-7f;
        }
        for (int k = 0; k < 4; ++k) lbf[8 + 4 * leg + k] = 0.f;
    }
}

// One robot: build, widen, solve, write X = -x.
QR_DEV void qr_fb_problem(const QrFbArgs& A, int prob) {
    float Gf[144], af[12], Cf[288], lbf[24];
    const float* inertia = A.inertia ? A.inertia + 9 * (size_t)prob : A.P.inertia;
    qr_fb_build(A.P, inertia, A.foot + 12 * (size_t)prob, A.acc + 6 * (size_t)prob, A.contact + 4 * (size_t)prob,
                A.gravity ? A.gravity + 3 * (size_t)prob : nullptr, A.frame ? A.frame + 9 * (size_t)prob : nullptr, Gf, af,
                Cf, lbf);
    double G[144], g0[12], C[288], c0[24], x[12];
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) G[12 * i + j] = (double)(i >= j ? Gf[12 * i + j] : Gf[12 * j + i]);
        g0[i] = -(double)af[i];
    }
    for (int i = 0; i < 288; ++i) C[i] = (double)Cf[i];
    for (int i = 0; i < 24; ++i) c0[i] = -(double)lbf[i];
    QrSmallQpWork W;
    int it = 0;
    int status = qr_small_qp_solve(12, 24, G, g0, C, c0, x, W, &it);
    bool bad = false;
    for (int i = 0; i < 12; ++i) bad |= !(fabs(x[i]) < 1e300);
    if (bad || status == 3) {   // the reference zeroes the forces when the solver hands back NaN (:279-292)
        status = 3;
        for (int i = 0; i < 12; ++i) x[i] = 0.0;
    }
    for (int i = 0; i < 12; ++i) A.force_out[12 * (size_t)prob + i] = -(float)x[i];
    if (A.status_out) A.status_out[prob] = status;
    if (A.iters_out) A.iters_out[prob] = it;
}
