// mpc_io.h -- the per-instance arithmetic on either side of the MPC solve, one thread per robot instance.
//
//   qr_mpc_contact_table    mpcTable of MPCStanceLegController::Run
//                           (/root/reference/quadruped/src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:282-303)
//   qr_mpc_reference_traj   trajAll of UpdateMPC (:345-376)
//   qr_mpc_grf_to_torque    f_ff = -R_base^T f (SolveDenseMPC :402-409) and tau = J_leg^T f_ff (GetAction :139-141,
//                           qrRobot::MapContactForceToJointTorques src/robots/qr_robot.cpp:241-251 with
//                           AnalyticalLegJacobian :148-172)
// All float32 with the reference's operation order (no FMA contraction); the table is bit-exact with the
// reference's x86-64 build.  sin / cos / sqrt of the torque path are evaluated in float64 and rounded once,
// which reproduces glibc's (almost always correctly rounded) float functions to the last bit except in rare
// ties.
#pragma once

#include "qr_team.h"

// progress[4], duty[4]: phaseInFullCycle / dutyFactor; early[4], contacts[4]: 0/1 (contacts may be null).
QR_DEV void qr_mpc_contact_table(int h, int num_horizon_l, const float* progress, const float* duty,
                                 const int32_t* early, const int32_t* contacts, float* table) {
    const float dPhase = (float)(1.0 / (double)(num_horizon_l * h));
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < 4; ++j) {
            float ph = QR_FADD(progress[j], QR_FMUL((float)i, dPhase));
            while (ph > 1.0f) ph = QR_FSUB(ph, 1.0f);
            table[4 * i + j] = (ph < duty[j] || (early && early[j])) ? 1.f : 0.f;
        }
    if (contacts)
        for (int j = 0; j < 4; ++j) table[j] = contacts[j] ? 1.f : 0.f;
}

// init[12] = {rollComp, pitchComp, yawDes, xDes, yDes, bodyHeight, 0, 0, yawRate, vxWorld, vyWorld, 0};
// pos_xy = actual base position (the start is clipped to +-0.1 m of it).
QR_DEV void qr_mpc_reference_traj(int h, float dt_mpc, const float* init, const float* pos_xy, float* traj) {
    float t0[12];
    for (int j = 0; j < 12; ++j) t0[j] = init[j];
    for (int a = 0; a < 2; ++a) {
        const float lo = QR_FSUB(pos_xy[a], 0.1f), hi = QR_FADD(pos_xy[a], 0.1f);
        const float x = t0[3 + a];
        t0[3 + a] = x < lo ? lo : (x > hi ? hi : x);
    }
    const float yaw_rate = t0[8], vx = t0[9], vy = t0[10];
    for (int i = 0; i < h; ++i) {
        for (int j = 0; j < 12; ++j) traj[12 * i + j] = t0[j];
        if (i > 0) {
            traj[12 * i + 2] = QR_FADD(traj[12 * (i - 1) + 2], QR_FMUL(dt_mpc, yaw_rate));
            traj[12 * i + 3] = QR_FADD(traj[12 * (i - 1) + 3], QR_FMUL(dt_mpc, vx));
            traj[12 * i + 4] = QR_FADD(traj[12 * (i - 1) + 4], QR_FMUL(dt_mpc, vy));
        }
    }
}

QR_DEV float qr_sinf_cr(float x) { return (float)sin((double)x); }
QR_DEV float qr_cosf_cr(float x) { return (float)cos((double)x); }
QR_DEV float qr_sqrtf_cr(float x) { return (float)sqrt((double)x); }

// One leg: f_ff = -R_base^T f (Rb = baseRMat row-major) and tau = J_leg^T f_ff.  t = the leg's three motor angles.
QR_DEV void qr_mpc_leg_force_torque(float hip_len, float upper_len, float lower_len, const float* Rb, int leg, const float* t,
                                    const float* f, float* ff, float* tau) {
#define QR_M(a, b) QR_FMUL(a, b)
#define QR_A(a, b) QR_FADD(a, b)
#define QR_S(a, b) QR_FSUB(a, b)
    for (int a = 0; a < 3; ++a)
        ff[a] = QR_A(QR_A(QR_M(-Rb[a], f[0]), QR_M(-Rb[3 + a], f[1])), QR_M(-Rb[6 + a], f[2]));
    const float sh = (leg & 1) ? hip_len : -hip_len;   // hipLength * pow(-1, leg_id + 1)
    const float s0 = qr_sinf_cr(t[0]), c0 = qr_cosf_cr(t[0]), s2 = qr_sinf_cr(t[2]), c2 = qr_cosf_cr(t[2]);
    const float uu = QR_M(upper_len, upper_len), ll = QR_M(lower_len, lower_len);
    const float lEff = qr_sqrtf_cr(QR_A(QR_A(uu, ll), QR_M(QR_M(QR_M(2.f, upper_len), lower_len), c2)));
    const float tEff = QR_A(t[1], QR_FDIV(t[2], 2.f));
    const float sE = qr_sinf_cr(tEff), cE = qr_cosf_cr(tEff);
    const float lu = QR_M(lower_len, upper_len);
    float J[9];
    J[0] = 0.f;
    J[1] = QR_M(-lEff, cE);
    J[2] = QR_S(QR_FDIV(QR_M(QR_M(lu, s2), sE), lEff), QR_FDIV(QR_M(lEff, cE), 2.f));
    J[3] = QR_A(QR_M(-sh, s0), QR_M(QR_M(lEff, c0), cE));
    J[4] = QR_M(QR_M(-lEff, s0), sE);
    J[5] = QR_S(QR_FDIV(QR_M(QR_M(QR_M(-lu, s0), s2), cE), lEff), QR_FDIV(QR_M(QR_M(lEff, s0), sE), 2.f));
    J[6] = QR_A(QR_M(sh, c0), QR_M(QR_M(lEff, s0), cE));
    J[7] = QR_M(QR_M(lEff, sE), c0);
    J[8] = QR_A(QR_FDIV(QR_M(QR_M(QR_M(lu, s2), c0), cE), lEff), QR_FDIV(QR_M(QR_M(lEff, sE), c0), 2.f));
    for (int a = 0; a < 3; ++a)
        tau[a] = QR_A(QR_A(QR_M(J[a], ff[0]), QR_M(J[3 + a], ff[1])), QR_M(J[6 + a], ff[2]));
#undef QR_M
#undef QR_A
#undef QR_S
}

// baseRMat = quaternionToRotationMatrix(q)^T (qr_robot.cpp:70, qr_se3.h:186-203), row-major
QR_DEV void qr_mpc_base_rmat(const float* quat, float* Rb) {
    const float e0 = quat[0], e1 = quat[1], e2 = quat[2], e3 = quat[3];
    Rb[0] = QR_FSUB(1.f, QR_FMUL(2.f, QR_FADD(QR_FMUL(e2, e2), QR_FMUL(e3, e3))));
    Rb[1] = QR_FMUL(2.f, QR_FSUB(QR_FMUL(e1, e2), QR_FMUL(e0, e3)));
    Rb[2] = QR_FMUL(2.f, QR_FADD(QR_FMUL(e1, e3), QR_FMUL(e0, e2)));
    Rb[3] = QR_FMUL(2.f, QR_FADD(QR_FMUL(e1, e2), QR_FMUL(e0, e3)));
    Rb[4] = QR_FSUB(1.f, QR_FMUL(2.f, QR_FADD(QR_FMUL(e1, e1), QR_FMUL(e3, e3))));
    Rb[5] = QR_FMUL(2.f, QR_FSUB(QR_FMUL(e2, e3), QR_FMUL(e0, e1)));
    Rb[6] = QR_FMUL(2.f, QR_FSUB(QR_FMUL(e1, e3), QR_FMUL(e0, e2)));
    Rb[7] = QR_FMUL(2.f, QR_FADD(QR_FMUL(e2, e3), QR_FMUL(e0, e1)));
    Rb[8] = QR_FSUB(1.f, QR_FMUL(2.f, QR_FADD(QR_FMUL(e1, e1), QR_FMUL(e2, e2))));
}

// quat (w,x,y,z), q[12], f_world[12] -> f_ff[12] (may be null), tau[12].
QR_DEV void qr_mpc_grf_to_torque(float hip_len, float upper_len, float lower_len, const float* quat, const float* q,
                                 const float* f_world, float* f_ff_out, float* tau) {
    float Rb[9];
    qr_mpc_base_rmat(quat, Rb);
    for (int leg = 0; leg < 4; ++leg) {
        float ff[3];
        qr_mpc_leg_force_torque(hip_len, upper_len, lower_len, Rb, leg, q + 3 * leg, f_world + 3 * leg, ff, tau + 3 * leg);
        if (f_ff_out)
            for (int a = 0; a < 3; ++a) f_ff_out[3 * leg + a] = ff[a];
    }
}
