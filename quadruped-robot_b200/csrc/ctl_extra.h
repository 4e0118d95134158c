// ctl_extra.h -- the controller arithmetic either side of the MPC / WBC kernels that the reference does per robot on the
// host, one thread per robot (or per leg).  float32 in the reference's operation order (no FMA contraction); the
// transcendental functions are evaluated in float64 and rounded once (see mpc_io.h).
//
//   qr_base_rmat            stateDataFlow.baseRMat = quaternionToRotationMatrix(q)^T   (src/robots/qr_robot.cpp:70,
//                           include/quadruped/utils/qr_se3.h:186-203)
//   qr_mpc_lever_arms       foot2ComInWorldFrame = baseRMat * (footPosInBaseFrame.colwise() - comOffset)
//                           (src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:396)
//   qr_leg_fk / qr_leg_jacobian / qr_leg_ik / qr_leg_ik_velocity
//                           qrRobot::FootPositionInHipFrame, FootPositionsInBaseFrame, AnalyticalLegJacobian,
//                           FootPositionInHipFrameToJointAngle, ComputeMotorAnglesFromFootLocalPosition,
//                           ComputeMotorVelocityFromFootLocalVelocity (src/robots/qr_robot.cpp:106-226)
//   qr_swing_targets_leg    case ADVANCED_TROT of qrRaibertSwingLegController::GetAction
//                           (src/controllers/qr_swing_leg_controller.cpp:361-409, joint targets :407-410): the
//                           pFoot_des / vFoot_des / aFoot_des rows of qrWbcCtrlData and the swing-leg joint targets
//   qr_gait_update          qrOpenLoopGaitGenerator::Update + Schedule (src/gait/qr_openloop_gait_generator.cpp:126-247)
#pragma once

#include "mpc_io.h"
#include "wbc_problem.h"   // qr_swing_parabola
#include "swing_extra.h"   // qr_mat3_vec, qr_mat3t_vec (dense 3x3 products in ascending column order)

#define QR_M(a, b) QR_FMUL(a, b)
#define QR_A(a, b) QR_FADD(a, b)
#define QR_S(a, b) QR_FSUB(a, b)
#define QR_D(a, b) QR_FDIV(a, b)

QR_DEV float qr_acosf_cr(float x) { return (float)acos((double)x); }
QR_DEV float qr_asinf_cr(float x) { return (float)asin((double)x); }
QR_DEV float qr_atan2f_cr(float y, float x) { return (float)atan2((double)y, (double)x); }

QR_DEV void qr_base_rmat(const float* quat, float* Rb) { qr_mpc_base_rmat(quat, Rb); }

// quat (w,x,y,z), foot_base[12] = footPosInBaseFrame (3x4 column-major: [3*leg + axis]), com_offset[3] -> r_feet[12]
QR_DEV void qr_mpc_lever_arms(const float* quat, const float* foot_base, const float* com_offset, float* r_feet) {
    float Rb[9];
    qr_base_rmat(quat, Rb);
    for (int leg = 0; leg < 4; ++leg) {
        float d[3];
        for (int a = 0; a < 3; ++a) d[a] = QR_S(foot_base[3 * leg + a], com_offset[a]);
        qr_mat3_vec(Rb, d, r_feet + 3 * leg);
    }
}

struct QrLegGeom {
    float hip_len, upper_len, lower_len;
    float hip_offset[12];   // qrRobot::hipOffset, 3x4 column-major
};

// FootPositionInHipFrame (qr_robot.cpp:125-145) + hipOffset column (FootPositionsInBaseFrame :175-184)
QR_DEV void qr_leg_fk(const QrLegGeom& G, int leg, const float* t, float* foot_base) {
    const float sh = (leg & 1) ? G.hip_len : -G.hip_len;   // hipLength * pow(-1, leg + 1)
    const float uu = QR_M(G.upper_len, G.upper_len), ll = QR_M(G.lower_len, G.lower_len);
    const float ld = qr_sqrtf_cr(QR_A(QR_A(uu, ll), QR_M(QR_M(QR_M(2.f, G.upper_len), G.lower_len), qr_cosf_cr(t[2]))));
    const float eff = QR_A(t[1], QR_D(t[2], 2.f));
    const float ox = QR_M(-ld, qr_sinf_cr(eff));
    const float ozh = QR_M(-ld, qr_cosf_cr(eff));
    const float oyh = sh;
    const float c0 = qr_cosf_cr(t[0]), s0 = qr_sinf_cr(t[0]);
    const float oy = QR_S(QR_M(c0, oyh), QR_M(s0, ozh));
    const float oz = QR_A(QR_M(s0, oyh), QR_M(c0, ozh));
    foot_base[0] = QR_A(ox, G.hip_offset[3 * leg]);
    foot_base[1] = QR_A(oy, G.hip_offset[3 * leg + 1]);
    foot_base[2] = QR_A(oz, G.hip_offset[3 * leg + 2]);
}

// AnalyticalLegJacobian (qr_robot.cpp:148-172), row-major
QR_DEV void qr_leg_jacobian(const QrLegGeom& G, int leg, const float* t, float* J) {
    const float sh = (leg & 1) ? G.hip_len : -G.hip_len;
    const float s0 = qr_sinf_cr(t[0]), c0 = qr_cosf_cr(t[0]), s2 = qr_sinf_cr(t[2]), c2 = qr_cosf_cr(t[2]);
    const float uu = QR_M(G.upper_len, G.upper_len), ll = QR_M(G.lower_len, G.lower_len);
    const float lEff = qr_sqrtf_cr(QR_A(QR_A(uu, ll), QR_M(QR_M(QR_M(2.f, G.upper_len), G.lower_len), c2)));
    const float tEff = QR_A(t[1], QR_D(t[2], 2.f));
    const float sE = qr_sinf_cr(tEff), cE = qr_cosf_cr(tEff);
    const float lu = QR_M(G.lower_len, G.upper_len);
    J[0] = 0.f;
    J[1] = QR_M(-lEff, cE);
    J[2] = QR_S(QR_D(QR_M(QR_M(lu, s2), sE), lEff), QR_D(QR_M(lEff, cE), 2.f));
    J[3] = QR_A(QR_M(-sh, s0), QR_M(QR_M(lEff, c0), cE));
    J[4] = QR_M(QR_M(-lEff, s0), sE);
    J[5] = QR_S(QR_D(QR_M(QR_M(QR_M(-lu, s0), s2), cE), lEff), QR_D(QR_M(QR_M(lEff, s0), sE), 2.f));
    J[6] = QR_A(QR_M(sh, c0), QR_M(QR_M(lEff, s0), cE));
    J[7] = QR_M(QR_M(lEff, sE), c0);
    J[8] = QR_A(QR_D(QR_M(QR_M(QR_M(lu, s2), c0), cE), lEff), QR_D(QR_M(QR_M(lEff, sE), c0), 2.f));
}

// ComputeMotorAnglesFromFootLocalPosition -> FootPositionInHipFrameToJointAngle (qr_robot.cpp:106-122, 200-208)
QR_DEV void qr_leg_ik(const QrLegGeom& G, int leg, const float* foot_local, float* t) {
    const float sh = (leg & 1) ? G.hip_len : -G.hip_len;
    const float x = QR_S(foot_local[0], G.hip_offset[3 * leg]), y = QR_S(foot_local[1], G.hip_offset[3 * leg + 1]),
                z = QR_S(foot_local[2], G.hip_offset[3 * leg + 2]);
    const float n2 = QR_A(QR_A(QR_M(x, x), QR_M(y, y)), QR_M(z, z));
    const float l2 = QR_A(QR_A(QR_M(sh, sh), QR_M(G.upper_len, G.upper_len)), QR_M(G.lower_len, G.lower_len));
    const float knee = -qr_acosf_cr(QR_D(QR_S(n2, l2), QR_M(QR_M(2.f, G.lower_len), G.upper_len)));
    const float l = qr_sqrtf_cr(QR_A(QR_A(QR_M(G.upper_len, G.upper_len), QR_M(G.lower_len, G.lower_len)),
                                     QR_M(QR_M(QR_M(2.f, G.upper_len), G.lower_len), qr_cosf_cr(knee))));
    const float hip = QR_S(qr_asinf_cr(QR_D(-x, l)), QR_D(knee, 2.f));
    const float ce = qr_cosf_cr(QR_A(hip, QR_D(knee, 2.f)));
    const float c1 = QR_S(QR_M(sh, y), QR_M(QR_M(l, ce), z));
    const float s1 = QR_A(QR_M(QR_M(l, ce), y), QR_M(sh, z));
    t[0] = qr_atan2f_cr(s1, c1);
    t[1] = hip;
    t[2] = knee;
}

// ComputeMotorVelocityFromFootLocalVelocity: AnalyticalLegJacobian(angles).inverse() * v (qr_robot.cpp:211-218); the 3x3
// inverse by cofactors over the determinant expanded along the first column, as Eigen evaluates fixed 3x3 inverses
QR_DEV void qr_leg_ik_velocity(const QrLegGeom& G, int leg, const float* t, const float* v, float* qd) {
    float J[9], inv[9];
    qr_leg_jacobian(G, leg, t, J);
#define QR_COF(i, j)                                                                                                  \
    QR_S(QR_M(J[3 * (((i) + 1) % 3) + ((j) + 1) % 3], J[3 * (((i) + 2) % 3) + ((j) + 2) % 3]),                        \
         QR_M(J[3 * (((i) + 1) % 3) + ((j) + 2) % 3], J[3 * (((i) + 2) % 3) + ((j) + 1) % 3]))
    const float c0 = QR_COF(0, 0), c1 = QR_COF(1, 0), c2 = QR_COF(2, 0);
    const float det = QR_A(QR_A(QR_M(c0, J[0]), QR_M(c1, J[3])), QR_M(c2, J[6]));
    const float invdet = QR_D(1.f, det);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) inv[3 * j + i] = QR_M(QR_COF(i, j), invdet);
#undef QR_COF
    qr_mat3_vec(inv, v, qd);
}

// p_w = t + R(q) p with R(q) as Eigen::Quaternion::toRotationMatrix evaluates it (robotics::math::invertRigidTransform,
// qr_se3.h:443-449: Isometry3 = translate(t) * rotate(quat))
QR_DEV void qr_invert_rigid_transform(const float* t, const float* quat, const float* p, float* o) {
    const float w = quat[0], x = quat[1], y = quat[2], z = quat[3];
    const float tx = QR_M(2.f, x), ty = QR_M(2.f, y), tz = QR_M(2.f, z);
    const float twx = QR_M(tx, w), twy = QR_M(ty, w), twz = QR_M(tz, w);
    const float txx = QR_M(tx, x), txy = QR_M(ty, x), txz = QR_M(tz, x);
    const float tyy = QR_M(ty, y), tyz = QR_M(tz, y), tzz = QR_M(tz, z);
    const float R[9] = {QR_S(1.f, QR_A(tyy, tzz)), QR_S(txy, twz), QR_A(txz, twy),
                        QR_A(txy, twz), QR_S(1.f, QR_A(txx, tzz)), QR_S(tyz, twx),
                        QR_S(txz, twy), QR_A(tyz, twx), QR_S(1.f, QR_A(txx, tyy))};
    float r[3];
    qr_mat3_vec(R, p, r);
    for (int a = 0; a < 3; ++a) o[a] = QR_A(r[a], t[a]);
}

// One swing leg in MPC mode.  foothold[3] = footholdPlanner->desiredFootholds.col(leg) (base frame), phase =
// footholdPlanner->phase[leg], switch_pos[3] = phaseSwitchFootGlobalPos.col(leg).  horizontal_terrain != 0: robotBaseR is
// the identity (qr_swing_leg_controller.cpp:262-265).  Returns 0 when the trajectory generator rejects the phase
// (outputs untouched).
QR_DEV int qr_swing_targets_leg(const QrLegGeom& G, int leg, const float* base_pos, const float* quat, const float* v_world,
                                const float* foothold, float phase, const float* switch_pos, float swing_duration,
                                int horizontal_terrain, float* p_foot_des, float* v_foot_des, float* a_foot_des,
                                float* foot_base_des, float* q_des, float* qd_des) {
    float Rb[9];
    if (horizontal_terrain) {
        for (int e = 0; e < 9; ++e) Rb[e] = (e % 4 == 0) ? 1.f : 0.f;
    } else {
        qr_base_rmat(quat, Rb);
    }
    float end_w[3], pos_w[3];
    qr_mat3_vec(Rb, foothold, end_w);
    if (!qr_swing_parabola(switch_pos, end_w, 0.1f, phase, 0, pos_w)) return 0;
    float pos_b[3], vel_b[3] = {0.f, 0.f, 0.f};
    qr_mat3t_vec(Rb, pos_w, pos_b);
    if ((double)phase < 1.0) {
        const float zero[3] = {0.f, 0.f, 0.f};   // the parabola generator's velocity output is identically 0
        qr_mat3t_vec(Rb, zero, vel_b);
        for (int a = 0; a < 3; ++a) vel_b[a] = QR_D(vel_b[a], swing_duration);
    }
    qr_invert_rigid_transform(base_pos, quat, pos_b, p_foot_des);
    float rv[3];
    qr_mat3_vec(Rb, vel_b, rv);
    for (int a = 0; a < 3; ++a) {
        v_foot_des[a] = QR_A(v_world[a], rv[a]);
        a_foot_des[a] = 0.f;   // robotBaseR * footAccInBaseFrame with footAccInBaseFrame = 0
        if (foot_base_des) foot_base_des[a] = pos_b[a];
    }
    if (q_des) {
        float t[3];
        qr_leg_ik(G, leg, pos_b, t);
        for (int a = 0; a < 3; ++a) q_des[a] = t[a];
        if (qd_des) qr_leg_ik_velocity(G, leg, t, vel_b, qd_des);
    }
    return 1;
}

// qrOpenLoopGaitGenerator::Update(currentTime) on caller-held state (one robot):
//   cfg [20]     per leg: initialLegPhase, fullCyclePeriod, initStateRadioInCycle, swingDuration, dutyFactor
//   istate [20]  in/out: curLegState[4], lastLegState[4], desiredLegState[4], legState[4], firstSwing | firstStance << 1
//   fstate [4]   in/out: resetTime, lastTime, cumDt, waitTime
//   out [12]     in/out: phaseInFullCycle[4], normalizedPhase[4], swingTimeRemaining[4]
//   allow [4]    out: allowSwitchLegState
// LegState: SWING 0, STANCE 1, EARLY_CONTACT 2, LOSE_CONTACT 3, USERDEFINED_SWING 4 (config/qr_enum_types.h:62-68).
QR_DEV void qr_gait_update(float current_time, const float* cfg, float contact_threshold, const int32_t* contacts, int stop,
                           int advanced_trot, int32_t* istate, float* fstate, float* out, int32_t* allow) {
    int32_t* cur = istate;
    int32_t* last = istate + 4;
    int32_t* des = istate + 8;
    int32_t* leg_state = istate + 12;
    int32_t* first = istate + 16;
    float time_since_reset = current_time;
    // ---- Schedule(currentTime), :211-247
    if (QR_A(fstate[0], cfg[1]) < time_since_reset) fstate[0] = time_since_reset;   // fullCyclePeriod[0] = cfg[5*0 + 1]
    time_since_reset = QR_S(time_since_reset, fstate[0]);
    for (int l = 0; l < 4; ++l) allow[l] = 1;
    if (advanced_trot) {
        int n_allowed = 4;
        for (int l = 0; l < 4; ++l)
            if (cur[l] == 0 && des[l] == 1 && !contacts[l]) { allow[l] = 0; --n_allowed; }
        if (n_allowed < 4) {
            const float dt_ = QR_S(current_time, fstate[1]);
            fstate[2] = QR_A(fstate[2], dt_);
            if (fstate[2] > fstate[3]) {
                for (int l = 0; l < 4; ++l) allow[l] = 1;
            } else {
                fstate[0] = QR_A(fstate[0], dt_);
            }
        } else {
            fstate[2] = 0.f;
        }
    }
    // ---- Update, :131-207
    const int all_allowed = allow[0] + allow[1] + allow[2] + allow[3] == 4;
    for (int l = 0; l < 4; ++l) {
        if (cur[l] == 4) continue;
        if (!all_allowed) continue;
        const float init_phase = cfg[5 * l], period = cfg[5 * l + 1], ratio = cfg[5 * l + 2], swing_dur = cfg[5 * l + 3];
        if (!stop || (stop && last[l] == 0)) {
            last[l] = cur[l];
            cur[l] = des[l];
        }
        const float augmented = QR_A(QR_M(init_phase, period), time_since_reset);
        const float ph = QR_D(fmodf(augmented, period), period);
        out[l] = ph;
        int first_swing = first[l] & 1, first_stance = (first[l] >> 1) & 1;
        if (ph < ratio) {
            des[l] = 1;
            out[4 + l] = QR_D(ph, ratio);
        } else {
            des[l] = 0;
            out[4 + l] = QR_D(QR_S(ph, ratio), QR_S(1.f, ratio));
            if (cur[l] == 1) {
                first_swing = 1;
                first_stance = 0;
                out[8 + l] = swing_dur;
            } else {
                first_swing = 0;
                out[8 + l] = QR_M(swing_dur, QR_S(1.f, out[4 + l]));
            }
        }
        first[l] = first_swing | (first_stance << 1);
        if (leg_state[l] == 2 && des[l] == 0) continue;
        leg_state[l] = des[l];
        if (out[4 + l] < contact_threshold) continue;
        if (leg_state[l] == 0 && contacts[l]) leg_state[l] = 2;
        if (cur[l] == 0 && (leg_state[l] == 2 || leg_state[l] == 1)) {
            first_stance = 1;
            first_swing = 0;
            first[l] = first_swing | (first_stance << 1);
        }
    }
    fstate[1] = current_time;
}

#undef QR_M
#undef QR_A
#undef QR_S
#undef QR_D
