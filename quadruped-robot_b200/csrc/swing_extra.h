// swing_extra.h -- swing-leg targets next to the hot path, float32, one thread per
// foot.  Paths relative to /root/reference/quadruped/.
//
//   qr_swing_bspline       qrFootBSplinePatternGenerator::SetParameters / UpdateSpline / GenerateTrajectory
//                          (src/controllers/qr_foot_trajectory_generator.cpp:30-163) on top of tinynurbs'
//                          curveDerivatives (extern/tinynurbs/include/tinynurbs/core/evaluate.h:66-100) =
//                          findSpan + bsplineDerBasis (core/basis.h:25-66, 163-272; The NURBS Book A2.1 / A2.3),
//                          degree 3, 9 control points, the knot vector of :45-48
//   qr_foothold_heuristic  qrFootholdPlanner::ComputeHeuristicFootHold (src/planner/qr_foothold_planner.cpp:112-240),
//                          both branches (leg allowed to switch state or not)
// Every float32 operation is issued in the reference's order (QR_FMUL / QR_FADD: no FMA contraction); the only
// library calls are atan2f / sinf / cosf, whose last-bit rounding differs between glibc and CUDA.
#pragma once

#include "qr_team.h"

QR_DEV int qr_bspline_find_span(const float* U, float u) {   // degree 3, 13 knots: n = 8
    const float eps = 1.1920929e-07f;
    if (u > QR_FSUB(U[9], eps)) return 8;
    if (u < QR_FADD(U[3], eps)) return 3;
    int low = 3, high = 9;
    int mid = (low + high) / 2;
    while (u < U[mid] || u >= U[mid + 1]) {
        if (u < U[mid]) high = mid; else low = mid;
        mid = (low + high) / 2;
    }
    return mid;
}

// Non-zero basis functions (row 0) and their first derivatives (row 1) at u: bsplineDerBasis(3, span, U, u, 1).
QR_DEV void qr_bspline_der_basis(const float* U, int span, float u, float ders[2][4]) {
    float ndu[4][4], left[4], right[4], a[2][4];
    ndu[0][0] = 1.f;
    for (int j = 1; j <= 3; ++j) {
        left[j] = QR_FSUB(u, U[span + 1 - j]);
        right[j] = QR_FSUB(U[span + j], u);
        float saved = 0.f;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = QR_FADD(right[r + 1], left[j - r]);
            const float temp = QR_FDIV(ndu[r][j - 1], ndu[j][r]);
            ndu[r][j] = QR_FADD(saved, QR_FMUL(right[r + 1], temp));
            saved = QR_FMUL(left[j - r], temp);
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= 3; ++j) ders[0][j] = ndu[j][3];
    for (int r = 0; r <= 3; ++r) {
        const int s1 = 0, s2 = 1;   // one derivative: no swap needed
        a[0][0] = 1.f;
        float d = 0.f;
        const int rk = r - 1, pk = 2;
        if (r >= 1) {
            a[s2][0] = QR_FDIV(a[s1][0], ndu[pk + 1][rk]);
            d = QR_FMUL(a[s2][0], ndu[rk][pk]);
        }
        const int j1 = rk >= -1 ? 1 : -rk;
        const int j2 = (r - 1 <= pk) ? 0 : 3 - r;
        for (int j = j1; j <= j2; ++j) {
            a[s2][j] = QR_FDIV(QR_FSUB(a[s1][j], a[s1][j - 1]), ndu[pk + 1][rk + j]);
            d = QR_FADD(d, QR_FMUL(a[s2][j], ndu[rk + j][pk]));
        }
        if (r <= pk) {
            a[s2][1] = QR_FDIV(-a[s1][0], ndu[pk + 1][r]);
            d = QR_FADD(d, QR_FMUL(a[s2][1], ndu[r][pk]));
        }
        ders[1][r] = QR_FMUL(d, 3.f);
    }
}

// One trajectory sample.  Returns 0 when GenerateTrajectory rejects the time (pos / vel untouched).
QR_DEV int qr_swing_bspline(const float* initial_pos, const float* target_pos, float height, float duration,
                            float initial_time, float time, float* pos, float* vel) {
    const float tx[9] = {-10.f, -10.3f, -13.f, -15.f, 0.f, 11.f, 10.5f, 10.2f, 10.f};
    const float tz[9] = {0.f, 0.2f, 2.f, 7.f, 7.8f, 8.f, 4.f, 1.f, 0.f};
    const float U[13] = {0.f, 0.f, 0.f, 0.f, (float)(0.3 / 6), (float)(1.3 / 6), (float)(2.5 / 6), (float)(3.0 / 6),
                         (float)(4.0 / 6), 1.f, 1.f, 1.f, 1.f};
    // ---- SetParameters (:53-85)
    const float sd0 = QR_FSUB(target_pos[0], initial_pos[0]), sd1 = QR_FSUB(target_pos[1], initial_pos[1]);
    const float sd2 = QR_FSUB(target_pos[2], initial_pos[2]);
    const float theta = atan2f(sd1, sd0);
    const float s = sinf(theta), c = cosf(theta);   // RTheta = [c s 0; -s c 0; 0 0 1]
    // ---- UpdateSpline (:88-136): everything in centimetres
    const float appex = QR_FMUL(height, 100.f);
    const float v0 = QR_FMUL(sd0, 100.f), v1 = QR_FMUL(sd1, 100.f), v2 = QR_FMUL(sd2, 100.f);
    const float e0 = QR_FADD(QR_FADD(QR_FMUL(c, v0), QR_FMUL(s, v1)), QR_FMUL(0.f, v2));
    const float e2 = QR_FADD(QR_FADD(QR_FMUL(0.f, v0), QR_FMUL(0.f, v1)), QR_FMUL(1.f, v2));
    const float xRatio = QR_FDIV(fabsf(QR_FSUB(e0, 0.f)), 20.f);
    const float x_mid = QR_FDIV(QR_FADD(e0, 0.f), 2.f);
    float cx[9], cz[9];
    if (e2 >= 0.f) {   // walk up
        const float z_left = appex, z_right = QR_FSUB(appex, QR_FSUB(e2, 0.f));
        const float zRatio = QR_FDIV(fabsf(z_left), 8.f);
        for (int i = 0; i < 9; ++i) {
            cx[i] = QR_FADD(QR_FMUL(tx[i], xRatio), x_mid);
            cz[i] = QR_FADD(QR_FMUL(tz[i], zRatio), 0.f);
        }
        cz[8] = e2;
        cz[7] = QR_FADD(cz[8], QR_FMUL(QR_FDIV(tz[7], 8.f), z_right));
        cz[6] = QR_FADD(cz[8], QR_FMUL(QR_FDIV(tz[6], 8.f), z_right));
        cz[5] = QR_FADD(cz[8], QR_FMUL(QR_FDIV(tz[5], 8.f), z_right));
    } else {           // walk down
        const float z_left = QR_FSUB(appex, QR_FSUB(0.f, e2)), z_right = appex;
        const float zRatio = QR_FDIV(fabsf(z_right), 8.f);
        for (int i = 0; i < 9; ++i) {
            cx[i] = QR_FADD(QR_FMUL(tx[i], xRatio), x_mid);
            cz[i] = QR_FADD(QR_FMUL(tz[i], zRatio), e2);
        }
        cz[0] = 0.f;   // the double literals of :131-134 (0.2/8, 2.0/8, 7.0/8), product in double, narrowed on assignment
        cz[1] = (float)QR_DADD((double)cz[0], QR_DMUL(0.2 / 8, (double)z_left));
        cz[2] = (float)QR_DADD((double)cz[0], QR_DMUL(2.0 / 8, (double)z_left));
        cz[3] = (float)QR_DADD((double)cz[0], QR_DMUL(7.0 / 8, (double)z_left));
    }
    // ---- GenerateTrajectory (:139-163)
    const float dt = QR_FSUB(time, initial_time);
    if ((double)dt < -1e-3 || (double)dt >= (double)duration + 1e-3) return 0;
    const int span = qr_bspline_find_span(U, dt);
    float ders[2][4];
    qr_bspline_der_basis(U, span, dt, ders);
    float pv[2][3];
    for (int k = 0; k < 2; ++k) {
        float px = 0.f, py = 0.f, pz = 0.f;
        for (int j = 0; j <= 3; ++j) {
            px = QR_FADD(px, QR_FMUL(ders[k][j], cx[span - 3 + j]));
            py = QR_FADD(py, QR_FMUL(ders[k][j], 0.f));
            pz = QR_FADD(pz, QR_FMUL(ders[k][j], cz[span - 3 + j]));
        }
        pv[k][0] = QR_FDIV(px, 100.f); pv[k][1] = QR_FDIV(py, 100.f); pv[k][2] = QR_FDIV(pz, 100.f);
    }
    // foot_pos = RTheta^T foot_pos + Tp ; foot_vel = RTheta^T foot_vel
    for (int k = 0; k < 2; ++k) {
        const float a0 = pv[k][0], a1 = pv[k][1], a2 = pv[k][2];
        const float r0 = QR_FADD(QR_FADD(QR_FMUL(c, a0), QR_FMUL(-s, a1)), QR_FMUL(0.f, a2));
        const float r1 = QR_FADD(QR_FADD(QR_FMUL(s, a0), QR_FMUL(c, a1)), QR_FMUL(0.f, a2));
        const float r2 = QR_FADD(QR_FADD(QR_FMUL(0.f, a0), QR_FMUL(0.f, a1)), QR_FMUL(1.f, a2));
        if (k == 0) {
            pos[0] = QR_FADD(r0, initial_pos[0]); pos[1] = QR_FADD(r1, initial_pos[1]); pos[2] = QR_FADD(r2, initial_pos[2]);
        } else {
            vel[0] = r0; vel[1] = r1; vel[2] = r2;
        }
    }
    return 1;
}

// Constants of the robot the foothold planner reads (qrRobot::hipOffset, GetDefaultHipPosition(), hipLength) and the
// swing gain swingKp (qr_foothold_planner.h).  3x4 matrices are column-major like Eigen's: m[3*leg + axis].
struct QrFootholdParams {
    float hip_offset[12];
    float hip_pos[12];
    float hip_len;
    float swing_kp[3];
};

QR_DEV void qr_mat3_vec(const float* R, const float* v, float* o) {           // R row-major
    for (int i = 0; i < 3; ++i) o[i] = QR_FADD(QR_FADD(QR_FMUL(R[3 * i], v[0]), QR_FMUL(R[3 * i + 1], v[1])), QR_FMUL(R[3 * i + 2], v[2]));
}
QR_DEV void qr_mat3t_vec(const float* R, const float* v, float* o) {          // R^T v
    for (int i = 0; i < 3; ++i) o[i] = QR_FADD(QR_FADD(QR_FMUL(R[i], v[0]), QR_FMUL(R[3 + i], v[1])), QR_FMUL(R[6 + i], v[2]));
}

// One swing leg.  Inputs are this robot's rows (see qr_gpu_foothold_heuristic_batch in include/qr_gpu.h).
QR_DEV void qr_foothold_heuristic(const QrFootholdParams& P, int leg, const float* com_vel, const float* w, const float* dR,
                                  const float* base_R, const float* rpy, const float* foot_base, const float* des_speed,
                                  float des_twist, float des_height, float swing_remain, int allow_switch,
                                  float norm_phase, float* foothold, float* phase) {
    const float side_sign = (leg % 2 == 0) ? -1.f : 1.f;
    const float* hip = P.hip_offset + 3 * leg;
    float target[3];
    if (!allow_switch) {
        // the leg keeps its place, slightly pulled in and pushed down (:161-183)
        float d[3], t[3];
        for (int k = 0; k < 3; ++k) d[k] = QR_FSUB(foot_base[3 * leg + k], P.hip_pos[3 * leg + k]);
        qr_mat3_vec(base_R, d, t);
        // the thresholds and offsets are double literals in the reference (0.01 + 0.00 * (-side_sign) etc.)
        if ((double)t[1] > 0.01 + 0.00 * (double)(-side_sign)) t[1] = (float)((double)t[1] - 0.005);
        else if ((double)t[1] < -0.01 + 0.00 * (double)side_sign) t[1] = (float)((double)t[1] + 0.005);
        t[2] = (float)((double)t[2] - 0.02);
        qr_mat3t_vec(base_R, t, target);
        for (int k = 0; k < 3; ++k) target[k] = QR_FADD(target[k], P.hip_pos[3 * leg + k]);
        *phase = 1.0f;
    } else {
        // hip velocity in the control frame, horizontal part (:143-147)
        const float cr[3] = {QR_FSUB(QR_FMUL(w[1], hip[2]), QR_FMUL(w[2], hip[1])), QR_FSUB(QR_FMUL(w[2], hip[0]), QR_FMUL(w[0], hip[2])),
                             QR_FSUB(QR_FMUL(w[0], hip[1]), QR_FMUL(w[1], hip[0]))};
        const float hv_b[3] = {QR_FADD(com_vel[0], cr[0]), QR_FADD(com_vel[1], cr[1]), QR_FADD(com_vel[2], cr[2])};
        float hv[3];
        qr_mat3_vec(dR, hv_b, hv);
        hv[2] = 0.f;
        const float twist[3] = {-hip[1], hip[0], 0.f};
        float tv[3], inner[3], dP[3];
        for (int k = 0; k < 3; ++k) tv[k] = QR_FADD(des_speed[k], QR_FMUL(des_twist, twist[k]));
        // dP = dR^T (tv * swingRemainTime - swingKp .* (tv - hv))   (:198-201)
        for (int k = 0; k < 3; ++k) inner[k] = QR_FSUB(QR_FMUL(tv[k], swing_remain), QR_FMUL(P.swing_kp[k], QR_FSUB(tv[k], hv[k])));
        qr_mat3t_vec(dR, inner, dP);
        const float thr = 0.2f;
        dP[0] = dP[0] < -thr ? -thr : (dP[0] > thr ? thr : dP[0]);
        dP[1] = dP[1] < -thr ? -thr : (dP[1] > thr ? thr : dP[1]);
        dP[2] = 0.f;
        // abad -> hip offset rotated by the roll angle: rollR = coordinateRotation(X, rpy[0]) (:211-218)
        const float sr = sinf(rpy[0]), cr_ = cosf(rpy[0]);
        const float iy = QR_FMUL(P.hip_len, side_sign);
        const float ro[3] = {QR_FADD(QR_FADD(QR_FMUL(1.f, 0.f), QR_FMUL(0.f, iy)), QR_FMUL(0.f, 0.f)),
                             QR_FADD(QR_FADD(QR_FMUL(0.f, 0.f), QR_FMUL(cr_, iy)), QR_FMUL(sr, 0.f)),
                             QR_FADD(QR_FADD(QR_FMUL(0.f, 0.f), QR_FMUL(-sr, iy)), QR_FMUL(cr_, 0.f))};
        const float ho[3] = {hip[0], hip[1], 0.f};
        for (int k = 0; k < 3; ++k) target[k] = QR_FADD(QR_FADD(dP[k], ho[k]), ro[k]);
        // (:219-222 write footTargetPosition(0,2) / (0,3) of a 3-vector: out of range in the reference, not mirrored)
        const float dh[3] = {0.f, 0.f, des_height};
        float sub[3];
        qr_mat3t_vec(base_R, dh, sub);
        for (int k = 0; k < 3; ++k) target[k] = QR_FSUB(target[k], sub[k]);
        *phase = norm_phase;
    }
    for (int k = 0; k < 3; ++k) foothold[3 * leg + k] = target[k];
}
