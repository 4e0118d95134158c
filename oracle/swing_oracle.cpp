// swing_oracle.cpp -- CPU ORACLE for the WALK-mode swing trajectory and the heuristic foothold.  TEST INFRASTRUCTURE ONLY.
//
//   qro_swing_bspline   qrFootBSplinePatternGenerator::SetParameters / UpdateSpline / GenerateTrajectory
//                       (/root/reference/quadruped/src/controllers/qr_foot_trajectory_generator.cpp:30-163) restated,
//                       evaluated with the reference's OWN vendored tinynurbs (extern/tinynurbs, header-only, compiled
//                       from where it lies; glm replaced by oracle/mini_glm): the spline mathematics is pinned on the
//                       reference's dependency, the generator glue is restated.
//   qro_foothold        qrFootholdPlanner::ComputeHeuristicFootHold (src/planner/qr_foothold_planner.cpp:112-240) restated.
// float arithmetic like the reference (this file is compiled with -ffp-contract=off).
#include <cmath>
#include <vector>

#include <tinynurbs/core/evaluate.h>

#include "qr_oracle.h"

namespace {
struct V3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
struct M3 { float m[3][3]; };
V3 mul(const M3& R, const V3& a) {
    V3 o;
    for (int i = 0; i < 3; ++i) o[i] = (R.m[i][0] * a[0] + R.m[i][1] * a[1]) + R.m[i][2] * a[2];
    return o;
}
M3 transpose(const M3& R) { M3 o; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[i][j] = R.m[j][i]; return o; }
M3 coordinateRotation(char axis, float theta) {   // utils/qr_se3.h:71-89
    const float s = std::sin(theta), c = std::cos(theta);
    M3 R;
    if (axis == 'X') R = {{{1, 0, 0}, {0, c, s}, {0, -s, c}}};
    else if (axis == 'Y') R = {{{c, 0, -s}, {0, 1, 0}, {s, 0, c}}};
    else R = {{{c, s, 0}, {-s, c, 0}, {0, 0, 1}}};
    return R;
}
}  // namespace

extern "C" int qro_swing_bspline(const float* initial_pos, const float* target_pos, float height, float duration,
                                 float initial_time, float time, float* pos, float* vel) {
    // constructor (:30-51)
    const std::vector<glm::vec3> tmpl = {glm::vec3(-10, 0, 0),   glm::vec3(-10.3, 0, 0.2), glm::vec3(-13, 0, 2),
                                         glm::vec3(-15, 0, 7),   glm::vec3(0, 0, 7.8),     glm::vec3(11, 0, 8),
                                         glm::vec3(10.5, 0, 4),  glm::vec3(10.2, 0, 1),    glm::vec3(10, 0, 0)};
    tinynurbs::Curve<float> crv;
    crv.control_points = tmpl;
    crv.knots = {0., 0., 0., 0., 0.3 / 6, 1.3 / 6, 2.5 / 6, 3.0 / 6, 4.0 / 6, 1, 1, 1, 1};
    crv.degree = 3;
    // SetParameters (:53-85)
    V3 step_delta;
    for (int k = 0; k < 3; ++k) step_delta[k] = target_pos[k] - initial_pos[k];
    const float dx = step_delta[0], dy = step_delta[1];
    const M3 RTheta = coordinateRotation('Z', std::atan2(dy, dx));
    V3 Tp = {{initial_pos[0], initial_pos[1], initial_pos[2]}};
    float target_appex = height;
    // UpdateSpline (:88-136)
    std::vector<glm::vec3>& cp = crv.control_points;
    target_appex *= 100.f;
    const V3 startPos = {{0.f, 0.f, 0.f}};
    V3 scaled = {{step_delta[0] * 100.f, step_delta[1] * 100.f, step_delta[2] * 100.f}};
    const V3 endPos = mul(RTheta, scaled);
    const float xRatio = std::abs(endPos[0] - startPos[0]) / 20.f;
    if (endPos[2] >= startPos[2]) {
        const float z_length_left = target_appex;
        const float z_length_right = target_appex - (endPos[2] - startPos[2]);
        const float zRatio = std::abs(z_length_left) / 8.f;
        const float x_mid = (endPos[0] + startPos[0]) / 2;
        const float z_offset = startPos[2];
        for (size_t i = 0; i < tmpl.size(); ++i) {
            cp[i].x = tmpl[i].x * xRatio + x_mid;
            cp[i].z = tmpl[i].z * zRatio + z_offset;
        }
        cp[8].z = endPos[2];
        cp[7].z = cp[8].z + tmpl[7].z / 8 * z_length_right;
        cp[6].z = cp[8].z + tmpl[6].z / 8 * z_length_right;
        cp[5].z = cp[8].z + tmpl[5].z / 8 * z_length_right;
    } else {
        const float z_length_left = target_appex - (startPos[2] - endPos[2]);
        const float z_length_right = target_appex;
        const float zRatio = std::abs(z_length_right) / 8.f;
        const float x_mid = (endPos[0] + startPos[0]) / 2;
        const float z_offset = endPos[2];
        for (size_t i = 0; i < tmpl.size(); ++i) {
            cp[i].x = tmpl[i].x * xRatio + x_mid;
            cp[i].z = tmpl[i].z * zRatio + z_offset;
        }
        cp[0].z = startPos[2];
        cp[1].z = cp[0].z + 0.2 / 8 * z_length_left;
        cp[2].z = cp[0].z + 2.0 / 8 * z_length_left;
        cp[3].z = cp[0].z + 7.0 / 8 * z_length_left;
    }
    // GenerateTrajectory (:139-163)
    const float dt = time - initial_time;
    if (dt < -1e-3 || dt >= duration + 1e-3) return 0;
    auto pv = tinynurbs::curveDerivatives(crv, 1, dt);
    V3 foot_pos = {{pv[0].x / 100, pv[0].y / 100, pv[0].z / 100}};
    V3 foot_vel = {{pv[1].x / 100, pv[1].y / 100, pv[1].z / 100}};
    const M3 Rt = transpose(RTheta);
    const V3 p = mul(Rt, foot_pos), v = mul(Rt, foot_vel);
    for (int k = 0; k < 3; ++k) { pos[k] = p[k] + Tp[k]; vel[k] = v[k]; }
    return 1;
}

extern "C" void qro_foothold(const qro_foothold_params* P, int legId, const float* com_vel, const float* w_, const float* dR_,
                             const float* base_R_, const float* rpy, const float* foot_base, const float* des_speed,
                             float desiredTwistingSpeed, float des_height, float swingRemainTime, int allowSwitch,
                             float normalizedPhase, float* foothold, float* phase) {
    M3 dR, robotBaseR;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { dR.m[i][j] = dR_[3 * i + j]; robotBaseR.m[i][j] = base_R_[3 * i + j]; }
    const V3 comVelocity = {{com_vel[0], com_vel[1], com_vel[2]}}, w = {{w_[0], w_[1], w_[2]}};
    const V3 desiredSpeed = {{des_speed[0], des_speed[1], des_speed[2]}};
    const V3 desiredHeight = {{0.f, 0.f, des_height}};
    const float side_sign[4] = {-1, 1, -1, 1};
    const V3 hipOffset = {{P->hip_offset[3 * legId], P->hip_offset[3 * legId + 1], P->hip_offset[3 * legId + 2]}};
    const V3 hipPos = {{P->hip_pos[3 * legId], P->hip_pos[3 * legId + 1], P->hip_pos[3 * legId + 2]}};
    const V3 twistingVector = {{-hipOffset[1], hipOffset[0], 0.f}};
    V3 cross = {{w[1] * hipOffset[2] - w[2] * hipOffset[1], w[2] * hipOffset[0] - w[0] * hipOffset[2], w[0] * hipOffset[1] - w[1] * hipOffset[0]}};
    V3 hipHorizontalVelocity = {{comVelocity[0] + cross[0], comVelocity[1] + cross[1], comVelocity[2] + cross[2]}};
    hipHorizontalVelocity = mul(dR, hipHorizontalVelocity);
    hipHorizontalVelocity[2] = 0.f;
    V3 targetHip;
    for (int k = 0; k < 3; ++k) targetHip[k] = desiredSpeed[k] + desiredTwistingSpeed * twistingVector[k];
    const float hipLen = P->hip_len;
    V3 footTargetPosition;
    if (!allowSwitch) {
        V3 d;
        for (int k = 0; k < 3; ++k) d[k] = foot_base[3 * legId + k] - hipPos[k];
        footTargetPosition = mul(robotBaseR, d);
        if (footTargetPosition[1] > 0.01 + 0.00 * (-side_sign[legId])) footTargetPosition[1] -= 0.005;
        else if (footTargetPosition[1] < -0.01 + 0.00 * side_sign[legId]) footTargetPosition[1] += 0.005;
        footTargetPosition[2] -= 0.02;
        footTargetPosition = mul(transpose(robotBaseR), footTargetPosition);
        for (int k = 0; k < 3; ++k) footTargetPosition[k] = footTargetPosition[k] + hipPos[k];
        *phase = 1.0f;
    } else {
        V3 inner;
        for (int k = 0; k < 3; ++k)
            inner[k] = targetHip[k] * swingRemainTime - P->swing_kp[k] * (targetHip[k] - hipHorizontalVelocity[k]);
        V3 dP = mul(transpose(dR), inner);
        const float thr = 0.2f;
        dP[0] = dP[0] < -thr ? -thr : (dP[0] > thr ? thr : dP[0]);
        dP[1] = dP[1] < -thr ? -thr : (dP[1] > thr ? thr : dP[1]);
        dP[2] = 0;
        const float interleave_y = hipLen * side_sign[legId];
        const M3 rollR = coordinateRotation('X', rpy[0]);
        const V3 off = mul(rollR, V3{{0, interleave_y, 0}});
        const V3 ho = {{hipOffset[0], hipOffset[1], 0}};
        for (int k = 0; k < 3; ++k) footTargetPosition[k] = (dP[k] + ho[k]) + off[k];
        const V3 sub = mul(transpose(robotBaseR), desiredHeight);
        for (int k = 0; k < 3; ++k) footTargetPosition[k] = footTargetPosition[k] - sub[k];
        *phase = normalizedPhase;
    }
    for (int k = 0; k < 3; ++k) foothold[3 * legId + k] = footTargetPosition[k];
}
