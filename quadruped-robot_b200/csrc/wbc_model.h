// wbc_model.h -- constants of the floating-base rigid-body tree, built once on the host in float64.
//
// Replaces qrRobotA1Sim::BuildDynamicModel / qrRobotLite3Sim::BuildDynamicModel
// (/root/reference/quadruped/src/robots/qr_robot_a1_sim.cpp:176-345, qr_robot_lite3_sim.cpp:176-345 --
// the two are identical: A1 link masses / inertias / abad location for every robot, only the link
// lengths and the body box come from the robot's yaml) together with FloatingBaseModel::addBase /
// addBody / addGroundContactPoint (src/dynamics/floating_base_model.cpp:277-420) and the SpatialInertia
// constructors / flipAlongAxis of include/quadruped/dynamics/spatial.hpp.
//
// Tree: body 0 = floating base (reference id 5); leg l in FR, FL, RR, RL has links 1+3l (abad, X axis),
// 2+3l (hip, Y), 3+3l (knee, Y); every link has a rotor twin (gear ratio 1, mass 1e-8) that the
// reference carries through every recursion, so it is kept.  Spatial vectors are [angular; linear],
// X = [R 0; -R[r]x R].  The float literals of the reference are rounded to float first, then widened.
#pragma once

#include <cmath>
#include <cstring>

#include "../../include/qr_gpu.h"

// Flat device layout (doubles), row-major 6x6 blocks.
struct QrWbcModelDev {
    double Xtree[12][36];   // parent -> link joint frame (before the joint rotation)
    double Xrot[12][36];    // parent -> rotor frame
    double Ibody[13][36];   // spatial inertias: [0] base, [1+j] link j
    double Irot[12][36];
    double foot[4][3];      // foot contact point in the knee link frame
    double gravity[3];
    double max_fz;          // totalNonRotorMass * 9.81 (qr_single_contact.cpp:31)
    double mu;              // 0.4f (qr_single_contact.cpp:35)
    double w_fb, w_fr;      // WBIC cost weights 0.1f / 1 (qr_wbc_locomotion_controller.cpp:45-46)
    double kp_ori, kd_ori, kp_pos, kd_pos, kp_foot, kd_foot;   // :59-73
};

namespace qr_wbc_host {

inline void zero(double* m, int n) { for (int i = 0; i < n; ++i) m[i] = 0.0; }
inline void matmul(const double* A, const double* B, double* C, int m, int k, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
            C[i * n + j] = s;
        }
}
inline void skew(const double* r, double* S) {
    S[0] = 0; S[1] = -r[2]; S[2] = r[1]; S[3] = r[2]; S[4] = 0; S[5] = -r[0]; S[6] = -r[1]; S[7] = r[0]; S[8] = 0;
}
// createSXform(R, r) = [R 0; -R [r]x R]
inline void sxform(const double* R, const double* r, double* X) {
    zero(X, 36);
    double S[9], RS[9];
    skew(r, S);
    matmul(R, S, RS, 3, 3, 3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            X[6 * i + j] = R[3 * i + j];
            X[6 * (3 + i) + 3 + j] = R[3 * i + j];
            X[6 * (3 + i) + j] = -RS[3 * i + j];
        }
}
// SpatialInertia(mass, com, I)
inline void spatial_inertia(double mass, const double* com, const double* I, double* M) {
    double c[9], cct[9];
    skew(com, c);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += c[3 * i + k] * c[3 * j + k];
            cct[3 * i + j] = s;
        }
    zero(M, 36);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            M[6 * i + j] = I[3 * i + j] + mass * cct[3 * i + j];
            M[6 * i + 3 + j] = mass * c[3 * i + j];
            M[6 * (3 + i) + j] = mass * c[3 * j + i];
        }
    for (int i = 0; i < 3; ++i) M[6 * (3 + i) + 3 + i] = mass;
}
// SpatialInertia::flipAlongAxis(Y): reflect through the 4x4 pseudo-inertia
inline void flip_y(const double* M, double* O) {
    const double h[3] = {0.5 * (M[6 * 2 + 4] - M[6 * 1 + 5]), 0.5 * (M[6 * 0 + 5] - M[6 * 2 + 3]), 0.5 * (M[6 * 1 + 3] - M[6 * 0 + 4])};
    const double trace = M[0] + M[7] + M[14];
    double P[16];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P[4 * i + j] = (i == j ? 0.5 * trace : 0.0) - M[6 * i + j];
    for (int i = 0; i < 3; ++i) { P[4 * i + 3] = h[i]; P[12 + i] = h[i]; }
    P[15] = M[35];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if ((i == 1) != (j == 1)) P[4 * i + j] = -P[4 * i + j];
    const double tE = P[0] + P[5] + P[10];
    const double hh[3] = {P[3], P[7], P[11]};
    double S[9];
    skew(hh, S);
    zero(O, 36);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            O[6 * i + j] = (i == j ? tE : 0.0) - P[4 * i + j];
            O[6 * i + 3 + j] = S[3 * i + j];
            O[6 * (3 + i) + j] = S[3 * j + i];
        }
    for (int i = 0; i < 3; ++i) O[6 * (3 + i) + 3 + i] = P[15];
}
inline void coord_rot(int axis, double th, double* R) {   // coordinateRotation (utils/qr_se3.h:72-89)
    const double s = std::sin(th), c = std::cos(th);
    for (int i = 0; i < 9; ++i) R[i] = 0.0;
    R[0] = R[4] = R[8] = 1.0;
    if (axis == 0) { R[4] = c; R[5] = s; R[7] = -s; R[8] = c; }
    else if (axis == 1) { R[0] = c; R[2] = -s; R[6] = s; R[8] = c; }
    else { R[0] = c; R[1] = s; R[3] = -s; R[4] = c; }
}

// qr_wbc_model (include/qr_gpu.h) -> device constants.
inline void build(const qr_wbc_model* g, QrWbcModelDev* M) {
    auto F = [](double v) { return (double)(float)v; };
    std::memset(M, 0, sizeof(*M));
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double rotor_scalar = (double)(float)((float)1e-2 * 1e-6);
    double rotorI[9] = {rotor_scalar, 0, 0, 0, rotor_scalar, 0, 0, 0, rotor_scalar};
    double RY[9], RX[9], t1[9], rotorIX[9], rotorIY[9];
    coord_rot(1, (double)(float)(M_PI / 2), RY);
    coord_rot(0, (double)(float)(M_PI / 2), RX);
    auto rot_inertia = [&](const double* R, double* out) {
        matmul(R, rotorI, t1, 3, 3, 3);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += t1[3 * i + k] * R[3 * j + k];
                out[3 * i + j] = s;
            }
    };
    rot_inertia(RY, rotorIX);
    rot_inertia(RX, rotorIY);
    auto scaled = [&](const double (&v)[9], double* out) { for (int i = 0; i < 9; ++i) out[i] = F(v[i]) * F(1e-6); };
    const double abad_raw[9] = {469.2, -9.4, -0.342, -9.4, 807.5, -0.466, -0.342, -0.466, 552.9};
    const double hip_raw[9] = {5529, 4.825, 343.9, 4.825, 5139.3, 22.4, 343.9, 22.4, 1367.8};
    const double knee_raw[9] = {2998, 0, -141.2, 0, 3014, 0, -141.2, 0, 32.4};
    const double body_raw[9] = {15853, 0, 0, 0, 37799, 0, 0, 0, 45654};
    double abadI[9], hipI[9], kneeI[9], bodyI[9];
    scaled(abad_raw, abadI); scaled(hip_raw, hipI); scaled(knee_raw, kneeI); scaled(body_raw, bodyI);
    const double abadC[3] = {F(-0.0033), 0, 0}, hipC[3] = {F(-0.003237), F(-0.022327), F(-0.027326)};
    const double kneeC[3] = {F(0.006435), 0, F(-0.107)}, zero3[3] = {0, 0, 0};
    double abadS[36], hipS[36], kneeS[36], rotX[36], rotY[36], abadSf[36], hipSf[36], rotXf[36], rotYf[36];
    spatial_inertia(F(0.696), abadC, abadI, abadS);
    spatial_inertia(F(1.013), hipC, hipI, hipS);
    spatial_inertia(F(0.166), kneeC, kneeI, kneeS);
    spatial_inertia(F(1e-8), zero3, rotorIX, rotX);
    spatial_inertia(F(1e-8), zero3, rotorIY, rotY);
    flip_y(abadS, abadSf); flip_y(hipS, hipSf); flip_y(rotX, rotXf); flip_y(rotY, rotYf);
    spatial_inertia(6.0, zero3, bodyI, M->Ibody[0]);
    double Rz[9];
    coord_rot(2, (double)(float)M_PI, Rz);
    double side = -1.0, total = M->Ibody[0][35];
    for (int leg = 0; leg < 4; ++leg) {
        const double sx = leg < 2 ? 1.0 : -1.0, sy = (leg % 2 == 0) ? -1.0 : 1.0;   // WithLegSigns
        const int a = 3 * leg, hjoint = 3 * leg + 1, k = 3 * leg + 2;
        const double r_abad[3] = {sx * F(0.1805), sy * F(0.047), 0}, r_abad_rot[3] = {sx * F(0.14), sy * F(0.047), 0};
        const double r_hip[3] = {0, sy * (double)g->hip_len, 0}, r_hip_rot[3] = {0, sy * F(0.04), 0};
        const double r_knee[3] = {0, 0, -(double)g->upper_len};
        sxform(I3, r_abad, M->Xtree[a]);      sxform(I3, r_abad_rot, M->Xrot[a]);
        sxform(I3, r_hip, M->Xtree[hjoint]);  sxform(Rz, r_hip_rot, M->Xrot[hjoint]);
        sxform(I3, r_knee, M->Xtree[k]);      sxform(I3, zero3, M->Xrot[k]);
        std::memcpy(M->Ibody[1 + a], side < 0 ? abadSf : abadS, sizeof(abadS));
        std::memcpy(M->Irot[a], side < 0 ? rotXf : rotX, sizeof(rotX));
        std::memcpy(M->Ibody[1 + hjoint], side < 0 ? hipSf : hipS, sizeof(hipS));
        std::memcpy(M->Irot[hjoint], side < 0 ? rotYf : rotY, sizeof(rotY));
        std::memcpy(M->Ibody[1 + k], kneeS, sizeof(kneeS));   // the knee link is not flipped (:320)
        std::memcpy(M->Irot[k], side < 0 ? rotYf : rotY, sizeof(rotY));
        M->foot[leg][0] = 0; M->foot[leg][1] = side < 0 ? F(0.004) : -F(0.004); M->foot[leg][2] = -(double)g->lower_len;
        total += M->Ibody[1 + a][35] + M->Ibody[1 + hjoint][35] + M->Ibody[1 + k][35];
        side = -side;
    }
    M->gravity[0] = 0; M->gravity[1] = 0; M->gravity[2] = (double)(float)-9.81;
    // totalNonRotorMass() * (T)9.81 (qr_single_contact.cpp:31), float literal widened
    M->max_fz = total * (double)(float)9.81;
    M->mu = (double)0.4f;
    M->w_fb = (double)0.1f; M->w_fr = 1.0;
    M->kp_ori = 100; M->kd_ori = 10; M->kp_pos = 100; M->kd_pos = 10; M->kp_foot = 500; M->kd_foot = 10;
}

}  // namespace qr_wbc_host
