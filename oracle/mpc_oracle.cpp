// mpc_oracle.cpp -- CPU ORACLE for the convex-MPC path.  TEST INFRASTRUCTURE ONLY.
//
// Eigen-free float32 restatement of the reference's QP construction, feeding the reference's own
// vendored qpOASES 3.2.0 (oracle/_ref/libqpOASES.a).  Every function cites the reference lines it
// follows (paths relative to /root/reference/quadruped/).  See qr_oracle.h for the pinning status.
//
// Arithmetic convention: everything the reference computes in float32 is computed in float32 here,
// matrix products are plain sequential dot products (k ascending) with a separate multiply and add
// (this file is compiled with -ffp-contract=off and the x86-64 baseline ISA has no FMA).  The GPU
// engine reproduces exactly this operation order, which is what makes (H, g) comparable bit for bit.
#include "qr_oracle.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include <qpOASES.hpp>

namespace {

// Dense row-major float32 matrix with the handful of operations the MPC build needs.
struct MatF {
    int r = 0, c = 0;
    std::vector<float> a;
    MatF() {}
    MatF(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, 0.f) {}
    float& operator()(int i, int j) { return a[size_t(i) * c + j]; }
    float operator()(int i, int j) const { return a[size_t(i) * c + j]; }
};

// C = A * B, sequential k, float32 multiply then add.
MatF matmul(const MatF& A, const MatF& B) {
    MatF C(A.r, B.c);
    for (int i = 0; i < A.r; ++i)
        for (int j = 0; j < B.c; ++j) {
            float s = 0.f;
            for (int k = 0; k < A.c; ++k) {
                float prod = A(i, k) * B(k, j);
                s = s + prod;
            }
            C(i, j) = s;
        }
    return C;
}

// Eigen::Quaternionf(w,x,y,z).toRotationMatrix() -- the formula of Eigen/src/Geometry/Quaternion.h,
// used by SolveMPCKernel (controllers/mpc/qr_mpc_interface.cpp:344-351).  Row-major 3x3 out.
void quat_to_rot(const float* q /*w,x,y,z*/, float R[9]) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
}

// 3x3 inverse by cofactors / determinant (what Eigen's fixed-size inverse() does for 3x3),
// qr_mpc_interface.cpp:324.
void inv3(const float M[9], float Mi[9]) {
    const float c00 = M[4] * M[8] - M[5] * M[7];
    const float c01 = M[5] * M[6] - M[3] * M[8];
    const float c02 = M[3] * M[7] - M[4] * M[6];
    const float det = (M[0] * c00 + M[1] * c01) + M[2] * c02;
    const float id = 1.f / det;
    Mi[0] = c00 * id;
    Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id;
    Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    Mi[3] = c01 * id;
    Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id;
    Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    Mi[6] = c02 * id;
    Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id;
    Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return (a0 * b0 + a1 * b1) + a2 * b2;
}

// ComputeContinuousTimeStateSpaceMatrices, qr_mpc_interface.cpp:296-331.
// R = "yawRotMat" (the full body rotation, :350-351), Iw = R diag(I) R^T (:365).
void continuous_model(const float R[9], const float inertia[3], float mass, const float* r_feet,
                      MatF& A, MatF& B) {
    A = MatF(13, 13);
    B = MatF(13, 12);
    A(3, 9) = 1.f; A(4, 10) = 1.f; A(5, 11) = 1.f; A(11, 12) = 1.f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A(i, 6 + j) = R[3 * j + i];  // R^T

    // I_world = (R * diag) * R^T
    float RI[9], Iw[9], Iwi[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) RI[3 * i + j] = R[3 * i + j] * inertia[j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Iw[3 * i + j] = dot3(RI[3 * i], R[3 * j], RI[3 * i + 1], R[3 * j + 1], RI[3 * i + 2], R[3 * j + 2]);
    inv3(Iw, Iwi);
    const float minv = 1.f / mass;
    for (int b = 0; b < 4; ++b) {
        const float rx = r_feet[3 * b], ry = r_feet[3 * b + 1], rz = r_feet[3 * b + 2];
        // crossMatrix(r), utils/qr_se3.h:95-103
        const float S[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                B(6 + i, 3 * b + j) = dot3(Iwi[3 * i], S[j], Iwi[3 * i + 1], S[3 + j], Iwi[3 * i + 2], S[6 + j]);
        for (int i = 0; i < 3; ++i) B(9 + i, 3 * b + i) = minv;
    }
}

// ConvertToDiscreteQP, qr_mpc_interface.cpp:257-293.  M = dt*[A B;0 0] is nilpotent of index 3
// (SURVEY.md section 0 fact 4), so exp(M) = I + M + M*M/2 exactly; Eigen's Pade-3 branch evaluates
// the same rational function.  powers[k] = Adt^k by repeated left multiplication (:272-276).
void discretise(const MatF& A, const MatF& B, float dt, int h, MatF& Adt, MatF& Bdt, MatF& Aqp,
                MatF& Bqp) {
    MatF M(25, 25);
    for (int i = 0; i < 13; ++i) {
        for (int j = 0; j < 13; ++j) M(i, j) = dt * A(i, j);
        for (int j = 0; j < 12; ++j) M(i, 13 + j) = dt * B(i, j);
    }
    MatF M2 = matmul(M, M);
    MatF E(25, 25);
    for (int i = 0; i < 25; ++i)
        for (int j = 0; j < 25; ++j) {
            float e = (i == j ? 1.f : 0.f) + M(i, j);
            E(i, j) = e + 0.5f * M2(i, j);
        }
    Adt = MatF(13, 13);
    Bdt = MatF(13, 12);
    for (int i = 0; i < 13; ++i) {
        for (int j = 0; j < 13; ++j) Adt(i, j) = E(i, j);
        for (int j = 0; j < 12; ++j) Bdt(i, j) = E(i, 13 + j);
    }
    std::vector<MatF> powers(h + 1);
    powers[0] = MatF(13, 13);
    for (int i = 0; i < 13; ++i) powers[0](i, i) = 1.f;
    for (int k = 1; k <= h; ++k) powers[k] = matmul(Adt, powers[k - 1]);

    Aqp = MatF(13 * h, 13);
    Bqp = MatF(13 * h, 12 * h);
    std::vector<MatF> G(h);  // G[k] = Adt^k * Bdt, shared by every block on the k-th sub-diagonal
    for (int k = 0; k < h; ++k) G[k] = matmul(powers[k], Bdt);
    for (int r = 0; r < h; ++r) {
        for (int i = 0; i < 13; ++i)
            for (int j = 0; j < 13; ++j) Aqp(13 * r + i, j) = powers[r + 1](i, j);
        for (int c = 0; c <= r; ++c)
            for (int i = 0; i < 13; ++i)
                for (int j = 0; j < 12; ++j) Bqp(13 * r + i, 12 * c + j) = G[r - c](i, j);
    }
}

}  // namespace

extern "C" int qro_mpc_build(const qro_mpc_params* P, const float* p, const float* v,
                             const float* quat, const float* w, const float* r_feet,
                             const float* rpy, const float* traj, const float* gait, float* H,
                             float* g, float* ub, float* Aqp_out, float* Bqp_out) {
    const int h = P->horizon;
    if (h < 1) return -1;
    const int n = 12 * h, s = 13 * h, m = 20 * h;

    // x0 = [rpy, p, w, v, -9.8]  (SolveMPC, qr_mpc_interface.cpp:362)
    float x0[13] = {rpy[0], rpy[1], rpy[2], p[0], p[1], p[2], w[0], w[1], w[2], v[0], v[1], v[2], -9.8f};

    float R[9];
    quat_to_rot(quat, R);
    MatF A, B, Adt, Bdt, Aqp, Bqp;
    continuous_model(R, P->inertia, P->mass, r_feet, A, B);
    discretise(A, B, P->dt, h, Adt, Bdt, Aqp, Bqp);

    // full_weight, X_d, U_b  (:376-390).  ResizeQPMats pre-fills U_b with 5e10 (:219-229).
    float wfull[13];
    for (int i = 0; i < 12; ++i) wfull[i] = P->weights[i];
    wfull[12] = 0.f;
    std::vector<float> Xd(s, 0.f);
    for (int i = 0, k = 0; i < h; ++i) {
        for (int j = 0; j < 12; ++j) Xd[13 * i + j] = traj[12 * i + j];
        for (int j = 0; j < 4; ++j, ++k) {
            for (int c = 0; c < 4; ++c) ub[5 * k + c] = 5e10f;
            ub[5 * k + 4] = gait[4 * i + j] * P->f_max;
        }
    }

    // temp = 2 Bqp^T L, filled block-wise for j >= i only (:396-406); other blocks stay zero.
    MatF T(n, s);
    for (int i = 0; i < h; ++i)
        for (int j = i; j < h; ++j)
            for (int nn = 0; nn < 13; ++nn) {
                const float w2 = 2 * wfull[nn];
                for (int a = 0; a < 12; ++a) T(12 * i + a, 13 * j + nn) = Bqp(13 * j + nn, 12 * i + a) * w2;
            }

    // qH = temp*Bqp + (2 alpha) I ;  qg = temp*(Aqp*x0 - X_d)   (:411-412)
    MatF TB = matmul(T, Bqp);
    const float two_alpha = 2 * P->alpha;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H[size_t(i) * n + j] = TB(i, j) + (i == j ? two_alpha : 0.f);

    std::vector<float> e(s);
    for (int i = 0; i < s; ++i) {
        float acc = 0.f;
        for (int k = 0; k < 13; ++k) {
            float prod = Aqp(i, k) * x0[k];
            acc = acc + prod;
        }
        e[i] = acc - Xd[i];
    }
    for (int i = 0; i < n; ++i) {
        float acc = 0.f;
        for (int k = 0; k < s; ++k) {
            float prod = T(i, k) * e[k];
            acc = acc + prod;
        }
        g[i] = acc;
    }
    if (Aqp_out) std::memcpy(Aqp_out, Aqp.a.data(), sizeof(float) * Aqp.a.size());
    if (Bqp_out) std::memcpy(Bqp_out, Bqp.a.data(), sizeof(float) * Bqp.a.size());
    (void)m;
    return 0;
}

// The reference's solver call: QProblem(n, m), Options::setToMPC(), printLevel none, cold init with
// lb = ub = NULL, lbA = 0, ubA = U_b  (qr_mpc_interface.cpp:414-438).
extern "C" int qro_qpoases_dense(int n, int m, const double* H, const double* g, const double* A,
                                 const double* lbA, const double* ubA, int nWSR_in, double* x,
                                 int* info, double* kkt, int* cstat) {
    qpOASES::int_t nWSR = nWSR_in;
    qpOASES::QProblem problem(n, m);
    qpOASES::Options option;
    option.setToMPC();
    option.printLevel = qpOASES::PL_NONE;
    problem.setOptions(option);
    int rval = problem.init(H, g, A, NULL, NULL, lbA, ubA, nWSR);
    int rval2 = problem.getPrimalSolution(x);
    if (info) {
        info[0] = rval;
        info[1] = int(nWSR);
    }
    if (kkt) {
        qpOASES::SolutionAnalysis sa;
        qpOASES::real_t st = 0, fe = 0, cm = 0;
        sa.getKktViolation(&problem, &st, &fe, &cm);
        kkt[0] = st; kkt[1] = fe; kkt[2] = cm;
    }
    if (cstat) {
        std::vector<qpOASES::real_t> ws(m);
        problem.getWorkingSetConstraints(ws.data());
        for (int i = 0; i < m; ++i) cstat[i] = ws[i] < -0.5 ? -1 : (ws[i] > 0.5 ? 1 : 0);
    }
    return rval2 == qpOASES::SUCCESSFUL_RETURN ? 0 : 1;
}

extern "C" int qro_mpc_qpoases(int h, float mu, const float* H, const float* g, const float* ub,
                               int nWSR, double* x, int* info, double* kkt, int* cstat) {
    const int n = 12 * h, m = 20 * h;
    // EigenToOASES (:127-136, :418-425): float -> double, row-major; lbA = 0.
    std::vector<double> Hd(size_t(n) * n), gd(n), Ad(size_t(m) * n, 0.0), lb(m, 0.0), ubd(m);
    for (size_t i = 0; i < Hd.size(); ++i) Hd[i] = H[i];
    for (int i = 0; i < n; ++i) gd[i] = g[i];
    for (int i = 0; i < m; ++i) ubd[i] = ub[i];
    // fmat (ResizeQPMats :230-240): 5x3 block [mu_ 0 1; -mu_ 0 1; 0 mu_ 1; 0 -mu_ 1; 0 0 1], mu_ = 1/mu
    const float mu_ = 1.f / mu;
    for (int k = 0; k < 4 * h; ++k) {
        double* blk = &Ad[size_t(5 * k) * n + 3 * k];
        blk[0 * n + 0] = mu_;  blk[0 * n + 2] = 1.f;
        blk[1 * n + 0] = -mu_; blk[1 * n + 2] = 1.f;
        blk[2 * n + 1] = mu_;  blk[2 * n + 2] = 1.f;
        blk[3 * n + 1] = -mu_; blk[3 * n + 2] = 1.f;
        blk[4 * n + 2] = 1.f;
    }
    return qro_qpoases_dense(n, m, Hd.data(), gd.data(), Ad.data(), lb.data(), ubd.data(), nWSR, x,
                             info, kkt, cstat);
}

extern "C" int qro_mpc_solve(const qro_mpc_params* P, const float* p, const float* v,
                             const float* quat, const float* w, const float* r_feet,
                             const float* rpy, const float* traj, const float* gait, int nWSR,
                             double* x, int* info) {
    const int h = P->horizon, n = 12 * h, m = 20 * h;
    std::vector<float> H(size_t(n) * n), g(n), ub(m);
    int rc = qro_mpc_build(P, p, v, quat, w, r_feet, rpy, traj, gait, H.data(), g.data(), ub.data(),
                           nullptr, nullptr);
    if (rc) return rc;
    return qro_mpc_qpoases(h, P->mu, H.data(), g.data(), ub.data(), nWSR, x, info, nullptr, nullptr);
}

// MPCStanceLegController::Run, qr_mpc_stance_leg_controller.cpp:282-303.
extern "C" void qro_mpc_contact_table(int h, int num_horizon_l, const float* progress,
                                      const float* duty, const int* early_contact,
                                      const int* contacts, float* table) {
    float dPhase = 1.0 / (num_horizon_l * h);  // double division narrowed to float (:284)
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < 4; ++j) {
            float ph = progress[j] + i * dPhase;
            while (ph > 1.0) ph -= 1.0;
            table[4 * i + j] = (ph < duty[j] || (early_contact && early_contact[j])) ? 1.f : 0.f;
        }
    if (contacts)
        for (int j = 0; j < 4; ++j) table[j] = float(contacts[j] != 0);
}

// MPCStanceLegController::UpdateMPC, qr_mpc_stance_leg_controller.cpp:345-376.
extern "C" void qro_mpc_reference_traj(int h, float dt_mpc, const float* init, const float* pos_xy,
                                       float* traj) {
    float t0[12];
    for (int j = 0; j < 12; ++j) t0[j] = init[j];
    // clip(x, p-0.1, p+0.1)  (:347-356)
    for (int a = 0; a < 2; ++a) {
        float lo = pos_xy[a] - 0.1f, hi = pos_xy[a] + 0.1f;
        float x = t0[3 + a];
        t0[3 + a] = x < lo ? lo : (x > hi ? hi : x);
    }
    const float yawRate = t0[8], vx = t0[9], vy = t0[10];
    for (int i = 0; i < h; ++i) {
        for (int j = 0; j < 12; ++j) traj[12 * i + j] = t0[j];
        if (i > 0) {
            traj[12 * i + 2] = traj[12 * (i - 1) + 2] + dt_mpc * yawRate;
            traj[12 * i + 3] = traj[12 * (i - 1) + 3] + dt_mpc * vx;
            traj[12 * i + 4] = traj[12 * (i - 1) + 4] + dt_mpc * vy;
        }
    }
}

// SolveDenseMPC post-processing, qr_mpc_stance_leg_controller.cpp:402-409.
extern "C" void qro_mpc_grf_to_leg_force(const float* Rb, const double* x, float* f_world,
                                         float* f_ff) {
    for (int leg = 0; leg < 4; ++leg) {
        float f[3];
        for (int a = 0; a < 3; ++a) f[a] = float(x[3 * leg + a]);
        for (int a = 0; a < 3; ++a) {
            f_world[3 * leg + a] = f[a];
            // (-R^T) f : negate the matrix first, then the product, as Eigen evaluates it
            f_ff[3 * leg + a] = ((-Rb[a]) * f[0] + (-Rb[3 + a]) * f[1]) + (-Rb[6 + a]) * f[2];
        }
    }
}

extern "C" double qro_mpc_time_batch(const qro_mpc_params* P, int count, const float* p,
                                     const float* v, const float* quat, const float* w,
                                     const float* r_feet, const float* rpy, const float* traj,
                                     const float* gait, int nWSR, double* x_all, double* lat,
                                     int* capped) {
    const int h = P->horizon, n = 12 * h;
    std::vector<double> x(n);
    int ncap = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < count; ++i) {
        int info[2];
        auto a = std::chrono::steady_clock::now();
        qro_mpc_solve(P, p + 3 * i, v + 3 * i, quat + 4 * i, w + 3 * i, r_feet + 12 * i, rpy + 3 * i,
                      traj + size_t(12) * h * i, gait + size_t(4) * h * i, nWSR, x.data(), info);
        auto b = std::chrono::steady_clock::now();
        if (lat) lat[i] = std::chrono::duration<double>(b - a).count();
        if (info[0] == qpOASES::RET_MAX_NWSR_REACHED) ++ncap;
        if (x_all) std::memcpy(x_all + size_t(n) * i, x.data(), sizeof(double) * n);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (capped) *capped = ncap;
    return std::chrono::duration<double>(t1 - t0).count();
}

// GRF -> joint torques: SolveDenseMPC (qr_mpc_stance_leg_controller.cpp:402-409, f_ff = -R_base^T f) followed by
// GetAction -> qrRobot::MapContactForceToJointTorques (qr_mpc_stance_leg_controller.cpp:139-141;
// src/robots/qr_robot.cpp:241-251) with the analytic leg Jacobian (qr_robot.cpp:148-172, UpdateDataFlow :62-66).
// quat = (w,x,y,z); R_base = quaternionToRotationMatrix(quat)^T (qr_robot.cpp:70).  Float32 throughout, with
// the float overloads of sqrt / sin / cos and the double pow(-1, leg+1) of the reference.
extern "C" void qro_mpc_grf_to_torque(float hip_len, float upper_len, float lower_len, const float* quat,
                                      const float* q, const float* f_world, float* f_ff_out, float* tau) {
    const float e0 = quat[0], e1 = quat[1], e2 = quat[2], e3 = quat[3];
    // utils/qr_se3.h:186-203 builds R (row-major below) and returns its transpose; baseRMat transposes back
    const float Rb[9] = {1 - 2 * (e2 * e2 + e3 * e3), 2 * (e1 * e2 - e0 * e3), 2 * (e1 * e3 + e0 * e2),
                         2 * (e1 * e2 + e0 * e3), 1 - 2 * (e1 * e1 + e3 * e3), 2 * (e2 * e3 - e0 * e1),
                         2 * (e1 * e3 - e0 * e2), 2 * (e2 * e3 + e0 * e1), 1 - 2 * (e1 * e1 + e2 * e2)};
    for (int leg = 0; leg < 4; ++leg) {
        const float* f = f_world + 3 * leg;
        float ff[3];
        for (int a = 0; a < 3; ++a) ff[a] = ((-Rb[a]) * f[0] + (-Rb[3 + a]) * f[1]) + (-Rb[6 + a]) * f[2];
        const float* t = q + 3 * leg;
        const float sh = hip_len * pow(-1, leg + 1);
        const float lEff = sqrt(upper_len * upper_len + lower_len * lower_len + 2 * upper_len * lower_len * cos(t[2]));
        const float tEff = t[1] + t[2] / 2;
        float J[9];
        J[0] = 0;
        J[1] = -lEff * cos(tEff);
        J[2] = lower_len * upper_len * sin(t[2]) * sin(tEff) / lEff - lEff * cos(tEff) / 2;
        J[3] = -sh * sin(t[0]) + lEff * cos(t[0]) * cos(tEff);
        J[4] = -lEff * sin(t[0]) * sin(tEff);
        J[5] = -lower_len * upper_len * sin(t[0]) * sin(t[2]) * cos(tEff) / lEff - lEff * sin(t[0]) * sin(tEff) / 2;
        J[6] = sh * cos(t[0]) + lEff * sin(t[0]) * cos(tEff);
        J[7] = lEff * sin(tEff) * cos(t[0]);
        J[8] = lower_len * upper_len * sin(t[2]) * cos(t[0]) * cos(tEff) / lEff + lEff * sin(tEff) * cos(t[0]) / 2;
        for (int a = 0; a < 3; ++a) {
            tau[3 * leg + a] = (J[a] * ff[0] + J[3 + a] * ff[1]) + J[6 + a] * ff[2];   // J^T f_ff
            if (f_ff_out) f_ff_out[3 * leg + a] = ff[a];
        }
    }
}
