// fb_oracle.cpp -- CPU ORACLE for the force-balance stance QP.  TEST INFRASTRUCTURE ONLY.
//
// Eigen-free restatement of Quadruped::ComputeContactForce and its helpers
// (/root/reference/quadruped/src/controllers/balance_controller/qr_qp_torque_optimizer.cpp):
//   ComputeMassMatrix        :31-64 and :403-427
//   ComputeConstraintMatrix  :67-112 (control frame) and :115-151 (world frame)
//   ComputeObjectiveMatrix   :154-182
//   ComputeWeightMatrix      :185-189
//   the QuadProg++ call      :236-276 / :343-383 -- solved by the reference's OWN vendored QuadProg++
//                            (oracle/_ref/libquadprog.a), with the same transposed copies and sign flips
// Arithmetic is float like the reference; matrix products are coefficient-wise with k ascending (Eigen's lazy
// product).  Parity caveat (same as the MPC oracle): a true Eigen build may route the 12x6 * 6x12 product through
// its GEMM kernel, whose summation order is not reproducible without Eigen -- "build unpinned by an Eigen build".
// The frame changes around the QP (Rcb / RigidTransform) are outside: inputs are the matrices the reference hands
// to ComputeMassMatrix, output is X = -x (4x3, leg-major).
#include <cmath>
#include <cstring>

#include "QuadProg++.hh"
#include "qr_oracle.h"

namespace {
struct Mat {
    int r, c;
    float v[24 * 12];
    Mat(int r_, int c_) : r(r_), c(c_) { for (int i = 0; i < r * c; ++i) v[i] = 0.f; }
    float& operator()(int i, int j) { return v[i * c + j]; }
    float operator()(int i, int j) const { return v[i * c + j]; }
};
Mat mul(const Mat& a, const Mat& b) {   // coefficient-wise product, k ascending
    Mat o(a.r, b.c);
    for (int i = 0; i < a.r; ++i)
        for (int j = 0; j < b.c; ++j) {
            float s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.c; ++k) s = s + a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}
Mat transpose(const Mat& a) {
    Mat o(a.c, a.r);
    for (int i = 0; i < a.r; ++i)
        for (int j = 0; j < a.c; ++j) o(j, i) = a(i, j);
    return o;
}
Mat inverse3(const Mat& m) {   // Eigen's 3x3 inverse: cofactors, determinant along the first column of cofactors
    auto cof = [&](int a, int b, int c, int d) { return m.v[a] * m.v[b] - m.v[c] * m.v[d]; };
    const float c00 = cof(4, 8, 5, 7), c01 = cof(5, 6, 3, 8), c02 = cof(3, 7, 4, 6);
    const float det = (m.v[0] * c00 + m.v[1] * c01) + m.v[2] * c02;
    const float id = 1.f / det;
    Mat o(3, 3);
    o.v[0] = c00 * id; o.v[1] = cof(2, 7, 1, 8) * id; o.v[2] = cof(1, 5, 2, 4) * id;
    o.v[3] = c01 * id; o.v[4] = cof(0, 8, 2, 6) * id; o.v[5] = cof(2, 3, 0, 5) * id;
    o.v[6] = c02 * id; o.v[7] = cof(1, 6, 0, 7) * id; o.v[8] = cof(0, 4, 1, 3) * id;
    return o;
}
}  // namespace

extern "C" int qro_force_balance(const qro_fb_params* P, const float* inertia, const float* foot, const float* acc,
                                 const int* contact, const float* gravity, const float* frame, float* force,
                                 float* G_out, float* a_out, float* C_out, float* lb_out, double* cost_out) {
    // ---- mass matrix
    Mat I3(3, 3);
    for (int i = 0; i < 9; ++i) I3.v[i] = inertia ? inertia[i] : P->inertia[i];
    const Mat invI = inverse3(I3);
    Mat M(6, 12);
    for (int leg = 0; leg < 4; ++leg) {
        for (int i = 0; i < 3; ++i) M(i, 3 * leg + i) = 1.f / P->mass;
        Mat S(3, 3);
        const float* x = foot + 3 * leg;
        S(0, 1) = -x[2]; S(0, 2) = x[1]; S(1, 0) = x[2]; S(1, 2) = -x[0]; S(2, 0) = -x[1]; S(2, 1) = x[0];
        const Mat B = mul(invI, S);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M(3 + i, 3 * leg + j) = B(i, j);
    }
    // ---- objective
    Mat Q(6, 6);
    for (int i = 0; i < 6; ++i) Q(i, i) = P->acc_weight[i];
    Mat quad = mul(mul(transpose(M), Q), M);
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) quad(i, j) = quad(i, j) + 1.f * P->reg_weight;   // Ones * regWeight
    Mat gv(1, 6);
    for (int k = 0; k < 6; ++k) {
        const float g = k < 3 ? (gravity ? gravity[k] : (k == 2 ? 9.8f : 0.f)) : 0.f;
        gv(0, k) = g + acc[k];
    }
    const Mat lin = mul(mul(gv, Q), M);   // 1 x 12
    for (int i = 0; i < 12; ++i) quad(i, i) = quad(i, i) + 1e-4f;   // + W
    // ---- constraints
    float n[3] = {0.f, 0.f, 1.f}, t1[3] = {1.f, 0.f, 0.f}, t2[3] = {0.f, 1.f, 0.f};
    if (frame) for (int k = 0; k < 3; ++k) { n[k] = frame[k]; t1[k] = frame[3 + k]; t2[k] = frame[6 + k]; }
    Mat A(24, 12);
    float lb[24];
    for (int i = 0; i < 24; ++i) lb[i] = 0.f;
    for (int leg = 0; leg < 4; ++leg) {
        for (int k = 0; k < 3; ++k) { A(2 * leg, 3 * leg + k) = n[k]; A(2 * leg + 1, 3 * leg + k) = -n[k]; }
        if (contact[leg] > 0) {
            if (P->world_frame) {
                lb[2 * leg] = P->fmin_ratio[leg] * P->mass * 9.8;        // float * float, * double, narrowed
                lb[2 * leg + 1] = -P->fmax_ratio[leg] * P->mass * 9.8;
            } else {
                const float fMin = P->fmin_ratio[leg] * P->mass * 9.8f, fMax = P->fmax_ratio[leg] * P->mass * 9.8f;
                lb[2 * leg] = fMin;
                lb[2 * leg + 1] = -fMax;
            }
        } else {
            lb[2 * leg] = 1e-7;
            lb[2 * leg + 1] = 1e-7;
        }
        const int row = 8 + 4 * leg;
        for (int k = 0; k < 3; ++k) {
            A(row, 3 * leg + k) = P->mu * n[k] + t1[k];
            A(row + 1, 3 * leg + k) = P->mu * n[k] - t1[k];
            A(row + 2, 3 * leg + k) = P->mu * n[k] + t2[k];
            A(row + 3, 3 * leg + k) = P->mu * n[k] - t2[k];
        }
    }
    if (G_out) memcpy(G_out, quad.v, 144 * sizeof(float));
    if (a_out) for (int i = 0; i < 12; ++i) a_out[i] = lin(0, i);
    if (C_out) memcpy(C_out, A.v, 288 * sizeof(float));
    if (lb_out) memcpy(lb_out, lb, sizeof(lb));
    // ---- QuadProg++ exactly as the reference calls it
    quadprogpp::Matrix<double> GG(12, 12), CICI(12, 24), CECE(12, 0);
    quadprogpp::Vector<double> aa(12), bb(24), ee(0), x(12);
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) GG[i][j] = double(quad(j, i));
    for (int i = 0; i < 12; ++i) aa[i] = double(-lin(0, i));
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 24; ++j) CICI[i][j] = double(A(j, i));   // Ci = A.transpose()
    for (int i = 0; i < 24; ++i) bb[i] = double(-lb[i]);
    const double cost = quadprogpp::solve_quadprog(GG, aa, CECE, ee, CICI, bb, x);
    if (cost_out) *cost_out = cost;
    int invalid = 0;
    for (int i = 0; i < 12; ++i) if (std::isnan(x[i])) ++invalid;
    for (int i = 0; i < 12; ++i) force[i] = invalid ? 0.f : -float(x[i]);
    return invalid ? 3 : (std::isinf(cost) ? 1 : 0);
}

// Generic inequality-constrained QP through the reference's QuadProg++: min 1/2 x'Gx + g0'x  s.t. C[i].x + c0[i] >= 0.
// G n x n row-major (symmetric), C m rows of n.  Returns QuadProg++'s cost (inf when it reports infeasibility).
extern "C" double qro_quadprog_ineq(int n, int m, const double* G, const double* g0, const double* C, const double* c0, double* x) {
    quadprogpp::Matrix<double> GG(n, n), CI(n, m), CE(n, 0);
    quadprogpp::Vector<double> gg(n), ci(m), ce(0), xx(n);
    for (int i = 0; i < n; ++i) { gg[i] = g0[i]; for (int j = 0; j < n; ++j) GG[i][j] = G[i * n + j]; }
    for (int i = 0; i < m; ++i) { ci[i] = c0[i]; for (int k = 0; k < n; ++k) CI[k][i] = C[i * n + k]; }
    const double cost = quadprogpp::solve_quadprog(GG, gg, CE, ce, CI, ci, xx);
    for (int i = 0; i < n; ++i) x[i] = xx[i];
    return cost;
}
