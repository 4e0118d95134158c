// wbc_oracle.cpp -- CPU ORACLE for the whole-body-control step.  TEST INFRASTRUCTURE ONLY.
//
// Eigen-free restatement of what qrWbcLocomotionController<float>::Run computes on the ticks where it
// recomputes (paths relative to /root/reference/quadruped/):
//   robot model constants   src/robots/qr_robot_a1_sim.cpp:176-345 (qr_robot_lite3_sim.cpp identical)
//   floating-base dynamics  src/dynamics/floating_base_model.cpp:469-524 (forwardKinematics), :587-600
//                           (biasAccelerations), :541-580 (contactJacobians), :750-767 (compositeInertias),
//                           :774-806 (massMatrix), :607-626 (gravity), :633-665 (Coriolis);
//                           include/quadruped/dynamics/spatial.hpp
//   tasks / contacts        src/controllers/wbc/task_set/*.cpp, src/controllers/wbc/qr_single_contact.cpp
//   kinematic WBC           src/controllers/wbc/qr_multitask_projection.cpp:38-106
//   WBIC                    src/controllers/wbc/qr_wholebody_impulse_ctrl.cpp:50-299
//   glue                    src/controllers/wbc/qr_wbc_locomotion_controller.cpp:108-219
// The QP is solved by the reference's own vendored QuadProg++ (oracle/_ref/libquadprog.a, double).
//
// Templated on the scalar: T = float follows the reference's arithmetic type (Eigen's summation
// order, LU pivoting and Jacobi sweeps are NOT reproducible without Eigen -- same caveat as the MPC
// oracle); T = double evaluates the same algorithm in double and is what the GPU engine (float64) is
// compared with.  The difference between the two is the reference's own rounding noise floor.
#include "qr_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "QuadProg++.hh"

namespace {

template <typename T>
struct Mat {
    int r = 0, c = 0;
    std::vector<T> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, T(0)) {}
    T& operator()(int i, int j) { return a[size_t(i) * c + j]; }
    T operator()(int i, int j) const { return a[size_t(i) * c + j]; }
    static Mat eye(int n) { Mat m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1; return m; }
};
template <typename T> Mat<T> operator*(const Mat<T>& A, const Mat<T>& B) {
    Mat<T> C(A.r, B.c);
    for (int i = 0; i < A.r; ++i)
        for (int j = 0; j < B.c; ++j) {
            T s = 0;
            for (int k = 0; k < A.c; ++k) s += A(i, k) * B(k, j);
            C(i, j) = s;
        }
    return C;
}
template <typename T> Mat<T> operator+(const Mat<T>& A, const Mat<T>& B) { Mat<T> C = A; for (size_t i = 0; i < C.a.size(); ++i) C.a[i] += B.a[i]; return C; }
template <typename T> Mat<T> operator-(const Mat<T>& A, const Mat<T>& B) { Mat<T> C = A; for (size_t i = 0; i < C.a.size(); ++i) C.a[i] -= B.a[i]; return C; }
template <typename T> Mat<T> operator*(T s, const Mat<T>& A) { Mat<T> C = A; for (auto& v : C.a) v *= s; return C; }
template <typename T> Mat<T> tr(const Mat<T>& A) { Mat<T> C(A.c, A.r); for (int i = 0; i < A.r; ++i) for (int j = 0; j < A.c; ++j) C(j, i) = A(i, j); return C; }
template <typename T> Mat<T> block(const Mat<T>& A, int i0, int j0, int nr, int nc) { Mat<T> C(nr, nc); for (int i = 0; i < nr; ++i) for (int j = 0; j < nc; ++j) C(i, j) = A(i0 + i, j0 + j); return C; }
template <typename T> void setblock(Mat<T>& A, int i0, int j0, const Mat<T>& B) { for (int i = 0; i < B.r; ++i) for (int j = 0; j < B.c; ++j) A(i0 + i, j0 + j) = B(i, j); }
template <typename T> Mat<T> colvec(std::initializer_list<T> v) { Mat<T> m((int)v.size(), 1); int i = 0; for (T x : v) m(i++, 0) = x; return m; }
template <typename T> T dot(const Mat<T>& a, const Mat<T>& b) { T s = 0; for (size_t i = 0; i < a.a.size(); ++i) s += a.a[i] * b.a[i]; return s; }

// ---- utils/qr_se3.h ------------------------------------------------------------------------------
template <typename T> Mat<T> skew(T x, T y, T z) {  // vectorToSkewMat / crossMatrix (:95-103, :122-131)
    Mat<T> m(3, 3);
    m(0, 1) = -z; m(0, 2) = y; m(1, 0) = z; m(1, 2) = -x; m(2, 0) = -y; m(2, 1) = x;
    return m;
}
template <typename T> Mat<T> skew(const Mat<T>& v) { return skew(v(0, 0), v(1, 0), v(2, 0)); }
template <typename T> Mat<T> unskew(const Mat<T>& m) {  // matToSkewVec (:135-140)
    return colvec<T>({T(0.5) * (m(2, 1) - m(1, 2)), T(0.5) * (m(0, 2) - m(2, 0)), T(0.5) * (m(1, 0) - m(0, 1))});
}
template <typename T> Mat<T> coord_rot(int axis, T th) {  // coordinateRotation (:72-89), passive
    T s = std::sin(th), c = std::cos(th);
    Mat<T> R = Mat<T>::eye(3);
    if (axis == 0) { R(1, 1) = c; R(1, 2) = s; R(2, 1) = -s; R(2, 2) = c; }
    else if (axis == 1) { R(0, 0) = c; R(0, 2) = -s; R(2, 0) = s; R(2, 2) = c; }
    else { R(0, 0) = c; R(0, 1) = s; R(1, 0) = -s; R(1, 1) = c; }
    return R;
}
template <typename T> Mat<T> quat_to_rot(const T* q) {  // quaternionToRotationMatrix (:186-203): world -> body
    T e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3];
    Mat<T> R(3, 3);
    R(0, 0) = 1 - 2 * (e2 * e2 + e3 * e3); R(0, 1) = 2 * (e1 * e2 - e0 * e3); R(0, 2) = 2 * (e1 * e3 + e0 * e2);
    R(1, 0) = 2 * (e1 * e2 + e0 * e3); R(1, 1) = 1 - 2 * (e1 * e1 + e3 * e3); R(1, 2) = 2 * (e2 * e3 - e0 * e1);
    R(2, 0) = 2 * (e1 * e3 - e0 * e2); R(2, 1) = 2 * (e2 * e3 + e0 * e1); R(2, 2) = 1 - 2 * (e1 * e1 + e2 * e2);
    return tr(R);
}
template <typename T> void rot_to_quat(const Mat<T>& r1, T* q) {  // rotationMatrixToQuaternion (:146-180)
    Mat<T> r = tr(r1);
    T t = r(0, 0) + r(1, 1) + r(2, 2);
    if (t > 0) {
        T S = std::sqrt(t + T(1)) * T(2);
        q[0] = T(0.25) * S; q[1] = (r(2, 1) - r(1, 2)) / S; q[2] = (r(0, 2) - r(2, 0)) / S; q[3] = (r(1, 0) - r(0, 1)) / S;
    } else if (r(0, 0) > r(1, 1) && r(0, 0) > r(2, 2)) {
        T S = std::sqrt(T(1) + r(0, 0) - r(1, 1) - r(2, 2)) * T(2);
        q[0] = (r(2, 1) - r(1, 2)) / S; q[1] = T(0.25) * S; q[2] = (r(0, 1) + r(1, 0)) / S; q[3] = (r(0, 2) + r(2, 0)) / S;
    } else if (r(1, 1) > r(2, 2)) {
        T S = std::sqrt(T(1) + r(1, 1) - r(0, 0) - r(2, 2)) * T(2);
        q[0] = (r(0, 2) - r(2, 0)) / S; q[1] = (r(0, 1) + r(1, 0)) / S; q[2] = T(0.25) * S; q[3] = (r(1, 2) + r(2, 1)) / S;
    } else {
        T S = std::sqrt(T(1) + r(2, 2) - r(0, 0) - r(1, 1)) * T(2);
        q[0] = (r(1, 0) - r(0, 1)) / S; q[1] = (r(0, 2) + r(2, 0)) / S; q[2] = (r(1, 2) + r(2, 1)) / S; q[3] = T(0.25) * S;
    }
}
template <typename T> void rpy_to_quat(const T* rpy, T* q) {  // rpyToQuat (:229-235) via rpyToRotMat (:108-116)
    Mat<T> R = coord_rot<T>(0, rpy[0]) * coord_rot<T>(1, rpy[1]) * coord_rot<T>(2, rpy[2]);
    rot_to_quat(R, q);
}
template <typename T> void quat_product(const T* a, const T* b, T* o) {  // quatProduct (:291-302)
    o[0] = a[0] * b[0] - (a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
    o[1] = a[0] * b[1] + b[0] * a[1] + (a[2] * b[3] - a[3] * b[2]);
    o[2] = a[0] * b[2] + b[0] * a[2] + (a[3] * b[1] - a[1] * b[3]);
    o[3] = a[0] * b[3] + b[0] * a[3] + (a[1] * b[2] - a[2] * b[1]);
}
template <typename T> void quat_to_so3(const T* q, T* so3) {  // quaternionToso3 (:383-397)
    so3[0] = q[1]; so3[1] = q[2]; so3[2] = q[3];
    T theta = T(2.0 * std::asin(std::sqrt(double(so3[0] * so3[0] + so3[1] * so3[1] + so3[2] * so3[2]))));
    if (std::fabs(theta) < T(0.0000001)) { so3[0] = so3[1] = so3[2] = 0; return; }
    T s = T(std::sin(double(theta) / 2.0));
    for (int i = 0; i < 3; ++i) { so3[i] /= s; so3[i] *= theta; }
}

// ---- dynamics/spatial.hpp ------------------------------------------------------------------------
template <typename T> Mat<T> sxform(const Mat<T>& R, const Mat<T>& r) {  // createSXform: [R 0; -R[r]x R]
    Mat<T> X(6, 6);
    setblock(X, 0, 0, R); setblock(X, 3, 3, R);
    setblock(X, 3, 0, T(-1) * (R * skew(r)));
    return X;
}
template <typename T> Mat<T> spatial_rot(int axis, T th) { Mat<T> R = coord_rot<T>(axis, th); Mat<T> X(6, 6); setblock(X, 0, 0, R); setblock(X, 3, 3, R); return X; }
template <typename T> Mat<T> sx_translation(const Mat<T>& X) {  // translationFromSXform
    Mat<T> R = block(X, 0, 0, 3, 3);
    return T(-1) * unskew(tr(R) * block(X, 3, 0, 3, 3));
}
template <typename T> Mat<T> invert_sxform(const Mat<T>& X) {
    Mat<T> R = block(X, 0, 0, 3, 3);
    Mat<T> r = sx_translation(X);
    return sxform(tr(R), T(-1) * (R * r));
}
template <typename T> Mat<T> motion_cross(const Mat<T>& a, const Mat<T>& b) {  // motionCrossProduct
    auto A = [&](int i) { return a(i, 0); };
    auto B = [&](int i) { return b(i, 0); };
    return colvec<T>({A(1) * B(2) - A(2) * B(1), A(2) * B(0) - A(0) * B(2), A(0) * B(1) - A(1) * B(0),
                      A(1) * B(5) - A(2) * B(4) + A(4) * B(2) - A(5) * B(1),
                      A(2) * B(3) - A(0) * B(5) - A(3) * B(2) + A(5) * B(0),
                      A(0) * B(4) - A(1) * B(3) + A(3) * B(1) - A(4) * B(0)});
}
template <typename T> Mat<T> force_cross(const Mat<T>& a, const Mat<T>& b) {  // forceCrossProduct
    auto A = [&](int i) { return a(i, 0); };
    auto B = [&](int i) { return b(i, 0); };
    return colvec<T>({B(2) * A(1) - B(1) * A(2) - B(4) * A(5) + B(5) * A(4),
                      B(0) * A(2) - B(2) * A(0) + B(3) * A(5) - B(5) * A(3),
                      B(1) * A(0) - B(0) * A(1) - B(3) * A(4) + B(4) * A(3),
                      B(5) * A(1) - B(4) * A(2), B(3) * A(2) - B(5) * A(0), B(4) * A(0) - B(3) * A(1)});
}
template <typename T> Mat<T> spatial_inertia(T mass, const Mat<T>& com, const Mat<T>& I) {
    Mat<T> c = skew(com);
    Mat<T> M(6, 6);
    setblock(M, 0, 0, I + mass * (c * tr(c)));
    setblock(M, 0, 3, mass * c);
    setblock(M, 3, 0, mass * tr(c));
    setblock(M, 3, 3, mass * Mat<T>::eye(3));
    return M;
}
template <typename T> Mat<T> flip_inertia_y(const Mat<T>& M) {  // flipAlongAxis(Y) through the pseudo-inertia
    Mat<T> h = unskew(block(M, 0, 3, 3, 3));
    Mat<T> Ibar = block(M, 0, 0, 3, 3);
    T m = M(5, 5);
    T trace = Ibar(0, 0) + Ibar(1, 1) + Ibar(2, 2);
    Mat<T> P(4, 4);
    setblock(P, 0, 0, T(0.5) * trace * Mat<T>::eye(3) - Ibar);
    for (int i = 0; i < 3; ++i) { P(i, 3) = h(i, 0); P(3, i) = h(i, 0); }
    P(3, 3) = m;
    Mat<T> X = Mat<T>::eye(4);
    X(1, 1) = -1;
    P = X * P * X;
    Mat<T> E = block(P, 0, 0, 3, 3);
    Mat<T> hh = colvec<T>({P(0, 3), P(1, 3), P(2, 3)});
    Mat<T> I(6, 6);
    setblock(I, 0, 0, (E(0, 0) + E(1, 1) + E(2, 2)) * Mat<T>::eye(3) - E);
    setblock(I, 0, 3, skew(hh));
    setblock(I, 3, 0, tr(skew(hh)));
    setblock(I, 3, 3, P(3, 3) * Mat<T>::eye(3));
    return I;
}

// ---- utils/qr_algebra.h:119-140: Jacobi-SVD pseudo-inverse with a singular-value threshold -----------
// One-sided (Hestenes) Jacobi on the columns of B = A' (or A when it is tall).
template <typename T> Mat<T> pinv(const Mat<T>& A, double thr) {
    if (A.r == 1 && A.c == 1) { Mat<T> o(1, 1); o(0, 0) = A(0, 0) > thr ? T(1) / A(0, 0) : T(0); return o; }
    const bool wide = A.r < A.c;
    Mat<T> B = wide ? tr(A) : A;   // tall: n x k
    const int n = B.r, k = B.c;
    Mat<T> V = Mat<T>::eye(k);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < k - 1; ++p)
            for (int q = p + 1; q < k; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < n; ++i) { alpha += double(B(i, p)) * B(i, p); beta += double(B(i, q)) * B(i, q); gamma += double(B(i, p)) * B(i, q); }
                if (gamma == 0) continue;
                off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta + 1e-300));
                double zeta = (beta - alpha) / (2 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                double c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int i = 0; i < n; ++i) { T bp = B(i, p), bq = B(i, q); B(i, p) = T(c * bp - s * bq); B(i, q) = T(s * bp + c * bq); }
                for (int i = 0; i < k; ++i) { T vp = V(i, p), vq = V(i, q); V(i, p) = T(c * vp - s * vq); V(i, q) = T(s * vp + c * vq); }
            }
        if (off < (sizeof(T) == 4 ? 1e-7 : 1e-15)) break;
    }
    // B = U S, original tall matrix = U S V'  ->  pinv = V S^-1 U' = V S^-2 B'
    Mat<T> P(k, n);
    for (int j = 0; j < k; ++j) {
        double s2 = 0;
        for (int i = 0; i < n; ++i) s2 += double(B(i, j)) * B(i, j);
        double sigma = std::sqrt(s2);
        if (sigma > thr)
            for (int i = 0; i < n; ++i)
                for (int l = 0; l < k; ++l) P(l, i) += T(V(l, j) * B(i, j) / s2);
    }
    return wide ? tr(P) : P;
}

// Eigen's A.inverse() on a dynamic matrix is a partial-pivoting LU solve of the identity.
template <typename T> Mat<T> inverse_lu(Mat<T> A) {
    const int n = A.r;
    Mat<T> I = Mat<T>::eye(n);
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(A(i, k)) > std::fabs(A(piv, k))) piv = i;
        if (piv != k) for (int j = 0; j < n; ++j) { std::swap(A(k, j), A(piv, j)); std::swap(I(k, j), I(piv, j)); }
        for (int i = k + 1; i < n; ++i) {
            T f = A(i, k) / A(k, k);
            for (int j = k; j < n; ++j) A(i, j) -= f * A(k, j);
            for (int j = 0; j < n; ++j) I(i, j) -= f * I(k, j);
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        for (int j = 0; j < n; ++j) I(k, j) /= A(k, k);
        for (int i = 0; i < k; ++i) for (int j = 0; j < n; ++j) I(i, j) -= A(i, k) * I(k, j);
    }
    return I;
}

// ---- the rigid-body tree (13 moving bodies: base = 5, legs 6..17) ---------------------------------
template <typename T>
struct Model {
    int parent[18];
    int axis[18];
    Mat<T> Xtree[18], Xrot[18], Ibody[18], Irot[18];
    int gc_parent[16];
    Mat<T> gc_loc[16];
    T gravity[3] = {0, 0, T(float(-9.81))};   // Vec3<float> g(0, 0, -9.81)
    T total_mass = 0;
};

template <typename T> Mat<T> leg_signs(T x, T y, T z, int leg) {  // qrRobot::WithLegSigns, src/robots/qr_robot.cpp:89-104
    switch (leg) {
        case 0: return colvec<T>({x, -y, z});
        case 1: return colvec<T>({x, y, z});
        case 2: return colvec<T>({-x, -y, z});
        default: return colvec<T>({-x, y, z});
    }
}

// qrRobotA1Sim::BuildDynamicModel, src/robots/qr_robot_a1_sim.cpp:176-345.  The constants are float
// literals in the reference; they are rounded to float first and then widened so that T = double sees
// the same model.
template <typename T> Model<T> build_model(const qro_wbc_model* cfg) {
    auto F = [](double v) { return T(float(v)); };
    Model<T> M;
    Mat<T> I3 = Mat<T>::eye(3);
    Mat<T> rotorI = T(float(float(1e-2) * 1e-6)) * I3;  // setIdentity() then scale_*1e-6
    {   // RY * I * RY' and RX * I * RX' with the float rotations of M_PI/2
        Mat<T> RY = coord_rot<T>(1, T(float(M_PI / 2))), RX = coord_rot<T>(0, T(float(M_PI / 2)));
        (void)RY; (void)RX;
    }
    Mat<T> RY = coord_rot<T>(1, T(float(M_PI / 2))), RX = coord_rot<T>(0, T(float(M_PI / 2)));
    Mat<T> rotorIX = RY * rotorI * tr(RY), rotorIY = RX * rotorI * tr(RX);
    auto M3 = [&](std::initializer_list<double> v) { Mat<T> m(3, 3); int i = 0; for (double x : v) { m.a[i++] = T(float(x)) ; } return T(float(1e-6)) * m; };
    Mat<T> abadI = M3({469.2, -9.4, -0.342, -9.4, 807.5, -0.466, -0.342, -0.466, 552.9});
    Mat<T> hipI = M3({5529, 4.825, 343.9, 4.825, 5139.3, 22.4, 343.9, 22.4, 1367.8});
    Mat<T> kneeI = M3({2998, 0, -141.2, 0, 3014, 0, -141.2, 0, 32.4});
    Mat<T> bodyI = M3({15853, 0, 0, 0, 37799, 0, 0, 0, 45654});
    Mat<T> abadS = spatial_inertia<T>(F(0.696), colvec<T>({F(-0.0033), 0, 0}), abadI);
    Mat<T> hipS = spatial_inertia<T>(F(1.013), colvec<T>({F(-0.003237), F(-0.022327), F(-0.027326)}), hipI);
    Mat<T> kneeS = spatial_inertia<T>(F(0.166), colvec<T>({F(0.006435), 0, F(-0.107)}), kneeI);
    Mat<T> zero3 = colvec<T>({0, 0, 0});
    Mat<T> rotX = spatial_inertia<T>(F(1e-8), zero3, rotorIX), rotY = spatial_inertia<T>(F(1e-8), zero3, rotorIY);
    Mat<T> bodyS = spatial_inertia<T>(T(6), zero3, bodyI);
    for (int i = 0; i < 6; ++i) { M.parent[i] = 0; M.axis[i] = 0; M.Xtree[i] = Mat<T>::eye(6); M.Xrot[i] = Mat<T>::eye(6); M.Ibody[i] = Mat<T>(6, 6); M.Irot[i] = Mat<T>(6, 6); }
    M.Ibody[5] = bodyS;
    // eight corners of the body box (floating_base_model.cpp:350-366), contact ids 0..7
    const T bx = T(cfg->body_size[0]), by = T(cfg->body_size[1]), bz = T(cfg->body_size[2]);
    const T sx[8] = {1, -1, 1, -1, 1, -1, 1, -1}, sy[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sz[8] = {1, 1, 1, 1, -1, -1, -1, -1};
    for (int k = 0; k < 8; ++k) { M.gc_parent[k] = 5; M.gc_loc[k] = colvec<T>({sx[k] * bx / 2, sy[k] * by / 2, sz[k] * bz / 2}); }
    const T hipL = T(cfg->hip_len), upL = T(cfg->upper_len), lowL = T(cfg->lower_len);
    T side = -1;
    int body = 5;
    for (int leg = 0; leg < 4; ++leg) {
        const int abad = ++body;
        M.parent[abad] = 5; M.axis[abad] = 0;
        M.Xtree[abad] = sxform(I3, leg_signs<T>(F(0.1805), F(0.047), 0, leg));
        M.Xrot[abad] = sxform(I3, leg_signs<T>(F(0.14), F(0.047), 0, leg));
        M.Ibody[abad] = side < 0 ? flip_inertia_y(abadS) : abadS;
        M.Irot[abad] = side < 0 ? flip_inertia_y(rotX) : rotX;
        const int hip = ++body;
        M.parent[hip] = abad; M.axis[hip] = 1;
        M.Xtree[hip] = sxform(I3, leg_signs<T>(0, hipL, 0, leg));
        M.Xrot[hip] = sxform(coord_rot<T>(2, T(float(M_PI))), leg_signs<T>(0, F(0.04), 0, leg));
        M.Ibody[hip] = side < 0 ? flip_inertia_y(hipS) : hipS;
        M.Irot[hip] = side < 0 ? flip_inertia_y(rotY) : rotY;
        M.gc_parent[8 + 2 * leg] = hip; M.gc_loc[8 + 2 * leg] = colvec<T>({0, 0, -upL});
        const int knee = ++body;
        M.parent[knee] = hip; M.axis[knee] = 1;
        M.Xtree[knee] = sxform(I3, colvec<T>({0, 0, -upL}));
        M.Xrot[knee] = sxform(I3, zero3);
        M.Ibody[knee] = kneeS;   // not flipped (:320)
        M.Irot[knee] = side < 0 ? flip_inertia_y(rotY) : rotY;
        M.gc_parent[9 + 2 * leg] = knee;
        M.gc_loc[9 + 2 * leg] = colvec<T>({0, side < 0 ? F(0.004) : -F(0.004), -lowL});
        side = -side;
    }
    M.total_mass = 0;
    for (int i = 0; i < 18; ++i) M.total_mass += M.Ibody[i](5, 5);   // totalNonRotorMass
    return M;
}

template <typename T>
struct Dyn {   // what UpdateModel leaves in the FloatingBaseModel (qr_wbc_locomotion_controller.cpp:138-168)
    Mat<T> H, G, C;           // 18x18, 18, 18
    Mat<T> Jc[16], Jcdqd[16]; // 3x18, 3
    Mat<T> pGC[16], vGC[16];
};

template <typename T>
Dyn<T> dynamics(const Model<T>& M, const T* quat, const T* pos, const T* bodyvel, const T* q, const T* qd) {
    Mat<T> Xup[18], Xuprot[18], Xa[18], S[18], Srot[18], v[18], vrot[18], c[18], crot[18], avp[18], avprot[18];
    // forwardKinematics
    Mat<T> R = quat_to_rot(quat);
    Xup[5] = sxform(R, colvec<T>({pos[0], pos[1], pos[2]}));
    v[5] = Mat<T>(6, 1);
    for (int i = 0; i < 6; ++i) v[5](i, 0) = bodyvel[i];
    for (int i = 6; i < 18; ++i) {
        Mat<T> XJ = spatial_rot<T>(M.axis[i], q[i - 6]);
        Xup[i] = XJ * M.Xtree[i];
        S[i] = Mat<T>(6, 1); S[i](M.axis[i], 0) = 1;
        Mat<T> vJ = qd[i - 6] * S[i];
        v[i] = Xup[i] * v[M.parent[i]] + vJ;
        Srot[i] = S[i];                    // gear ratio 1
        Xuprot[i] = XJ * M.Xrot[i];
        vrot[i] = Xuprot[i] * v[M.parent[i]] + vJ;
        c[i] = motion_cross(v[i], vJ);
        crot[i] = motion_cross(vrot[i], vJ);
    }
    Xa[5] = Xup[5];
    for (int i = 6; i < 18; ++i) Xa[i] = Xup[i] * Xa[M.parent[i]];
    Dyn<T> D;
    for (int k = 0; k < 16; ++k) {
        const int i = M.gc_parent[k];
        Mat<T> Xai = invert_sxform(Xa[i]);
        Mat<T> vs = Xai * v[i];
        Mat<T> Rr = block(Xai, 0, 0, 3, 3);
        D.pGC[k] = Rr * (M.gc_loc[k] - sx_translation(Xai));   // sXFormPoint
        Mat<T> w = block(vs, 0, 0, 3, 1), vl = block(vs, 3, 0, 3, 1);
        D.vGC[k] = vl + skew(w) * D.pGC[k];                    // spatialToLinearVelocity
    }
    // biasAccelerations
    avp[5] = Mat<T>(6, 1);
    for (int i = 6; i < 18; ++i) { avp[i] = Xup[i] * avp[M.parent[i]] + c[i]; avprot[i] = Xuprot[i] * avp[M.parent[i]] + crot[i]; }
    // contactJacobians
    for (int k = 0; k < 16; ++k) {
        int i = M.gc_parent[k];
        Mat<T> Rai = tr(block(Xa[i], 0, 0, 3, 3));
        Mat<T> Xc = sxform(Rai, M.gc_loc[k]);
        Mat<T> ac = Xc * avp[i], vc = Xc * v[i];
        D.Jcdqd[k] = block(ac, 3, 0, 3, 1) + skew(block(vc, 0, 0, 3, 1)) * block(vc, 3, 0, 3, 1);
        D.Jc[k] = Mat<T>(3, 18);
        Mat<T> Xout = block(Xc, 3, 0, 3, 6);
        while (i > 5) {
            Mat<T> col = Xout * S[i];
            for (int r = 0; r < 3; ++r) D.Jc[k](r, i) = col(r, 0);
            Xout = Xout * Xup[i];
            i = M.parent[i];
        }
        setblock(D.Jc[k], 0, 0, Xout);
    }
    // compositeInertias
    Mat<T> IC[18];
    for (int i = 5; i < 18; ++i) IC[i] = M.Ibody[i];
    for (int i = 17; i > 5; --i) {
        IC[M.parent[i]] = IC[M.parent[i]] + tr(Xup[i]) * IC[i] * Xup[i];
        IC[M.parent[i]] = IC[M.parent[i]] + tr(Xuprot[i]) * M.Irot[i] * Xuprot[i];
    }
    // massMatrix
    D.H = Mat<T>(18, 18);
    setblock(D.H, 0, 0, IC[5]);
    for (int j = 6; j < 18; ++j) {
        Mat<T> f = IC[j] * S[j], frot = M.Irot[j] * Srot[j];
        D.H(j, j) = dot(S[j], f) + dot(Srot[j], frot);
        f = tr(Xup[j]) * f + tr(Xuprot[j]) * frot;
        int i = M.parent[j];
        while (i > 5) {
            D.H(i, j) = dot(S[i], f); D.H(j, i) = D.H(i, j);
            f = tr(Xup[i]) * f;
            i = M.parent[i];
        }
        for (int r = 0; r < 6; ++r) { D.H(r, j) = f(r, 0); D.H(j, r) = f(r, 0); }
    }
    // generalizedGravityForce
    D.G = Mat<T>(18, 1);
    Mat<T> ag[18], agrot[18];
    ag[5] = Xup[5] * colvec<T>({0, 0, 0, M.gravity[0], M.gravity[1], M.gravity[2]});
    { Mat<T> g6 = T(-1) * (IC[5] * ag[5]); for (int r = 0; r < 6; ++r) D.G(r, 0) = g6(r, 0); }
    for (int i = 6; i < 18; ++i) {
        ag[i] = Xup[i] * ag[M.parent[i]];
        agrot[i] = Xuprot[i] * ag[M.parent[i]];
        D.G(i, 0) = -dot(S[i], IC[i] * ag[i]) - dot(Srot[i], M.Irot[i] * agrot[i]);
    }
    // generalizedCoriolisForce
    D.C = Mat<T>(18, 1);
    Mat<T> fvp[18], fvprot[18];
    fvp[5] = M.Ibody[5] * avp[5] + force_cross(v[5], M.Ibody[5] * v[5]);
    for (int i = 6; i < 18; ++i) {
        fvp[i] = M.Ibody[i] * avp[i] + force_cross(v[i], M.Ibody[i] * v[i]);
        fvprot[i] = M.Irot[i] * avprot[i] + force_cross(vrot[i], M.Irot[i] * vrot[i]);
    }
    for (int i = 17; i > 5; --i) {
        D.C(i, 0) = dot(S[i], fvp[i]) + dot(Srot[i], fvprot[i]);
        fvp[M.parent[i]] = fvp[M.parent[i]] + tr(Xup[i]) * fvp[i];
        fvp[M.parent[i]] = fvp[M.parent[i]] + tr(Xuprot[i]) * fvprot[i];
    }
    for (int r = 0; r < 6; ++r) D.C(r, 0) = fvp[5](r, 0);
    return D;
}

template <typename T> struct Task { Mat<T> Jt, JtDotQdot, xddot, posErr, desVel; };

template <typename T> T clip10(T v) { return std::min(std::max(v, T(-10)), T(10)); }

// WeightedInverse, qr_wholebody_impulse_ctrl.cpp:291-299 (threshold 1e-4 from the header default)
template <typename T> Mat<T> weighted_inverse(const Mat<T>& J, const Mat<T>& Winv, double thr = 0.0001) {
    Mat<T> temp = Winv * tr(J);
    Mat<T> lambda = J * temp;
    return temp * pinv(lambda, thr);
}

template <typename T>
int wbc_step(const qro_wbc_model* cfg, const float* state_f, const float* cmd_f, const int* contact,
             T* tau, T* fr, T* qdes, T* qddes, T* dbg) {
    const int FOOT[4] = {9, 11, 13, 15};   // linkID::FR, FL, HR, HL (config/qr_enum_types.h:35-45)
    Model<T> M = build_model<T>(cfg);
    T st[37], cmd[66];
    for (int i = 0; i < 37; ++i) st[i] = T(state_f[i]);
    for (int i = 0; i < 66; ++i) cmd[i] = T(cmd_f[i]);
    const T *quat = st, *pos = st + 4, *bodyvel = st + 7, *q = st + 13, *qd = st + 25;
    Dyn<T> D = dynamics(M, quat, pos, bodyvel, q, qd);
    const T *pBody_des = cmd, *vBody_des = cmd + 3, *aBody_des = cmd + 6, *rpy_des = cmd + 9, *vOri_des = cmd + 12,
            *pFoot = cmd + 15, *vFoot = cmd + 27, *aFoot = cmd + 39, *Fr_des = cmd + 51, *prevOriVel = cmd + 63;

    // ---- tasks and contacts (ContactTaskUpdate, qr_wbc_locomotion_controller.cpp:172-201)
    std::vector<Task<T>> tasks;
    Mat<T> Rot = quat_to_rot(quat);   // world -> body
    {   // qrTaskBodyOrientation (Kp 100, Kd 10)
        Task<T> t;
        t.Jt = Mat<T>(3, 18); setblock(t.Jt, 0, 0, tr(Rot));
        t.JtDotQdot = Mat<T>(3, 1);
        T qdes_[4], qinv[4] = {quat[0], -quat[1], -quat[2], -quat[3]}, qe[4], so3[3];
        rpy_to_quat(rpy_des, qdes_);
        quat_product(qdes_, qinv, qe);
        if (qe[0] < 0) for (int i = 0; i < 4; ++i) qe[i] *= T(-1);
        quat_to_so3(qe, so3);
        Mat<T> dv = colvec<T>({prevOriVel[0] - bodyvel[0], prevOriVel[1] - bodyvel[1], prevOriVel[2] - bodyvel[2]});
        Mat<T> ve = tr(Rot) * dv;   // uses the desiredVel stored by the PREVIOUS call (qr_task_body_orientation.cpp:68)
        t.posErr = Mat<T>(3, 1); t.desVel = Mat<T>(3, 1); t.xddot = Mat<T>(3, 1);
        for (int i = 0; i < 3; ++i) {
            t.posErr(i, 0) = so3[i];
            t.desVel(i, 0) = vOri_des[i];
            t.xddot(i, 0) = clip10<T>(T(100) * so3[i] + T(10) * ve(i, 0) + T(0));
        }
        tasks.push_back(t);
    }
    {   // qrTaskBodyPosition (Kp 100, Kd 10)
        Task<T> t;
        t.Jt = Mat<T>(3, 18); setblock(t.Jt, 0, 3, tr(Rot));
        t.JtDotQdot = Mat<T>(3, 1);
        Mat<T> vw = tr(Rot) * colvec<T>({bodyvel[3], bodyvel[4], bodyvel[5]});
        t.posErr = Mat<T>(3, 1); t.desVel = Mat<T>(3, 1); t.xddot = Mat<T>(3, 1);
        for (int i = 0; i < 3; ++i) {
            t.posErr(i, 0) = pBody_des[i] - pos[i];
            t.desVel(i, 0) = vBody_des[i];
            t.xddot(i, 0) = clip10<T>(T(100) * (pBody_des[i] - pos[i]) + T(10) * (vBody_des[i] - vw(i, 0)) + aBody_des[i]);
        }
        tasks.push_back(t);
    }
    std::vector<int> stance;
    for (int leg = 0; leg < 4; ++leg) {
        if (contact[leg]) { stance.push_back(leg); continue; }
        Task<T> t;   // qrTaskLinkPosition (Kp 500, Kd 10), virtual_depend = true
        t.Jt = D.Jc[FOOT[leg]];
        t.JtDotQdot = D.Jcdqd[FOOT[leg]];
        t.posErr = Mat<T>(3, 1); t.desVel = Mat<T>(3, 1); t.xddot = Mat<T>(3, 1);
        for (int i = 0; i < 3; ++i) {
            t.posErr(i, 0) = pFoot[3 * leg + i] - D.pGC[FOOT[leg]](i, 0);
            t.desVel(i, 0) = vFoot[3 * leg + i];
            t.xddot(i, 0) = T(500) * t.posErr(i, 0) + T(10) * (vFoot[3 * leg + i] - D.vGC[FOOT[leg]](i, 0)) + aFoot[3 * leg + i];
        }
        tasks.push_back(t);
    }
    const int nc = (int)stance.size(), dimFr = 3 * nc;
    Mat<T> JC(dimFr, 18), JCdqd(dimFr, 1), fdes(dimFr, 1);
    for (int k = 0; k < nc; ++k) {
        setblock(JC, 3 * k, 0, D.Jc[FOOT[stance[k]]]);
        setblock(JCdqd, 3 * k, 0, D.Jcdqd[FOOT[stance[k]]]);
        for (int i = 0; i < 3; ++i) fdes(3 * k + i, 0) = Fr_des[3 * stance[k] + i];
    }
    Mat<T> I18 = Mat<T>::eye(18);

    // ---- kinematic WBC: qrMultitaskProjection::FindConfiguration (qr_multitask_projection.cpp:38-106)
    {
        const double thr = 0.001;
        Mat<T> Nc = I18;
        if (nc > 0) Nc = I18 - pinv(JC, thr) * JC;
        Mat<T> JtPre = tasks[0].Jt * Nc;
        Mat<T> Jp = pinv(JtPre, thr);
        Mat<T> dq = Jp * tasks[0].posErr, qdot = Jp * tasks[0].desVel;
        Mat<T> prev_dq = dq, prev_qdot = qdot;
        Mat<T> Npre = Nc * (I18 - pinv(JtPre, thr) * JtPre);
        for (size_t i = 1; i < tasks.size(); ++i) {
            JtPre = tasks[i].Jt * Npre;
            Jp = pinv(JtPre, thr);
            dq = prev_dq + Jp * (tasks[i].posErr - tasks[i].Jt * prev_dq);
            qdot = prev_qdot + Jp * (tasks[i].desVel - tasks[i].Jt * prev_qdot);
            if (i < tasks.size() - 1) {
                Npre = Npre * (I18 - pinv(JtPre, thr) * JtPre);
                prev_dq = dq; prev_qdot = qdot;
            }
        }
        for (int i = 0; i < 12; ++i) { qdes[i] = q[i] + dq(6 + i, 0); qddes[i] = qdot(6 + i, 0); }
    }

    // ---- WBIC: GetModelRes + MakeTorque (qr_wholebody_impulse_ctrl.cpp:50-126)
    Mat<T> A = D.H, Ainv = inverse_lu(D.H);
    Mat<T> qdd(18, 1), Npre = I18;
    if (dimFr > 0) {
        Mat<T> JcBar = weighted_inverse(JC, Ainv);
        qdd = JcBar * (T(-1) * JCdqd);
        Npre = I18 - JcBar * JC;
    }
    for (size_t i = 0; i < tasks.size(); ++i) {
        Mat<T> JtPre = tasks[i].Jt * Npre;
        Mat<T> JtBar = weighted_inverse(JtPre, Ainv);
        qdd = qdd + JtBar * (tasks[i].xddot - tasks[i].JtDotQdot - tasks[i].Jt * qdd);
        if (i < tasks.size() - 1) Npre = Npre * (I18 - JtBar * JtPre);
    }
    // QP (SetCost :232-247, SetEqualityConstraint :129-148, SetInequalityConstraint :152-167)
    const int nz = 6 + dimFr, mi = dimFr > 0 ? 6 * nc : 1;
    const float muf = 0.4f;
    const T mu = T(muf);
    const T maxFz = M.total_mass * T(float(9.81));   // totalNonRotorMass() * (T)9.81
    Mat<T> UF(6 * nc, dimFr), ineq(6 * nc, 1);
    for (int k = 0; k < nc; ++k) {   // qrSingleContact (qr_single_contact.cpp:29-111)
        const int r0 = 6 * k, c0 = 3 * k;
        UF(r0, c0 + 2) = 1;
        UF(r0 + 1, c0) = 1; UF(r0 + 1, c0 + 2) = mu;
        UF(r0 + 2, c0) = -1; UF(r0 + 2, c0 + 2) = mu;
        UF(r0 + 3, c0 + 1) = 1; UF(r0 + 3, c0 + 2) = mu;
        UF(r0 + 4, c0 + 1) = -1; UF(r0 + 4, c0 + 2) = mu;
        UF(r0 + 5, c0 + 2) = -1;
        ineq(r0 + 5, 0) = -maxFz;
    }
    Mat<T> tot = A * qdd + D.C + D.G;
    if (dimFr > 0) tot = tot - tr(JC) * fdes;
    quadprogpp::Matrix<double> G(0.0, nz, nz), CE(0.0, nz, 6), CI(0.0, nz, mi);
    quadprogpp::Vector<double> g0(0.0, nz), ce0(0.0, 6), ci0(0.0, mi), z(0.0, nz);
    for (int i = 0; i < 6; ++i) G[i][i] = T(0.1f);   // weightFb = 0.1 stored in float (DVec<float>)
    for (int i = 0; i < dimFr; ++i) G[6 + i][6 + i] = 1.0;
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) CE[j][i] = A(i, j);
        for (int j = 0; j < dimFr; ++j) CE[6 + j][i] = -JC(j, i);   // -(Sf JC')(i, j)
        ce0[i] = tot(i, 0);                                         // qpce0 = -ce0 = +Sf(...)
    }
    if (dimFr > 0) {
        Mat<T> ci = ineq - UF * fdes;
        for (int i = 0; i < 6 * nc; ++i) {
            for (int j = 0; j < dimFr; ++j) CI[6 + j][i] = UF(i, j);
            ci0[i] = -ci(i, 0);
        }
    }
    double cost = quadprogpp::solve_quadprog(G, g0, CE, ce0, CI, ci0, z);
    for (int i = 0; i < 6; ++i) qdd(i, 0) += T(z[i]);
    Mat<T> fopt(dimFr, 1);
    for (int i = 0; i < dimFr; ++i) fopt(i, 0) = T(z[6 + i]) + fdes(i, 0);
    Mat<T> tt = A * qdd + D.C + D.G;
    if (dimFr > 0) tt = tt - tr(JC) * fopt;
    for (int i = 0; i < 12; ++i) tau[i] = tt(6 + i, 0);
    for (int i = 0; i < 12; ++i) fr[i] = 0;
    for (int k = 0; k < nc; ++k) for (int i = 0; i < 3; ++i) fr[3 * stance[k] + i] = fopt(3 * k + i, 0);
    if (dbg) {   // H(324) G(18) C(18) Jc feet(4*54) Jcdqd(12) pGC(12) vGC(12) qdd(18)
        T* o = dbg;
        for (T x : D.H.a) *o++ = x;
        for (T x : D.G.a) *o++ = x;
        for (T x : D.C.a) *o++ = x;
        for (int l = 0; l < 4; ++l) for (T x : D.Jc[FOOT[l]].a) *o++ = x;
        for (int l = 0; l < 4; ++l) for (T x : D.Jcdqd[FOOT[l]].a) *o++ = x;
        for (int l = 0; l < 4; ++l) for (T x : D.pGC[FOOT[l]].a) *o++ = x;
        for (int l = 0; l < 4; ++l) for (T x : D.vGC[FOOT[l]].a) *o++ = x;
        for (T x : qdd.a) *o++ = x;
    }
    return std::isfinite(cost) ? 0 : 1;   // QuadProg++ returns inf when infeasible (ignored by the reference, :113)
}

}  // namespace

extern "C" int qro_wbc_step_f32(const qro_wbc_model* cfg, const float* state, const float* cmd, const int* contact,
                                float* tau, float* fr, float* qdes, float* qddes, float* dbg) {
    return wbc_step<float>(cfg, state, cmd, contact, tau, fr, qdes, qddes, dbg);
}
extern "C" int qro_wbc_step_f64(const qro_wbc_model* cfg, const float* state, const float* cmd, const int* contact,
                                double* tau, double* fr, double* qdes, double* qddes, double* dbg) {
    return wbc_step<double>(cfg, state, cmd, contact, tau, fr, qdes, qddes, dbg);
}

// Swing foot in MPC mode: SwingFootTrajectory::GenerateTrajectoryPoint (src/controllers/
// qr_foot_trajectory_generator.cpp:322-343) -> qrFootParabolaPatternGenerator::GenerateTrajectory (:188-215)
// -> qrQuadraticSpline::getPoint(t, mid, out) (src/utils/qr_geometry.cpp:157-190), with the reference's
// mixed float/double evaluation (pow() is the double overload).  Velocity and acceleration outputs of the
// reference are identically zero.  Returns 0 when the generator rejects the phase (outside [0, 1+1e-3)).
extern "C" int qro_swing_parabola(const float* start, const float* end, float height, float t_in,
                                  int phase_module, float* pos) {
    float phase;
    if (phase_module) {
        if (t_in <= 0.5) phase = 0.8 * std::sin(t_in * M_PI);
        else phase = 0.8 + (t_in - 0.5) * 0.4;
    } else {
        phase = t_in;
    }
    const float initialTime = 0.f, duration = 1.f;   // SetParameters(0., start, end, stepParams(duration = 1 phase unit))
    if (phase < initialTime - 1e-3) return 0;
    if (phase >= initialTime + duration + 1e-3) return 0;
    const float x = (1 - phase) * start[0] + phase * end[0];
    const float y = (1 - phase) * start[1] + phase * end[1];
    const float mid = std::max(end[2], start[2]) + height;
    const float mid_phase = 0.5;
    float deltaOne, deltaTwo, deltaThree, coefa, coefb, coefc;
    deltaOne = mid - start[2];
    deltaTwo = end[2] - start[2];
    deltaThree = pow(mid_phase, 2) - mid_phase;
    coefa = (deltaOne - deltaTwo * mid_phase) / deltaThree;
    coefb = (deltaTwo * pow(mid_phase, 2) - deltaOne) / deltaThree;
    coefc = start[2];
    float z = coefa * pow(phase, 2) + coefb * phase + coefc;
    if (phase - initialTime < 0.) z = 0.f;   // getPoint returns false for dt < 0 and leaves `out` at its default (x = 0), qr_geometry.cpp:171-173
    pos[0] = x; pos[1] = y; pos[2] = z;
    return 1;
}
