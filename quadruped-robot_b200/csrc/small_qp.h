// small_qp.h -- dense strictly convex QP with a handful of variables, one THREAD per problem, float64.
//
//     min 1/2 x'Gx + g0'x     s.t.   c_i'x + c0_i >= 0,  i = 0..m-1          (n <= 12, m <= 24)
//
// Replaces quadprogpp::solve_quadprog as the force-balance controller calls it
// (/root/reference/quadruped/src/controllers/balance_controller/qr_qp_torque_optimizer.cpp:273-276, 380-383;
// solver: extern/QuadProgpp/src/QuadProg++.cc:453-...), i.e. no equality rows.  Same method -- the dual
// active-set iteration of Goldfarb & Idnani (Math. Programming 27, 1983) -- and the same FORMULATION as
// QuadProg++ (its variable names np, z, r, d, u, R_norm, psi, t1, t2 are the paper's; its termination constant
// m * eps * c1 * c2 * 100 is kept so that both stop on the same instances), but not its code: no gotos, bitmask
// active sets, a restart on dependent rows, fixed row-major arrays.  G = LL', J = L^-T, the active normals are kept as N = J[:, :q] R with R upper triangular, and
// constraints enter / leave through Givens rotations of J and R.  The most violated constraint is processed
// first, and -- like the reference's solver -- an infeasible constraint ends the iteration with the CURRENT
// iterate in x (the force-balance QP of a leg in swing asks for n.f >= 1e-7 and -n.f >= 1e-7 at once; the
// reference keeps whatever the solver holds at that point, see fb_problem.h).
//
// The problems are tiny and independent, so the GPU mapping is one thread per problem with the work arrays in
// local memory (L1-resident); the batch supplies the parallelism.
#pragma once

#include "qr_team.h"

#define QR_SQP_N 12
#define QR_SQP_M 24

struct QrSmallQpWork {
    double L[QR_SQP_N * QR_SQP_N];   // Cholesky factor of G (lower, row-major)
    double J[QR_SQP_N * QR_SQP_N];   // L^-T rotated; J J' = G^-1
    double R[QR_SQP_N * QR_SQP_N];   // upper triangular, q x q used
    double d[QR_SQP_N], z[QR_SQP_N], r[QR_SQP_N], np[QR_SQP_N];
    double u[QR_SQP_N + 1], u_old[QR_SQP_N + 1], x_old[QR_SQP_N];
    double s[QR_SQP_M];
    int A[QR_SQP_N + 1], A_old[QR_SQP_N + 1];
    unsigned active;                 // bit i: constraint i is in the active set
    unsigned excluded;               // bit i: found linearly dependent, do not try again
};

// Rotation that maps (a, b) to (rho, 0); returns 0 when b is already zero.
QR_DEV int qr_sqp_givens(double a, double b, double& c, double& s, double& rho) {
    if (b == 0.0) return 0;
    rho = sqrt(a * a + b * b);
    c = a / rho;
    s = b / rho;
    return 1;
}

// Status: 0 optimal, 1 infeasible constraint met (x = iterate at that point), 2 iteration cap, 3 G not positive definite.
QR_DEV int qr_small_qp_solve(int n, int m, const double* G /*n x n, lower triangle read*/, const double* g0,
                             const double* C /*m rows of n: c_i = C + i*n*/, const double* c0, double* x,
                             QrSmallQpWork& W, int* iters) {
    const int N = QR_SQP_N;
    const double eps = 2.220446049250313e-16, inf = 1e300;
    double* L = W.L;
    double* J = W.J;
    double* R = W.R;
    // ---- Cholesky G = L L'
    double c1 = 0.0;
    for (int i = 0; i < n; ++i) c1 += G[i * n + i];
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j <= i; ++j) {
            double a = G[i * n + j];
            for (int k = 0; k < j; ++k) a -= L[i * N + k] * L[j * N + k];
            if (i == j) {
                if (!(a > 0.0)) { *iters = 0; return 3; }
                L[i * N + i] = sqrt(a);
            } else {
                L[i * N + j] = a / L[j * N + j];
            }
        }
    }
    // ---- J = L^-T (upper triangular): column i solves L' J[:,i] = e_i
    double c2 = 0.0;
    for (int i = 0; i < n; ++i) {
        for (int row = n - 1; row >= 0; --row) {
            double a = (row == i) ? 1.0 : 0.0;
            for (int k = row + 1; k < n; ++k) a -= L[k * N + row] * J[k * N + i];
            J[row * N + i] = (row > i) ? 0.0 : a / L[row * N + row];
        }
        c2 += J[i * N + i];
    }
    // ---- unconstrained minimiser x = -G^-1 g0
    for (int i = 0; i < n; ++i) {
        double a = -g0[i];
        for (int k = 0; k < i; ++k) a -= L[i * N + k] * W.z[k];
        W.z[i] = a / L[i * N + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double a = W.z[i];
        for (int k = i + 1; k < n; ++k) a -= L[k * N + i] * x[k];
        x[i] = a / L[i * N + i];
    }
    int q = 0;
    double R_norm = 1.0;
    W.active = 0u;
    W.excluded = 0u;
    const double psi_tol = (double)m * eps * c1 * c2 * 100.0;
    int it = 0;
    const int max_it = 20 * (n + m);
    for (;;) {
        // ---- step 1: constraint values, termination test
        double psi = 0.0;
        for (int i = 0; i < m; ++i) {
            double a = c0[i];
            for (int k = 0; k < n; ++k) a += C[i * n + k] * x[k];
            W.s[i] = a;
            if (!((W.active >> i) & 1u) && a < 0.0) psi += a;
        }
        if (fabs(psi) <= psi_tol) { *iters = it; return 0; }
        for (int k = 0; k < q; ++k) { W.u_old[k] = W.u[k]; W.A_old[k] = W.A[k]; }
        for (int k = 0; k < n; ++k) W.x_old[k] = x[k];
        const unsigned active_old = W.active;
        const int q_old = q;
        bool restart = false;
        // ---- step 2: most violated constraint
        for (;;) {
            int ip = -1;
            double ss = 0.0;
            for (int i = 0; i < m; ++i)
                if (!((W.active >> i) & 1u) && !((W.excluded >> i) & 1u) && W.s[i] < ss) { ss = W.s[i]; ip = i; }
            if (ip < 0) { *iters = it; return 0; }
            for (int k = 0; k < n; ++k) W.np[k] = C[ip * n + k];
            W.u[q] = 0.0;
            W.A[q] = ip;
            bool dependent = false;
            for (;;) {   // step 2a .. 2c for this constraint
                if (++it > max_it) { *iters = it; return 2; }
                // d = J' np ; z = J[:, q:] d[q:] ; r = R^-1 d[:q]
                for (int k = 0; k < n; ++k) {
                    double a = 0.0;
                    for (int i = 0; i < n; ++i) a += J[i * N + k] * W.np[i];
                    W.d[k] = a;
                }
                for (int i = 0; i < n; ++i) {
                    double a = 0.0;
                    for (int k = q; k < n; ++k) a += J[i * N + k] * W.d[k];
                    W.z[i] = a;
                }
                for (int i = q - 1; i >= 0; --i) {
                    double a = W.d[i];
                    for (int k = i + 1; k < q; ++k) a -= R[i * N + k] * W.r[k];
                    W.r[i] = a / R[i * N + i];
                }
                // step lengths
                double t1 = inf, t2 = inf;
                int lpos = -1;
                for (int k = 0; k < q; ++k)
                    if (W.r[k] > 0.0 && W.u[k] / W.r[k] < t1) { t1 = W.u[k] / W.r[k]; lpos = k; }
                double zz = 0.0, znp = 0.0;
                for (int k = 0; k < n; ++k) { zz += W.z[k] * W.z[k]; znp += W.z[k] * W.np[k]; }
                if (fabs(zz) > eps) t2 = -W.s[ip] / znp;
                const double t = t1 < t2 ? t1 : t2;
                if (t >= inf) { *iters = it; return 1; }   // infeasible: keep the current iterate
                if (t2 >= inf) {
                    // dual step only, then drop the blocking constraint
                    for (int k = 0; k < q; ++k) W.u[k] -= t * W.r[k];
                    W.u[q] += t;
                } else {
                    for (int k = 0; k < n; ++k) x[k] += t * W.z[k];
                    for (int k = 0; k < q; ++k) W.u[k] -= t * W.r[k];
                    W.u[q] += t;
                    if (t == t2) {
                        // full step: the constraint joins the active set.  Rotate d[q+1..n-1] into d[q].
                        for (int j = n - 1; j > q; --j) {
                            double c, sn, rho;
                            if (!qr_sqp_givens(W.d[j - 1], W.d[j], c, sn, rho)) continue;
                            W.d[j - 1] = rho;
                            W.d[j] = 0.0;
                            for (int i = 0; i < n; ++i) {
                                const double a = J[i * N + j - 1], b = J[i * N + j];
                                J[i * N + j - 1] = c * a + sn * b;
                                J[i * N + j] = -sn * a + c * b;
                            }
                        }
                        if (fabs(W.d[q]) <= eps * R_norm) { dependent = true; break; }
                        for (int i = 0; i <= q; ++i) R[i * N + q] = W.d[i];
                        if (fabs(W.d[q]) > R_norm) R_norm = fabs(W.d[q]);
                        W.active |= 1u << ip;
                        ++q;
                        break;   // back to step 1
                    }
                }
                // partial (or pure dual) step: constraint A[lpos] leaves the active set
                {
                    const int l = W.A[lpos];
                    W.active &= ~(1u << l);
                    for (int k = lpos; k < q; ++k) {   // also moves the candidate's slot (index q) down
                        W.A[k] = W.A[k + 1];
                        W.u[k] = W.u[k + 1];
                    }
                    for (int k = lpos; k < q - 1; ++k)
                        for (int i = 0; i <= k + 1; ++i) R[i * N + k] = R[i * N + k + 1];
                    --q;
                    for (int j = lpos; j < q; ++j) {   // restore the triangle: zero R[j+1][j]
                        double c, sn, rho;
                        if (!qr_sqp_givens(R[j * N + j], R[(j + 1) * N + j], c, sn, rho)) continue;
                        R[j * N + j] = rho;
                        R[(j + 1) * N + j] = 0.0;
                        for (int k = j + 1; k < q; ++k) {
                            const double a = R[j * N + k], b = R[(j + 1) * N + k];
                            R[j * N + k] = c * a + sn * b;
                            R[(j + 1) * N + k] = -sn * a + c * b;
                        }
                        for (int i = 0; i < n; ++i) {
                            const double a = J[i * N + j], b = J[i * N + j + 1];
                            J[i * N + j] = c * a + sn * b;
                            J[i * N + j + 1] = -sn * a + c * b;
                        }
                    }
                    double a = c0[ip];
                    for (int k = 0; k < n; ++k) a += C[ip * n + k] * x[k];
                    W.s[ip] = a;
                }
            }
            if (!dependent) break;
            // the candidate is linearly dependent on the active normals: forget it and restore the state of step 1
            W.excluded |= 1u << ip;
            restart = true;
            break;
        }
        if (restart) {
            // Re-establish the factorisation of the old active set from scratch (rare path): reset J, R and re-add.
            for (int k = 0; k < n; ++k) x[k] = W.x_old[k];
            for (int i = 0; i < n; ++i)
                for (int row = n - 1; row >= 0; --row) {
                    double a = (row == i) ? 1.0 : 0.0;
                    for (int k = row + 1; k < n; ++k) a -= L[k * N + row] * J[k * N + i];
                    J[row * N + i] = (row > i) ? 0.0 : a / L[row * N + row];
                }
            q = 0;
            W.active = 0u;
            for (int a_i = 0; a_i < q_old; ++a_i) {
                const int ia = W.A_old[a_i];
                for (int k = 0; k < n; ++k) {
                    double a = 0.0;
                    for (int i = 0; i < n; ++i) a += J[i * N + k] * C[ia * n + i];
                    W.d[k] = a;
                }
                for (int j = n - 1; j > q; --j) {
                    double c, sn, rho;
                    if (!qr_sqp_givens(W.d[j - 1], W.d[j], c, sn, rho)) continue;
                    W.d[j - 1] = rho;
                    W.d[j] = 0.0;
                    for (int i = 0; i < n; ++i) {
                        const double a = J[i * N + j - 1], b = J[i * N + j];
                        J[i * N + j - 1] = c * a + sn * b;
                        J[i * N + j] = -sn * a + c * b;
                    }
                }
                for (int i = 0; i <= q; ++i) R[i * N + q] = W.d[i];
                W.A[q] = ia;
                W.u[q] = W.u_old[a_i];
                W.active |= 1u << ia;
                ++q;
            }
            (void)active_old;
            // recompute the constraint values at the restored point and choose again (the dependent one is excluded)
            continue;
        }
    }
}
