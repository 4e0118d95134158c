// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Pulls the reference's own convex-MPC translation unit in, UNMODIFIED and from where it lies under
// /root/reference (quadruped/src/controllers/mpc/qr_mpc_interface.cpp, found through the -I path the
// Makefile passes), and exposes its public entry points SetupProblem / SolveMPCKernel /
// GetMPCSolution (:160, :325, :445) plus the file-static QP arrays it hands to qpOASES (:404-412)
// through a C interface, so tests can hold the restatement in mpc_oracle.cpp against the real thing.
// Eigen is replaced by oracle/mini_eigen (see its header for what that does and does not pin).
#include "qr_mpc_interface.cpp"

#include <cstdio>
#include <time.h>
#include <unistd.h>

namespace {
// SetupProblem prints a banner every call; keep test / bench output readable.
void quiet_setup(int horizon, const double* params, const float* inertia, const float* weights) {
    float inertia_[3] = {inertia[0], inertia[1], inertia[2]};
    float weights_[12];
    for (int i = 0; i < 12; ++i) weights_[i] = weights[i];
    fflush(stdout);
    int saved = dup(1);
    FILE* devnull = fopen("/dev/null", "w");
    if (devnull) dup2(fileno(devnull), 1);
    Quadruped::SetupProblem(params[0], horizon, params[1], params[2], params[3], inertia_, weights_,
                            (float)params[4]);
    fflush(stdout);
    if (devnull) {
        dup2(saved, 1);
        fclose(devnull);
    }
    close(saved);
}

void solve_one(int horizon, const float* p, const float* v, const float* quat, const float* w,
               const float* r_feet, const float* rpy, const float* traj, const float* gait) {
    Vec3<float> p_(p[0], p[1], p[2]), v_(v[0], v[1], v[2]), w_(w[0], w[1], w[2]), rpy_(rpy[0], rpy[1], rpy[2]);
    Quat<float> q_(quat[0], quat[1], quat[2], quat[3]);
    Eigen::Matrix<float, 3, 4> r_;
    for (int leg = 0; leg < 4; ++leg)
        for (int a = 0; a < 3; ++a) r_(a, leg) = r_feet[3 * leg + a];
    std::vector<float> traj_(traj, traj + 12 * horizon), gait_(gait, gait + 4 * horizon);
    Quadruped::SolveMPCKernel(p_, v_, q_, w_, r_, rpy_, traj_.data(), gait_.data());
}
}   // namespace

extern "C" {

// Same argument layout as qro_mpc_build / qro_mpc_solve (qr_oracle.h).  params = {dt, mu, f_max,
// mass, alpha}.  Outputs are the reference's own qpOASES buffers (double): H[n*n] row-major, g[n],
// ub[m], x[n] with n = 12h, m = 20h; any may be NULL.
int qr_ref_mpc_solve(int horizon, const double* params, const float* inertia, const float* weights,
                     const float* p, const float* v, const float* quat, const float* w,
                     const float* r_feet, const float* rpy, const float* traj, const float* gait,
                     double* H, double* g, double* ub, double* x) {
    quiet_setup(horizon, params, inertia, weights);
    solve_one(horizon, p, v, quat, w, r_feet, rpy, traj, gait);

    const int n = 12 * horizon, m = 20 * horizon;
    if (H) for (int i = 0; i < n * n; ++i) H[i] = H_qpoases[i];
    if (g) for (int i = 0; i < n; ++i) g[i] = g_qpoases[i];
    if (ub) for (int i = 0; i < m; ++i) ub[i] = ub_qpoases[i];
    if (x) for (int i = 0; i < n; ++i) x[i] = Quadruped::GetMPCSolution(i);
    return 0;
}

// Wall-clock `count` consecutive control ticks of the reference in this process: SetupProblem once (as
// MPCStanceLegController::Reset does, qr_mpc_stance_leg_controller.cpp:90), then per problem
// SolveMPCKernel + the twelve GetMPCSolution reads of SolveDenseMPC (:399-405).  Arrays hold `count`
// consecutive problems.  Returns seconds; per-problem seconds in lat[count], forces in x12[count*12]
// (either may be NULL).
double qr_ref_mpc_time_batch(int horizon, const double* params, const float* inertia, const float* weights,
                             int count, const float* p, const float* v, const float* quat, const float* w,
                             const float* r_feet, const float* rpy, const float* traj, const float* gait,
                             double* x12, double* lat) {
    quiet_setup(horizon, params, inertia, weights);
    auto now = [] {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    const double t0 = now();
    double prev = t0;
    for (int i = 0; i < count; ++i) {
        solve_one(horizon, p + 3 * i, v + 3 * i, quat + 4 * i, w + 3 * i, r_feet + 12 * i, rpy + 3 * i,
                  traj + 12 * horizon * i, gait + 4 * horizon * i);
        for (int k = 0; k < 12; ++k) {
            double f = Quadruped::GetMPCSolution(k);
            if (x12) x12[12 * i + k] = f;
        }
        const double t = now();
        if (lat) lat[i] = t - prev;
        prev = t;
    }
    return prev - t0;
}

}   // extern "C"
