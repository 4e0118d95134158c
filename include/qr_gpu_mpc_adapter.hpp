// qr_gpu_mpc_adapter.hpp -- header-only drop-in for the reference's MPC seam on top of libqr_gpu.so.
//
// Re-implements, with the same names, argument meaning and (absence of) error behaviour,
//     void   Quadruped::SetupProblem(double dt, int horizon, double frictionCoeff, double fMax,
//                                    double totalMass, float* inertia, float* weight, float alpha);
//     void   Quadruped::SolveMPCKernel(Vec3<float>& p, Vec3<float>& v, Quat<float>& q, Vec3<float>& w,
//                                      Eigen::Matrix<float,3,4>& r, Vec3<float>& rpy,
//                                      float* state_trajectory, float* gait);
//     double Quadruped::GetMPCSolution(int index);
// of /root/reference/quadruped/include/quadruped/controllers/mpc/qr_mpc_interface.h:157, 200, 215
// (definitions: src/controllers/mpc/qr_mpc_interface.cpp:160-175, 334-356, 446-451), so that
// MPCStanceLegController::Reset / SolveDenseMPC (qr_mpc_stance_leg_controller.cpp:90, 399-405) compile
// unchanged when this header replaces qr_mpc_interface.h and the target links libqr_gpu.so instead of
// qpOASES.  Batch = 1 here; controllers that own many robot instances call qr_gpu_mpc_solve_batch_host
// directly.
//
// The vector arguments are templates: anything with a contiguous float `data()` works -- Eigen's
// Vec3<float>, Quat<float> (w,x,y,z) and Eigen::Matrix<float,3,4> (column-major = r_feet[3*leg+axis])
// in the reference tree; Eigen itself is not required to compile this header.
//
// Like the reference (file-scope statics, qr_mpc_interface.cpp:35-104) this adapter keeps process-wide
// state and is not re-entrant.  Differences a caller can observe: swing-leg forces are exactly 0 (the
// reference returns ~1e-13), the answer is the exact QP optimum (the reference returns qpOASES' iterate,
// truncated at nWSR = 100 on ~10 % of trot instances), and solver failures are available through
// Quadruped::GetMPCStatus() instead of being dropped.
#ifndef QR_GPU_MPC_ADAPTER_HPP
#define QR_GPU_MPC_ADAPTER_HPP

#include <cstdio>
#include <cstring>
#include <vector>

#include "qr_gpu.h"

namespace Quadruped {

namespace gpu_detail {
struct State {
    qr_mpc_params params{};
    bool configured = false;
    bool solved = false;
    int status = 0;
    std::vector<float> solution;  // 12 * horizon
    float grf[12] = {0};
};
inline State& state() {
    static State s;
    return s;
}
}  // namespace gpu_detail

inline void SetupProblem(double dt, int horizon, double frictionCoeff, double fMax, double totalMass,
                         float* inertia, float* weight, float alpha) {
    std::printf("SetupProblem: f_max = %f, mass = %f, horizon = %d\n", fMax, totalMass, horizon);
    gpu_detail::State& s = gpu_detail::state();
    // the narrowing conversions of ProblemConfig (qr_mpc_interface.cpp:163-173)
    s.params.horizon = horizon;
    s.params.dt = static_cast<float>(dt);
    s.params.mu = static_cast<float>(frictionCoeff);
    s.params.f_max = static_cast<float>(fMax);
    s.params.mass = static_cast<float>(totalMass);
    std::memcpy(s.params.inertia, inertia, 3 * sizeof(float));
    std::memcpy(s.params.weights, weight, 12 * sizeof(float));
    s.params.alpha = alpha;
    s.solution.assign(static_cast<size_t>(12) * horizon, 0.f);
    s.solved = false;
    if (qr_gpu_init(-1) != QR_OK) {
        std::printf("qr_gpu_init failed: %s\n", qr_gpu_last_error());
        s.configured = false;
        return;
    }
    s.configured = true;
}

template <class V3, class Q4, class M34>
inline void SolveMPCKernel(V3& p, V3& v, Q4& q, V3& w, M34& r, V3& rpy, float* state_trajectory, float* gait) {
    gpu_detail::State& s = gpu_detail::state();
    if (!s.configured) {
        std::printf("failed to solve!\n");
        return;
    }
    int32_t status = 0;
    const int rc = qr_gpu_mpc_solve_batch_host(&s.params, nullptr, 1, p.data(), v.data(), q.data(), w.data(),
                                               r.data(), rpy.data(), state_trajectory, gait, nullptr, nullptr,
                                               s.grf, s.solution.data(), &status, nullptr);
    if (rc != QR_OK) {
        std::printf("failed to solve! (%s)\n", qr_gpu_last_error());
        return;
    }
    s.status = status;
    s.solved = true;
}

inline double GetMPCSolution(int index) {
    gpu_detail::State& s = gpu_detail::state();
    if (!s.solved) return 0.f;
    return s.solution[static_cast<size_t>(index)];
}

// Additive: per-instance outcome of the last solve (0 verified optimum, 1 interior-point iterate,
// 2 inconsistent bounds, 3 non-finite input) -- the reference ignores qpOASES' return code.
inline int GetMPCStatus() { return gpu_detail::state().status; }

}  // namespace Quadruped

#endif  // QR_GPU_MPC_ADAPTER_HPP
