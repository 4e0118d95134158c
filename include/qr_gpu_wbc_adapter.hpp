// qr_gpu_wbc_adapter.hpp -- header-only host-side mirror of the reference's WBC seam on top of libqr_gpu.so.
//
// The reference's seam is qrWbcLocomotionController<float>::Run(void* precomputeData)
// (/root/reference/quadruped/src/controllers/wbc/qr_wbc_locomotion_controller.cpp:108-139): it pulls the robot
// state from controlFSMData->quadruped (UpdateModel :141-157), the commands from the qrWbcCtrlData the pointer
// refers to (include/quadruped/controllers/qr_state_dataflow.h:133-193), computes on every second call
// (:111) and writes legCmd[i].tua for stance legs (UpdateLegCMD :205-219).  Those classes need the robot / FSM
// headers, so this adapter keeps the same three steps as plain functions over the same fields:
//
//     Quadruped::gpu::WbcController wbc(model);                       // ctor(FloatingBaseModel&, fsmData*)
//     wbc.UpdateModel(quat, pos, bodyAngVel, bodyLinVel, q, qd);       // == UpdateModel(...)
//     wbc.Run(ctrlData);                                              // == Run(void*), same every-2nd-call gating
//     wbc.jointTorqueCmd[12], desiredJPos[12], desiredJVel[12], optimalFr[12]
//
// Every vector argument is a template on anything with a contiguous float data() (Eigen's Vec3<float> etc.).
// Batch = 1; controllers owning many robots call qr_gpu_wbc_solve_batch[_host] directly.
#ifndef QR_GPU_WBC_ADAPTER_HPP
#define QR_GPU_WBC_ADAPTER_HPP

#include <cstdint>
#include <cstring>

#include "qr_gpu.h"

namespace Quadruped {
namespace gpu {

// The qrWbcCtrlData fields Run reads (qr_state_dataflow.h:133-193), as plain arrays.
struct WbcCtrlData {
    float pBody_des[3], vBody_des[3], aBody_des[3], pBody_RPY_des[3], vBody_Ori_des[3];
    float pFoot_des[4][3], vFoot_des[4][3], aFoot_des[4][3], Fr_des[4][3];
    bool contact_state[4];
    bool allowAfterMPC;
};

class WbcController {
public:
    explicit WbcController(const qr_wbc_model& m) : model_(m) {
        std::memset(state_, 0, sizeof(state_));
        std::memset(prevOriVelDes_, 0, sizeof(prevOriVelDes_));
        std::memset(jointTorqueCmd, 0, sizeof(jointTorqueCmd));
        std::memset(desiredJPos, 0, sizeof(desiredJPos));
        std::memset(desiredJVel, 0, sizeof(desiredJVel));
        std::memset(optimalFr, 0, sizeof(optimalFr));
    }

    // FBModelState as UpdateModel fills it (:141-157): bodyOrientation (w,x,y,z), bodyPosition, bodyVelocity =
    // [angular (body frame), linear (body frame)], q, qd.
    template <class Q4, class V3a, class V3b, class V3c, class V12a, class V12b>
    void UpdateModel(const Q4& quat, const V3a& pos, const V3b& angVelBody, const V3c& linVelBody, const V12a& q,
                     const V12b& qd) {
        std::memcpy(state_, quat.data(), 4 * sizeof(float));
        std::memcpy(state_ + 4, pos.data(), 3 * sizeof(float));
        std::memcpy(state_ + 7, angVelBody.data(), 3 * sizeof(float));
        std::memcpy(state_ + 10, linVelBody.data(), 3 * sizeof(float));
        std::memcpy(state_ + 13, q.data(), 12 * sizeof(float));
        std::memcpy(state_ + 25, qd.data(), 12 * sizeof(float));
    }

    // Run(void* precomputeData): recomputes on every second call like the reference (iteration % 2, :111), otherwise
    // keeps the previous command.  Returns the per-robot status of qr_gpu.h (0 ok), or a negative QR_E* code.
    int Run(const WbcCtrlData& d) {
        if (iteration_++ % 2 != 0) return status_;   // same gating as :111 / :133 of the reference
        float cmd[66];
        std::memcpy(cmd, d.pBody_des, 12);       std::memcpy(cmd + 3, d.vBody_des, 12);
        std::memcpy(cmd + 6, d.aBody_des, 12);   std::memcpy(cmd + 9, d.pBody_RPY_des, 12);
        std::memcpy(cmd + 12, d.vBody_Ori_des, 12);
        std::memcpy(cmd + 15, d.pFoot_des, 48);  std::memcpy(cmd + 27, d.vFoot_des, 48);
        std::memcpy(cmd + 39, d.aFoot_des, 48);  std::memcpy(cmd + 51, d.Fr_des, 48);
        std::memcpy(cmd + 63, prevOriVelDes_, 12);   // the orientation task keeps the previous command (:68 of qr_task_body_orientation.cpp)
        int32_t contact[4];
        for (int l = 0; l < 4; ++l) contact[l] = d.contact_state[l] ? 1 : 0;
        int32_t st = 0;
        const int rc = qr_gpu_wbc_solve_batch_host(&model_, 1, state_, cmd, contact, jointTorqueCmd, optimalFr, desiredJPos,
                                                   desiredJVel, &st);
        if (rc != QR_OK) return rc;
        std::memcpy(prevOriVelDes_, d.vBody_Ori_des, 12);
        status_ = st;
        return st;
    }

    float jointTorqueCmd[12], desiredJPos[12], desiredJVel[12], optimalFr[12];

private:
    qr_wbc_model model_;
    float state_[37];
    float prevOriVelDes_[3];
    long iteration_ = 0;
    int status_ = 0;
};

}  // namespace gpu
}  // namespace Quadruped
#endif
