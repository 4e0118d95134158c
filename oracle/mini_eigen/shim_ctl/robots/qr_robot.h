// Stand-in for the reference's robots/qr_robot.h -- TEST INFRASTRUCTURE ONLY.
//
// The real header needs yaml-cpp, the Unitree SDK, ROS timers and the estimator / visualisation headers, none of
// which exist in this image.  This file declares a plain-data `Quadruped::qrRobot` (and the state / WBC-command
// structs it carries) with exactly the member names the reference functions pinned by oracle/ref_ctl_shim.cpp read,
// so that those functions compile from where they lie under /root/reference WITHOUT EDITING:
//   * whole translation units that only use a qrRobot*: controllers/balance_controller/qr_qp_torque_optimizer.cpp;
//   * line ranges of member functions cut out by oracle/Makefile (target `refctl`) into _ref/gen/*.inc: the leg
//     kinematics of src/robots/qr_robot.cpp:89-251 are compiled as members of THIS class.
// Nothing here is arithmetic: every function body that computes something comes from the reference.
#ifndef MINI_QR_ROBOT_H
#define MINI_QR_ROBOT_H
#include <cmath>
#include <iostream>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include <Eigen/Dense>

#include "config/qr_config.h"
#include "config/qr_enum_types.h"
#include "utils/qr_algebra.h"
#include "utils/qr_cpptypes.h"
#include "utils/qr_se3.h"

// controllers/qr_state_dataflow.h:133-193 (qrWbcCtrlData) and :195-... (qrStateDataFlow): same field names and types
struct qrWbcCtrlData {
    Vec3<float> pBody_des, vBody_des, aBody_des, pBody_RPY_des, vBody_Ori_des;
    Vec3<float> pFoot_des[4], vFoot_des[4], aFoot_des[4], Fr_des[4];
    Vec4<bool> contact_state;
    bool allowAfterMPC = true;
};
struct qrStateDataFlow {
    Eigen::Matrix<float, 3, 4> footPositionsInBaseFrame, footVelocitiesInBaseFrame;
    Vec3<float> baseVInWorldFrame, baseWInWorldFrame, baseLinearAcceleration;
    std::vector<Mat3<float>> footJvs = std::vector<Mat3<float>>(4);
    Eigen::Matrix<float, 3, 4> estimatedFootForce;
    Vec3<float> estimatedMoment;
    float heightInControlFrame = 0.27;
    Vec3<float> zmp;
    Mat3<float> baseRMat, groundRMat, baseRInControlFrame;
    Vec4<float> groundOrientation;
    qrWbcCtrlData wbcData;
};

namespace Quadruped {

class qrRobot {
public:
    // geometry / inertia (qr_robot.h: hipLength .. totalInertia)
    float hipLength = 0, upperLegLength = 0, lowerLegLength = 0, totalMass = 0;
    Mat3<float> totalInertia;
    Vec3<float> comOffset;
    Eigen::Matrix<float, 3, 4> hipOffset, defaultHipPosition;
    // state
    Vec3<float> basePosition, baseRollPitchYaw, baseRollPitchYawRate, baseVelocityInBaseFrame;
    Quat<float> baseOrientation;
    Eigen::Matrix<float, 12, 1> motorAngles, motorVelocities, motortorque, standUpMotorAngles;
    Eigen::Matrix<bool, 4, 1> footContact;
    bool stop = false, isSim = true;
    std::unordered_map<std::string, int> controlParams;
    qrStateDataFlow stateDataFlow;

    Vec3<float> GetBasePosition() const { return basePosition; }
    Quat<float> GetBaseOrientation() const { return baseOrientation; }
    Vec3<float> GetBaseRollPitchYaw() const { return baseRollPitchYaw; }
    Vec3<float> GetBaseRollPitchYawRate() const { return baseRollPitchYawRate; }
    Vec3<float> GetBaseVelocityInBaseFrame() const { return baseVelocityInBaseFrame; }
    Eigen::Matrix<float, 12, 1> GetMotorAngles() const { return motorAngles; }
    Eigen::Matrix<float, 12, 1> GetMotorVelocities() const { return motorVelocities; }
    Eigen::Matrix<bool, 4, 1> GetFootContact() const { return footContact; }
    Eigen::Matrix<float, 3, 4> GetDefaultHipPosition() const { return defaultHipPosition; }
    Eigen::Matrix<float, 3, 4> GetFootPositionsInBaseFrame() { return stateDataFlow.footPositionsInBaseFrame; }

    // bodies: src/robots/qr_robot.cpp:89-251 (cut out by the Makefile, compiled in ref_ctl_shim.cpp)
    Vec3<float> WithLegSigns(const Vec3<float>& v, int leg_id);
    Vec3<float> FootPositionInHipFrameToJointAngle(Vec3<float>& foot_position, int hip_sign = 1);
    Vec3<float> FootPositionInHipFrame(Vec3<float>& angles, int hip_sign = 1);
    Eigen::Matrix<float, 3, 3> AnalyticalLegJacobian(Vec3<float>& leg_angles, int leg_id);
    Mat34<float> FootPositionsInBaseFrame(Eigen::Matrix<float, 12, 1> foot_angles);
    Mat34<float> ComputeFootVelocitiesInBaseFrame();
    void ComputeMotorAnglesFromFootLocalPosition(int leg_id, Vec3<float> foot_local_position,
                                                 Eigen::Matrix<int, 3, 1>& joint_idx, Vec3<float>& joint_angles);
    Vec3<float> ComputeMotorVelocityFromFootLocalVelocity(int leg_id, Vec3<float> leg_angles, Vec3<float> foot_local_velocity);
    Mat34<float> GetFootPositionsInWorldFrame(bool use_input = false, Vec3<float> base_position = {0.f, 0.f, 0.f},
                                              Quat<float> base_orientation = {1.f, 0.f, 0.f, 0.f});
    Eigen::Matrix<float, 3, 3> ComputeJacobian(int leg_id);
    std::map<int, float> MapContactForceToJointTorques(int leg_id, Vec3<float> contact_force);
};

}   // namespace Quadruped
#endif
