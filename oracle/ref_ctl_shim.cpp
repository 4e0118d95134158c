// ref_ctl_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Pins the restatements around the MPC / WBC core against code compiled from the reference itself
// (oracle/Makefile target `refctl` -> oracle/_ref/libqr_ctl_ref.so).  Two techniques, neither of which edits or copies
// a reference source into this repository:
//
//  (1) whole translation units compiled from where they lie under /root/reference/quadruped, against oracle/mini_eigen:
//        src/controllers/qr_foot_trajectory_generator.cpp   qrFootParabolaPatternGenerator, SwingFootTrajectory,
//                                                           qrFootBSplinePatternGenerator (+ vendored tinynurbs)
//        src/utils/qr_geometry.cpp                          qrQuadraticSpline::getPoint ...
//        src/controllers/balance_controller/qr_qp_torque_optimizer.cpp   ComputeContactForce (+ vendored QuadProg++)
//        src/controllers/mpc/qr_mpc_interface.cpp           SolveMPCKernel / GetMPCSolution (+ vendored qpOASES)
//      (the last two see the plain-data stand-ins robots/qr_robot.h and estimators/qr_ground_surface_estimator.h of
//      oracle/mini_eigen/shim_ctl instead of the real headers, which need yaml-cpp / ROS / the vendor SDKs);
//
//  (2) member functions whose classes cannot be compiled here (they hang off the robot / estimator / ROS object
//      graph): the Makefile cuts the LINE RANGE of the function out of the reference source with sed into
//      oracle/_ref/gen/<name>.inc (git-ignored build output) and this file #includes it inside a stand-in class that
//      has the member names the body reads.  The ranges (file:lines) are listed next to every #include.
//
// The extern "C" functions at the bottom set those members from plain arrays, call the reference body, and copy
// the results out.
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "robots/qr_robot.h"   // oracle/mini_eigen/shim_ctl
#include "estimators/qr_ground_surface_estimator.h"
#include "controllers/qr_foot_trajectory_generator.h"
#include "controllers/balance_controller/qr_qp_torque_optimizer.h"
#include "controllers/mpc/qr_mpc_interface.h"

using std::map;
using Eigen::Matrix;

namespace Quadruped {

using robotics::math::clip;   // the reference sources see these through `using namespace robotics::math` in their headers
using robotics::math::CoordinateAxis;

// ---- src/robots/qr_robot.cpp:89-251: WithLegSigns, FootPositionInHipFrameToJointAngle, FootPositionInHipFrame,
//      AnalyticalLegJacobian, FootPositionsInBaseFrame, ComputeFootVelocitiesInBaseFrame,
//      ComputeMotorAnglesFromFootLocalPosition, ComputeMotorVelocityFromFootLocalVelocity,
//      GetFootPositionsInWorldFrame, ComputeJacobian, MapContactForceToJointTorques
#include "gen/robot_kin.inc"

// ---- gait generator: include/quadruped/gait/qr_gait.h members; bodies of
//      src/gait/qr_openloop_gait_generator.cpp:126-208 (Update) and :211-247 (Schedule)
class qrGaitGenerator {
public:
    qrRobot* robot = nullptr;
    std::string gait;
    float resetTime = 0, timeSinceReset = 0, lastTime = 0;
    Eigen::Matrix<float, 4, 1> stanceDuration, swingDuration, dutyFactor, phaseInFullCycle, initialLegPhase, offset;
    Eigen::Matrix<int, 4, 1> initialLegState, nextLegState, legState, desiredLegState, lastLegState, curLegState, detectedLegState;
    Eigen::Matrix<float, 4, 1> normalizedPhase, initStateRadioInCycle;
    Vec4<float> fullCyclePeriod;
    Vec4<bool> allowSwitchLegState;
    float contactDetectionPhaseThreshold = 0.1f;
    unsigned long gaitCycle = 0;
    Vec4<bool> firstSwing = {false, false, false, false};
    Vec4<float> contactStartPhase;
    Vec4<float> swingTimeRemaining = {0.f, 0.f, 0.f, 0.f};
    Vec4<bool> firstStance = {false, false, false, false};
    Eigen::Matrix<float, 12, 1> firstStanceAngles;
    virtual ~qrGaitGenerator() = default;
};
class qrOpenLoopGaitGenerator : public qrGaitGenerator {
public:
    float cumDt = 0, waitTime = 1.0;
    void Update(float currentTime);
    void Schedule(float currentTime);
};
#include "gen/gait_update.inc"

// ---- MPC stance-leg controller: members of include/quadruped/controllers/mpc/qr_mpc_stance_leg_controller.h
struct qrDesiredStateCommand {
    Vec12<float> stateDes;
    Eigen::Matrix<float, 3, 4> footTargetPositionsInWorldFrame;
};
struct qrUserParameters {
    float footClearance = 0.01f;
};
class MPCStanceLegController {
public:
    qrOpenLoopGaitGenerator* gaitGenerator = nullptr;
    int numHorizonL = 2, horizonLength = 10;
    float dtMPC = 0.03f, bodyHeight = 0.27f, yawTurnRate = 0, yawDesTrue = 0;
    Eigen::Matrix<float, Eigen::Dynamic, 4> mpcTable_storage;   // not used: see below
    Vec4<bool> contactState;
    Vec3<float> posDesiredinWorld, vDesWorld, rpyComp;
    float trajAll[12 * 36];
    Eigen::Matrix<float, 3, 4> f, f_ff;
    // mpcTable is `Eigen::Matrix<float, Dynamic, 4, RowMajor>` in the reference (.h:181); SolveMPCKernel gets its
    // .data().  mini_eigen matrices are column-major, so the table is this small row-major holder with the two
    // operations the cut-out lines use.
    struct RowMajorTable {
        std::vector<float> d;
        float& operator()(int i, int j) { return d[4 * i + j]; }
        float* data() { return d.data(); }
    } mpcTable;

    void FillTable() {
        // src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:282-303
#include "gen/mpc_table.inc"
    }
    void FillTrajectory(qrRobot* robot) {
        // src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:345-376 (body of UpdateMPC up to the SolveDenseMPC call)
#include "gen/mpc_traj.inc"
    }
    void SolveDenseMPC(qrRobot* robot);
};
// src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:385-410
#include "gen/mpc_solve_dense.inc"

// ---- foothold planner: members of include/quadruped/planner/qr_foothold_planner.h; body of
//      src/planner/qr_foothold_planner.cpp:112-240 (ComputeHeuristicFootHold)
class qrFootholdPlanner {
public:
    qrRobot* robot = nullptr;
    qrGaitGenerator* gaitGenerator = nullptr;
    qrUserParameters* userParameters = nullptr;
    qrDesiredStateCommand* desiredStateCommand = nullptr;
    Eigen::Matrix<float, 3, 4> desiredFootholds;
    Vec4<float> phase;
    Vec3<float> swingKp;
    Eigen::Matrix<int, 4, 1> moveDown;
    void ComputeHeuristicFootHold(std::vector<u8> swingFootIds);
};
#include "gen/foothold.inc"

// ---- swing-leg controller in MPC mode: src/controllers/qr_swing_leg_controller.cpp:361-409 (case ADVANCED_TROT of
//      GetAction) and :417-420 (joint targets), with the locals of :238-268 declared as the reference declares them
struct SwingTrot {
    qrRobot* robot = nullptr;
    qrOpenLoopGaitGenerator* gaitGenerator = nullptr;
    qrFootholdPlanner* footholdPlanner = nullptr;
    qrDesiredStateCommand* desiredStateCommand = nullptr;
    // the controller builds one trajectory object per leg from the configured spline type (XYLinear_ZParabola for the
    // trot gaits); the default constructor of SwingFootTrajectory leaves its generator pointer unset
    struct PerLeg {
        std::unique_ptr<SwingFootTrajectory> t[4];
        PerLeg() {
            qrSplineInfo info;
            info.splineType = SplineType::XYLinear_ZParabola;
            for (auto& p : t) p.reset(new SwingFootTrajectory(info));
        }
        SwingFootTrajectory& operator[](int i) { return *t[i]; }
    } swingFootTrajectories;
    Eigen::Matrix<float, 3, 4> phaseSwitchFootGlobalPos, desiredFootPositionsInBaseFrame, foot_pos_rel_last_time,
        foot_pos_target_last_time;
    bool horizontal_terrain = true;
    Vec3<float> jointAnglesOut[4], motorVelocityOut[4];

    void Run(const std::vector<u8>& swingFootIds) {
        auto& stateData = robot->stateDataFlow;
        Matrix<float, 3, 1> footTargetPosition;
        Matrix<float, 3, 1> footPositionInBaseFrame, footVelocityInBaseFrame, footAccInBaseFrame;
        Matrix<float, 3, 1> footPositionInWorldFrame, footVelocityInWorldFrame, footAccInWorldFrame;
        Matrix<int, 3, 1> jointIdx;
        Matrix<float, 3, 1> jointAngles;
        Eigen::Matrix<float, 3, 4> footVCurrent;
        Eigen::Matrix<float, 3, 4> footPositionsInBaseFrame = robot->GetFootPositionsInBaseFrame();
        Eigen::Matrix<float, 3, 4> footVelocitysInBaseFrame = robot->stateDataFlow.footVelocitiesInBaseFrame;
        Quat<float> robotComOrientation = robot->GetBaseOrientation();
        Mat3<float> robotBaseR = robot->stateDataFlow.baseRMat;
        float phase;
        if (horizontal_terrain) robotBaseR.setIdentity();   // :262-265 (groundEstimator->terrain.terrainType < 2)
        for (u8 legId : swingFootIds) {
            footVelocityInBaseFrame.setZero();
            footVelocityInWorldFrame.setZero();
            footAccInBaseFrame.setZero();
            footAccInWorldFrame.setZero();
            {
#include "gen/swing_trot.inc"
            }
#include "gen/swing_joint.inc"
            jointAnglesOut[legId] = jointAngles;
            motorVelocityOut[legId] = motorVelocity;
        }
    }
};

}   // namespace Quadruped

using namespace Quadruped;

namespace {
void set_robot_geometry(qrRobot& r, const float* geom /* hip_len upper lower, hipOffset[12] col-major, comOffset[3] */) {
    r.hipLength = geom[0];
    r.upperLegLength = geom[1];
    r.lowerLegLength = geom[2];
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) r.hipOffset(a, l) = geom[3 + 3 * l + a];
    for (int a = 0; a < 3; ++a) r.comOffset[a] = geom[15 + a];
}
}   // namespace

extern "C" {

// Leg kinematics of qrRobot (qr_robot.cpp:106-251) for one robot.
//   geom[18]   hip_len, upper_len, lower_len, hipOffset (3x4 column-major), comOffset
//   q, qd[12]  motor angles / velocities;  f_leg[12] contact force per leg (the f_ff of the MPC controller)
// Outputs (any may be NULL): foot_base[12] FootPositionsInBaseFrame (3x4 column-major), jac[36] the four analytic
// Jacobians row-major, foot_vel[12] ComputeFootVelocitiesInBaseFrame, tau[12] MapContactForceToJointTorques,
// ik_q[12] ComputeMotorAnglesFromFootLocalPosition of the FK result, ik_qd[12] ComputeMotorVelocityFromFootLocalVelocity.
int qr_ref_leg_kinematics(const float* geom, const float* q, const float* qd, const float* f_leg, float* foot_base,
                          float* jac, float* foot_vel, float* tau, float* ik_q, float* ik_qd) {
    qrRobot r;
    set_robot_geometry(r, geom);
    for (int i = 0; i < 12; ++i) {
        r.motorAngles[i] = q[i];
        r.motorVelocities[i] = qd ? qd[i] : 0.f;
    }
    Mat34<float> fp = r.FootPositionsInBaseFrame(r.motorAngles);
    r.stateDataFlow.footPositionsInBaseFrame = fp;
    for (int l = 0; l < 4; ++l) r.stateDataFlow.footJvs[l] = r.ComputeJacobian(l);   // UpdateDataFlow, qr_robot.cpp:62-64
    Mat34<float> fv = r.ComputeFootVelocitiesInBaseFrame();
    for (int l = 0; l < 4; ++l) {
        for (int a = 0; a < 3; ++a) {
            if (foot_base) foot_base[3 * l + a] = fp(a, l);
            if (foot_vel) foot_vel[3 * l + a] = fv(a, l);
            if (jac)
                for (int b = 0; b < 3; ++b) jac[9 * l + 3 * a + b] = r.stateDataFlow.footJvs[l](a, b);
        }
        if (tau && f_leg) {
            std::map<int, float> t = r.MapContactForceToJointTorques(l, Vec3<float>(f_leg[3 * l], f_leg[3 * l + 1], f_leg[3 * l + 2]));
            for (auto& kv : t) tau[kv.first] = kv.second;
        }
        if (ik_q) {
            Eigen::Matrix<int, 3, 1> idx;
            Vec3<float> ang;
            Vec3<float> local(fp(0, l), fp(1, l), fp(2, l));
            r.ComputeMotorAnglesFromFootLocalPosition(l, local, idx, ang);
            for (int a = 0; a < 3; ++a) ik_q[idx[a]] = ang[a];
            if (ik_qd) {
                Vec3<float> v = r.ComputeMotorVelocityFromFootLocalVelocity(l, ang, Vec3<float>(fv(0, l), fv(1, l), fv(2, l)));
                for (int a = 0; a < 3; ++a) ik_qd[3 * l + a] = v[a];
            }
        }
    }
    return 0;
}

// Contact table + reference trajectory of MPCStanceLegController (qr_mpc_stance_leg_controller.cpp:282-303, 345-376).
//   progress, duty [4]; leg_state [4] (LegState values); contacts [4] 0/1
//   init [12] = rpyComp0, rpyComp1, yawDesTrue, posDesiredinWorld x y, bodyHeight, -, -, yawTurnRate, vDesWorld x y, -
//   base_xy [2] actual base position
int qr_ref_mpc_inputs(int h, int num_horizon_l, float dt_mpc, const float* progress, const float* duty,
                      const int* leg_state, const int* contacts, const float* init, const float* base_xy,
                      float* table_out /*[4h] row-major*/, float* traj_out /*[12h]*/) {
    qrRobot robot;
    qrOpenLoopGaitGenerator gg;
    MPCStanceLegController c;
    c.gaitGenerator = &gg;
    c.numHorizonL = num_horizon_l;
    c.horizonLength = h;
    c.dtMPC = dt_mpc;
    c.mpcTable.d.assign(4 * h, 0.f);
    for (int j = 0; j < 4; ++j) {
        gg.phaseInFullCycle[j] = progress[j];
        gg.dutyFactor[j] = duty[j];
        gg.legState[j] = leg_state ? leg_state[j] : LegState::STANCE;
        c.contactState[j] = contacts ? contacts[j] != 0 : false;
    }
    c.FillTable();   // (row 0 is always overwritten with contactState, :301-303: callers pass the contact flags)
    for (int k = 0; k < 4 * h; ++k) table_out[k] = c.mpcTable.d[k];
    if (traj_out) {
        c.rpyComp = Vec3<float>(init[0], init[1], 0.f);
        c.yawDesTrue = init[2];
        c.posDesiredinWorld = Vec3<float>(init[3], init[4], 0.f);
        c.bodyHeight = init[5];
        c.yawTurnRate = init[8];
        c.vDesWorld = Vec3<float>(init[9], init[10], 0.f);
        robot.basePosition = Vec3<float>(base_xy[0], base_xy[1], 0.f);
        c.FillTrajectory(&robot);
        for (int k = 0; k < 12 * h; ++k) traj_out[k] = c.trajAll[k];
    }
    return 0;
}

// SolveDenseMPC (qr_mpc_stance_leg_controller.cpp:385-410): lever arms R (foot - comOffset), SolveMPCKernel, f, f_ff,
// Fr_des.  setup[5+3+12] = dt, mu, f_max, mass, alpha, inertia[3], weights[12].
//   lever_out [12] foot2ComInWorldFrame (3x4 column-major) -- recomputed by the same expression after the call
int qr_ref_solve_dense_mpc(int h, const double* setup, const float* geom, const float* rpy, const float* pos,
                           const float* quat, const float* v_world, const float* w_world, const float* foot_base,
                           const float* traj, const float* table, float* lever_out, float* f_out, float* f_ff_out,
                           float* fr_des_out) {
    qrRobot robot;
    set_robot_geometry(robot, geom);
    float inertia[3], weights[12];
    for (int i = 0; i < 3; ++i) inertia[i] = (float)setup[5 + i];
    for (int i = 0; i < 12; ++i) weights[i] = (float)setup[8 + i];
    SetupProblem(setup[0], h, setup[1], setup[2], setup[3], inertia, weights, (float)setup[4]);
    MPCStanceLegController c;
    c.horizonLength = h;
    c.mpcTable.d.assign(table, table + 4 * h);
    std::memcpy(c.trajAll, traj, sizeof(float) * 12 * h);
    for (int a = 0; a < 3; ++a) {
        robot.baseRollPitchYaw[a] = rpy[a];
        robot.basePosition[a] = pos[a];
        robot.stateDataFlow.baseVInWorldFrame[a] = v_world[a];
        robot.stateDataFlow.baseWInWorldFrame[a] = w_world[a];
    }
    for (int a = 0; a < 4; ++a) robot.baseOrientation[a] = quat[a];
    robot.stateDataFlow.baseRMat = robotics::math::quaternionToRotationMatrix(robot.baseOrientation).transpose();   // qr_robot.cpp:70
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) robot.stateDataFlow.footPositionsInBaseFrame(a, l) = foot_base[3 * l + a];
    c.SolveDenseMPC(&robot);
    Eigen::Matrix<float, 3, 4> lever =
        robot.stateDataFlow.baseRMat * (robot.GetFootPositionsInBaseFrame().colwise() - robot.comOffset);   // :396
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) {
            if (lever_out) lever_out[3 * l + a] = lever(a, l);
            if (f_out) f_out[3 * l + a] = c.f(a, l);
            if (f_ff_out) f_ff_out[3 * l + a] = c.f_ff(a, l);
            if (fr_des_out) fr_des_out[3 * l + a] = robot.stateDataFlow.wbcData.Fr_des[l][a];
        }
    return 0;
}

// SwingFootTrajectory in MPC mode (qr_foot_trajectory_generator.cpp:281-343 -> qrFootParabolaPatternGenerator :166-215
// -> qrQuadraticSpline, qr_geometry.cpp:127-190): ResetFootTrajectory(duration, start, end, height) +
// GenerateTrajectoryPoint(pos, vel, acc, t, phase_module).
int qr_ref_swing_parabola(const float* start, const float* end, float height, float t, int phase_module, float* pos,
                          float* vel, float* acc) {
    qrSplineInfo info;
    info.splineType = SplineType::XYLinear_ZParabola;
    SwingFootTrajectory traj(info);
    Vec3<float> s(start[0], start[1], start[2]), e(end[0], end[1], end[2]);
    traj.ResetFootTrajectory(1.f, s, e, height);
    Vec3<float> p, v, a;
    p.setZero(); v.setZero(); a.setZero();
    bool ok = traj.GenerateTrajectoryPoint(p, v, a, t, phase_module != 0);
    for (int i = 0; i < 3; ++i) {
        pos[i] = p[i];
        if (vel) vel[i] = v[i];
        if (acc) acc[i] = a[i];
    }
    return ok ? 1 : 0;
}

// qrFootholdPlanner::ComputeHeuristicFootHold (qr_foothold_planner.cpp:112-240) for one robot.
//   hip_pos [12] GetDefaultHipPosition (3x4 column-major); state arrays as in qr_gpu_foothold_heuristic_batch
int qr_ref_foothold(const float* geom, const float* hip_pos, const float* swing_kp, const float* com_vel,
                    const float* rpy_rate, const float* dR, const float* base_R, const float* rpy, const float* foot_base,
                    const float* q, const float* state_des /*[12]*/, float foot_clearance, const float* swing_remain,
                    const float* norm_phase, const int* allow_switch, const int* swing_mask, float* foothold_io,
                    float* phase_io) {
    qrRobot robot;
    set_robot_geometry(robot, geom);
    qrGaitGenerator gg;
    qrUserParameters up;
    up.footClearance = foot_clearance;
    qrDesiredStateCommand cmd;
    for (int i = 0; i < 12; ++i) cmd.stateDes[i] = state_des[i];
    qrFootholdPlanner fp;
    fp.robot = &robot;
    fp.gaitGenerator = &gg;
    fp.userParameters = &up;
    fp.desiredStateCommand = &cmd;
    robot.controlParams["mode"] = LocomotionMode::ADVANCED_TROT;
    for (int a = 0; a < 3; ++a) {
        fp.swingKp[a] = swing_kp[a];
        robot.baseVelocityInBaseFrame[a] = com_vel[a];
        robot.baseRollPitchYawRate[a] = rpy_rate[a];
        robot.baseRollPitchYaw[a] = rpy[a];
        for (int b = 0; b < 3; ++b) {
            robot.stateDataFlow.baseRInControlFrame(a, b) = dR[3 * a + b];
            robot.stateDataFlow.baseRMat(a, b) = base_R[3 * a + b];
        }
    }
    robot.stateDataFlow.groundRMat.setIdentity();
    robot.baseOrientation = Quat<float>(1.f, 0.f, 0.f, 0.f);
    robot.basePosition.setZero();
    std::vector<u8> ids;
    for (int l = 0; l < 4; ++l) {
        for (int a = 0; a < 3; ++a) {
            robot.defaultHipPosition(a, l) = hip_pos[3 * l + a];
            robot.stateDataFlow.footPositionsInBaseFrame(a, l) = foot_base[3 * l + a];
            fp.desiredFootholds(a, l) = foothold_io[3 * l + a];
            robot.motorAngles[3 * l + a] = q ? q[3 * l + a] : 0.f;
        }
        gg.swingTimeRemaining[l] = swing_remain[l];
        gg.normalizedPhase[l] = norm_phase[l];
        gg.allowSwitchLegState[l] = allow_switch[l] != 0;
        gg.stanceDuration[l] = 0.3f;
        fp.phase[l] = phase_io[l];
        if (swing_mask[l]) ids.push_back((u8)l);
    }
    fp.ComputeHeuristicFootHold(ids);
    for (int l = 0; l < 4; ++l) {
        for (int a = 0; a < 3; ++a) foothold_io[3 * l + a] = fp.desiredFootholds(a, l);
        phase_io[l] = fp.phase[l];
    }
    return 0;
}

// Swing-leg targets of the MPC mode (qr_swing_leg_controller.cpp:361-409, 417-420) for one robot: per swing leg the
// world-frame foot target handed to the WBC (pFoot_des / vFoot_des / aFoot_des of qrWbcCtrlData), the base-frame foot
// position and the joint targets from the leg inverse kinematics.
int qr_ref_swing_targets(const float* geom, const float* base_pos, const float* quat, const float* v_world,
                         const float* foothold /*[12] desiredFootholds*/, const float* planner_phase /*[4]*/,
                         const float* switch_pos /*[12] phaseSwitchFootGlobalPos*/, const float* swing_duration /*[4]*/,
                         const int* swing_mask, int horizontal_terrain, float* p_foot_des, float* v_foot_des,
                         float* a_foot_des, float* foot_base_des, float* q_des, float* qd_des) {
    qrRobot robot;
    set_robot_geometry(robot, geom);
    qrOpenLoopGaitGenerator gg;
    qrFootholdPlanner fp;
    qrDesiredStateCommand cmd;
    SwingTrot st;
    st.robot = &robot;
    st.gaitGenerator = &gg;
    st.footholdPlanner = &fp;
    st.desiredStateCommand = &cmd;
    st.horizontal_terrain = horizontal_terrain != 0;
    for (int a = 0; a < 3; ++a) {
        robot.basePosition[a] = base_pos[a];
        robot.stateDataFlow.baseVInWorldFrame[a] = v_world[a];
    }
    for (int a = 0; a < 4; ++a) robot.baseOrientation[a] = quat[a];
    robot.stateDataFlow.baseRMat = robotics::math::quaternionToRotationMatrix(robot.baseOrientation).transpose();
    std::vector<u8> ids;
    for (int l = 0; l < 4; ++l) {
        for (int a = 0; a < 3; ++a) {
            fp.desiredFootholds(a, l) = foothold[3 * l + a];
            st.phaseSwitchFootGlobalPos(a, l) = switch_pos[3 * l + a];
        }
        fp.phase[l] = planner_phase[l];
        gg.swingDuration[l] = swing_duration[l];
        if (swing_mask[l]) ids.push_back((u8)l);
    }
    st.Run(ids);
    for (int l = 0; l < 4; ++l) {
        if (!swing_mask[l]) continue;
        for (int a = 0; a < 3; ++a) {
            p_foot_des[3 * l + a] = robot.stateDataFlow.wbcData.pFoot_des[l][a];
            v_foot_des[3 * l + a] = robot.stateDataFlow.wbcData.vFoot_des[l][a];
            a_foot_des[3 * l + a] = robot.stateDataFlow.wbcData.aFoot_des[l][a];
            if (foot_base_des) foot_base_des[3 * l + a] = st.desiredFootPositionsInBaseFrame(a, l);
            if (q_des) q_des[3 * l + a] = st.jointAnglesOut[l][a];
            if (qd_des) qd_des[3 * l + a] = st.motorVelocityOut[l][a];
        }
    }
    return 0;
}

// One Update(currentTime) of qrOpenLoopGaitGenerator (qr_openloop_gait_generator.cpp:126-247) on caller-held state.
//   cfg [4][5]  per leg: initialLegPhase, fullCyclePeriod, initStateRadioInCycle, swingDuration, dutyFactor (unused here)
//   istate [5][4] in/out: curLegState, lastLegState, desiredLegState, legState, (firstSwing | firstStance << 1)
//   fstate [4] in/out: resetTime, lastTime, cumDt, waitTime
//   out [3][4] in/out: phaseInFullCycle, normalizedPhase, swingTimeRemaining;  allow [4] out: allowSwitchLegState
int qr_ref_gait_update(float current_time, const float* cfg, float contact_threshold, const int* contacts, int stop,
                       int advanced_trot, int* istate, float* fstate, float* out, int* allow) {
    qrRobot robot;
    qrOpenLoopGaitGenerator g;
    g.robot = &robot;
    g.gait = advanced_trot ? "advanced_trot" : "trot";
    robot.stop = stop != 0;
    g.contactDetectionPhaseThreshold = contact_threshold;
    for (int l = 0; l < 4; ++l) {
        g.initialLegPhase[l] = cfg[5 * l];
        g.fullCyclePeriod[l] = cfg[5 * l + 1];
        g.initStateRadioInCycle[l] = cfg[5 * l + 2];
        g.swingDuration[l] = cfg[5 * l + 3];
        g.dutyFactor[l] = cfg[5 * l + 4];
        robot.footContact[l] = contacts[l] != 0;
        g.curLegState[l] = istate[l];
        g.lastLegState[l] = istate[4 + l];
        g.desiredLegState[l] = istate[8 + l];
        g.legState[l] = istate[12 + l];
        g.firstSwing[l] = (istate[16 + l] & 1) != 0;
        g.firstStance[l] = (istate[16 + l] & 2) != 0;
        g.phaseInFullCycle[l] = out[l];
        g.normalizedPhase[l] = out[4 + l];
        g.swingTimeRemaining[l] = out[8 + l];
        g.contactStartPhase[l] = 0;
    }
    g.resetTime = fstate[0];
    g.lastTime = fstate[1];
    g.cumDt = fstate[2];
    g.waitTime = fstate[3];
    g.Update(current_time);
    for (int l = 0; l < 4; ++l) {
        istate[l] = g.curLegState[l];
        istate[4 + l] = g.lastLegState[l];
        istate[8 + l] = g.desiredLegState[l];
        istate[12 + l] = g.legState[l];
        istate[16 + l] = (g.firstSwing[l] ? 1 : 0) | (g.firstStance[l] ? 2 : 0);
        out[l] = g.phaseInFullCycle[l];
        out[4 + l] = g.normalizedPhase[l];
        out[8 + l] = g.swingTimeRemaining[l];
        if (allow) allow[l] = g.allowSwitchLegState[l] ? 1 : 0;
    }
    fstate[0] = g.resetTime;
    fstate[1] = g.lastTime;
    fstate[2] = g.cumDt;
    fstate[3] = g.waitTime;
    return 0;
}

// Quadruped::ComputeContactForce, world-frame overload (qr_qp_torque_optimizer.cpp:304-400), through the reference's
// own ComputeMassMatrix / ComputeObjectiveMatrix / ComputeWeightMatrix / ComputeConstraintMatrix + QuadProg++.
//   out [12] = the 3x4 result, column-major (leg columns)
int qr_ref_contact_force_world(float mass, const float* inertia9, const float* quat, const float* foot_base,
                               const float* desired_acc, const int* contacts, const float* n, const float* t1,
                               const float* t2, const float* acc_weight, const float* fmin_ratio,
                               const float* fmax_ratio, float reg_weight, float mu, float* out) {
    qrRobot robot;
    robot.totalMass = mass;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) robot.totalInertia(a, b) = inertia9[3 * a + b];
    for (int a = 0; a < 4; ++a) robot.baseOrientation[a] = quat[a];
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) robot.stateDataFlow.footPositionsInBaseFrame(a, l) = foot_base[3 * l + a];
    Eigen::Matrix<float, 6, 1> acc, w;
    for (int i = 0; i < 6; ++i) {
        acc[i] = desired_acc[i];
        w[i] = acc_weight[i];
    }
    Eigen::Matrix<bool, 4, 1> c;
    Vec4<float> fmin, fmax;
    for (int l = 0; l < 4; ++l) {
        c[l] = contacts[l] != 0;
        fmin[l] = fmin_ratio[l];
        fmax[l] = fmax_ratio[l];
    }
    Vec3<float> nn(n[0], n[1], n[2]), tt1(t1[0], t1[1], t1[2]), tt2(t2[0], t2[1], t2[2]);
    Eigen::Matrix<float, 3, 4> X = ComputeContactForce(&robot, acc, c, w, nn, tt1, tt2, fmin, fmax, reg_weight, mu);
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) out[3 * l + a] = X(a, l);
    return 0;
}

// Quadruped::ComputeContactForce, control-frame overload (qr_qp_torque_optimizer.cpp:190-301), through the reference's
// own helpers + QuadProg++.  terrain_type: TerrainType (PLANE 0 ... SLOPE 3); control_rpy [3], aligned [9] row-major =
// groundEstimator->GetControlFrameRPY() / GetAlignedDirections().
//   out [12] = the returned 3x4 matrix (X * Rcb)^T, column-major (leg columns)
//   derived [9 + 9 + 12 + 3 + 3] (optional): Rcb, Rcb I Rcb^T, (Rcb footPos)^T (4x3 row-major), g.head(3), surfaceNormal --
//   the quantities the reference derives at :203-225 before it calls its helpers, recomputed here with the same
//   expressions so that a test can hand them to the restatement / the kernel, whose inputs they are
int qr_ref_contact_force_control(float mass, const float* inertia9, const float* quat, const float* foot_base,
                                 const float* desired_acc, const int* contacts, int terrain_type, const float* control_rpy,
                                 const float* aligned9, const float* acc_weight, float fmin_ratio, float fmax_ratio,
                                 float reg_weight, float mu, float* out, float* derived) {
    qrRobot robot;
    qrGroundSurfaceEstimator ge;
    robot.totalMass = mass;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            robot.totalInertia(a, b) = inertia9[3 * a + b];
            ge.alignedDirections(a, b) = aligned9[3 * a + b];
        }
    for (int a = 0; a < 3; ++a) ge.controlFrameRPY[a] = control_rpy[a];
    ge.terrain.terrainType = (TerrainType)terrain_type;
    for (int a = 0; a < 4; ++a) robot.baseOrientation[a] = quat[a];
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) robot.stateDataFlow.footPositionsInBaseFrame(a, l) = foot_base[3 * l + a];
    Eigen::Matrix<float, 6, 1> acc, w;
    for (int i = 0; i < 6; ++i) {
        acc[i] = desired_acc[i];
        w[i] = acc_weight[i];
    }
    Eigen::Matrix<bool, 4, 1> c;
    for (int l = 0; l < 4; ++l) c[l] = contacts[l] != 0;
    Eigen::Matrix<float, 3, 4> F = ComputeContactForce(&robot, &ge, acc, c, w, reg_weight, mu, fmin_ratio, fmax_ratio);
    for (int l = 0; l < 4; ++l)
        for (int a = 0; a < 3; ++a) out[3 * l + a] = F(a, l);
    if (derived) {
        Mat3<float> Rcb;
        Vec3<float> g3(0.f, 0.f, 9.8f), normal(0.f, 0.f, 1.f);
        if (terrain_type == TerrainType::PLANE || terrain_type == TerrainType::PLUM_PILES) {
            Rcb = Mat3<float>::Identity();
        } else {   // :217-219
            Rcb = ge.alignedDirections.transpose() * robotics::math::quaternionToRotationMatrix(robot.baseOrientation).transpose();
            g3 = ge.alignedDirections.transpose() * g3;
            normal << -std::sin(control_rpy[1]), 0.f, std::cos(control_rpy[1]);
        }
        Mat3<float> inertia = Rcb * robot.totalInertia * Rcb.transpose();                                             // :224
        Eigen::Matrix<float, 4, 3> footPosition = (Rcb * robot.GetFootPositionsInBaseFrame()).transpose();           // :225
        float* o = derived;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) *o++ = Rcb(a, b);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) *o++ = inertia(a, b);
        for (int l = 0; l < 4; ++l) for (int a = 0; a < 3; ++a) *o++ = footPosition(l, a);
        for (int a = 0; a < 3; ++a) *o++ = g3[a];
        for (int a = 0; a < 3; ++a) *o++ = normal[a];
    }
    return 0;
}

}   // extern "C"
