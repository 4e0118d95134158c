// ref_wbc_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's OWN whole-body-control classes, compiled UNMODIFIED from where they lie under
// /root/reference/quadruped (oracle/Makefile target `refwbc`, against oracle/mini_eigen):
//   src/dynamics/floating_base_model.cpp                FloatingBaseModel<float>
//   src/controllers/wbc/qr_single_contact.cpp           qrSingleContact<float>
//   src/controllers/wbc/task_set/qr_task_*.cpp          body orientation / body position / link position
//   src/controllers/wbc/qr_multitask_projection.cpp     qrMultitaskProjection<float>
//   src/controllers/wbc/qr_wholebody_impulse_ctrl.cpp   qrWholeBodyImpulseCtrl<float> (+ QuadProg++)
// What cannot be compiled here is the glue around them -- qrWbcLocomotionController needs the robot /
// estimator / ROS classes -- so this file plays that part: it runs the reference's own BuildDynamicModel
// (cut out of src/robots/qr_robot_a1_sim.cpp:176-345, see below) and then issues the call sequence of
// qr_wbc_locomotion_controller.cpp:29-73 (construction, gains), :138-168 (UpdateModel),
// :172-201 (ContactTaskUpdate), :122-124 (FindConfiguration, MakeTorque).
// Same C signature as qro_wbc_step_f32 (qr_oracle.h) so tests can hold the two side by side.
#include "controllers/wbc/qr_multitask_projection.hpp"
#include "controllers/wbc/qr_wholebody_impulse_ctrl.hpp"
#include "controllers/wbc/task_set/qr_task_body_orientation.hpp"
#include "controllers/wbc/task_set/qr_task_body_position.hpp"
#include "controllers/wbc/task_set/qr_task_link_position.hpp"

#include "qr_oracle.h"

#include <chrono>
#include <cmath>
#include <memory>
#include <vector>

using robotics::math::coordinateRotation;
using robotics::math::CoordinateAxis;

// qrRobotA1Sim::BuildDynamicModel (src/robots/qr_robot_a1_sim.cpp:176-345; qr_robot_lite3_sim.cpp:176-345 is the same
// text) compiled from the reference: the Makefile cuts its body (lines 178-344, up to the dormant self-test) into
// _ref/gen/build_model.inc, and qrRobot::WithLegSigns (qr_robot.cpp:89-104) into _ref/gen/robot_signs.inc; the class
// below only supplies the members the body reads (the yaml node is a three-number stand-in).
namespace Quadruped {
struct CfgNode {
    std::vector<float> body_size;
    CfgNode operator[](const char*) const { return *this; }
    template <class T> T as() const { return body_size; }
};
class qrRobot {
public:
    CfgNode robotConfig;
    float hipLength = 0, upperLegLength = 0, lowerLegLength = 0;
    FloatingBaseModel<float> model;
    Vec3<float> WithLegSigns(const Vec3<float>& v, int leg_id);
    bool BuildDynamicModel();
};
#include "gen/robot_signs.inc"
bool qrRobot::BuildDynamicModel() {
#include "gen/build_model.inc"
    return true;
}
}   // namespace Quadruped

namespace {
void build_model(FloatingBaseModel<float>& model, const qro_wbc_model* cfg) {
    Quadruped::qrRobot r;
    r.robotConfig.body_size = {cfg->body_size[0], cfg->body_size[1], cfg->body_size[2]};
    r.hipLength = cfg->hip_len;
    r.upperLegLength = cfg->upper_len;
    r.lowerLegLength = cfg->lower_len;
    r.BuildDynamicModel();
    model = r.model;
}
}   // namespace

namespace {

// The controller objects qrWbcLocomotionController owns (construction + gains, qr_wbc_locomotion_controller.cpp:29-73),
// built once per robot model like the reference does at start-up; step() is one recomputing tick.
struct RefWbc {
    static constexpr size_t dimConfig = 18;
    const int FOOT[4] = {9, 11, 13, 15};   // Quadruped::linkID::FR, FL, HR, HL (config/qr_enum_types.h:35-45)
    FloatingBaseModel<float> fb;
    std::vector<qrTask<float>*> taskList;
    std::vector<qrSingleContact<float>*> contactList;
    qrMultitaskProjection<float> multitask;
    qrWholeBodyImpulseCtrl<float> wbic;
    qrWBICExtraData<float> extra;
    std::unique_ptr<qrTaskBodyOrientation<float>> taskOriP;
    std::unique_ptr<qrTaskBodyPosition<float>> taskPosP;
    std::unique_ptr<qrSingleContact<float>> footContact[4];
    std::unique_ptr<qrTaskLinkPosition<float>> taskFoot[4];

    explicit RefWbc(const qro_wbc_model* cfg) : multitask(dimConfig), wbic(dimConfig, &contactList, &taskList) {
        build_model(fb, cfg);
        extra.weightFb = DVec<float>::Constant(6, 0.1);
        extra.weightFr = DVec<float>::Constant(12, 1);
        taskOriP.reset(new qrTaskBodyOrientation<float>(&fb));
        taskPosP.reset(new qrTaskBodyPosition<float>(&fb));
        for (int l = 0; l < 4; ++l) {
            footContact[l].reset(new qrSingleContact<float>(&fb, FOOT[l]));
            taskFoot[l].reset(new qrTaskLinkPosition<float>(&fb, FOOT[l]));
        }
        for (int i = 0; i < 3; ++i) {
            taskPosP->Kp[i] = 100.;
            taskPosP->Kd[i] = 10.;
            taskOriP->Kp[i] = 100.;
            taskOriP->Kd[i] = 10.;
            for (int l = 0; l < 4; ++l) {
                taskFoot[l]->Kp[i] = 500;
                taskFoot[l]->Kd[i] = 10.;
            }
        }
    }

    int step(const float* state, const float* cmd, const int* contact, float* tau, float* fr, float* qdes, float* qddes,
             float* dbg) {
    qrTaskBodyOrientation<float>& taskOri = *taskOriP;
    qrTaskBodyPosition<float>& taskPos = *taskPosP;
    taskList.clear();      // qr_wbc_locomotion_controller.cpp:175-176 (ContactTaskUpdate starts from empty lists)
    contactList.clear();
    // UpdateModel, :138-168
    FBModelState<float> ms;
    ms.q = DVec<float>::Zero(12);
    ms.qd = DVec<float>::Zero(12);
    DVec<float> fullConfig(12 + 7);
    fullConfig.setZero();
    for (int i = 0; i < 4; ++i) ms.bodyOrientation[i] = state[i];
    for (int i = 0; i < 3; ++i) ms.bodyPosition[i] = state[4 + i];
    for (int i = 0; i < 6; ++i) ms.bodyVelocity[i] = state[7 + i];
    for (int i = 0; i < 12; ++i) {
        ms.q[i] = state[13 + i];
        ms.qd[i] = state[25 + i];
        fullConfig[i + 6] = ms.q[i];
    }
    fb.setState(ms);
    fb.contactJacobians();
    fb.massMatrix();
    fb.generalizedGravityForce();
    fb.generalizedCoriolisForce();
    wbic.GetModelRes(fb);

    // ContactTaskUpdate, :172-201.  The orientation task reads the desiredVel its PREVIOUS call stored
    // (qr_task_body_orientation.cpp:68); a first call seeds that state with cmd[63..65].
    const float *pBody_des = cmd, *vBody_des = cmd + 3, *aBody_des = cmd + 6, *rpy_des = cmd + 9, *vOri_des = cmd + 12,
                *pFoot = cmd + 15, *vFoot = cmd + 27, *aFoot = cmd + 39, *Fr_des = cmd + 51, *prevOriVel = cmd + 63;
    auto v3 = [](const float* p) { return Vec3<float>(p[0], p[1], p[2]); };
    Vec3<float> zeroVec3;
    zeroVec3.setZero();
    Vec3<float> rpyDes = v3(rpy_des);
    Quat<float> quatDes = robotics::math::rpyToQuat(rpyDes);
    taskOri.UpdateTask(&quatDes, v3(prevOriVel), zeroVec3);
    taskOri.UpdateTask(&quatDes, v3(vOri_des), zeroVec3);
    Vec3<float> pBody = v3(pBody_des);
    taskPos.UpdateTask(&pBody, v3(vBody_des), v3(aBody_des));
    taskList.push_back(&taskOri);
    taskList.push_back(&taskPos);
    Vec3<float> pFootDes[4];
    for (int leg = 0; leg < 4; ++leg) {
        if (contact[leg]) {
            footContact[leg]->SetDesiredFr((DVec<float>)(v3(Fr_des + 3 * leg)));
            footContact[leg]->UpdateContactSpec();
            contactList.push_back(footContact[leg].get());
        } else {
            pFootDes[leg] = v3(pFoot + 3 * leg);
            taskFoot[leg]->UpdateTask(&pFootDes[leg], v3(vFoot + 3 * leg), v3(aFoot + 3 * leg));
            taskList.push_back(taskFoot[leg].get());
        }
    }

    // Run, :122-124
    DVec<float> jointTorqueCmd(12), desiredJPos(12), desiredJVel(12);
    multitask.FindConfiguration(fullConfig, taskList, contactList, desiredJPos, desiredJVel);
    wbic.MakeTorque(jointTorqueCmd, &extra);

    for (int i = 0; i < 12; ++i) {
        tau[i] = jointTorqueCmd[i];
        qdes[i] = desiredJPos[i];
        qddes[i] = desiredJVel[i];
        fr[i] = 0;
    }
    int k = 0;
    for (int leg = 0; leg < 4; ++leg)
        if (contact[leg]) {
            for (int i = 0; i < 3; ++i) fr[3 * leg + i] = extra.optimalFr[3 * k + i];
            ++k;
        }
    if (dbg) {   // H(324, row-major) G(18) C(18) Jc feet(4*54) Jcdqd(12) pGC(12) vGC(12) -- qdd(18) not exposed
        float* o = dbg;
        const DMat<float>& H = fb.getMassMatrix();
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) *o++ = H(i, j);
        for (int i = 0; i < 18; ++i) *o++ = fb.getGravityForce()[i];
        for (int i = 0; i < 18; ++i) *o++ = fb.getCoriolisForce()[i];
        for (int l = 0; l < 4; ++l)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 18; ++j) *o++ = fb._Jc[FOOT[l]](i, j);
        for (int l = 0; l < 4; ++l)
            for (int i = 0; i < 3; ++i) *o++ = fb._Jcdqd[FOOT[l]][i];
        for (int l = 0; l < 4; ++l)
            for (int i = 0; i < 3; ++i) *o++ = fb._pGC[FOOT[l]][i];
        for (int l = 0; l < 4; ++l)
            for (int i = 0; i < 3; ++i) *o++ = fb._vGC[FOOT[l]][i];
        for (int i = 0; i < 18; ++i) *o++ = 0.f;
    }
    return 0;
    }   // step
};

}   // namespace

extern "C" int qr_ref_wbc_step(const qro_wbc_model* cfg, const float* state, const float* cmd, const int* contact,
                               float* tau, float* fr, float* qdes, float* qddes, float* dbg) {
    RefWbc ctl(cfg);
    return ctl.step(state, cmd, contact, tau, fr, qdes, qddes, dbg);
}

// Times `count` ticks (rows of state [37], cmd [66], contact [4]) on controller objects built once, like the running
// reference: per tick UpdateModel + ContactTaskUpdate + FindConfiguration + MakeTorque.  tau_out [count][12] or NULL,
// lat_out [count] seconds or NULL.  Returns the total seconds.

extern "C" double qr_ref_wbc_time_batch(const qro_wbc_model* cfg, int count, const float* state, const float* cmd,
                                        const int* contact, float* tau_out, double* lat_out) {
    RefWbc ctl(cfg);
    float tau[12], fr[12], qdes[12], qddes[12];
    double total = 0;
    for (int i = 0; i < count; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        ctl.step(state + 37 * (size_t)i, cmd + 66 * (size_t)i, contact + 4 * (size_t)i, tau, fr, qdes, qddes, nullptr);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        total += dt;
        if (lat_out) lat_out[i] = dt;
        if (tau_out)
            for (int k = 0; k < 12; ++k) tau_out[12 * (size_t)i + k] = tau[k];
    }
    return total;
}
