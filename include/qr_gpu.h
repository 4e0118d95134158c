/* qr_gpu.h -- C ABI of the B200-native batched locomotion-control engine (libqr_gpu.so).
 *
 * Drop-in boundary for the convex-MPC hot path of TopHillRobotics/quadruped-robot.  Each entry point
 * names the reference interface it replaces (paths relative to /root/reference/quadruped/).
 * Plain pointers and sizes only; no C++/torch types.  All batched arrays are "problem rows":
 * element [i][k] of an array with row length K lives at base[i*K + k] (one robot instance per row),
 * which is what one CTA per problem reads with coalesced loads.
 *
 * Return value of every function: 0 on success, a negative QR_E* code on an API error (bad argument,
 * CUDA failure).  Numerical outcomes are reported per instance in status_out:
 *   0 converged and verified optimal (KKT conditions hold on the identified active set)
 *   1 interior-point iterate returned (active-set verification did not settle / iteration cap)
 *   2 infeasible or inconsistent bounds (f_max < 0)
 *   3 non-finite input or breakdown
 * (The reference swallows solver failures: qr_mpc_interface.cpp:436-442 ignores qpOASES' return code.)
 */
#ifndef QR_GPU_H
#define QR_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QR_OK 0
#define QR_EINVAL (-22)
#define QR_ECUDA (-5)
#define QR_ENOMEM (-12)

/* The reference's arrays stop at K_MAX_GAIT_SEGMENTS = 16 (controllers/mpc/qr_mpc_interface.h:33); the engine
 * goes to 32 so that the long-preview configuration of BASELINE.json (horizon 30, 360 variables) runs too. */
#define QR_MAX_HORIZON 32

/* Replaces Quadruped::ProblemConfig + the inertia/mass SetupProblem stores in MPCRobotState
 * (include/quadruped/controllers/mpc/qr_mpc_interface.h:107-144; SetupProblem :157,
 * src/controllers/mpc/qr_mpc_interface.cpp:160-175).  Values are the float32 the reference keeps. */
typedef struct {
    int32_t horizon;  /* 1..QR_MAX_HORIZON */
    float dt;         /* dtMPC */
    float mu;         /* frictionCoeff (0.45 in qr_mpc_stance_leg_controller.cpp:90) */
    float f_max;      /* totalMass * 9.81 */
    float mass;
    float inertia[3]; /* body-frame diagonal */
    float weights[12];/* rpy, xyz, omega, v */
    float alpha;      /* 4e-6 */
} qr_mpc_params;

#define QR_QP_NO_PREDICTION 1
#define QR_QP_SCALAR_FACTOR 2   /* long-horizon size classes: factorise every reduced system with the scalar 3x3-block
                                   LDL' instead of the tensor-core blocked Cholesky (csrc/chol8.h); diagnostics / tests */

/* Solver knobs; pass NULL for the defaults written next to each field. */
typedef struct {
    int32_t max_as_rounds;    /* 32    cold-start active-set rounds before the interior-point fallback */
    int32_t max_ipm_iter;     /* 40    interior-point iteration cap (fallback path)               */
    int32_t max_polish_rounds;/* 12    active-set verification / correction rounds after it       */
    int32_t flags;            /* 0     QR_QP_NO_PREDICTION: skip the coarse active-set prediction (diagnostics / tests);
                                       QR_QP_SCALAR_FACTOR: see above;
                                       occupies what used to be alignment padding: the layout is unchanged */
    double ipm_tol;           /* 1e-7  fallback: scaled stationarity and complementarity-gap tolerance */
    double act_kappa;         /* 1e3   constraint i is guessed active when s_i < kappa*lambda_i   */
    double feas_tol;          /* 1e-9  admissible constraint violation after the polish [N]      */
    double mult_tol;          /* 1e-11 admissible negative multiplier after the polish           */
} qr_qp_options;

/* Device selection / context creation.  qr_gpu_init(device) makes `device` the calling thread's current CUDA
 * device (device < 0 keeps the current one) and creates the engine's context on it; it may be called for several
 * devices of one process and again from every thread that wants to use a device.  qr_gpu_shutdown frees the
 * contexts of all devices.
 *
 * Threading / stream contract.  Every entry point works on the CALLING THREAD'S CURRENT DEVICE, which must have
 * been initialised (QR_ECUDA otherwise), and all device pointers passed to it must belong to that device.  The
 * device-pointer calls only enqueue work on the caller's stream and may be issued concurrently from several host
 * threads and on several streams: workspaces are pooled per device and keyed by stream (a launch sequence never
 * shares its scratch, counters or work lists with one that may run at the same time).  The *_host calls are
 * synchronous, take a private staging slot for their duration and are equally safe to call from several threads.
 * qr_gpu_last_error() returns the last error text of the calling thread. */
int qr_gpu_init(int device);
void qr_gpu_shutdown(void);
const char* qr_gpu_last_error(void);

/* Launch geometry of the fused kernel for instances with `stance_footsteps` non-swing entries in their
 * contact table (the workspace is sized by classes of 8 foot-steps): SM count, resident CTAs per SM,
 * threads per CTA, dynamic shared memory per CTA.  For bench / roofline reporting. */
int qr_gpu_mpc_occupancy(int horizon, int stance_footsteps, int* sm_count, int* ctas_per_sm,
                         int* threads_per_cta, int* smem_bytes);

/* qr_gpu_mpc_solve_batch -- replaces SolveMPCKernel + GetMPCSolution
 * (qr_mpc_interface.h:200,215; qr_mpc_interface.cpp:334-356, 359-443, 446-451) for `batch`
 * independent robot instances: state -> SRB model -> discretisation -> condensed QP -> solve.
 *
 *   p, v, w, rpy   [batch][3]   base position, world velocity, world angular velocity, roll-pitch-yaw
 *   quat           [batch][4]   (w,x,y,z)
 *   r_feet         [batch][12]  foot - CoM, world aligned, r_feet[3*leg+axis]  (Eigen 3x4 column-major)
 *   traj           [batch][12h] reference trajectory (state_trajectory)
 *   gait           [batch][4h]  contact table, row-major h x 4 floats (mpcTable.data())
 *   mu_i, fmax_i   [batch] or NULL  per-instance friction / force limit (NULL: P->mu, P->f_max)
 *   grf_out        [batch][12]  step-0 ground reaction forces, world frame (GetMPCSolution(0..11))
 *   u_out          [batch][12h] or NULL  full solution vector
 *   status_out     [batch] or NULL
 *   iters_out      [batch][2] or NULL    {interior-point iterations (0 unless the fallback ran),
 *                                         full-size active-set rounds incl. the verification rounds after a
 *                                         fallback; the rounds on the coarse prediction problem are not counted}
 * All pointers are DEVICE pointers; the call is asynchronous on `cuda_stream` (a cudaStream_t) and safe to issue
 * concurrently on other streams / from other threads (see the contract above). */
int qr_gpu_mpc_solve_batch(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                           const float* p, const float* v, const float* quat, const float* w,
                           const float* r_feet, const float* rpy, const float* traj,
                           const float* gait, const float* mu_i, const float* fmax_i,
                           float* grf_out, float* u_out, int32_t* status_out, int32_t* iters_out,
                           void* cuda_stream);

/* Same call with HOST buffers (what a reference-side caller holds): inputs are copied host->device,
 * the fused kernel runs, results are copied device->host, and the call returns after the stream has
 * drained.  Pageable or pinned memory both work; pinned avoids a staging copy. */
int qr_gpu_mpc_solve_batch_host(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                                const float* p, const float* v, const float* quat, const float* w,
                                const float* r_feet, const float* rpy, const float* traj,
                                const float* gait, const float* mu_i, const float* fmax_i,
                                float* grf_out, float* u_out, int32_t* status_out,
                                int32_t* iters_out);

/* Fused post-processing of the first 12 forces, done by the solve kernel itself while the instance's rows are still in
 * shared memory (qr_gpu_mpc_solve_batch_ex): what SolveDenseMPC and GetAction do after GetMPCSolution
 * (qr_mpc_stance_leg_controller.cpp:402-409, 139-141): f_ff = -R_base^T f (the force the leg exerts, base frame),
 * tau_leg = J_leg^T f_ff (qrRobot::MapContactForceToJointTorques, src/robots/qr_robot.cpp:241-251, analytic Jacobian
 * :148-172) and wbcData.Fr_des = f. */
typedef struct {
    float hip_len, upper_len, lower_len;
    const float* q;       /* [batch][12] motor angles (needed for f_ff_out / tau_out) */
    float* f_ff_out;      /* [batch][12] or NULL */
    float* tau_out;       /* [batch][12] or NULL */
    float* wbc_cmd_io;    /* [batch][66] qrWbcCtrlData rows of qr_gpu_wbc_solve_batch, or NULL: Fr_des (entries 51..62) is written */
} qr_mpc_epilogue;

/* qr_gpu_mpc_solve_batch with the epilogue above (epilogue may be NULL: identical to qr_gpu_mpc_solve_batch). */
int qr_gpu_mpc_solve_batch_ex(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                              const float* p, const float* v, const float* quat, const float* w,
                              const float* r_feet, const float* rpy, const float* traj,
                              const float* gait, const float* mu_i, const float* fmax_i,
                              float* grf_out, float* u_out, int32_t* status_out, int32_t* iters_out,
                              const qr_mpc_epilogue* epilogue, void* cuda_stream);

/* The same host-buffer call sharded over several GPUs of this process (BASELINE.json configs[2]: "batch 65536
 * sharded across 8 B200"; the reference's seam is the single SolveDenseMPC call,
 * qr_mpc_stance_leg_controller.cpp:385-410).  The batch is cut into contiguous shards [g*B/G, (g+1)*B/G), one per
 * entry of devices[]; one host thread per device uploads its shard, runs the fused kernel and downloads the
 * results straight into the caller's arrays (that device->host copy is the "final gather").  There is no
 * collective on the path.  Contexts of devices not yet initialised are created on the fly; the per-device host threads
 * are persistent (created on first use, joined by qr_gpu_shutdown) and the calling thread's current device is not
 * touched.  One multi-device call runs at a time (further callers wait).  Returns the first failing shard's code. */
int qr_gpu_mpc_solve_batch_host_multi(int n_devices, const int* devices, const qr_mpc_params* P,
                                      const qr_qp_options* opt, int batch, const float* p, const float* v,
                                      const float* quat, const float* w, const float* r_feet, const float* rpy,
                                      const float* traj, const float* gait, const float* mu_i, const float* fmax_i,
                                      float* grf_out, float* u_out, int32_t* status_out, int32_t* iters_out);

/* qr_gpu_mpc_condense_batch -- replaces ComputeContinuousTimeStateSpaceMatrices + ConvertToDiscreteQP
 * + the H/g/U_b build of SolveMPC (qr_mpc_interface.cpp:296-331, 257-293, 359-412): float32 QP data
 * with the reference's operation order.  H_out [batch][n*n] row-major, g_out [batch][n],
 * ub_out [batch][20h] (n = 12h).  Device pointers, asynchronous on the stream. */
int qr_gpu_mpc_condense_batch(const qr_mpc_params* P, int batch, const float* p, const float* v,
                              const float* quat, const float* w, const float* r_feet,
                              const float* rpy, const float* traj, const float* gait,
                              const float* fmax_i, float* H_out, float* g_out, float* ub_out,
                              void* cuda_stream);

/* qr_gpu_qp_solve_batch -- replaces the qpOASES call of SolveMPC (qr_mpc_interface.cpp:414-438) on
 * caller-supplied QP data: min 1/2 x'Hx + g'x  s.t. the friction-pyramid rows of ResizeQPMats
 * (:230-240) with 0 <= A x <= ub.  H [batch][n*n] float32 row-major (symmetrised as (H+H')/2),
 * g [batch][n], ub [batch][20h] (only entries 5k+4 are read), mu_i [batch] or NULL (then `mu`).
 * x_out [batch][n].  Device pointers, asynchronous on the stream. */
int qr_gpu_qp_solve_batch(int horizon, float mu, const qr_qp_options* opt, int batch,
                          const float* H, const float* g, const float* ub, const float* mu_i,
                          float* x_out, double* x_out_f64, int32_t* status_out, int32_t* iters_out,
                          void* cuda_stream);

/* qr_gpu_mpc_inputs_batch -- replaces the per-tick input preparation of MPCStanceLegController: the contact
 * table mpcTable (Run, qr_mpc_stance_leg_controller.cpp:282-303; bit-exact masks) and the reference trajectory
 * trajAll (UpdateMPC, :345-376), written straight into the `gait` / `traj` rows qr_gpu_mpc_solve_batch reads.
 *   progress, duty [batch][4]   gaitGenerator->phaseInFullCycle / dutyFactor
 *   early_contact  [batch][4] or NULL   legState == EARLY_CONTACT flags
 *   contacts       [batch][4] or NULL   measured contact flags overwriting row 0 (:301-303)
 *   traj_init      [batch][12]  {rollComp, pitchComp, yawDes, xDes, yDes, bodyHeight, 0, 0, yawRate, vxW, vyW, 0}
 *   pos_xy         [batch][2]   actual base x, y (the start is clipped to +-0.1 m of it, :347-356)
 *   num_horizon_l = max(2, int(fullCyclePeriod / 0.4)) (:50).  Either output may be NULL. */
int qr_gpu_mpc_inputs_batch(int horizon, int num_horizon_l, float dt_mpc, int batch, const float* progress,
                            const float* duty, const int32_t* early_contact, const int32_t* contacts,
                            const float* traj_init, const float* pos_xy, float* gait_out, float* traj_out,
                            void* cuda_stream);

/* qr_gpu_mpc_leg_torque_batch -- replaces the post-processing of the MPC forces: f_ff = -R_base^T f
 * (SolveDenseMPC, qr_mpc_stance_leg_controller.cpp:402-409) and tau_leg = J_leg^T f_ff (GetAction :139-141 ->
 * qrRobot::MapContactForceToJointTorques, src/robots/qr_robot.cpp:241-251, analytic Jacobian :148-172).
 *   quat [batch][4] (w,x,y,z), q [batch][12] motor angles, grf [batch][12] world-frame forces
 *   f_ff_out [batch][12] or NULL, tau_out [batch][12] */
int qr_gpu_mpc_leg_torque_batch(float hip_len, float upper_len, float lower_len, int batch, const float* quat,
                                const float* q, const float* grf, float* f_ff_out, float* tau_out, void* cuda_stream);

/* qr_gpu_mpc_lever_arms_batch -- the r_feet rows of qr_gpu_mpc_solve_batch from what the controller holds:
 * foot2ComInWorldFrame = baseRMat * (footPosInBaseFrame.colwise() - comOffset)  (SolveDenseMPC,
 * qr_mpc_stance_leg_controller.cpp:396; baseRMat = quaternionToRotationMatrix(q)^T, src/robots/qr_robot.cpp:70).
 *   quat [batch][4] (w,x,y,z), foot_base [batch][12] (3x4 column-major), com_offset [3] HOST values, r_feet_out [batch][12] */
int qr_gpu_mpc_lever_arms_batch(int batch, const float* quat, const float* foot_base, const float* com_offset,
                                float* r_feet_out, void* cuda_stream);

/* Leg geometry of qrRobot: hipLength, upperLegLength, lowerLegLength and hipOffset (3x4, column-major m[3*leg + axis]). */
typedef struct {
    float hip_len, upper_len, lower_len;
    float hip_offset[12];
} qr_leg_geometry;

/* qr_gpu_leg_kinematics_batch -- replaces qrRobot::FootPositionsInBaseFrame / FootPositionInHipFrame, ComputeJacobian /
 * AnalyticalLegJacobian and ComputeFootVelocitiesInBaseFrame (src/robots/qr_robot.cpp:125-197, 230-236) for `batch`
 * robots.  q, qd [batch][12]; foot_base_out [batch][12] (3x4 column-major), jac_out [batch][4][9] row-major,
 * foot_vel_out [batch][12]; any output may be NULL (qd may be NULL when foot_vel_out is). */
int qr_gpu_leg_kinematics_batch(const qr_leg_geometry* geom, int batch, const float* q, const float* qd,
                                float* foot_base_out, float* jac_out, float* foot_vel_out, void* cuda_stream);

/* qr_gpu_leg_ik_batch -- replaces qrRobot::ComputeMotorAnglesFromFootLocalPosition (FootPositionInHipFrameToJointAngle)
 * and ComputeMotorVelocityFromFootLocalVelocity (src/robots/qr_robot.cpp:106-122, 200-218).  foot_base [batch][12] foot
 * positions in the base frame, foot_vel [batch][12] or NULL, leg_mask [batch][4] or NULL (0: leave that leg's outputs
 * untouched); q_out [batch][12], qd_out [batch][12] or NULL. */
int qr_gpu_leg_ik_batch(const qr_leg_geometry* geom, int batch, const float* foot_base, const float* foot_vel,
                        const int32_t* leg_mask, float* q_out, float* qd_out, void* cuda_stream);

/* qr_gpu_swing_targets_batch -- replaces the MPC-mode (LocomotionMode::ADVANCED_TROT) branch of
 * qrRaibertSwingLegController::GetAction (src/controllers/qr_swing_leg_controller.cpp:361-409) and its joint targets
 * (:407-410) for the swing legs of `batch` robots: parabola from the lift-off position to the planned foothold
 * (SwingFootTrajectory, height 0.1), and the pFoot_des / vFoot_des / aFoot_des rows of qrWbcCtrlData written straight
 * into the WBC command rows qr_gpu_wbc_solve_batch reads.
 *   base_pos [batch][3], quat [batch][4], v_world [batch][3]  robot->basePosition, GetBaseOrientation(), baseVInWorldFrame
 *   foothold [batch][12]      footholdPlanner->desiredFootholds (output of qr_gpu_foothold_heuristic_batch)
 *   planner_phase [batch][4]  footholdPlanner->phase;  switch_pos [batch][12] phaseSwitchFootGlobalPos
 *   swing_duration [batch][4] gaitGenerator->swingDuration;  swing_mask [batch][4] int32
 *   horizontal_terrain        != 0: robotBaseR is the identity (:262-265)
 *   wbc_cmd_io [batch][66]    entries 15+3*leg.., 27+3*leg.., 39+3*leg.. of the swing legs are written
 *   foot_base_des_out, q_des_out, qd_des_out [batch][12] or NULL; valid_out [batch][4] or NULL (0: stance leg, or the
 *   trajectory generator rejected the phase and nothing was written) */
int qr_gpu_swing_targets_batch(const qr_leg_geometry* geom, int batch, const float* base_pos, const float* quat,
                               const float* v_world, const float* foothold, const float* planner_phase,
                               const float* switch_pos, const float* swing_duration, const int32_t* swing_mask,
                               int horizontal_terrain, float* wbc_cmd_io, float* foot_base_des_out, float* q_des_out,
                               float* qd_des_out, int32_t* valid_out, void* cuda_stream);

/* qr_gpu_gait_update_batch -- replaces qrOpenLoopGaitGenerator::Update + Schedule
 * (src/gait/qr_openloop_gait_generator.cpp:126-247) for `batch` robots, one call per control tick, on caller-held state:
 *   time [batch]         currentTime
 *   cfg [batch][20]      per leg: initialLegPhase, fullCyclePeriod, initStateRadioInCycle, swingDuration, dutyFactor
 *   contacts [batch][4]  robot->GetFootContact();  stop [batch] or NULL  robot->stop;  advanced_trot: gait == "advanced_trot"
 *   istate_io [batch][20] curLegState[4], lastLegState[4], desiredLegState[4], legState[4], firstSwing | firstStance << 1
 *                         (LegState values of config/qr_enum_types.h:62-68)
 *   fstate_io [batch][4]  resetTime, lastTime, cumDt, waitTime
 *   phase_full_io, norm_phase_io, swing_remain_io [batch][4]  phaseInFullCycle, normalizedPhase, swingTimeRemaining -- the
 *                         rows qr_gpu_mpc_inputs_batch (progress) and qr_gpu_foothold_heuristic_batch read
 *   allow_out [batch][4] or NULL  allowSwitchLegState
 *   early_out [batch][4] or NULL  legState == EARLY_CONTACT (the early_contact rows of qr_gpu_mpc_inputs_batch)
 *   swing_mask_out [batch][4] or NULL  the legs the swing-leg controller moves this tick (swingFootIds,
 *                         src/controllers/qr_swing_leg_controller.cpp:218-228)
 *   stance_mask_out [batch][4] or NULL  its complement: the contact_state rows of qr_gpu_wbc_solve_batch and the
 *                         `contacts` rows (row 0 of the contact table) of qr_gpu_mpc_inputs_batch */
int qr_gpu_gait_update_batch(int batch, const float* time, const float* cfg, float contact_threshold,
                             const int32_t* contacts, const int32_t* stop, int advanced_trot, int32_t* istate_io,
                             float* fstate_io, float* phase_full_io, float* norm_phase_io, float* swing_remain_io,
                             int32_t* allow_out, int32_t* early_out, int32_t* swing_mask_out, int32_t* stance_mask_out,
                             void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------
 * Whole-body control
 * ------------------------------------------------------------------------------------------------- */

/* Geometry the reference reads from the robot's yaml in BuildDynamicModel
 * (src/robots/qr_robot_a1_sim.cpp:176-186; config/<robot>/*.yaml: body_size, hip_l, upper_l, lower_l).
 * Link masses, inertias and joint locations are the constants hard-coded there. */
typedef struct {
    float body_size[3];
    float hip_len, upper_len, lower_len;
} qr_wbc_model;

/* qr_gpu_wbc_solve_batch -- replaces one recomputing tick of qrWbcLocomotionController<float>::Run
 * (src/controllers/wbc/qr_wbc_locomotion_controller.cpp:108-219): FloatingBaseModel::setState /
 * contactJacobians / massMatrix / generalizedGravityForce / generalizedCoriolisForce
 * (src/dynamics/floating_base_model.cpp:469-806), the task / contact updates (task_set/*.cpp,
 * qr_single_contact.cpp), qrMultitaskProjection::FindConfiguration (qr_multitask_projection.cpp:38-106)
 * and qrWholeBodyImpulseCtrl::GetModelRes / MakeTorque (qr_wholebody_impulse_ctrl.cpp:50-126) including
 * the QuadProg++ solve (:113), for `batch` independent robots.  Rows:
 *   state   [batch][37]  quat(w,x,y,z) pos(3) body twist (omega_body, v_body) q(12) qd(12)
 *                        (FBModelState as filled by UpdateModel :141-157)
 *   cmd     [batch][66]  qrWbcCtrlData (controllers/qr_state_dataflow.h:133-193): pBody_des vBody_des
 *                        aBody_des pBody_RPY_des vBody_Ori_des (3 each), pFoot_des[4] vFoot_des[4]
 *                        aFoot_des[4] Fr_des[4] (12 each), then the vBody_Ori_des of the PREVIOUS call (3):
 *                        the orientation task's velocity error uses it (qr_task_body_orientation.cpp:68)
 *   contact [batch][4]   contact_state (int32 0/1)
 *   tau_out [batch][12]  joint torques jointTorqueCmd (the reference applies those of stance legs, :205-219)
 *   fr_out  [batch][12] or NULL   optimal reaction forces (qrWBICExtraData::optimalFr scattered per leg)
 *   qdes_out, qddes_out [batch][12] or NULL   desiredJPos / desiredJVel of the kinematic WBC
 *   status_out [batch] or NULL   0 ok, 1 QP fallback iterate, 3 non-finite
 * The every-second-call gating of Run and the stance-only write-back stay with the caller.
 * Device pointers, asynchronous on the stream.  Arithmetic is float64 (the reference: float32 with a
 * float64 QP). */
int qr_gpu_wbc_solve_batch(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                           const int32_t* contact, float* tau_out, float* fr_out, float* qdes_out,
                           float* qddes_out, int32_t* status_out, void* cuda_stream);

/* Same call with HOST buffers (what the reference-side controller holds for its one robot): inputs are copied
 * host->device, the kernel runs, tau / fr / qdes / qddes / status come back, and the call returns after the
 * stream has drained.  Any output pointer except tau_out may be NULL. */
int qr_gpu_wbc_solve_batch_host(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                                const int32_t* contact, float* tau_out, float* fr_out, float* qdes_out,
                                float* qddes_out, int32_t* status_out);

/* Same with float64 outputs and the intermediate model quantities (tests): dbg [batch][630] =
 * H(324) G(18) Cqd(18) Jc of the four feet (216) Jcdqd(12) pFoot(12) vFoot(12) qddot(18); any may be NULL. */
int qr_gpu_wbc_solve_batch_f64(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                               const int32_t* contact, double* tau_out, double* fr_out, double* qdes_out,
                               double* qddes_out, double* dbg_out, int32_t* status_out, void* cuda_stream);

/* qr_gpu_swing_parabola_batch -- replaces SwingFootTrajectory::GenerateTrajectoryPoint in MPC mode
 * (src/controllers/qr_foot_trajectory_generator.cpp:322-343 -> qrFootParabolaPatternGenerator :188-215 ->
 * qrQuadraticSpline::getPoint, src/utils/qr_geometry.cpp:157-190): start,end [batch][3], height, phase
 * [batch], phase_module 0/1; pos_out [batch][3] (velocity / acceleration are identically zero in the
 * reference), valid_out [batch] (0 when the generator rejects the phase). */
int qr_gpu_swing_parabola_batch(int batch, const float* start, const float* end, const float* height,
                                const float* phase, int phase_module, float* pos_out, int32_t* valid_out,
                                void* cuda_stream);

/* qr_gpu_swing_bspline_batch -- replaces qrFootBSplinePatternGenerator (the WALK-mode swing trajectory):
 * SetParameters + UpdateSpline + GenerateTrajectory (src/controllers/qr_foot_trajectory_generator.cpp:53-163) with
 * tinynurbs::curveDerivatives (extern/tinynurbs/include/tinynurbs/core/evaluate.h:66-100, core/basis.h:25-66, 163-272)
 * for `batch` feet.  initial_pos, target_pos [batch][3]; height, duration, initial_time, time [batch];
 * pos_out, vel_out [batch][3] (metres, metres per unit of spline parameter); valid_out [batch] or NULL (0 where
 * GenerateTrajectory returns false and the outputs are left untouched). */
int qr_gpu_swing_bspline_batch(int batch, const float* initial_pos, const float* target_pos, const float* height,
                               const float* duration, const float* initial_time, const float* time, float* pos_out,
                               float* vel_out, int32_t* valid_out, void* cuda_stream);

/* Robot constants the foothold planner reads: qrRobot::hipOffset, GetDefaultHipPosition() (3x4, column-major:
 * m[3*leg + axis]), hipLength, and the planner's swingKp. */
typedef struct {
    float hip_offset[12];
    float hip_pos[12];
    float hip_len;
    float swing_kp[3];
} qr_foothold_params;

/* qr_gpu_foothold_heuristic_batch -- replaces qrFootholdPlanner::ComputeHeuristicFootHold
 * (src/planner/qr_foothold_planner.cpp:112-240) for `batch` robots.  Rows:
 *   com_vel [batch][3]  GetBaseVelocityInBaseFrame();  rpy_rate [batch][3]  GetBaseRollPitchYawRate()
 *   dR, base_R [batch][9] row-major  stateDataFlow.baseRInControlFrame, baseRMat;  rpy [batch][3]
 *   foot_base [batch][12]  GetFootPositionsInBaseFrame() (3x4 column-major)
 *   des_speed [batch][3] = stateDes.segment(6,3), des_twist [batch] = stateDes(11),
 *   des_height [batch] = stateDes(2) - footClearance
 *   swing_remain, norm_phase [batch][4]  gaitGenerator->swingTimeRemaining / normalizedPhase
 *   allow_switch, swing_mask [batch][4] int32  allowSwitchLegState; 1 for the legs in swingFootIds
 *   foothold_io [batch][12]  desiredFootholds (3x4 column-major, base frame): columns of swing legs are overwritten
 *   phase_io [batch][4]      planner phase of the swing legs
 * (The reference's footTargetPosition(0,2) / (0,3) writes at :219-222 index a 3-vector out of range and are not
 * mirrored.) */
int qr_gpu_foothold_heuristic_batch(const qr_foothold_params* P, int batch, const float* com_vel, const float* rpy_rate,
                                    const float* dR, const float* base_R, const float* rpy, const float* foot_base,
                                    const float* des_speed, const float* des_twist, const float* des_height,
                                    const float* swing_remain, const float* norm_phase, const int32_t* allow_switch,
                                    const int32_t* swing_mask, float* foothold_io, float* phase_io, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------
 * Force-balance stance controller
 * ------------------------------------------------------------------------------------------------- */

/* Parameters of Quadruped::ComputeContactForce
 * (include/quadruped/controllers/balance_controller/qr_qp_torque_optimizer.h:145-153, 172-183; defaults there:
 * regWeight 1e-4, frictionCoef 0.5 (control frame) / 0.45, fMinRatio 0.01, fMaxRatio 10). */
typedef struct {
    float mass;           /* robot->totalMass */
    float inertia[9];     /* row-major 3x3, used when the per-robot `inertia` array is NULL */
    float acc_weight[6];  /* accWeight */
    float reg_weight;     /* regWeight */
    float mu;             /* frictionCoef */
    float fmin_ratio[4];  /* per leg (the control-frame overload passes one scalar: repeat it) */
    float fmax_ratio[4];
    int32_t world_frame;  /* 0: arithmetic of the control-frame overload (qr_qp_torque_optimizer.cpp:192-301),
                             1: of the world-frame overload (:304-400); they differ in how lb is rounded */
} qr_fb_params;

/* qr_gpu_force_balance_batch -- replaces ComputeContactForce (qr_qp_torque_optimizer.cpp:192-400: ComputeMassMatrix,
 * ComputeObjectiveMatrix, ComputeWeightMatrix, ComputeConstraintMatrix and the QuadProg++ solve :273-276 / :380-383),
 * called by TorqueStanceLegController::GetAction (qr_torque_stance_leg_controller.cpp:496-500), for `batch` robots.
 *   inertia  [batch][9] or NULL   the 3x3 matrix the reference inverts (Rcb I Rcb' resp. rotMat I rotMat')
 *   foot     [batch][12]          footPositions.row(leg): foot positions in the frame of the computation
 *   acc      [batch][6]           desiredAcc / desiredDdq
 *   contact  [batch][4]           int32 0/1
 *   gravity  [batch][3] or NULL   g.head(3) when the control frame is tilted (NULL: 0, 0, 9.8)
 *   frame    [batch][9] or NULL   normal, tangent1, tangent2 (NULL: e_z, e_x, e_y -- the PLANE terrain)
 *   force_out [batch][12]         X(leg, axis) = -x: the matrix the reference then rotates back ((X*Rcb)' resp.
 *                                 RigidTransform); that frame change stays with the caller
 *   status_out [batch] or NULL    0 optimal; 1 the solver met an infeasible row and the iterate at that point is
 *                                 returned (what the reference does with QuadProg++'s result: a leg in swing asks for
 *                                 n.x >= 1e-7 and -n.x >= 1e-7 at once); 2 iteration cap; 3 not solvable, forces zeroed
 * Device pointers, asynchronous on the stream. */
int qr_gpu_force_balance_batch(const qr_fb_params* P, int batch, const float* inertia, const float* foot,
                               const float* acc, const int32_t* contact, const float* gravity, const float* frame,
                               float* force_out, int32_t* status_out, int32_t* iters_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* QR_GPU_H */
