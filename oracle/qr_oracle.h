/* qr_oracle.h -- C interface of the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product library (libqr_gpu.so) never links or calls anything declared here.
 *
 * The oracle is an Eigen-free float32 restatement of the reference's convex-MPC build
 *   quadruped/src/controllers/mpc/qr_mpc_interface.cpp:160-451
 *   quadruped/src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:282-303, 345-376, 396-409
 * that hands the QP to the reference's own vendored qpOASES 3.2.0 (oracle/_ref/libqpOASES.a, compiled
 * from /root/reference/quadruped/extern/qpOASES/src) exactly as SolveMPC does (:428-438).
 *
 * Pinning status: the reference repository holds NO test, golden vector or fixture for this path
 * (SURVEY.md section 4 / 8c).  The MPC restatement is pinned instead against the reference ITSELF run
 * here: its own qr_mpc_interface.cpp compiled unmodified against oracle/mini_eigen
 * (oracle/_ref/libqr_mpc_ref.so, ref_shim.cpp) -- H, g, U_b and the stock qpOASES solution agree bit
 * for bit when both sides evaluate the matrix exponential by its finite series, and to float32
 * rounding (3e-7) with Eigen's Pade evaluation; tests/test_oracle.py holds both checks.  The rounding
 * of a true Eigen build is not reproducible here (Eigen is not installed) and stays UNPINNED; all
 * float32 products in the oracle are plain sequential sums without FMA contraction.  The WBC
 * restatement (wbc_oracle.cpp) is pinned the same way: floating_base_model.cpp, qr_single_contact.cpp,
 * qr_multitask_projection.cpp, qr_wholebody_impulse_ctrl.cpp and task_set/ of the reference compiled
 * unmodified (oracle/_ref/libqr_wbc_ref.so, ref_wbc_shim.cpp) agree with it to float32 rounding
 * (tests/test_wbc.py).
 */
#ifndef QR_ORACLE_H
#define QR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors Quadruped::ProblemConfig (+ the body inertia SetupProblem stores in MPCRobotState),
 * include/quadruped/controllers/mpc/qr_mpc_interface.h:107-144.  The narrowing double->float of
 * dt / frictionCoeff / fMax / totalMass that SetupProblem performs (:163-168) is the caller's job. */
typedef struct {
    int   horizon;
    float dt;
    float mu;        /* frictionCoeff */
    float f_max;
    float mass;
    float inertia[3];
    float weights[12];
    float alpha;
} qro_mpc_params;

/* Float32 QP data exactly as SolveMPC builds them (qr_mpc_interface.cpp:359-425).
 *   p,v,w,rpy : 3 floats each;  quat = (w,x,y,z);  r_feet[3*leg+axis] (column-major 3x4, foot - CoM,
 *   world aligned);  traj[12*h];  gait[4*h] row-major h x 4.
 * Outputs (caller-allocated): H[n*n] row-major, g[n], ub[m]  with n = 12h, m = 20h.
 * Optional outputs (may be NULL): Aqp[13h*13], Bqp[13h*12h] row-major.                           */
int qro_mpc_build(const qro_mpc_params* P, const float* p, const float* v, const float* quat,
                  const float* w, const float* r_feet, const float* rpy, const float* traj,
                  const float* gait, float* H, float* g, float* ub, float* Aqp, float* Bqp);

/* The reference's qpOASES call (qr_mpc_interface.cpp:414-438) on float32 data widened to double.
 * nWSR = 100 reproduces the stock cap; a large value gives the converged oracle.
 *   x[n]            primal solution (getPrimalSolution)
 *   info[0]         qpOASES return value of init()
 *   info[1]         working-set recalculations used
 *   kkt[3]          stationarity / feasibility / complementarity (SolutionAnalysis::getKktViolation)
 *   cstat[m]        constraint status (-1 lower active, 0 inactive, +1 upper active); may be NULL   */
int qro_mpc_qpoases(int horizon, float mu, const float* H, const float* g, const float* ub,
                    int nWSR, double* x, int* info, double* kkt, int* cstat);

/* Same solver on double data supplied by the caller (used to study H symmetrisation etc.). */
int qro_qpoases_dense(int n, int m, const double* H, const double* g, const double* A,
                      const double* lbA, const double* ubA, int nWSR, double* x, int* info,
                      double* kkt, int* cstat);

/* SolveMPCKernel + GetMPCSolution in one call: build + qpOASES; x[12h] out. */
int qro_mpc_solve(const qro_mpc_params* P, const float* p, const float* v, const float* quat,
                  const float* w, const float* r_feet, const float* rpy, const float* traj,
                  const float* gait, int nWSR, double* x, int* info);

/* Contact table, qr_mpc_stance_leg_controller.cpp:282-303.  progress[4], duty[4] float32;
 * early_contact[4] / contacts[4] are 0/1 flags (contacts may be NULL: row 0 is then left as computed).
 * table[h*4] row-major float 0/1. */
void qro_mpc_contact_table(int horizon, int num_horizon_l, const float* progress, const float* duty,
                           const int* early_contact, const int* contacts, float* table);

/* Reference trajectory, qr_mpc_stance_leg_controller.cpp:345-376.
 * init[12] = {rpyComp0, rpyComp1, yawDes, xDes, yDes, bodyHeight, 0,0,yawRate, vxW, vyW, 0};
 * pos_xy[2] = actual base x,y used to clip the start to +-0.1 m.  traj[12*h] out. */
void qro_mpc_reference_traj(int horizon, float dt_mpc, const float* init, const float* pos_xy,
                            float* traj);

/* Post-processing, qr_mpc_stance_leg_controller.cpp:402-409: f_ff(leg) = -R_base^T f_world(leg).
 * R_base[9] row-major (base->world).  f_world[12] = solution entries 0..11; f_ff[12] out (3*leg+axis). */
void qro_mpc_grf_to_leg_force(const float* R_base, const double* x, float* f_world, float* f_ff);

/* GRF -> leg force in the base frame -> joint torques through the analytic leg Jacobian
 * (qr_mpc_stance_leg_controller.cpp:402-409, 139-141; src/robots/qr_robot.cpp:148-172, 241-251).
 * quat (w,x,y,z), q[12], f_world[12]; f_ff_out[12] (may be NULL), tau[12]. */
void qro_mpc_grf_to_torque(float hip_len, float upper_len, float lower_len, const float* quat, const float* q,
                           const float* f_world, float* f_ff_out, float* tau);

/* Wall-clock a batch of `count` independent SolveMPC calls (cold QProblem each) in this process.
 * Inputs are SoA-free: arrays of `count` consecutive problems.  Returns seconds; per-problem
 * seconds in lat[count] (may be NULL); number of RET_MAX_NWSR_REACHED in *capped. */
double qro_mpc_time_batch(const qro_mpc_params* P, int count, const float* p, const float* v,
                          const float* quat, const float* w, const float* r_feet, const float* rpy,
                          const float* traj, const float* gait, int nWSR, double* x_all,
                          double* lat, int* capped);

/* ------------------------------------------------------------------------------------------------
 * Whole-body control step (wbc_oracle.cpp): floating-base dynamics + tasks/contacts + kinematic WBC +
 * WBIC with the reference's QuadProg++.  One recomputing tick of qrWbcLocomotionController<float>::Run.
 *   model       link lengths / body box of the robot (A1: a1_sim.yaml, Lite3: lite3_sim/robot.yaml); the
 *               masses/inertias/locations hard-coded in BuildDynamicModel are inside the oracle
 *   state[37]   quat(w,x,y,z) pos(3) body twist(6: omega_body, v_body) q(12) qd(12)
 *   cmd[66]     pBody_des vBody_des aBody_des pBody_RPY_des vBody_Ori_des (3 each) pFoot_des[4] vFoot_des[4]
 *               aFoot_des[4] Fr_des[4] (12 each) prev vBody_Ori_des (3: the desiredVel the orientation task
 *               kept from the previous call, qr_task_body_orientation.cpp:68)
 *   contact[4]  contact_state
 * Outputs: tau[12] (all legs; the reference applies the stance ones), fr[12] optimal reaction forces
 * (zeros for swing legs), qdes[12] / qddes[12] from the kinematic WBC; dbg (may be NULL):
 * H(324) G(18) C(18) Jc of the 4 feet (216) Jcdqd(12) pGC(12) vGC(12) qddot(18).
 * _f32 computes in float like the reference, _f64 the same algorithm in double.  Returns 1 when
 * QuadProg++ reports infeasibility (the reference ignores it). */
typedef struct {
    float body_size[3];
    float hip_len, upper_len, lower_len;
} qro_wbc_model;
int qro_wbc_step_f32(const qro_wbc_model* model, const float* state, const float* cmd, const int* contact,
                     float* tau, float* fr, float* qdes, float* qddes, float* dbg);
int qro_wbc_step_f64(const qro_wbc_model* model, const float* state, const float* cmd, const int* contact,
                     double* tau, double* fr, double* qdes, double* qddes, double* dbg);

/* Swing-foot position in MPC mode (parabola generator); returns 0 when the phase is rejected. */
int qro_swing_parabola(const float* start, const float* end, float height, float t, int phase_module,
                       float* pos);

/* ------------------------------------------------------------------------------------------------
 * Force-balance stance QP (fb_oracle.cpp): ComputeContactForce of qr_qp_torque_optimizer.cpp with the reference's
 * QuadProg++.  force[12] = X(leg, axis) = -x.  Optional outputs: the float32 QP data G(144) a(12) C(24x12) lb(24)
 * and QuadProg++'s return value (inf when it met an infeasible row).  Returns 0, 1 (infeasible row met, iterate
 * kept -- what the reference does) or 3 (NaN: forces zeroed). */
typedef struct {
    float mass;
    float inertia[9];
    float acc_weight[6];
    float reg_weight;
    float mu;
    float fmin_ratio[4];
    float fmax_ratio[4];
    int world_frame;
} qro_fb_params;
int qro_force_balance(const qro_fb_params* P, const float* inertia, const float* foot, const float* acc,
                      const int* contact, const float* gravity, const float* frame, float* force,
                      float* G_out, float* a_out, float* C_out, float* lb_out, double* cost_out);

double qro_quadprog_ineq(int n, int m, const double* G, const double* g0, const double* C, const double* c0, double* x);

/* WALK-mode swing trajectory (cubic B-spline through the reference's vendored tinynurbs) and heuristic foothold
 * (swing_oracle.cpp).  qro_swing_bspline returns 0 when GenerateTrajectory rejects the time. */
int qro_swing_bspline(const float* initial_pos, const float* target_pos, float height, float duration,
                      float initial_time, float time, float* pos, float* vel);
typedef struct {
    float hip_offset[12];
    float hip_pos[12];
    float hip_len;
    float swing_kp[3];
} qro_foothold_params;
void qro_foothold(const qro_foothold_params* P, int legId, const float* com_vel, const float* w, const float* dR,
                  const float* base_R, const float* rpy, const float* foot_base, const float* des_speed,
                  float des_twist, float des_height, float swing_remain, int allow_switch, float norm_phase,
                  float* foothold, float* phase);

#ifdef __cplusplus
}
#endif
#endif
