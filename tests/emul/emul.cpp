// emul.cpp -- HOST EMULATION of the device code in quadruped-robot_b200/csrc (TEST INFRASTRUCTURE ONLY).
//
// The kernels are written as barrier-separated phases (csrc/qr_team.h); compiled by g++ the phases run
// one after another with a loop over the thread index.  This lets CPU-only tests (`-m "not gpu"`)
// exercise the exact control flow and arithmetic of the CUDA path against the oracle.  It is not a
// product path: libqr_gpu.so never contains or calls this code, and nothing outside tests/ loads it.
#include <cstring>
#include <vector>

#include "../../quadruped-robot_b200/csrc/mpc_problem.h"
#include "../../quadruped-robot_b200/csrc/wbc_problem.h"
#include "../../quadruped-robot_b200/csrc/mpc_io.h"
#include "../../quadruped-robot_b200/csrc/fb_problem.h"
#include "../../quadruped-robot_b200/csrc/swing_extra.h"
#include "../../quadruped-robot_b200/csrc/ctl_extra.h"

static qr_qp_options emul_default_options() {
    qr_qp_options o;
    o.max_as_rounds = 32; o.max_ipm_iter = 40; o.max_polish_rounds = 12; o.flags = 0; o.ipm_tol = 1e-7; o.act_kappa = 1e3;
    o.feas_tol = 1e-9; o.mult_tol = 1e-11;
    return o;
}

// Workspace of one emulated team, sized like the CUDA launch of the same size class would be.
struct EmulTeam {
    std::vector<unsigned char> smem;
    std::vector<double> fallback, coarse;
    QrMpcSmem S;
    EmulTeam(int nfcap, int horizon) : smem(qr_mpc_smem_bytes(nfcap, horizon) + 64), fallback(qr_fallback_doubles(nfcap)),
                                       coarse(9 * qr_ntri(nfcap)) {
        qr_mpc_carve(S, smem.data(), nfcap, horizon, fallback.data());
        S.Hc = coarse.data();
        S.Hc2 = qr_coarse2_cap(nfcap) > 0 ? S.Hc + 9 * qr_ntri(qr_coarse_cap(nfcap)) : nullptr;
        qr_mpc_init_tables<128>(S, nfcap);
    }
};

static int class_cap_of(const float* gait, float fmax, int horizon) {
    int nf = 0;
    for (int k = 0; k < 4 * horizon; ++k) nf += (gait[k] * fmax > 0.f) ? 1 : 0;
    // the instantiated CUDA size classes: 8 .. 72 in steps of 8, then 96 and 128
    if (nf <= 72) { const int cap = ((nf + 7) / 8) * 8; return cap < 8 ? 8 : cap; }
    return nf <= 96 ? 96 : 128;
}

extern "C" int qr_emul_mpc_solve_batch(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                                       const float* p, const float* v, const float* quat, const float* w,
                                       const float* r_feet, const float* rpy, const float* traj,
                                       const float* gait, const float* mu_i, const float* fmax_i,
                                       float* grf_out, float* u_out, double* u_out_f64,
                                       int32_t* status_out, int32_t* iters_out) {
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P; A.opt = opt ? *opt : emul_default_options(); A.batch = batch;
    A.p = p; A.v = v; A.quat = quat; A.w = w; A.r_feet = r_feet; A.rpy = rpy; A.traj = traj; A.gait = gait;
    A.mu_i = mu_i; A.fmax_i = fmax_i;
    A.grf_out = grf_out; A.u_out = u_out; A.x_out_f64 = u_out_f64; A.status_out = status_out; A.iters_out = iters_out;
    for (int i = 0; i < batch; ++i) {
        // same size classification as qr_mpc_classify_kernel
        A.nfcap = class_cap_of(gait + (size_t)i * 4 * P->horizon, fmax_i ? fmax_i[i] : P->f_max, P->horizon);
        EmulTeam team(A.nfcap, P->horizon);
        qr_mpc_solve_problem<128, true>(A, i, team.S);
    }
    return 0;
}

extern "C" int qr_emul_mpc_condense_batch(const qr_mpc_params* P, int batch, const float* p, const float* v,
                                          const float* quat, const float* w, const float* r_feet,
                                          const float* rpy, const float* traj, const float* gait,
                                          const float* fmax_i, float* H_out, float* g_out, float* ub_out) {
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P; A.opt = emul_default_options(); A.batch = batch; A.nfcap = 4 * P->horizon;
    A.p = p; A.v = v; A.quat = quat; A.w = w; A.r_feet = r_feet; A.rpy = rpy; A.traj = traj; A.gait = gait;
    A.fmax_i = fmax_i; A.H_out = H_out; A.g_out = g_out; A.ub_out = ub_out;
    EmulTeam team(A.nfcap, P->horizon);
    for (int i = 0; i < batch; ++i) qr_mpc_condense_problem<128>(A, i, team.S);
    return 0;
}

// Number of (row, column) pairs of qH for which qr_condense_h_pair (the interleaved evaluation the fused path uses)
// differs in any bit from two calls of qr_condense_h_entry (the evaluation that is checked against the oracle).
extern "C" int qr_emul_condense_pair_mismatches(const qr_mpc_params* P, int batch, const float* p, const float* v,
                                                const float* quat, const float* w, const float* r_feet,
                                                const float* rpy, const float* traj, const float* gait) {
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P; A.opt = emul_default_options(); A.batch = batch; A.nfcap = 4 * P->horizon;
    A.p = p; A.v = v; A.quat = quat; A.w = w; A.r_feet = r_feet; A.rpy = rpy; A.traj = traj; A.gait = gait;
    EmulTeam team(A.nfcap, P->horizon);
    QrMpcSmem& S = team.S;
    const int h = P->horizon, n = 12 * h;
    int bad = 0;
    for (int i = 0; i < batch; ++i) {
        qr_mpc_stage<128>(A, i, S);
        QrCondenseTables& T = *S.T;
        qr_condense_model(A.P, S.state, S.state + 3, S.state + 6, S.state + 10, S.state + 13, S.state + 25, T);
        qr_condense_tables<128>(A.P, S.traj, T);
        for (int r = 0; r < n; ++r)
            for (int c = 0; c <= r; ++c) {
                float a, b;
                qr_condense_h_pair(T, h, r / 12, (r % 12) / 3, r % 3, c / 12, (c % 12) / 3, c % 3, &a, &b);
                const float a0 = qr_condense_h_entry(T, h, r / 12, (r % 12) / 3, r % 3, c / 12, (c % 12) / 3, c % 3);
                const float b0 = qr_condense_h_entry(T, h, c / 12, (c % 12) / 3, c % 3, r / 12, (r % 12) / 3, r % 3);
                bad += (memcmp(&a, &a0, 4) != 0) + (memcmp(&b, &b0, 4) != 0);
            }
    }
    return bad;
}

extern "C" int qr_emul_qp_solve_batch(int horizon, float mu, const qr_qp_options* opt, int batch,
                                      const float* H, const float* g, const float* ub, const float* mu_i,
                                      float* x_out, double* x_out_f64, int32_t* status_out, int32_t* iters_out) {
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P.horizon = horizon; A.P.mu = mu;
    A.opt = opt ? *opt : emul_default_options(); A.batch = batch; A.nfcap = 4 * horizon;
    A.mu_i = mu_i; A.H_in = H; A.g_in = g; A.ub_in = ub;
    A.x_out = x_out; A.x_out_f64 = x_out_f64; A.status_out = status_out; A.iters_out = iters_out;
    EmulTeam team(A.nfcap, horizon);
    for (int i = 0; i < batch; ++i) qr_qp_solve_problem<128>(A, i, team.S);
    return 0;
}

extern "C" int qr_emul_wbc_solve_batch(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                                       const int32_t* contact, double* tau, double* fr, double* qdes, double* qddes,
                                       double* dbg, int32_t* status) {
    QrWbcModelDev M;
    qr_wbc_host::build(model, &M);
    std::vector<unsigned char> smem(qr_wbc_smem_bytes() + 64);
    QrWbcWork W;
    qr_wbc_carve(W, smem.data());
    qr_wbc_init_tables<32>(W);
    QrWbcArgs A;
    memset(&A, 0, sizeof(A));
    A.model = &M; A.opt = emul_default_options(); A.batch = batch;
    A.state = state; A.cmd = cmd; A.contact = contact;
    A.tau64 = tau; A.fr64 = fr; A.qdes64 = qdes; A.qddes64 = qddes; A.dbg = dbg; A.status = status;
    for (int i = 0; i < batch; ++i) qr_wbc_problem<32>(A, i, W);
    return 0;
}

extern "C" int qr_emul_swing_parabola(const float* start, const float* end, float height, float t, int phase_module, float* pos) {
    return qr_swing_parabola(start, end, height, t, phase_module, pos);
}

extern "C" void qr_emul_mpc_contact_table(int h, int nhl, const float* progress, const float* duty, const int32_t* early,
                                          const int32_t* contacts, float* table) {
    qr_mpc_contact_table(h, nhl, progress, duty, early, contacts, table);
}
extern "C" void qr_emul_mpc_reference_traj(int h, float dt, const float* init, const float* pos_xy, float* traj) {
    qr_mpc_reference_traj(h, dt, init, pos_xy, traj);
}
extern "C" void qr_emul_mpc_grf_to_torque(float hip, float up, float low, const float* quat, const float* q, const float* f,
                                          float* ff, float* tau) {
    qr_mpc_grf_to_torque(hip, up, low, quat, q, f, ff, tau);
}

extern "C" int qr_emul_force_balance_batch(const qr_fb_params* P, int batch, const float* inertia, const float* foot,
                                           const float* acc, const int32_t* contact, const float* gravity,
                                           const float* frame, float* force_out, int32_t* status_out, int32_t* iters_out) {
    QrFbArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P; A.batch = batch;
    A.inertia = inertia; A.foot = foot; A.acc = acc; A.contact = contact; A.gravity = gravity; A.frame = frame;
    A.force_out = force_out; A.status_out = status_out; A.iters_out = iters_out;
    for (int i = 0; i < batch; ++i) qr_fb_problem(A, i);
    return 0;
}
// QP data exactly as handed to the solver (float32), for the bit-exactness test against the oracle
extern "C" void qr_emul_fb_build(const qr_fb_params* P, const float* inertia, const float* foot, const float* acc,
                                 const int32_t* contact, const float* gravity, const float* frame, float* G, float* a,
                                 float* C, float* lb) {
    qr_fb_build(*P, inertia ? inertia : P->inertia, foot, acc, contact, gravity, frame, G, a, C, lb);
}

extern "C" int qr_emul_swing_bspline(const float* ip, const float* tp, float height, float duration, float t0, float t,
                                     float* pos, float* vel) {
    return qr_swing_bspline(ip, tp, height, duration, t0, t, pos, vel);
}
extern "C" void qr_emul_foothold(const qr_foothold_params* P, int leg, const float* com_vel, const float* w, const float* dR,
                                 const float* base_R, const float* rpy, const float* foot_base, const float* des_speed,
                                 float des_twist, float des_height, float swing_remain, int allow_switch, float norm_phase,
                                 float* foothold, float* phase) {
    QrFootholdParams Q;
    memcpy(&Q, P, sizeof(Q));
    qr_foothold_heuristic(Q, leg, com_vel, w, dR, base_R, rpy, foot_base, des_speed, des_twist, des_height, swing_remain,
                          allow_switch, norm_phase, foothold, phase);
}

extern "C" int qr_emul_small_qp(int n, int m, const double* G, const double* g0, const double* C, const double* c0, double* x,
                                int* iters) {
    QrSmallQpWork W;
    return qr_small_qp_solve(n, m, G, g0, C, c0, x, W, iters);
}

// ---- csrc/ctl_extra.h: lever arms, leg kinematics, MPC-mode swing targets, gait phase (one robot per call)
static QrLegGeom emul_geom(const qr_leg_geometry* g) {
    QrLegGeom G;
    G.hip_len = g->hip_len; G.upper_len = g->upper_len; G.lower_len = g->lower_len;
    for (int k = 0; k < 12; ++k) G.hip_offset[k] = g->hip_offset[k];
    return G;
}
extern "C" void qr_emul_lever_arms(const float* quat, const float* foot_base, const float* com, float* r_feet) {
    qr_mpc_lever_arms(quat, foot_base, com, r_feet);
}
extern "C" void qr_emul_leg_kinematics(const qr_leg_geometry* g, const float* q, const float* qd, float* foot_base, float* jac,
                                       float* foot_vel, float* ik_q, float* ik_qd) {
    const QrLegGeom G = emul_geom(g);
    for (int leg = 0; leg < 4; ++leg) {
        qr_leg_fk(G, leg, q + 3 * leg, foot_base + 3 * leg);
        qr_leg_jacobian(G, leg, q + 3 * leg, jac + 9 * leg);
        qr_mat3_vec(jac + 9 * leg, qd + 3 * leg, foot_vel + 3 * leg);
        qr_leg_ik(G, leg, foot_base + 3 * leg, ik_q + 3 * leg);
        qr_leg_ik_velocity(G, leg, ik_q + 3 * leg, foot_vel + 3 * leg, ik_qd + 3 * leg);
    }
}
extern "C" void qr_emul_swing_targets(const qr_leg_geometry* g, const float* base_pos, const float* quat, const float* v_world,
                                      const float* foothold, const float* planner_phase, const float* switch_pos,
                                      const float* swing_duration, const int32_t* swing_mask, int horizontal, float* cmd,
                                      float* foot_base_des, float* q_des, float* qd_des, int32_t* valid) {
    const QrLegGeom G = emul_geom(g);
    for (int leg = 0; leg < 4; ++leg) {
        valid[leg] = 0;
        if (!swing_mask[leg]) continue;
        valid[leg] = qr_swing_targets_leg(G, leg, base_pos, quat, v_world, foothold + 3 * leg, planner_phase[leg], switch_pos + 3 * leg,
                                          swing_duration[leg], horizontal, cmd + 15 + 3 * leg, cmd + 27 + 3 * leg, cmd + 39 + 3 * leg,
                                          foot_base_des + 3 * leg, q_des + 3 * leg, qd_des + 3 * leg);
    }
}
extern "C" void qr_emul_gait_update(float t, const float* cfg, float thr, const int32_t* contacts, int stop, int advanced,
                                    int32_t* istate, float* fstate, float* out, int32_t* allow) {
    qr_gait_update(t, cfg, thr, contacts, stop, advanced, istate, fstate, out, allow);
}

// the fused solve with its post-processing epilogue (qr_gpu_mpc_solve_batch_ex)
extern "C" int qr_emul_mpc_solve_batch_ex(const qr_mpc_params* P, int batch, const float* p, const float* v, const float* quat,
                                          const float* w, const float* r_feet, const float* rpy, const float* traj,
                                          const float* gait, float hip, float upper, float lower, const float* q,
                                          float* grf_out, float* ff_out, float* tau_out, float* cmd_io, int32_t* status_out) {
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P; A.opt = emul_default_options(); A.batch = batch;
    A.p = p; A.v = v; A.quat = quat; A.w = w; A.r_feet = r_feet; A.rpy = rpy; A.traj = traj; A.gait = gait;
    A.grf_out = grf_out; A.status_out = status_out;
    A.ep_q = q; A.ep_ff = ff_out; A.ep_tau = tau_out; A.ep_cmd = cmd_io; A.ep_hip = hip; A.ep_upper = upper; A.ep_lower = lower;
    for (int i = 0; i < batch; ++i) {
        A.nfcap = class_cap_of(gait + (size_t)i * 4 * P->horizon, P->f_max, P->horizon);
        EmulTeam team(A.nfcap, P->horizon);
        qr_mpc_solve_problem<128, true>(A, i, team.S);
    }
    return 0;
}
