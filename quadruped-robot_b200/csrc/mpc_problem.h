// mpc_problem.h -- one MPC instance end to end on one thread team: stage inputs, condense (float32,
// reference operation order), eliminate swing foot-steps, solve (float64), scatter the forces.
// Shared by the CUDA kernels (mpc_kernels.cu) and by the host emulation used in CPU-only tests.
#pragma once

#include "mpc_condense.h"
#include "mpc_io.h"
#include "qp_solver.h"

// Arguments common to all kernels (passed by value).
struct QrMpcArgs {
    qr_mpc_params P;
    qr_qp_options opt;
    int batch;
    int nfcap;               // capacity (stance foot-steps) this launch's workspace is carved for
    const float *p, *v, *quat, *w, *r_feet, *rpy, *traj, *gait, *mu_i, *fmax_i;
    float* grf_out;          // [batch][12]
    float* u_out;            // [batch][12h] or null
    int32_t* status_out;     // [batch] or null
    int32_t* iters_out;      // [batch][2] or null
    double* scratch;         // [teams][qr_fallback_doubles(nfcap)] vectors of the interior-point fallback
    double* hs_global;       // [teams][9*ntri(nfcap)] Hessian blocks when they are kept out of shared memory, else null
    double* k_global;        // [teams][9*ntri(nfcap)] matrix under factorisation when it does not fit in shared memory, else null
    double* hc_global;       // [teams][9*ntri(nfcap)] Hessian of the coarse (move-blocked) problem that predicts the active set, or null
    int coarse_rounds;       // rounds spent on the coarse problem at most (0: the default QR_COARSE_MAX_ROUNDS)
    // work list of this launch (size class): problems list[0 .. *count), handed out through *next.
    // list == null: problems 0 .. batch-1.
    const int* list;
    const int* count;
    int* next;
    // condense-only outputs
    float *H_out, *g_out, *ub_out;
    // qp-only inputs / outputs
    const float *H_in, *g_in, *ub_in;
    float* x_out;
    double* x_out_f64;
    // fused post-processing of the step-0 forces (qr_gpu_mpc_solve_batch_ex): leg forces and joint torques of
    // MPCStanceLegController (f_ff = -R_base^T f, tau = J_leg^T f_ff) and Fr_des of the WBC command rows
    const float* ep_q;       // [batch][12] motor angles, or null: no epilogue
    float* ep_ff;            // [batch][12] or null
    float* ep_tau;           // [batch][12] or null
    float* ep_cmd;           // [batch][66] qrWbcCtrlData rows: Fr_des (entries 51..62) is written, or null
    float ep_hip, ep_upper, ep_lower;
};

QR_HD int qr_ntri(int nf) { return (nf * (nf + 1)) / 2; }
QR_HD size_t qr_fallback_doubles(int nfcap) { return (size_t)46 * nfcap + 8; }
// The coarse problem pairs consecutive foot-steps of a leg; it is only used when that removes at least a quarter
// of the foot-steps.
QR_HD int qr_coarse_cap(int nfcap) { return (3 * nfcap + 3) / 4; }
// A second level (pairs of pairs) in front of it pays only for long horizons, where a factorisation of the first
// level is itself expensive: workspaces above QR_COARSE2_MIN_CAP foot-steps carry its vectors.
#ifndef QR_COARSE2_MIN_CAP
#define QR_COARSE2_MIN_CAP 48
#endif
// round caps of the two levels of those classes (pairs / pairs of pairs)
#ifndef QR_COARSE_MAX_ROUNDS_2L_PAIR
#define QR_COARSE_MAX_ROUNDS_2L_PAIR 2
#endif
#ifndef QR_COARSE_MAX_ROUNDS_2L_QUAD
#define QR_COARSE_MAX_ROUNDS_2L_QUAD 2
#endif
QR_HD int qr_coarse2_cap(int nfcap) { return nfcap >= QR_COARSE2_MIN_CAP ? (3 * qr_coarse_cap(nfcap) + 3) / 4 : 0; }
QR_HD size_t qr_kbytes(int nfcap) {
    size_t kb = (size_t)9 * qr_ntri(nfcap) * sizeof(double);
    if (qr_k8_bytes(nfcap) > kb) kb = qr_k8_bytes(nfcap);   // long-horizon classes: room for the 8x8-tile layout of chol8.h
    return kb > sizeof(QrCondenseTables) ? kb : ((sizeof(QrCondenseTables) + 15) & ~(size_t)15);
}

// Shared-memory footprint in bytes for a workspace able to hold nfcap stance foot-steps.
QR_HD size_t qr_mpc_smem_bytes(int nfcap, int horizon, bool hs_in_smem = true, bool k_in_smem = true) {
    size_t bytes = hs_in_smem ? (size_t)9 * qr_ntri(nfcap) * sizeof(double) : 0;   // Hs
    bytes += k_in_smem ? qr_kbytes(nfcap)                          // K (aliased by the condense tables)
                       : ((sizeof(QrCondenseTables) + 15) & ~(size_t)15);   // the tables alone
    bytes += (size_t)(9 + 9 + 6 * 3 + 1) * nfcap * sizeof(double) + 8 * sizeof(double);
    bytes += (size_t)(3 + 1) * qr_coarse_cap(nfcap) * sizeof(double);   // coarse problem: g, ub
    bytes += (size_t)(3 + 1) * qr_coarse2_cap(nfcap) * sizeof(double);  // second coarse level (long horizons only): g, ub
    if (qr_coarse2_cap(nfcap) > 0) bytes += (size_t)(3 + 1 + 2) * qr_coarse_cap(nfcap) * sizeof(int);  // .. descriptors of the first level, grp2, gmem2
    bytes += (size_t)(16 * horizon + 32) * sizeof(float);          // staged traj + gait + state rows
    bytes += (size_t)(3 * nfcap + (nfcap + 1) + 3 * nfcap + 2 * 4 * horizon + 8 + 24 + nfcap + 3 * nfcap) * sizeof(int);   // .. + grp, gmem
    bytes += (size_t)((qr_ntri(nfcap) + 1) / 2) * sizeof(int);         // tri (unsigned short, padded to ints)
    return (bytes + 15) & ~(size_t)15;
}

struct QrMpcSmem {
    QrQpWork W;
    QrCondenseTables* T;  // aliases W.K (dead before the first factorisation)
    float* traj;   // [12h]
    float* gait;   // [4h]
    float* state;  // [32]: p v quat w r_feet rpy
    int* fs;       // [4h] stance list: fs[s] = foot-step id 4*i + leg
    int* slot;     // [4h] inverse map (or -1)
    int* misc;     // [8]
    double* scal;  // [8] scalars broadcast through shared memory (scal[0] = mu_)
    // coarse (move-blocked) problem used to predict the active set: see qr_mpc_predict_active_set
    double* Hc;    // [9*ntri(coarse cap)] global (L2-resident) or null: prediction disabled
    double* gc;    // [3*coarse cap]
    double* ubc;   // [coarse cap]
    int* grp;      // [nfcap] coarse foot-step of every stance foot-step
    int* gmem;     // [2*nfcap] members of every coarse foot-step (second = -1 for a single)
    // second coarse level (pairs of first-level foot-steps), only when qr_coarse2_cap(nfcap) > 0
    double* Hc2;   // global, behind Hc
    double* gc2;   // [3*coarse2 cap]
    double* ubc2;  // [coarse2 cap]
    int* cdesc;    // [3*coarse cap] leg, first step, last step of every first-level foot-step
    int* grp2;     // [coarse cap]
    int* gmem2;    // [2*coarse cap]
};

QR_DEV void qr_mpc_carve(QrMpcSmem& S, unsigned char* base, int nfcap, int horizon, double* fallback,
                         double* hs_global = nullptr, double* k_global = nullptr) {
    double* d = reinterpret_cast<double*>(base);
    QrQpWork& W = S.W;
    const int n = 3 * nfcap, m = 5 * nfcap;
    if (hs_global) W.Hs = hs_global;
    else { W.Hs = d; d += 9 * qr_ntri(nfcap); }
    W.k8 = (!k_global && qr_k8_bytes(nfcap) > 0) ? 1 : 0;
    if (k_global) {
        W.K = k_global;
        S.T = reinterpret_cast<QrCondenseTables*>(d);
        d = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d) + ((sizeof(QrCondenseTables) + 15) & ~(size_t)15));
    } else {
        W.K = d; d = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d) + qr_kbytes(nfcap));
        S.T = reinterpret_cast<QrCondenseTables*>(W.K);
    }
    W.Dinv = d; d += 9 * nfcap;
    W.zv = d; d += 9 * nfcap;
    W.ps = d; d += n; W.g = d; d += n; W.xn = d; d += n; W.q = d; d += n; W.wv = d; d += n;
    W.dx = d; d += n;
    W.ubz = d; d += nfcap;
    S.gc = d; d += 3 * qr_coarse_cap(nfcap);
    S.ubc = d; d += qr_coarse_cap(nfcap);
    S.gc2 = d; d += 3 * qr_coarse2_cap(nfcap);
    S.ubc2 = d; d += qr_coarse2_cap(nfcap);
    S.Hc = nullptr;
    S.Hc2 = nullptr;
    S.scal = d; d += 8;
    // fixed-size integer tables first, the horizon-dependent rows last: with a compile-time capacity every
    // pointer of the solver is then a constant offset from the shared-memory base
    int* ip = reinterpret_cast<int*>(d);
    W.act = ip; ip += nfcap;
    W.flag = ip; ip += nfcap;
    W.vert = ip; ip += nfcap;
    W.foff = ip; ip += nfcap + 1;
    W.rfoot = ip; ip += 3 * nfcap;
    S.misc = ip; ip += 8;
    W.hist = ip; ip += 24 + nfcap;
    S.grp = ip; ip += nfcap;
    S.gmem = ip; ip += 2 * nfcap;
    S.cdesc = S.grp2 = S.gmem2 = nullptr;
    if (qr_coarse2_cap(nfcap) > 0) {
        S.cdesc = ip; ip += 3 * qr_coarse_cap(nfcap);
        S.grp2 = ip; ip += qr_coarse_cap(nfcap);
        S.gmem2 = ip; ip += 2 * qr_coarse_cap(nfcap);
    }
    W.tri = reinterpret_cast<unsigned short*>(ip);
    ip += (qr_ntri(nfcap) + 1) / 2;
    float* f = reinterpret_cast<float*>(ip);
    S.state = f; f += 32;
    S.traj = f; f += 12 * horizon;
    S.gait = f; f += 4 * horizon;
    ip = reinterpret_cast<int*>(f);
    S.fs = ip; ip += 4 * horizon;
    S.slot = ip; ip += 4 * horizon;
    // interior-point fallback vectors (global)
    double* gsc = fallback;
    W.x = gsc; gsc += n; W.dxa = gsc; gsc += n; W.rd = gsc; gsc += n; W.yv = gsc; gsc += n;
    W.s = gsc; gsc += m; W.lam = gsc; gsc += m; W.dsa = gsc; gsc += m; W.dla = gsc; gsc += m;
    W.rc = gsc; gsc += m; W.dl = gsc; gsc += m;
    W.red = gsc;
}

// Once per launch: the triangular index decode table.
template <int NT>
QR_DEV void qr_mpc_init_tables(QrMpcSmem& S, int nfcap) {
    QR_FOR(idx, qr_ntri(nfcap)) {
        int I, J;
        qr_tri_decode(idx, I, J);
        S.W.tri[idx] = (unsigned short)((I << 8) | J);
    }
    QR_SYNC();
}

// Stage this problem's rows and build the stance list.  After return (and its trailing barrier):
// S.state/traj/gait hold the inputs, S.fs/slot the stance map, W.nf, W.ubz, W.mu_ are set, and
// S.misc[0] = per-instance status so far (0 ok, 2 negative bound, 3 non-finite input / over capacity).
template <int NT>
QR_DEV void qr_mpc_stage(const QrMpcArgs& A, int prob, QrMpcSmem& S) {
    const int h = A.P.horizon;
    QR_FOR(i, 12 * h) S.traj[i] = A.traj[(size_t)prob * 12 * h + i];
    QR_FOR(i, 4 * h) S.gait[i] = A.gait[(size_t)prob * 4 * h + i];
    QR_FOR(i, 28) {
        float val;
        if (i < 3) val = A.p[(size_t)prob * 3 + i];
        else if (i < 6) val = A.v[(size_t)prob * 3 + i - 3];
        else if (i < 10) val = A.quat[(size_t)prob * 4 + i - 6];
        else if (i < 13) val = A.w[(size_t)prob * 3 + i - 10];
        else if (i < 25) val = A.r_feet[(size_t)prob * 12 + i - 13];
        else val = A.rpy[(size_t)prob * 3 + i - 25];
        S.state[i] = val;
    }
    QR_SYNC();
    QR_THREADS(t) {
        if (t == 0) {
            const float fmax = A.fmax_i ? A.fmax_i[prob] : A.P.f_max;
            const float mu = A.mu_i ? A.mu_i[prob] : A.P.mu;
            S.scal[0] = (double)QR_FDIV(1.f, mu);        // mu_ = 1.f / frictionCoeff (qr_mpc_interface.cpp:230)
            int nf = 0, st = 0;
            QR_UNROLL_SMALL
            for (int k = 0; k < 4 * h; ++k) {
                const float ub = QR_FMUL(S.gait[k], fmax);   // U_b(5k+4) = gait * fMax (:387)
                if (ub > 0.f && nf < A.nfcap) {
                    S.fs[nf] = k; S.slot[k] = nf; S.W.ubz[nf] = (double)ub; ++nf;
                } else {
                    S.slot[k] = -1;
                    if (ub < 0.f) st = 2;
                    if (!(ub == ub)) st = 3;
                    if (ub > 0.f) st = 3;  // over capacity: the size classification makes this unreachable
                }
            }
            QR_UNROLL_SMALL
            for (int i = 0; i < 28; ++i) if (!(fabsf(S.state[i]) < 3.0e38f)) st = 3;
            if (!(mu > 0.f)) st = 3;
            S.misc[0] = st;
            S.misc[1] = nf;
        }
    }
    QR_SYNC();
    S.W.nf = S.misc[1];
    S.W.mu_ = S.scal[0];
}

// Condense into the solver workspace: symmetric Hessian blocks of the stance variables -> W.Hs,
// g -> W.g.  H_sym = (H + H')/2 evaluated in float64 from the two float32 entries.
template <int NT>
QR_DEV void qr_mpc_condense_to_work(const QrMpcArgs& A, QrMpcSmem& S) {
    const int h = A.P.horizon;
    QrCondenseTables& T = *S.T;
    QR_THREADS(t) {
        if (t == 0) qr_condense_model(A.P, S.state, S.state + 3, S.state + 6, S.state + 10, S.state + 13, S.state + 25, T);
    }
    QR_SYNC();
    qr_condense_tables<NT>(A.P, S.traj, T);
    QR_SYNC();
    const int nf = S.W.nf;
    QR_FOR(idx, 9 * qr_ntri(nf)) {
        const int b = idx / 9, e = idx - 9 * b;
        const int code = S.W.tri[b];
        const int ks = S.fs[code >> 8], kt = S.fs[code & 255];
        const int r = e / 3, c = e - 3 * r;
        float hst, hts;
        qr_condense_h_pair(T, h, ks >> 2, ks & 3, r, kt >> 2, kt & 3, c, &hst, &hts);
        S.W.Hs[idx] = 0.5 * ((double)hst + (double)hts);
    }
    QR_FOR(i, 3 * nf) {
        const int k = S.fs[i / 3];
        S.W.g[i] = (double)qr_condense_g_entry(T, h, k >> 2, k & 3, i % 3);
    }
    QR_SYNC();
}


// Active-set prediction on a coarse problem.
//
// The block active-set iteration spends two thirds of its factorisation work in the first three rounds (72, 61, 52
// free variables for the A1 trot at h = 10) while the set of active rows spreads along the horizon like a wave, one
// or two time steps per round.  Here the wave is run on a QUARTER of the work first: consecutive foot-steps of the
// same leg (same stance phase) are tied pairwise to one force ("move blocking": x = T z, H_c = T'HT, g_c = T'g, same
// pyramid and cap per pair), that half-size QP is solved by the same iteration, and every foot-step starts the
// full-size iteration from its pair's active rows (qr_qp_solve, stage 0).  The full-size iteration then needs 4-5
// rounds instead of 7-8, all of them on the small final systems, and still ends only on verified KKT conditions --
// the prediction changes the starting guess, never the result.  This routine builds the coarse problem.
// H_c = T'HT by 3x3 blocks (block-packed like W.Hs, diagonal blocks in full), g_c = T'g, ub_c for the ng tied foot-steps
// whose members gmem lists.  The four loads of an entry are issued together (a missing second member reads the first
// one's entry with weight 0): the fine Hessian lives in L2, so the latency of a dependent or branchy load sequence
// would dominate this phase.
template <int NT>
QR_DEV void qr_coarse_reduce(const unsigned short* tri, const double* __restrict__ Hs, const double* g, const double* ubz,
                             int ng, const int* gmem, double* __restrict__ Hc, double* gc, double* ubc) {
    QR_FOR(idx, 9 * qr_ntri(ng)) {
        const int b = idx / 9, e = idx - 9 * b;
        const int code = tri[b];
        const int Ac = code >> 8, Bc = code & 255;
        const int r = e / 3, c = e - 3 * r;
        const int f0 = gmem[2 * Ac], f1r = gmem[2 * Ac + 1], g0 = gmem[2 * Bc], g1r = gmem[2 * Bc + 1];
        const int f1 = f1r < 0 ? f0 : f1r, g1 = g1r < 0 ? g0 : g1r;
        const double wf = f1r < 0 ? 0.0 : 1.0, wg = g1r < 0 ? 0.0 : 1.0;
        const int erc = 3 * r + c, ecr = 3 * c + r;
        const double v00 = f0 >= g0 ? Hs[qr_blk(f0, g0) + erc] : Hs[qr_blk(g0, f0) + ecr];
        const double v01 = f0 >= g1 ? Hs[qr_blk(f0, g1) + erc] : Hs[qr_blk(g1, f0) + ecr];
        const double v10 = f1 >= g0 ? Hs[qr_blk(f1, g0) + erc] : Hs[qr_blk(g0, f1) + ecr];
        const double v11 = f1 >= g1 ? Hs[qr_blk(f1, g1) + erc] : Hs[qr_blk(g1, f1) + ecr];
        Hc[idx] = (v00 + wg * v01) + wf * (v10 + wg * v11);
    }
    QR_FOR(i, 3 * ng) {
        const int gidx = i / 3, a = i - 3 * gidx;
        const int f = gmem[2 * gidx], f2 = gmem[2 * gidx + 1];
        gc[i] = g[3 * f + a] + (f2 >= 0 ? g[3 * f2 + a] : 0.0);
        if (a == 0) ubc[gidx] = ubz[f];
    }
    QR_SYNC();
}

template <int NT, bool L2>
QR_DEV void qr_mpc_build_coarse(QrMpcSmem& S, const qr_qp_options& opt, int max_rounds, QrCoarse* C) {
    QrQpWork& W = S.W;
    const int nf = W.nf;
    C[0].ng = 0; C[0].Hs = S.Hc; C[0].g = S.gc; C[0].ubz = S.ubc; C[0].grp = S.grp;
    C[0].max_rounds = max_rounds > 0 ? max_rounds : (L2 ? QR_COARSE_MAX_ROUNDS_2L_PAIR : QR_COARSE_MAX_ROUNDS);
    if (L2) {
        C[1].ng = 0; C[1].Hs = S.Hc2; C[1].g = S.gc2; C[1].ubz = S.ubc2; C[1].grp = S.grp2;
        C[1].max_rounds = max_rounds > 0 ? max_rounds : QR_COARSE_MAX_ROUNDS_2L_QUAD;
    }
    if (!S.Hc || nf < 8 || (opt.flags & QR_QP_NO_PREDICTION)) return;
    const bool want2 = L2 && S.Hc2 != nullptr && nf >= QR_COARSE2_MIN_CAP - 7;
    // Pairing: consecutive foot-steps of one leg with the same force cap are tied to one force; a foot-step that finds
    // no partner stays single.  The second level pairs the first level's foot-steps by the same rule (a first-level
    // foot-step covers the steps first..last of its leg).
    QR_THREADS(t) {
        if (t == 0) {
            int open[4] = {-1, -1, -1, -1}, last[4] = {-9, -9, -9, -9};
            int ng = 0;
            for (int s = 0; s < nf; ++s) {
                const int k = S.fs[s], step = k >> 2, leg = k & 3;
                if (open[leg] >= 0 && last[leg] == step - 1 && W.ubz[S.gmem[2 * open[leg]]] == W.ubz[s]) {
                    S.grp[s] = open[leg];
                    S.gmem[2 * open[leg] + 1] = s;
                    if (want2) S.cdesc[3 * open[leg] + 2] = step;
                    open[leg] = -1;
                } else {
                    S.grp[s] = ng;
                    S.gmem[2 * ng] = s;
                    S.gmem[2 * ng + 1] = -1;
                    if (want2) { S.cdesc[3 * ng] = leg; S.cdesc[3 * ng + 1] = step; S.cdesc[3 * ng + 2] = step; }
                    open[leg] = ng;
                    ++ng;
                }
                last[leg] = step;
            }
            S.misc[2] = ng;
            int ng2 = 0;
            if (want2 && 4 * ng <= 3 * nf) {
                for (int l = 0; l < 4; ++l) { open[l] = -1; last[l] = -9; }
                for (int s = 0; s < ng; ++s) {
                    const int leg = S.cdesc[3 * s], first = S.cdesc[3 * s + 1], lst = S.cdesc[3 * s + 2];
                    const double ub = W.ubz[S.gmem[2 * s]];
                    if (open[leg] >= 0 && last[leg] == first - 1 && W.ubz[S.gmem[2 * S.gmem2[2 * open[leg]]]] == ub) {
                        S.grp2[s] = open[leg];
                        S.gmem2[2 * open[leg] + 1] = s;
                        open[leg] = -1;
                    } else {
                        S.grp2[s] = ng2;
                        S.gmem2[2 * ng2] = s;
                        S.gmem2[2 * ng2 + 1] = -1;
                        open[leg] = ng2;
                        ++ng2;
                    }
                    last[leg] = lst;
                }
            }
            S.misc[3] = ng2;
        }
    }
    QR_SYNC();
    const int ng = S.misc[2], ng2 = S.misc[3];
    if (4 * ng > 3 * nf) return;   // hardly anything to pair
    qr_coarse_reduce<NT>(W.tri, W.Hs, W.g, W.ubz, ng, S.gmem, S.Hc, S.gc, S.ubc);
    C[0].ng = ng;
    if (L2 && ng2 > 0 && 4 * ng2 <= 3 * ng) {
        qr_coarse_reduce<NT>(W.tri, S.Hc, S.gc, S.ubc, ng2, S.gmem2, S.Hc2, S.gc2, S.ubc2);
        C[1].ng = ng2;
    }
}

// Scatter the stance solution into the 12h force vector (swing foot-steps are exactly zero).
template <int NT>
QR_DEV void qr_mpc_scatter(const QrMpcArgs& A, int prob, QrMpcSmem& S, const double* x, int status,
                           int ipm_iters, int as_rounds) {
    const int h = A.P.horizon;
    const bool valid = status != 2 && status != 3;
    QR_FOR(i, 12 * h) {
        const int k = i / 3, sl = S.slot[k];
        const double v64 = (sl >= 0 && valid) ? x[3 * sl + (i - 3 * k)] : 0.0;
        const float val = (float)v64;
        if (A.u_out) A.u_out[(size_t)prob * 12 * h + i] = val;
        if (i < 12 && A.grf_out) A.grf_out[(size_t)prob * 12 + i] = val;
        if (A.x_out) A.x_out[(size_t)prob * 12 * h + i] = val;
        if (A.x_out_f64) A.x_out_f64[(size_t)prob * 12 * h + i] = v64;
    }
    // epilogue, one thread per leg: what SolveDenseMPC / GetAction do with the first 12 forces
    // (qr_mpc_stance_leg_controller.cpp:402-409, 139-141; qr_robot.cpp:241-251)
    if (A.ep_q || A.ep_cmd) {
        QR_FOR(leg, 4) {
            const int sl = S.slot[leg];   // foot-step `leg` of step 0
            float f[3];
            for (int a = 0; a < 3; ++a) f[a] = (sl >= 0 && valid) ? (float)x[3 * sl + a] : 0.f;
            if (A.ep_cmd)
                for (int a = 0; a < 3; ++a) A.ep_cmd[(size_t)prob * 66 + 51 + 3 * leg + a] = f[a];
            if (A.ep_q && (A.ep_ff || A.ep_tau)) {
                float Rb[9], ff[3], tau[3];
                qr_mpc_base_rmat(S.state + 6, Rb);
                qr_mpc_leg_force_torque(A.ep_hip, A.ep_upper, A.ep_lower, Rb, leg, A.ep_q + (size_t)prob * 12 + 3 * leg, f, ff, tau);
                for (int a = 0; a < 3; ++a) {
                    if (A.ep_ff) A.ep_ff[(size_t)prob * 12 + 3 * leg + a] = ff[a];
                    if (A.ep_tau) A.ep_tau[(size_t)prob * 12 + 3 * leg + a] = tau[a];
                }
            }
        }
    }
    QR_THREADS(t) {
        if (t == 0) {
            if (A.status_out) A.status_out[prob] = status;
            if (A.iters_out) { A.iters_out[2 * prob] = ipm_iters; A.iters_out[2 * prob + 1] = as_rounds; }
        }
    }
}

// breakdown guard: a non-finite result is reported, never returned as a force
template <int NT>
QR_DEV int qr_result_status(QrMpcSmem& S, const double* x, int status) {
    int bad = 0;
    QR_FOR(i, 3 * S.W.nf) bad |= !(fabs(x[i]) < 1e300);
    return QR_ANY(bad) ? 3 : status;
}

// The fused path: SolveMPCKernel + GetMPCSolution for one instance.
template <int NT, bool L2>
QR_DEV void qr_mpc_solve_problem(const QrMpcArgs& A, int prob, QrMpcSmem& S) {
    QR_PROF_DECL;
    qr_mpc_stage<NT>(A, prob, S);
    QR_PROF(20);
    int status = S.misc[0];
    int it = 0, rounds = 0;
    const double* x = S.W.xn;
    if (status == 0) {
        qr_mpc_condense_to_work<NT>(A, S);
        QR_PROF(21);
        QrCoarse C[L2 ? 2 : 1];
        qr_mpc_build_coarse<NT, L2>(S, A.opt, A.coarse_rounds, C);
        QR_PROF(23);
        status = qr_qp_solve<NT, L2>(S.W, A.opt, &it, &rounds, &x, C QR_PROF_PASS);
        status = qr_result_status<NT>(S, x, status);
    }
    qr_mpc_scatter<NT>(A, prob, S, x, status, it, rounds);
    QR_SYNC();
    QR_PROF(22);
}

// Condense only: float32 H (n x n), g (n), ub (20h) exactly as SolveMPC hands them to qpOASES.
template <int NT>
QR_DEV void qr_mpc_condense_problem(const QrMpcArgs& A, int prob, QrMpcSmem& S) {
    qr_mpc_stage<NT>(A, prob, S);
    const int h = A.P.horizon, n = 12 * h;
    QrCondenseTables& T = *S.T;
    QR_THREADS(t) {
        if (t == 0) qr_condense_model(A.P, S.state, S.state + 3, S.state + 6, S.state + 10, S.state + 13, S.state + 25, T);
    }
    QR_SYNC();
    qr_condense_tables<NT>(A.P, S.traj, T);
    QR_SYNC();
    float* H = A.H_out + (size_t)prob * n * n;
    QR_FOR(idx, n * n) {
        const int r = idx / n, c = idx - r * n;
        H[idx] = qr_condense_h_entry(T, h, r / 12, (r % 12) / 3, r % 3, c / 12, (c % 12) / 3, c % 3);
    }
    QR_FOR(i, n) A.g_out[(size_t)prob * n + i] = qr_condense_g_entry(T, h, i / 12, (i % 12) / 3, i % 3);
    const float fmax = A.fmax_i ? A.fmax_i[prob] : A.P.f_max;
    QR_FOR(k, 4 * h) {
        float* ub = A.ub_out + (size_t)prob * 20 * h + 5 * k;
        ub[0] = ub[1] = ub[2] = ub[3] = 5e10f;            // BIG_NUMBER prefill (ResizeQPMats :219-229)
        ub[4] = QR_FMUL(S.gait[k], fmax);
    }
    QR_SYNC();
}

// QP only: caller-supplied float32 H, g, ub (the qpOASES call's inputs).
template <int NT>
QR_DEV void qr_qp_solve_problem(const QrMpcArgs& A, int prob, QrMpcSmem& S) {
    const int h = A.P.horizon, n = 12 * h;
    const float* H = A.H_in + (size_t)prob * n * n;
    const float* g = A.g_in + (size_t)prob * n;
    const float* ub = A.ub_in + (size_t)prob * 20 * h;
    QR_THREADS(t) {
        if (t == 0) {
            const float mu = A.mu_i ? A.mu_i[prob] : A.P.mu;
            S.scal[0] = (double)QR_FDIV(1.f, mu);
            int nf = 0, st = 0;
            for (int k = 0; k < 4 * h; ++k) {
                const float u = ub[5 * k + 4];
                if (u > 0.f && nf < A.nfcap) { S.fs[nf] = k; S.slot[k] = nf; S.W.ubz[nf] = (double)u; ++nf; }
                else { S.slot[k] = -1; if (u < 0.f) st = 2; if (!(u == u)) st = 3; if (u > 0.f) st = 3; }
            }
            if (!(mu > 0.f)) st = 3;
            S.misc[0] = st; S.misc[1] = nf;
        }
    }
    QR_SYNC();
    S.W.nf = S.misc[1];
    S.W.mu_ = S.scal[0];
    int status = S.misc[0], it = 0, rounds = 0;
    const double* x = S.W.xn;
    if (status == 0) {
        const int nf = S.W.nf;
        QR_FOR(idx, 9 * qr_ntri(nf)) {
            const int b = idx / 9, e = idx - 9 * b;
            const int code = S.W.tri[b];
            const int ri = 3 * S.fs[code >> 8] + e / 3, ci = 3 * S.fs[code & 255] + e % 3;
            S.W.Hs[idx] = 0.5 * ((double)H[(size_t)ri * n + ci] + (double)H[(size_t)ci * n + ri]);
        }
        QR_FOR(i, 3 * nf) S.W.g[i] = (double)g[3 * S.fs[i / 3] + i % 3];
        QR_SYNC();
        QR_PROF_DECL;
        QrCoarse C[2];
        qr_mpc_build_coarse<NT, true>(S, A.opt, A.coarse_rounds, C);
        status = qr_qp_solve<NT, true>(S.W, A.opt, &it, &rounds, &x, C QR_PROF_PASS);
        status = qr_result_status<NT>(S, x, status);
    }
    qr_mpc_scatter<NT>(A, prob, S, x, status, it, rounds);
    QR_SYNC();
}
