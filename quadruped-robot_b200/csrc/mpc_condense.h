// mpc_condense.h -- single-rigid-body model, closed-form discretisation and QP condensing, float32.
//
// Replaces (paths relative to /root/reference/quadruped/src/controllers/mpc/):
//   ComputeContinuousTimeStateSpaceMatrices  qr_mpc_interface.cpp:296-331
//   ConvertToDiscreteQP                      qr_mpc_interface.cpp:257-293
//   the H / g / U_b build of SolveMPC        qr_mpc_interface.cpp:359-412
//
// Design.  The reference forms dense 13h x 12h matrices and multiplies them.  Here nothing of that
// size is ever materialised: with M = dt*[A B;0 0] nilpotent of index 3 the discrete model has the
// closed form  Adt = I + dt*A + dt^2/2*A^2,  Bdt = dt*B + dt^2/2*A*B, the powers Adt^k differ from
// the identity in 9 + 3 + 2 entries, and every block of Bqp on the k-th sub-diagonal is the same
// 13 x 12 matrix G_k = Adt^k * Bdt of which only rows 0..5 depend on k.  A team keeps the small
// tables (G_k, G_k scaled by the weights, Aqp*x0 - X_d) in shared memory and each thread
// accumulates whole entries of H and g from them.
//
// Bit-exactness.  The reference does all of this in float32.  Rebuilding the QP in another
// precision or summation order moves the optimal forces by more than the parity tolerance
// (SURVEY.md section 0 fact 2), so every float32 operation below is issued in the order the
// reference's dense products visit their NON-ZERO terms (k ascending, multiply then add, no FMA
// contraction: QR_FMUL / QR_FADD); terms that are structurally zero are skipped, which changes
// nothing because x + 0 == x.  tests/ check (H, g, ub) against the oracle for equality, and against the
// reference's own qr_mpc_interface.cpp compiled from /root/reference (oracle/_ref/libqr_mpc_ref.so) for equality
// when that build evaluates exp() of the nilpotent matrix by its finite series (the test hook
// MINI_EIGEN_EXP_NILPOTENT3); against the scaling-and-squaring Pade evaluation of Eigen's MatrixFunctions the
// entries agree to 3e-7 relative (float32 rounding of the Pade steps), not bitwise.
#pragma once
// Loops of the once-per-instance phases are kept rolled: the throughput kernels are bound by instruction fetch, and an
// unrolled copy of a loop body is code that evicts the active-set round loop of the other CTAs on the SM.
#ifndef QR_UNROLL_SMALL
#define QR_UNROLL_SMALL _Pragma("unroll 1")
#endif

#include "qr_team.h"
#include "../../include/qr_gpu.h"

#define QR_HMAX QR_MAX_HORIZON

struct alignas(16) QrF4 { float x, y, z, w; };

// Shared-memory tables of one problem (about 8 KB).
struct QrCondenseTables {
    QrF4 Gx[QR_HMAX][12];        // {G_k(0,c), G_k(1,c), G_k(2,c), G_k(3+a, 3b+a)}: rows 0..2 of column c, and gpos[k]
    QrF4 TGx[QR_HMAX][12];       // the same times 2*w[row]: {TG02[k][0..2][c], tgpos[k][c % 3]}
    float G68[3][12];            // rows 6..8 of G_k (k independent) = dt * Iw^-1 [r]x
    float TG68[3][12];
    float gvel;                  // G_k(9+a, 3b+a) = dt/m
    float tgvel[3];
    float e[QR_HMAX][12];        // (Aqp x0 - X_d), rows 0..11 of every step
    float prot[QR_HMAX + 1][9];  // Adt^k(0:3, 6:9)
    float ppos[QR_HMAX + 1];     // Adt^k(3+a, 9+a)
    float p1112[QR_HMAX + 1];    // Adt^k(11, 12)
    float p512[QR_HMAX + 1];     // Adt^k(5, 12)
    float bdt02[3][12];          // Bdt rows 0..2
    float bpos;                  // Bdt(3+a, 3b+a)
    float dtRt[9];               // dt * R^T
    float x0[13];
    float w2[12];
    float two_alpha;
    float h2;                    // Adt(5,12) = dt^2/2
    float dt;
};

// Phase A (one thread): continuous model and the discrete blocks.
// p,v,quat,w,r_feet,rpy are this problem's rows.
QR_DEV void qr_condense_model(const qr_mpc_params& P, const float* p, const float* v,
                              const float* quat, const float* w, const float* r_feet,
                              const float* rpy, QrCondenseTables& T) {
    const float dt = P.dt;
    // x0 = [rpy, p, w, v, -9.8]   (qr_mpc_interface.cpp:362)
    for (int i = 0; i < 3; ++i) {
        T.x0[i] = rpy[i]; T.x0[3 + i] = p[i]; T.x0[6 + i] = w[i]; T.x0[9 + i] = v[i];
    }
    T.x0[12] = -9.8f;
    for (int i = 0; i < 12; ++i) T.w2[i] = QR_FMUL(2.f, P.weights[i]);
    T.two_alpha = QR_FMUL(2.f, P.alpha);
    T.dt = dt;

    // Quaternion (w,x,y,z) -> rotation, Eigen's toRotationMatrix() sequence (:344-351)
    const float qw = quat[0], qx = quat[1], qy = quat[2], qz = quat[3];
    const float tx = QR_FMUL(2.f, qx), ty = QR_FMUL(2.f, qy), tz = QR_FMUL(2.f, qz);
    const float twx = QR_FMUL(tx, qw), twy = QR_FMUL(ty, qw), twz = QR_FMUL(tz, qw);
    const float txx = QR_FMUL(tx, qx), txy = QR_FMUL(ty, qx), txz = QR_FMUL(tz, qx);
    const float tyy = QR_FMUL(ty, qy), tyz = QR_FMUL(tz, qy), tzz = QR_FMUL(tz, qz);
    float R[9];
    R[0] = QR_FSUB(1.f, QR_FADD(tyy, tzz)); R[1] = QR_FSUB(txy, twz); R[2] = QR_FADD(txz, twy);
    R[3] = QR_FADD(txy, twz); R[4] = QR_FSUB(1.f, QR_FADD(txx, tzz)); R[5] = QR_FSUB(tyz, twx);
    R[6] = QR_FSUB(txz, twy); R[7] = QR_FADD(tyz, twx); R[8] = QR_FSUB(1.f, QR_FADD(txx, tyy));

    // I_world = (R diag(I)) R^T and its inverse by cofactors (:365, :324)
    float RI[9], Iw[9], Ii[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) RI[3 * i + j] = QR_FMUL(R[3 * i + j], P.inertia[j]);
#define QR_DOT3(a0, b0, a1, b1, a2, b2) \
    QR_FADD(QR_FADD(QR_FMUL(a0, b0), QR_FMUL(a1, b1)), QR_FMUL(a2, b2))
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Iw[3 * i + j] = QR_DOT3(RI[3 * i], R[3 * j], RI[3 * i + 1], R[3 * j + 1], RI[3 * i + 2], R[3 * j + 2]);
#define QR_COF(a, b, c, d) QR_FSUB(QR_FMUL(Iw[a], Iw[b]), QR_FMUL(Iw[c], Iw[d]))
    const float c00 = QR_COF(4, 8, 5, 7), c01 = QR_COF(5, 6, 3, 8), c02 = QR_COF(3, 7, 4, 6);
    const float det = QR_FADD(QR_FADD(QR_FMUL(Iw[0], c00), QR_FMUL(Iw[1], c01)), QR_FMUL(Iw[2], c02));
    const float id = QR_FDIV(1.f, det);
    Ii[0] = QR_FMUL(c00, id);
    Ii[1] = QR_FMUL(QR_COF(2, 7, 1, 8), id);
    Ii[2] = QR_FMUL(QR_COF(1, 5, 2, 4), id);
    Ii[3] = QR_FMUL(c01, id);
    Ii[4] = QR_FMUL(QR_COF(0, 8, 2, 6), id);
    Ii[5] = QR_FMUL(QR_COF(2, 3, 0, 5), id);
    Ii[6] = QR_FMUL(c02, id);
    Ii[7] = QR_FMUL(QR_COF(1, 6, 0, 7), id);
    Ii[8] = QR_FMUL(QR_COF(0, 4, 1, 3), id);
#undef QR_COF
    const float minv = QR_FDIV(1.f, P.mass);

    // dt*A: rows 0..2 hold dt*R^T  (A.block(0,6,3,3) = R^T, :313)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T.dtRt[3 * i + j] = QR_FMUL(dt, R[3 * j + i]);
    // dt*B rows 6..8 = dt * Iw^-1 [r_b]x (:327-330); they are rows 6..8 of Bdt and of every G_k.
    for (int b = 0; b < 4; ++b) {
        const float rx = r_feet[3 * b], ry = r_feet[3 * b + 1], rz = r_feet[3 * b + 2];
        const float S[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const float bij = QR_DOT3(Ii[3 * i], S[j], Ii[3 * i + 1], S[3 + j], Ii[3 * i + 2], S[6 + j]);
                T.G68[i][3 * b + j] = QR_FMUL(dt, bij);
            }
    }
    const float dtminv = QR_FMUL(dt, minv);
    T.gvel = dtminv;
    T.bpos = QR_FMUL(0.5f, QR_FMUL(dt, dtminv));   // Bdt(3+a,3b+a) = 1/2 * (dtA*dtB)
    T.h2 = QR_FMUL(0.5f, QR_FMUL(dt, dt));         // Adt(5,12)
    // Bdt rows 0..2 = 1/2 * (dt R^T)(dt B_rows6..8), sequential over the three non-zero terms
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 12; ++j) {
            const float s = QR_DOT3(T.dtRt[3 * i], T.G68[0][j], T.dtRt[3 * i + 1], T.G68[1][j],
                                    T.dtRt[3 * i + 2], T.G68[2][j]);
            T.bdt02[i][j] = QR_FMUL(0.5f, s);
        }
#undef QR_DOT3
    // Powers of Adt by repeated left multiplication (:272-276): each non-trivial entry is a running sum.
    for (int i = 0; i < 9; ++i) T.prot[0][i] = 0.f;
    T.ppos[0] = 0.f; T.p1112[0] = 0.f; T.p512[0] = 0.f;
    for (int k = 1; k <= P.horizon; ++k) {
        for (int i = 0; i < 9; ++i) T.prot[k][i] = QR_FADD(T.prot[k - 1][i], T.dtRt[i]);
        T.ppos[k] = QR_FADD(T.ppos[k - 1], dt);
        T.p1112[k] = QR_FADD(T.p1112[k - 1], dt);
        T.p512[k] = QR_FADD(QR_FADD(T.p512[k - 1], QR_FMUL(dt, T.p1112[k - 1])), T.h2);
    }
}

// Phase B (whole team): the k-dependent tables.  Call after a barrier following qr_condense_model.
template <int NT>
QR_DEV void qr_condense_tables(const qr_mpc_params& P, const float* traj, QrCondenseTables& T) {
    const int h = P.horizon;
    QR_FOR(idx, h * 36) {
        const int k = idx / 36, i = (idx % 36) / 12, j = idx % 12;
        // G_k(i,j) = Bdt(i,j) + sum_c Adt^k(i,6+c) * Bdt(6+c,j)
        float s = T.bdt02[i][j];
        s = QR_FADD(s, QR_FMUL(T.prot[k][3 * i + 0], T.G68[0][j]));
        s = QR_FADD(s, QR_FMUL(T.prot[k][3 * i + 1], T.G68[1][j]));
        s = QR_FADD(s, QR_FMUL(T.prot[k][3 * i + 2], T.G68[2][j]));
        (&T.Gx[k][j].x)[i] = s;
        (&T.TGx[k][j].x)[i] = QR_FMUL(s, T.w2[i]);
    }
    QR_FOR(idx, 36) {
        const int i = idx / 12, j = idx % 12;
        T.TG68[i][j] = QR_FMUL(T.G68[i][j], T.w2[6 + i]);
    }
    QR_FOR(idx, h * 12) {
        const int k = idx / 12, c = idx % 12;
        const float gp = QR_FADD(T.bpos, QR_FMUL(T.ppos[k], T.gvel));
        T.Gx[k][c].w = gp;
        T.TGx[k][c].w = QR_FMUL(gp, T.w2[3 + c % 3]);
    }
    QR_FOR(a, 3) { T.tgvel[a] = QR_FMUL(T.gvel, T.w2[9 + a]); }
    // e = Aqp*x0 - X_d, rows 0..11 of step j use Adt^(j+1)   (:411-412, :382-385)
    QR_FOR(idx, h * 12) {
        const int j = idx / 12, nrow = idx % 12, k = j + 1;
        const float* x0 = T.x0;
        float acc;
        if (nrow < 3) {
            acc = x0[nrow];
            acc = QR_FADD(acc, QR_FMUL(T.prot[k][3 * nrow + 0], x0[6]));
            acc = QR_FADD(acc, QR_FMUL(T.prot[k][3 * nrow + 1], x0[7]));
            acc = QR_FADD(acc, QR_FMUL(T.prot[k][3 * nrow + 2], x0[8]));
        } else if (nrow < 6) {
            acc = QR_FADD(x0[nrow], QR_FMUL(T.ppos[k], x0[nrow + 6]));
            if (nrow == 5) acc = QR_FADD(acc, QR_FMUL(T.p512[k], x0[12]));
        } else if (nrow < 11) {
            acc = x0[nrow];
        } else {
            acc = QR_FADD(x0[11], QR_FMUL(T.p1112[k], x0[12]));
        }
        T.e[j][nrow] = QR_FSUB(acc, traj[12 * j + nrow]);
    }
}

// One float32 entry of qH: row 12*i + 3*la + aa, column 12*j + 3*lb + ab  (:411).  The terms of rows 6..11
// do not depend on r: their float32 products are formed once and added in the reference's position.
QR_DEV float qr_condense_h_entry(const QrCondenseTables& T, int h, int i, int la, int aa, int j,
                                 int lb, int ab) {
    const int ca = 3 * la + aa, cb = 3 * lb + ab;
    const bool same_axis = (aa == ab);
    const float p6 = QR_FMUL(T.TG68[0][ca], T.G68[0][cb]);
    const float p7 = QR_FMUL(T.TG68[1][ca], T.G68[1][cb]);
    const float p8 = QR_FMUL(T.TG68[2][ca], T.G68[2][cb]);
    const float pv = QR_FMUL(T.tgvel[aa], T.gvel);
    float s = 0.f;
    const int r0 = (i > j ? i : j);
    const QrF4* tg = &T.TGx[r0 - i][ca];
    const QrF4* gg = &T.Gx[r0 - j][cb];
    for (int r = r0; r < h; ++r, tg += 12, gg += 12) {
        const QrF4 a = *tg, b = *gg;
        s = QR_FADD(s, QR_FMUL(a.x, b.x));
        s = QR_FADD(s, QR_FMUL(a.y, b.y));
        s = QR_FADD(s, QR_FMUL(a.z, b.z));
        if (same_axis) s = QR_FADD(s, QR_FMUL(a.w, b.w));
        s = QR_FADD(s, p6);
        s = QR_FADD(s, p7);
        s = QR_FADD(s, p8);
        if (same_axis) s = QR_FADD(s, pv);
    }
    if (i == j && ca == cb) s = QR_FADD(s, T.two_alpha);
    return s;
}

// Entry (12*i + 3*la + aa, 12*j + 3*lb + ab) of qH and its transposed entry in ONE loop: the same two float32
// summation chains as two calls of qr_condense_h_entry (bit-identical results), interleaved so that the dependent
// FADD chain of one hides the latency of the other (the symmetrised Hessian needs both, see qr_mpc_condense_to_work).
QR_DEV void qr_condense_h_pair(const QrCondenseTables& T, int h, int i, int la, int aa, int j, int lb, int ab,
                               float* hst, float* hts) {
    const int ca = 3 * la + aa, cb = 3 * lb + ab;
    const bool same_axis = (aa == ab);
    const float p6 = QR_FMUL(T.TG68[0][ca], T.G68[0][cb]), q6 = QR_FMUL(T.TG68[0][cb], T.G68[0][ca]);
    const float p7 = QR_FMUL(T.TG68[1][ca], T.G68[1][cb]), q7 = QR_FMUL(T.TG68[1][cb], T.G68[1][ca]);
    const float p8 = QR_FMUL(T.TG68[2][ca], T.G68[2][cb]), q8 = QR_FMUL(T.TG68[2][cb], T.G68[2][ca]);
    const float pv = QR_FMUL(T.tgvel[aa], T.gvel), qv = QR_FMUL(T.tgvel[ab], T.gvel);
    float s = 0.f, u = 0.f;
    const int r0 = (i > j ? i : j);
    const QrF4* tga = &T.TGx[r0 - i][ca];
    const QrF4* gga = &T.Gx[r0 - i][ca];
    const QrF4* tgb = &T.TGx[r0 - j][cb];
    const QrF4* ggb = &T.Gx[r0 - j][cb];
    QR_UNROLL_SMALL
    for (int r = r0; r < h; ++r, tga += 12, gga += 12, tgb += 12, ggb += 12) {
        const QrF4 a = *tga, b = *ggb;     // entry (s, t): TGx[r - i][ca] . Gx[r - j][cb]
        const QrF4 c = *tgb, d = *gga;     // entry (t, s): TGx[r - j][cb] . Gx[r - i][ca]
        s = QR_FADD(s, QR_FMUL(a.x, b.x));
        u = QR_FADD(u, QR_FMUL(c.x, d.x));
        s = QR_FADD(s, QR_FMUL(a.y, b.y));
        u = QR_FADD(u, QR_FMUL(c.y, d.y));
        s = QR_FADD(s, QR_FMUL(a.z, b.z));
        u = QR_FADD(u, QR_FMUL(c.z, d.z));
        if (same_axis) { s = QR_FADD(s, QR_FMUL(a.w, b.w)); u = QR_FADD(u, QR_FMUL(c.w, d.w)); }
        s = QR_FADD(s, p6);
        u = QR_FADD(u, q6);
        s = QR_FADD(s, p7);
        u = QR_FADD(u, q7);
        s = QR_FADD(s, p8);
        u = QR_FADD(u, q8);
        if (same_axis) { s = QR_FADD(s, pv); u = QR_FADD(u, qv); }
    }
    if (i == j && ca == cb) { s = QR_FADD(s, T.two_alpha); u = QR_FADD(u, T.two_alpha); }
    *hst = s;
    *hts = u;
}

// One float32 entry of qg: row 12*i + 3*la + aa  (:412).
QR_DEV float qr_condense_g_entry(const QrCondenseTables& T, int h, int i, int la, int aa) {
    const int ca = 3 * la + aa;
    float s = 0.f;
    QR_UNROLL_SMALL
    for (int j = i; j < h; ++j) {
        const QrF4 a = T.TGx[j - i][ca];
        s = QR_FADD(s, QR_FMUL(a.x, T.e[j][0]));
        s = QR_FADD(s, QR_FMUL(a.y, T.e[j][1]));
        s = QR_FADD(s, QR_FMUL(a.z, T.e[j][2]));
        s = QR_FADD(s, QR_FMUL(a.w, T.e[j][3 + aa]));
        s = QR_FADD(s, QR_FMUL(T.TG68[0][ca], T.e[j][6]));
        s = QR_FADD(s, QR_FMUL(T.TG68[1][ca], T.e[j][7]));
        s = QR_FADD(s, QR_FMUL(T.TG68[2][ca], T.e[j][8]));
        s = QR_FADD(s, QR_FMUL(T.tgvel[aa], T.e[j][9 + aa]));
    }
    return s;
}
