// qp_solver.h -- batched dense friction-pyramid QP, one thread team per problem, float64.
//
// Replaces the qpOASES call of SolveMPC (/root/reference/quadruped/src/controllers/mpc/
// qr_mpc_interface.cpp:414-438; constraint rows from ResizeQPMats :230-240):
//     min 1/2 x'Hx + g'x   s.t.  0 <= mu_*fx + fz, 0 <= -mu_*fx + fz, 0 <= mu_*fy + fz,
//                                0 <= -mu_*fy + fz, 0 <= fz <= ub        per foot-step (mu_ = 1/mu)
// Foot-steps with ub == 0 (swing) are pinned to f = 0 by these rows, so their variables are
// eliminated up front; the row 0 <= fz is implied by the pyramid rows and is dropped.  What is left
// are nf "stance" foot-steps, n = 3*nf variables and 5 inequalities on each 3-vector.
//
// Method (not qpOASES' online active set -- that is inherently sequential):
//   1. Mehrotra predictor-corrector interior point.  The constraint matrix is block diagonal
//      (5 rows on 3 variables), so A'DA is block-diagonal 3x3 and every iteration is one Cholesky
//      of K = H + blkdiag and two solves.  Run to a loose tolerance: it only has to identify the
//      active set.
//   2. Active-set polish.  Each foot-step's active rows define f = Z y + p with Z (3 x d, d<=3);
//      Z is padded to 3x3 so the reduced KKT matrix Z'HZ (+ identity on padded slots) keeps the
//      3x3 block structure and reuses the same factorisation.  After the equality-constrained solve
//      the multipliers and the inactive rows are checked; wrong guesses are corrected and the solve
//      repeated.  On exit the KKT conditions hold to feas_tol / mult_tol, i.e. the point is THE
//      optimum of the strictly convex QP (H >= 2*alpha*I), not an approximation of it.
//
// Storage: symmetric matrices are lower block-triangular with full 3x3 blocks, block (S,T), S >= T,
// at 9*(S*(S+1)/2 + T).  H lives in a per-CTA global scratch that stays L2-resident; K (the matrix
// being factorised) lives in shared memory.
#pragma once

#include "qr_team.h"
#include "../../include/qr_gpu.h"

struct QrQpWork {
    int nf;             // stance foot-steps
    double mu_;         // 1/mu as the reference rounds it (float32 value)
    const double* Hs;   // [9*ntri] symmetric block-packed Hessian (global scratch)
    double* K;          // [9*ntri] shared: matrix under factorisation
    double* Dinv;       // [9*nf]   inverse of the diagonal Cholesky blocks
    double* Zs;         // [9*nf]   polish bases (3x3, zero-padded columns)
    double* ps;         // [3*nf]   polish offsets
    double* g;          // [n]
    double* x;          // [n]  interior-point iterate
    double* xn;         // [n]  polished iterate
    double* q;          // [n]  H x + g
    double* wv;         // [n]  solve work vector
    double* yv;         // [n]  forward-solve result
    double* dxa;        // [n]
    double* dx;         // [n]
    double* rd;         // [n]
    double* s;          // [5*nf]
    double* lam;        // [5*nf]
    double* dsa;        // [5*nf]
    double* dla;        // [5*nf]
    double* rc;         // [5*nf]
    double* dl;         // [5*nf]
    double* ubz;        // [nf]
    double* red;        // [4*nf] reduction scratch
    int* act;           // [nf] active-row bit masks (bit c = row c; bit 4 = cap)
    int* flag;          // [nf]
    int* vert;          // [nf] polish: foot-step pinned to the apex f = 0
};

QR_DEV int qr_blk(int S, int T) { return 9 * ((S * (S + 1)) / 2 + T); }

// Decode a lower-triangular linear index idx -> (I, J), I >= J.
QR_DEV void qr_tri_decode(int idx, int& I, int& J) {
    int i = (int)((sqrtf(8.f * (float)idx + 1.f) - 1.f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= idx) ++i;
    while (i * (i + 1) / 2 > idx) --i;
    I = i;
    J = idx - i * (i + 1) / 2;
}

QR_DEV double qr_rsqrt(double v) {
#ifdef QR_ON_DEVICE
    return rsqrt(v);
#else
    return 1.0 / sqrt(v);
#endif
}

// Row i of (Hs * v), Hs symmetric block-packed.
QR_DEV double qr_sym_matvec_row(const double* Hs, const double* v, int nf, int i) {
    const int S = i / 3, a = i - 3 * S;
    double acc = 0.0;
    const double* row = Hs + qr_blk(S, 0) + 3 * a;
    for (int T = 0; T <= S; ++T, row += 9)
        acc += row[0] * v[3 * T] + row[1] * v[3 * T + 1] + row[2] * v[3 * T + 2];
    for (int T = S + 1; T < nf; ++T) {
        const double* col = Hs + qr_blk(T, S) + a;
        acc += col[0] * v[3 * T] + col[3] * v[3 * T + 1] + col[6] * v[3 * T + 2];
    }
    return acc;
}

// Cholesky factor of a 3x3 SPD block given by its lower part, as the reciprocal pivots and the
// off-diagonal entries (all a thread needs for its own forward substitution).
struct QrChol3 {
    double i00, i11, i22, l10, l20, l21;
};
QR_DEV QrChol3 qr_chol3(const double* D, int* bad) {
    QrChol3 c;
    double d00 = D[0];
    if (!(d00 > 1e-300)) { d00 = 1e-300; *bad = 1; }
    c.i00 = qr_rsqrt(d00);
    c.l10 = D[3] * c.i00;
    c.l20 = D[6] * c.i00;
    double t = D[4] - c.l10 * c.l10;
    if (!(t > 1e-300)) { t = 1e-300; *bad = 1; }
    c.i11 = qr_rsqrt(t);
    c.l21 = (D[7] - c.l20 * c.l10) * c.i11;
    double u = D[8] - c.l20 * c.l20 - c.l21 * c.l21;
    if (!(u > 1e-300)) { u = 1e-300; *bad = 1; }
    c.i22 = qr_rsqrt(u);
    return c;
}

// In-place blocked Cholesky K = L L'.  Off-diagonal blocks of K are overwritten with L; the diagonal
// blocks are left untouched and their inverse factors go to Dinv (that is all the solves need).
// Returns non-zero (uniformly) if a pivot had to be clamped.
template <int NT>
QR_DEV int qr_blk_cholesky(QrQpWork& W) {
    const int nf = W.nf;
    double* K = W.K;
    QR_FOR(f, nf) W.flag[f] = 0;
    for (int Kc = 0; Kc < nf; ++Kc) {
        // panel: every row of the block column solves against the (redundantly factorised) diagonal block
        QR_FOR(idx, 3 * (nf - Kc)) {
            const int I = Kc + idx / 3, a = idx % 3;
            int bad = 0;
            const QrChol3 c = qr_chol3(K + qr_blk(Kc, Kc), &bad);
            if (I == Kc) {
                double* Di = W.Dinv + 9 * Kc;
                if (a == 0) {
                    Di[0] = c.i00; Di[1] = 0.0; Di[2] = 0.0;
                } else if (a == 1) {
                    Di[3] = -c.l10 * c.i00 * c.i11; Di[4] = c.i11; Di[5] = 0.0;
                } else {
                    const double inv10 = -c.l10 * c.i00 * c.i11;
                    Di[6] = -(c.l20 * c.i00 + c.l21 * inv10) * c.i22;
                    Di[7] = -c.l21 * c.i11 * c.i22;
                    Di[8] = c.i22;
                    if (bad) W.flag[Kc] = 1;
                }
            } else {
                double* r = K + qr_blk(I, Kc) + 3 * a;
                const double x0 = r[0] * c.i00;
                const double x1 = (r[1] - x0 * c.l10) * c.i11;
                const double x2 = (r[2] - x0 * c.l20 - x1 * c.l21) * c.i22;
                r[0] = x0; r[1] = x1; r[2] = x2;
            }
        }
        QR_SYNC();
        // trailing update A_IJ -= L_IK L_JK'
        const int nb = nf - Kc - 1;
        QR_FOR(idx, (nb * (nb + 1)) / 2) {
            int I, J;
            qr_tri_decode(idx, I, J);
            I += Kc + 1; J += Kc + 1;
            const double* li = K + qr_blk(I, Kc);
            const double* lj = K + qr_blk(J, Kc);
            double* a = K + qr_blk(I, J);
            double l[9], m[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) { l[e] = li[e]; m[e] = lj[e]; }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c2 = 0; c2 < 3; ++c2)
                    a[3 * r + c2] -= l[3 * r] * m[3 * c2] + l[3 * r + 1] * m[3 * c2 + 1] + l[3 * r + 2] * m[3 * c2 + 2];
        }
        QR_SYNC();
    }
    int bad = 0;
    for (int f = 0; f < nf; ++f) bad |= W.flag[f];
    QR_SYNC();
    return bad;
}

// Solve (L L') out = W.wv.  W.wv and W.yv are destroyed.
template <int NT>
QR_DEV void qr_blk_solve(QrQpWork& W, double* out) {
    const int nf = W.nf;
    const double* K = W.K;
    double* v = W.wv;
    double* y = W.yv;
    // forward: L y = v
    for (int Kc = 0; Kc < nf; ++Kc) {
        QR_FOR(idx, 3 * (nf - Kc)) {
            const int I = Kc + idx / 3, a = idx % 3;
            const double* Di = W.Dinv + 9 * Kc;
            const double b0 = v[3 * Kc], b1 = v[3 * Kc + 1], b2 = v[3 * Kc + 2];
            const double y0 = Di[0] * b0;
            const double y1 = Di[3] * b0 + Di[4] * b1;
            const double y2 = Di[6] * b0 + Di[7] * b1 + Di[8] * b2;
            if (I == Kc) {
                y[3 * Kc + a] = (a == 0 ? y0 : (a == 1 ? y1 : y2));
            } else {
                const double* r = K + qr_blk(I, Kc) + 3 * a;
                v[3 * I + a] -= r[0] * y0 + r[1] * y1 + r[2] * y2;
            }
        }
        QR_SYNC();
    }
    // backward: L' out = y, in place on y (a step reads block Kc of y and updates blocks J < Kc)
    for (int Kc = nf - 1; Kc >= 0; --Kc) {
        QR_FOR(idx, 3 * (Kc + 1)) {
            const int J = idx / 3, a = idx % 3;
            const double* Di = W.Dinv + 9 * Kc;
            const double b0 = y[3 * Kc], b1 = y[3 * Kc + 1], b2 = y[3 * Kc + 2];
            // x_K = Linv' b
            const double x0 = Di[0] * b0 + Di[3] * b1 + Di[6] * b2;
            const double x1 = Di[4] * b1 + Di[7] * b2;
            const double x2 = Di[8] * b2;
            if (J == Kc) {
                out[3 * Kc + a] = (a == 0 ? x0 : (a == 1 ? x1 : x2));
            } else {
                const double* blk = K + qr_blk(Kc, J);   // L_KJ, column a of it
                y[3 * J + a] -= blk[a] * x0 + blk[3 + a] * x1 + blk[6 + a] * x2;
            }
        }
        QR_SYNC();
    }
}

// Constraint values of one foot-step: c0..c3 pyramid faces, c4 = ub - fz.
QR_DEV void qr_foot_constraints(double mu_, double ub, const double* f, double* c) {
    c[0] = mu_ * f[0] + f[2];
    c[1] = -mu_ * f[0] + f[2];
    c[2] = mu_ * f[1] + f[2];
    c[3] = -mu_ * f[1] + f[2];
    c[4] = ub - f[2];
}
// A_f' t for one foot-step (t has 5 entries).
QR_DEV void qr_foot_At(double mu_, const double* t, double* o) {
    o[0] = mu_ * (t[0] - t[1]);
    o[1] = mu_ * (t[2] - t[3]);
    o[2] = t[0] + t[1] + t[2] + t[3] - t[4];
}

template <int NT>
QR_DEV double qr_red_max(const double* r, int cnt) {
    double m = 0.0;
    for (int i = 0; i < cnt; ++i) m = qr_max(m, r[i]);
    return m;
}
template <int NT>
QR_DEV double qr_red_sum(const double* r, int cnt) {
    double m = 0.0;
    for (int i = 0; i < cnt; ++i) m += r[i];
    return m;
}
template <int NT>
QR_DEV double qr_red_min1(const double* r, int cnt) {
    double m = 1.0;
    for (int i = 0; i < cnt; ++i) m = qr_min(m, r[i]);
    return m;
}

// Copy Hs -> K and add the barrier blocks A'DA (D = lam/s) on the diagonal.
template <int NT>
QR_DEV void qr_build_kkt(QrQpWork& W) {
    const int nf = W.nf, ntri = (nf * (nf + 1)) / 2;
    QR_FOR(idx, 9 * ntri) W.K[idx] = W.Hs[idx];
    QR_SYNC();
    QR_FOR(f, nf) {
        const double* s = W.s + 5 * f;
        const double* l = W.lam + 5 * f;
        const double d0 = l[0] / s[0], d1 = l[1] / s[1], d2 = l[2] / s[2], d3 = l[3] / s[3], d4 = l[4] / s[4];
        const double m = W.mu_;
        double* D = W.K + qr_blk(f, f);
        const double xx = m * m * (d0 + d1), yy = m * m * (d2 + d3);
        const double xz = m * (d0 - d1), yz = m * (d2 - d3);
        D[0] += xx; D[4] += yy; D[8] += d0 + d1 + d2 + d3 + d4;
        D[2] += xz; D[6] += xz; D[5] += yz; D[7] += yz;
    }
    QR_SYNC();
}

// Interior-point phase.  On exit W.x, W.s, W.lam hold the final iterate.  Returns iterations used;
// *converged tells whether the tolerance was met.
template <int NT>
QR_DEV int qr_ipm(QrQpWork& W, const qr_qp_options& opt, int* converged) {
    const int nf = W.nf, n = 3 * nf, m = 5 * nf;
    const double mu_ = W.mu_;
    double* red = W.red;
    // ---- starting point: strictly interior, fz = ub/4
    QR_FOR(f, nf) {
        W.x[3 * f] = 0.0; W.x[3 * f + 1] = 0.0; W.x[3 * f + 2] = 0.25 * W.ubz[f];
    }
    QR_SYNC();
    QR_FOR(i, n) W.q[i] = qr_sym_matvec_row(W.Hs, W.x, nf, i) + W.g[i];
    QR_FOR(f, nf) qr_foot_constraints(mu_, W.ubz[f], W.x + 3 * f, W.s + 5 * f);
    QR_FOR(i, n) red[nf + i] = fabs(W.g[i]);
    QR_SYNC();
    QR_FOR(f, nf) {
        red[f] = qr_max(fabs(W.q[3 * f]), qr_max(fabs(W.q[3 * f + 1]), fabs(W.q[3 * f + 2])));
    }
    QR_SYNC();
    const double gscale = qr_max(1.0, qr_red_max<NT>(red + nf, n));
    {
        const double lam0 = qr_red_max<NT>(red, nf) + 1e-3;
        QR_SYNC();
        QR_FOR(f, nf) {
            const double* s = W.s + 5 * f;
            red[f] = s[0] + s[1] + s[2] + s[3] + s[4];
        }
        QR_SYNC();
        const double mu0 = lam0 * qr_red_sum<NT>(red, nf) / (double)m;
        QR_SYNC();
        QR_FOR(c, m) W.lam[c] = mu0 / W.s[c];
        QR_SYNC();
    }
    *converged = 0;
    int it = 0;
    for (; it < opt.max_ipm_iter; ++it) {
        // residuals (q, s are current)
        QR_FOR(f, nf) {
            double at[3];
            qr_foot_At(mu_, W.lam + 5 * f, at);
            double rmax = 0.0, gap = 0.0;
            for (int a = 0; a < 3; ++a) {
                const double r = W.q[3 * f + a] - at[a];
                W.rd[3 * f + a] = r;
                rmax = qr_max(rmax, fabs(r));
            }
            for (int c = 0; c < 5; ++c) gap += W.s[5 * f + c] * W.lam[5 * f + c];
            red[f] = rmax;
            red[nf + f] = gap;
        }
        QR_SYNC();
        const double rdmax = qr_red_max<NT>(red, nf);
        const double gap = qr_red_sum<NT>(red + nf, nf) / (double)m;
        QR_SYNC();
        if (rdmax < opt.ipm_tol * gscale && gap < opt.ipm_tol) { *converged = 1; break; }

        qr_build_kkt<NT>(W);
        qr_blk_cholesky<NT>(W);

        // predictor: K dxa = -(Hx + g)
        QR_FOR(i, n) W.wv[i] = -W.q[i];
        QR_SYNC();
        qr_blk_solve<NT>(W, W.dxa);
        QR_FOR(f, nf) {
            const double* dxf = W.dxa + 3 * f;
            double ds[5];
            ds[0] = mu_ * dxf[0] + dxf[2]; ds[1] = -mu_ * dxf[0] + dxf[2];
            ds[2] = mu_ * dxf[1] + dxf[2]; ds[3] = -mu_ * dxf[1] + dxf[2];
            ds[4] = -dxf[2];
            double ap = 1e30, ad = 1e30;
            for (int c = 0; c < 5; ++c) {
                const double s = W.s[5 * f + c], l = W.lam[5 * f + c];
                const double dl = -l - (l / s) * ds[c];
                W.dsa[5 * f + c] = ds[c];
                W.dla[5 * f + c] = dl;
                if (ds[c] < 0.0) ap = qr_min(ap, -s / ds[c]);
                if (dl < 0.0) ad = qr_min(ad, -l / dl);
            }
            red[f] = ap;
            red[nf + f] = ad;
        }
        QR_SYNC();
        const double apa = qr_red_min1<NT>(red, nf);
        const double ada = qr_red_min1<NT>(red + nf, nf);
        QR_SYNC();
        QR_FOR(f, nf) {
            double acc = 0.0;
            for (int c = 0; c < 5; ++c)
                acc += (W.s[5 * f + c] + apa * W.dsa[5 * f + c]) * (W.lam[5 * f + c] + ada * W.dla[5 * f + c]);
            red[f] = acc;
        }
        QR_SYNC();
        const double mu_aff = qr_red_sum<NT>(red, nf) / (double)m;
        QR_SYNC();
        double sigma = mu_aff / gap;
        sigma = sigma * sigma * sigma;
        // corrector right-hand side
        QR_FOR(f, nf) {
            double t[5];
            for (int c = 0; c < 5; ++c) {
                const double rc = W.s[5 * f + c] * W.lam[5 * f + c] + W.dsa[5 * f + c] * W.dla[5 * f + c] - sigma * gap;
                W.rc[5 * f + c] = rc;
                t[c] = rc / W.s[5 * f + c];
            }
            double at[3];
            qr_foot_At(mu_, t, at);
            for (int a = 0; a < 3; ++a) W.wv[3 * f + a] = -W.rd[3 * f + a] - at[a];
        }
        QR_SYNC();
        qr_blk_solve<NT>(W, W.dx);
        QR_FOR(f, nf) {
            const double* dxf = W.dx + 3 * f;
            double ds[5];
            ds[0] = mu_ * dxf[0] + dxf[2]; ds[1] = -mu_ * dxf[0] + dxf[2];
            ds[2] = mu_ * dxf[1] + dxf[2]; ds[3] = -mu_ * dxf[1] + dxf[2];
            ds[4] = -dxf[2];
            double ap = 1e30, ad = 1e30;
            for (int c = 0; c < 5; ++c) {
                const double s = W.s[5 * f + c], l = W.lam[5 * f + c];
                const double dl = -(W.rc[5 * f + c] + l * ds[c]) / s;
                W.dl[5 * f + c] = dl;
                if (ds[c] < 0.0) ap = qr_min(ap, -s / ds[c]);
                if (dl < 0.0) ad = qr_min(ad, -l / dl);
            }
            red[f] = ap;
            red[nf + f] = ad;
        }
        QR_SYNC();
        double ap = 1e30, ad = 1e30;
        for (int f = 0; f < nf; ++f) { ap = qr_min(ap, red[f]); ad = qr_min(ad, red[nf + f]); }
        QR_SYNC();
        ap = ap < 1.0 ? 0.995 * ap : 1.0;
        ad = ad < 1.0 ? 0.995 * ad : 1.0;
        QR_FOR(i, n) W.x[i] += ap * W.dx[i];
        QR_FOR(c, m) W.lam[c] += ad * W.dl[c];
        QR_SYNC();
        QR_FOR(i, n) W.q[i] = qr_sym_matvec_row(W.Hs, W.x, nf, i) + W.g[i];
        QR_FOR(f, nf) qr_foot_constraints(mu_, W.ubz[f], W.x + 3 * f, W.s + 5 * f);
        QR_SYNC();
    }
    return it;
}

// Basis of one foot-step's active face: f = Z y + p.  act bits 0..3 = pyramid faces, bit 4 = cap.
// Returns 1 when the active rows pin f = 0 (apex of the pyramid).
QR_DEV int qr_foot_basis(int act, double mu_, double ub, double* Z /*3x3 row-major*/, double* p) {
    const int a0 = act & 1, a1 = (act >> 1) & 1, a2 = (act >> 2) & 1, a3 = (act >> 3) & 1, cap = (act >> 4) & 1;
    for (int e = 0; e < 9; ++e) Z[e] = 0.0;
    p[0] = p[1] = p[2] = 0.0;
    if (a0 + a1 == 2 || a2 + a3 == 2 || a0 + a1 + a2 + a3 >= 3) return 1;
    const double im = 1.0 / mu_;
    // active face mu_*fx + fz = 0 -> fx = -fz/mu_ ; face -mu_*fx + fz = 0 -> fx = +fz/mu_
    const double kx = a0 ? -im : (a1 ? im : 0.0);
    const double ky = a2 ? -im : (a3 ? im : 0.0);
    const int fx_free = !(a0 | a1), fy_free = !(a2 | a3);
    if (fx_free) Z[0] = 1.0;            // column 0: e_x
    if (fy_free) Z[4] = 1.0;            // column 1: e_y
    if (cap) {
        p[0] = kx * ub; p[1] = ky * ub; p[2] = ub;
    } else {
        Z[2] = kx; Z[5] = ky; Z[8] = 1.0;  // column 2: (kx, ky, 1) * fz
    }
    return 0;
}

// Active-set polish.  Starts from the interior-point iterate (W.x, W.s, W.lam); result in W.xn.
// Returns rounds used; *ok = 1 when the KKT conditions were verified.
template <int NT>
QR_DEV int qr_polish(QrQpWork& W, const qr_qp_options& opt, int* ok) {
    const int nf = W.nf, n = 3 * nf, ntri = (nf * (nf + 1)) / 2;
    const double mu_ = W.mu_;
    QR_FOR(f, nf) {
        int a = 0;
        for (int c = 0; c < 5; ++c)
            if (W.s[5 * f + c] < opt.act_kappa * W.lam[5 * f + c]) a |= (1 << c);
        W.act[f] = a;
    }
    QR_SYNC();
    *ok = 0;
    int round = 0;
    for (; round < opt.max_polish_rounds; ++round) {
        QR_FOR(f, nf) { W.vert[f] = qr_foot_basis(W.act[f], mu_, W.ubz[f], W.Zs + 9 * f, W.ps + 3 * f); }
        QR_SYNC();
        // hp = H p + g  -> q
        QR_FOR(i, n) W.q[i] = qr_sym_matvec_row(W.Hs, W.ps, nf, i) + W.g[i];
        // reduced matrix Z_S' H_ST Z_T (+ identity on padded slots)
        QR_FOR(idx, ntri) {
            int S, T;
            qr_tri_decode(idx, S, T);
            const double* Hb = W.Hs + 9 * idx;
            const double* ZS = W.Zs + 9 * S;
            const double* ZT = W.Zs + 9 * T;
            double tmp[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    tmp[3 * r + c] = Hb[3 * r] * ZT[c] + Hb[3 * r + 1] * ZT[3 + c] + Hb[3 * r + 2] * ZT[6 + c];
            double* Kb = W.K + 9 * idx;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    Kb[3 * r + c] = ZS[r] * tmp[c] + ZS[3 + r] * tmp[3 + c] + ZS[6 + r] * tmp[6 + c];
            if (S == T) {
                // a padded slot has an all-zero basis column
                for (int c = 0; c < 3; ++c)
                    if (ZS[c] == 0.0 && ZS[3 + c] == 0.0 && ZS[6 + c] == 0.0) Kb[4 * c] = 1.0;
            }
        }
        QR_SYNC();
        QR_FOR(i, n) {
            const int f = i / 3, c = i - 3 * f;
            const double* Z = W.Zs + 9 * f;
            W.wv[i] = -(Z[c] * W.q[3 * f] + Z[3 + c] * W.q[3 * f + 1] + Z[6 + c] * W.q[3 * f + 2]);
        }
        QR_SYNC();
        qr_blk_cholesky<NT>(W);
        qr_blk_solve<NT>(W, W.dx);
        QR_FOR(i, n) {
            const int f = i / 3, a = i - 3 * f;
            const double* Z = W.Zs + 9 * f;
            const double* y = W.dx + 3 * f;
            W.xn[i] = Z[3 * a] * y[0] + Z[3 * a + 1] * y[1] + Z[3 * a + 2] * y[2] + W.ps[i];
        }
        QR_SYNC();
        QR_FOR(i, n) W.q[i] = qr_sym_matvec_row(W.Hs, W.xn, nf, i) + W.g[i];
        QR_SYNC();
        // verify / correct the active sets
        QR_FOR(f, nf) {
            const double* r = W.q + 3 * f;
            const int act = W.act[f];
            int nact = act, changed = 0;
            const double im = 1.0 / mu_;
            if (W.vert[f]) {
                // apex: the gradient must lie in the cone spanned by the four face normals
                if (r[2] < (fabs(r[0]) + fabs(r[1])) * im - opt.mult_tol) {
                    const double l0 = r[0] > 0.0 ? r[0] * im : 0.0, l1 = r[0] < 0.0 ? -r[0] * im : 0.0;
                    const double l2 = r[1] > 0.0 ? r[1] * im : 0.0, l3 = r[1] < 0.0 ? -r[1] * im : 0.0;
                    nact = (l0 > 0.0 ? 1 : 0) | (l1 > 0.0 ? 2 : 0) | (l2 > 0.0 ? 4 : 0) | (l3 > 0.0 ? 8 : 0);
                    const int hasx = nact & 3, hasy = nact & 12;
                    if (hasx && hasy) {
                        // least-squares multipliers on the edge; drop the face that would pull inward
                        const double sx = (nact & 1) ? 1.0 : -1.0, sy = (nact & 4) ? 1.0 : -1.0;
                        const double b0 = sx * mu_ * r[0] + r[2], b1 = sy * mu_ * r[1] + r[2];
                        const double dd = mu_ * mu_ + 1.0, det = dd * dd - 1.0;
                        const double lx = (dd * b0 - b1) / det, ly = (dd * b1 - b0) / det;
                        if (lx < 0.0 || ly < 0.0) {
                            if (lx < ly) nact &= ~3; else nact &= ~12;
                        }
                    }
                    changed = 1;
                }
            } else {
                double c[5];
                qr_foot_constraints(mu_, W.ubz[f], W.xn + 3 * f, c);
                int viol = 0;
                for (int k = 0; k < 5; ++k)
                    if (!((act >> k) & 1) && c[k] < -opt.feas_tol) viol |= (1 << k);
                if (viol) {
                    nact = act | viol;
                    changed = 1;
                } else if (act) {
                    // multipliers of the (independent) active rows: r = sum lambda_c a_c
                    const double lx = (act & 1) ? r[0] * im : ((act & 2) ? -r[0] * im : 0.0);
                    const double ly = (act & 4) ? r[1] * im : ((act & 8) ? -r[1] * im : 0.0);
                    const double lc = (act & 16) ? (lx + ly - r[2]) : 0.0;
                    double worst = -opt.mult_tol;
                    int drop = 0;
                    if ((act & 3) && lx < worst) { worst = lx; drop = act & 3; }
                    if ((act & 12) && ly < worst) { worst = ly; drop = act & 12; }
                    if ((act & 16) && lc < worst) { worst = lc; drop = 16; }
                    if (drop) { nact = act & ~drop; changed = 1; }
                }
            }
            W.act[f] = nact;
            W.red[f] = changed ? 1.0 : 0.0;
        }
        QR_SYNC();
        const double any = qr_red_max<NT>(W.red, nf);
        QR_SYNC();
        if (any == 0.0) { *ok = 1; ++round; break; }
    }
    return round;
}

// Full solve on a prepared workspace (Hs, g, ubz, mu_ set).  Result in W.xn (verified) or W.x.
// Returns the per-instance status code of qr_gpu.h.
template <int NT>
QR_DEV int qr_qp_solve(QrQpWork& W, const qr_qp_options& opt, int* ipm_iters, int* polish_rounds,
                       const double** result) {
    int conv = 0, ok = 0;
    *ipm_iters = 0; *polish_rounds = 0;
    *result = W.xn;
    if (W.nf == 0) return 0;
    *ipm_iters = qr_ipm<NT>(W, opt, &conv);
    *polish_rounds = qr_polish<NT>(W, opt, &ok);
    if (!ok) { *result = W.x; return 1; }
    return 0;
}
