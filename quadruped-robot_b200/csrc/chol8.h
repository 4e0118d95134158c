// chol8.h -- blocked Cholesky of the LARGE reduced systems on the FP64 tensor cores (device only).
//
// The long-horizon size classes (h = 30: up to 216 reduced variables) spend half their time in the factorisation of
// Z'HZ.  With 3x3 blocks (qr_ldl_factor) that is 45..72 elimination steps, each a read-modify-write pass over the whole
// trailing matrix in shared memory by scalar DFMA code: on one 256-thread CTA per SM the pass is bound by instruction
// issue and shared-memory latency (FP64 pipe 12 % busy, profiles/r02_fused_h30_kernel_ncu_metrics.txt).  Here the same
// matrix is held as 8x8 tiles and factorised as K = L L' with one tile column per step:
//     diag:     L_pp = chol(A_pp), R_p = L_pp^-1 (one warp, registers; overwrites the tile -- the solves only need R_p)
//     panel:    L_Ip = A_Ip R_p'                        2 x mma.sync.m8n8k4.f64 per tile
//     trailing: A_IJ -= L_Ip L_Jp'   (I >= J > p)       2 x mma.sync.m8n8k4.f64 per tile
// i.e. 8 columns per pass over the trailing matrix instead of 3, an eighth of the instructions per flop, and two CTA
// barriers per step (the next step's diagonal factor is computed by the warp that updates that tile while the other
// warps finish the trailing update).  The forward substitution of the right-hand side rides along, the backward
// substitution is a second sweep over the tile rows.  Summation order differs from the 3x3-block code, so results agree
// with it to rounding (1e-13 relative on the solves), not bit for bit; the active-set iteration verifies the KKT
// conditions with the exact Hessian either way.  The host emulation (tests/emul) has no tensor cores and keeps the
// 3x3-block code for every size.
//
// Layout: tile (I, J), I >= J, of an nt x nt tile matrix at 64 * (J*nt - J(J-1)/2 + I - J) (packed by columns, like
// qr_kblk); inside a tile entry (r, c) sits at 8r + (c ^ 4*((r >> 1) & 1)).  The XOR makes the fragment loads of
// m8n8k4 -- lane (g, q) reads (g, 4h + q) -- hit 16 distinct 8-byte bank pairs per half-warp, and keeps the pairs
// (g, 2q), (g, 2q + 1) of the accumulator fragment adjacent and 16-byte aligned.  Diagonal tiles are stored in full.
#pragma once

#ifndef QR_CHOL8_MIN_NB
#define QR_CHOL8_MIN_NB 16   // reduced systems of at least this many 3x3 block columns take this path (when the class supports it)
#endif
#ifndef QR_CHOL8_MIN_CAP
#define QR_CHOL8_MIN_CAP 56  // workspace classes (stance foot-steps) whose K region is sized for the 8x8 layout
#endif
#ifndef QR_CHOL8_MIN_NT
#define QR_CHOL8_MIN_NT 256  // team size from which the path is compiled in (one CTA per SM: the classes bound by step latency)
#endif
#ifndef QR_CHOL8_BIAS
#define QR_CHOL8_BIAS 14     // trailing tiles each other warp takes before warp 0 (busy with the next diagonal factor) joins in
#endif

#ifndef QR_C8P
#define QR_C8P(tag) ((void)0)   // cycle marks of tests/cpp/chol8_test.cu
#endif

QR_HD int qr_k8_nt(int nred) { return (nred + 7) >> 3; }
// bytes of the 8x8-tile layout for a workspace of nfcap stance foot-steps (0: class below QR_CHOL8_MIN_CAP)
QR_HD size_t qr_k8_bytes(int nfcap) {
    if (nfcap < QR_CHOL8_MIN_CAP) return 0;
    const size_t nt = (size_t)qr_k8_nt(3 * nfcap);
    return nt * (nt + 1) / 2 * 64 * sizeof(double);
}

#if defined(__CUDACC__)
__device__ __forceinline__ int qr_k8_swz(int r, int c) { return r * 8 + (c ^ (((r >> 1) & 1) << 2)); }
__device__ __forceinline__ int qr_k8_tile(int nt, int I, int J) { return 64 * (J * nt - ((J * (J - 1)) >> 1) + (I - J)); }
// entry (i, j) of the lower triangle (i >= j, or any pair inside one diagonal tile)
__device__ __forceinline__ int qr_k8_idx(int nt, int i, int j) {
    return qr_k8_tile(nt, i >> 3, j >> 3) + qr_k8_swz(i & 7, j & 7);
}

__device__ __forceinline__ void qr_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Cholesky factor of one 8x8 diagonal tile, returned as its inverse R = L^-1 (lower triangular) written over the tile
// with zeros above the diagonal.  Called by a whole warp: every lane runs the same register-resident recurrence (no
// shuffle or shared-memory round trip on the dependent chain) and stores every entry (same value, same address: uniform
// control flow -- a per-lane split of the 64 stores compiled to a 32-way divergent tail that cost more than the
// factorisation itself, 7.5 k cycles against 1.6 k).  R is eliminated alongside L (Gauss-Jordan on [A | I]): row k of R
// is E_k / l_kk and E_i -= l_ik R_k for i > k.
// The routine is bound by its pivot chain (rsqrt -> scale -> update, ~190 cycles per pivot), not by its instruction
// count: measured with tests/cpp/chol8_test.cu (cycles per tile, lone warp) this form takes 1.55 k; the inverse split over
// the lanes (lane j solves L r = e_j, 45 instructions instead of 120) 1.56 k; [A | I] spread over the lanes with shuffles
// 1.61 k; square-root-free elimination with the products formed under the reciprocal 2.45 k; a float-seeded Newton rsqrt
// instead of the library's 2.5 k (the conversions are slow); rsqrt.approx.ftz.f64 + 2 / 3 Newton steps 1.38 / 1.56 k.
// Warm, the routine is 1.3 k cycles = 8 x (89 for the library rsqrt + 85 for scale -> update -> clamp); the inverse and the
// stores are free (hidden under the chain).  QR_C8_DIAG / QR_C8_RSQRT keep the variants buildable.
// A non-positive pivot is clamped; the caller's verification and finiteness checks catch a breakdown, as in qr_inv3_sym.
#ifndef QR_C8_NEWTON
#define QR_C8_NEWTON 3
#endif
__device__ __forceinline__ double qr_rsqrt_pos(double p) {   // 1/sqrt(p) for a positive normal p: hardware seed + Newton
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double hp = 0.5 * p;
#pragma unroll
    for (int it = 0; it < QR_C8_NEWTON; ++it) {
        const double e = fma(-hp * r, r, 0.5);
        r = fma(r, e, r);
    }
    return r;
}
#ifndef QR_C8_NEWTON
#define QR_C8_NEWTON 3
#endif
#ifndef QR_C8_DIAG
#define QR_C8_DIAG 0
#endif
#ifndef QR_C8_RSQRT
#define QR_C8_RSQRT(p) rsqrt(p)
#endif
__device__ __noinline__ void qr_chol8_diag(double* T) {
#if QR_C8_DIAG == 0
    double a[36], E[36];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            a[i * (i + 1) / 2 + j] = T[qr_k8_swz(i, j)];
            E[i * (i + 1) / 2 + j] = i == j ? 1.0 : 0.0;
        }
    __syncwarp();   // every lane has read the tile before any lane overwrites it
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double p = a[k * (k + 1) / 2 + k];
        if (!(p > 1e-300)) p = 1e-300;
        const double r = QR_C8_RSQRT(p);
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // row k of R (zeros above the diagonal)
            const double v = j < k ? E[k * (k + 1) / 2 + j] * r : (j == k ? r : 0.0);
            if (j <= k) E[k * (k + 1) / 2 + j] = v;
            T[qr_k8_swz(k, j)] = v;
        }
#pragma unroll
        for (int i = k + 1; i < 8; ++i) a[i * (i + 1) / 2 + k] *= r;
#pragma unroll
        for (int i = k + 1; i < 8; ++i) {
#pragma unroll
            for (int j = k + 1; j <= i; ++j) a[i * (i + 1) / 2 + j] -= a[i * (i + 1) / 2 + k] * a[j * (j + 1) / 2 + k];
#pragma unroll
            for (int j = 0; j <= k; ++j) E[i * (i + 1) / 2 + j] -= a[i * (i + 1) / 2 + k] * E[k * (k + 1) / 2 + j];
        }
    }
#else
    double a[36], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = T[qr_k8_swz(i, j)];
    __syncwarp();   // every lane has read the tile before any lane overwrites it
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double p = a[k * (k + 1) / 2 + k];
        if (!(p > 1e-300)) p = 1e-300;
        const double r = QR_C8_RSQRT(p);
        d[k] = r;
#pragma unroll
        for (int i = k + 1; i < 8; ++i) a[i * (i + 1) / 2 + k] *= r;
#pragma unroll
        for (int i = k + 1; i < 8; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) a[i * (i + 1) / 2 + j] -= a[i * (i + 1) / 2 + k] * a[j * (j + 1) / 2 + k];
    }
    const int j = threadIdx.x & 7;
    double col[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int m = 0; m + 1 < i; m += 2) {
            s0 += a[i * (i + 1) / 2 + m] * col[m];
            s1 += a[i * (i + 1) / 2 + m + 1] * col[m + 1];
        }
        if (i & 1) s0 += a[i * (i + 1) / 2 + i - 1] * col[i - 1];
        col[i] = d[i] * ((i == j ? 1.0 : 0.0) - (s0 + s1));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) T[qr_k8_swz(i, j)] = col[i];
#endif
}

// K = L L' in place (tiles below the diagonal: L_Ip; diagonal tiles: R_p = L_pp^-1), fused with the forward
// substitution y <- L^-1 y when with_rhs != 0.  nt tile rows, y has 8*nt entries, tri is the lower-triangular decode
// table of the workspace (QrQpWork::tri).  All NT threads call it; ends with a CTA barrier.
// A warp's tiles are processed four at a time -- operand loads of all four first, then the eight mma.sync -- because a
// single tile is a dependent chain (decode -> fragment loads -> mma -> mma -> store, ~110 cycles) that nothing else hides
// with two warps per scheduler.
struct QrC8Frag { int oa0, oa1, oc; };

__device__ __forceinline__ void qr_chol8_update_tile(const double* col, double* trail, const unsigned short* tri, int nt,
                                                     int p, int ntile, int idx, const QrC8Frag& F, double2& c,
                                                     double& a0, double& a1, double& b0, double& b1) {
    const int code = tri[ntile - 1 - idx];
    const int I = nt - 1 - (code & 255), J = nt - 1 - (code >> 8);
    const double* VI = col + 64 * (I - p);
    const double* VJ = col + 64 * (J - p);
    a0 = -VI[F.oa0]; a1 = -VI[F.oa1]; b0 = VJ[F.oa0]; b1 = VJ[F.oa1];
    c = *reinterpret_cast<const double2*>(trail + 64 * idx + F.oc);
}

// tiles first, first + stride, ... < end of the trailing matrix (A_IJ -= L_Ip L_Jp')
__device__ __forceinline__ void qr_chol8_update_range(const double* col, double* trail, const unsigned short* tri, int nt,
                                                      int p, int ntile, int first, int end, int stride, const QrC8Frag& F) {
    int idx = first;
    for (; idx + 3 * stride < end; idx += 4 * stride) {
        double2 c[4];
        double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) qr_chol8_update_tile(col, trail, tri, nt, p, ntile, idx + u * stride, F, c[u], a0[u], a1[u], b0[u], b1[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) qr_dmma(c[u].x, c[u].y, a0[u], b0[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) qr_dmma(c[u].x, c[u].y, a1[u], b1[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<double2*>(trail + 64 * (idx + u * stride) + F.oc) = c[u];
    }
    for (; idx < end; idx += stride) {
        double2 c;
        double a0, a1, b0, b1;
        qr_chol8_update_tile(col, trail, tri, nt, p, ntile, idx, F, c, a0, a1, b0, b1);
        qr_dmma(c.x, c.y, a0, b0);
        qr_dmma(c.x, c.y, a1, b1);
        *reinterpret_cast<double2*>(trail + 64 * idx + F.oc) = c;
    }
}

template <int NT>
__device__ __forceinline__ void qr_chol8_factor(double* K, double* y, const unsigned short* tri, int nt, int with_rhs) {
    constexpr int NW = NT / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    QrC8Frag F;
    F.oa0 = qr_k8_swz(g, q); F.oa1 = qr_k8_swz(g, 4 + q); F.oc = qr_k8_swz(g, 2 * q);
    if (warp == 0) qr_chol8_diag(K);
    __syncthreads();
    QR_C8P(0);
    for (int p = 0; p < nt; ++p) {
        const int nrem = nt - p - 1;
        double* const col = K + qr_k8_tile(nt, p, p);   // R_p, then the tiles (p+1 .. nt-1, p) at +64 each
        {   // ---- panel: L_Ip = A_Ip R_p'   (B fragment: B[k][n] = R[n][k], lane (k = q, n = g) reads R(g, 4h + q))
            const double b0 = col[F.oa0], b1 = col[F.oa1];
            int t = warp;
            for (; t + NW < nrem; t += 2 * NW) {
                double* W0 = col + 64 * (t + 1);
                double* W1 = col + 64 * (t + NW + 1);
                const double a00 = W0[F.oa0], a01 = W0[F.oa1], a10 = W1[F.oa0], a11 = W1[F.oa1];
                double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
                qr_dmma(c00, c01, a00, b0);
                qr_dmma(c10, c11, a10, b0);
                qr_dmma(c00, c01, a01, b1);
                qr_dmma(c10, c11, a11, b1);
                *reinterpret_cast<double2*>(W0 + F.oc) = make_double2(c00, c01);
                *reinterpret_cast<double2*>(W1 + F.oc) = make_double2(c10, c11);
            }
            for (; t < nrem; t += NW) {
                double* Wt = col + 64 * (t + 1);
                const double a0 = Wt[F.oa0], a1 = Wt[F.oa1];
                double c0 = 0.0, c1 = 0.0;
                qr_dmma(c0, c1, a0, b0);
                qr_dmma(c0, c1, a1, b1);
                *reinterpret_cast<double2*>(Wt + F.oc) = make_double2(c0, c1);
            }
            if (with_rhs && warp == NW - 1) {   // z_p = R_p y_p (R_p has zeros above the diagonal: fixed-length sums)
                double s = 0.0;
                const int r = lane & 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += col[qr_k8_swz(r, j)] * y[8 * p + j];
                __syncwarp();
                if (lane < 8) y[8 * p + lane] = s;
            }
        }
        __syncthreads();
        QR_C8P(1);
        if (nrem == 0) break;
        // ---- trailing update A_IJ -= L_Ip L_Jp' over the tiles in memory order (tile 0 = the next diagonal tile).
        // Work split: warp 0 updates the next diagonal tile and factors it; meanwhile the first BIAS * (NW - 1) other
        // tiles go to warps 1 .. NW-1 only, the rest round-robin over all warps.
        const int ntile = (nrem * (nrem + 1)) >> 1;
        double* const trail = K + qr_k8_tile(nt, p + 1, p + 1);
        constexpr int BIAS = QR_CHOL8_BIAS * (NW - 1);
        const int split = ntile - 1 < BIAS ? ntile : BIAS + 1;   // tiles 1 .. split-1: warps 1 .. NW-1
        if (warp == 0) {
            qr_chol8_update_range(col, trail, tri, nt, p, ntile, 0, 1, 1, F);
            __syncwarp();
            qr_chol8_diag(trail);
            QR_C8P(4);
        } else {
            qr_chol8_update_range(col, trail, tri, nt, p, ntile, warp, split, NW - 1, F);
        }
        qr_chol8_update_range(col, trail, tri, nt, p, ntile, split + warp, ntile, NW, F);
        if (with_rhs) {   // y_I -= L_Ip z_p: one warp pass per tile row, lane (g, q) takes row g, columns 2q, 2q + 1
            const double z0 = y[8 * p + 2 * q], z1 = y[8 * p + 2 * q + 1];
            for (int t = NW - 1 - warp; t < nrem; t += NW) {
                const double2 v = *reinterpret_cast<const double2*>(col + 64 * (t + 1) + F.oc);
                double s = v.x * z0 + v.y * z1;
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (q == 0) y[8 * (p + 1 + t) + g] -= s;
            }
        }
        __syncthreads();
        QR_C8P(2);
    }
}

// x = L'^-1 z after qr_chol8_factor (z in y): x_p = R_p' (z_p - sum_{I > p} L_Ip' x_I), evaluated right-looking -- once
// x_p is known every z_J, J < p, loses L_pJ' x_p (one thread per entry, tile row p).  The first nout entries of x go to
// out.  xs: 8 doubles of scratch.  All NT threads call it; ends with a CTA barrier.  (One barrier per step -- the eight
// threads that finish z_{p-1} exchanging it by shuffle and forming x_{p-1} on the spot -- was measured slower: 18.4 k
// against 14.5 k cycles at 27 tile rows; the shuffles and the divergent branch cost more than the barrier.)
template <int NT>
__device__ __forceinline__ void qr_chol8_backward(const double* K, double* y, double* xs, int nt, double* out, int nout) {
    for (int p = nt - 1; p >= 0; --p) {
        if (threadIdx.x < 8) {
            const int c = threadIdx.x;
            const double* R = K + qr_k8_tile(nt, p, p);
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) s += R[qr_k8_swz(r, c)] * y[8 * p + r];   // R(r, c) = 0 for r < c
            xs[c] = s;
            if (8 * p + c < nout) out[8 * p + c] = s;
        }
        __syncthreads();
        if (p == 0) break;
        for (int t = threadIdx.x; t < 8 * p; t += NT) {
            const int J = t >> 3, c = t & 7;
            const double* L = K + qr_k8_tile(nt, p, J);
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int r = 0; r < 8; r += 2) {
                s0 += L[qr_k8_swz(r, c)] * xs[r];
                s1 += L[qr_k8_swz(r + 1, c)] * xs[r + 1];
            }
            y[t] -= s0 + s1;
        }
        __syncthreads();
    }
    QR_C8P(3);
}

#endif   // __CUDACC__
