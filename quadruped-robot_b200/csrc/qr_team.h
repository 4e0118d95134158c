// qr_team.h -- "one thread team per problem" programming layer.
//
// Every solver routine in csrc/ is written as a sequence of PHASES separated by team barriers.
// Inside a phase a thread touches only its own loop indices; data crosses threads only through
// the workspace (shared memory on the GPU) and only across a barrier.  Compiled by nvcc for sm_100a
// a phase is a strided loop over threadIdx.x and QR_SYNC() is a CTA (or warp) barrier.  Compiled by
// g++ (tests/emul, test infrastructure only) the same source runs the phases one after another with
// a plain loop over the thread index, which lets the control flow and the arithmetic be checked on
// a machine without a GPU.  There is no CPU product path: libqr_gpu.so contains only the nvcc build.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define QR_DEV __device__ __forceinline__
#define QR_DEV_NOINLINE __device__ __noinline__
#define QR_HD __host__ __device__ inline
#else
#define QR_DEV inline
#define QR_DEV_NOINLINE inline
#define QR_HD inline
#endif

#ifndef QR_MULTI_TEAM_NT
#define QR_MULTI_TEAM_NT 64   // team size that may share a CTA with other teams (device: see qr_tid below)
#endif

#if defined(__CUDA_ARCH__)
// ---- device ------------------------------------------------------------------------------
#define QR_ON_DEVICE 1
// Teams of QR_MULTI_TEAM_NT threads (the WBC team size; no MPC size class uses it) may share a CTA: thread t belongs to
// team t / NT, its index in the team is t % NT, and the team's barrier is the named barrier 1 + team.  Several robots per
// CTA run the same code at nearly the same time, so the CTA fetches an instruction line once for all of them -- the WBC
// kernel's body is 20 k instructions executed once per robot, and the instruction caches are what bounds it.
template <int NT>
__device__ __forceinline__ int qr_tid() {
    return NT == QR_MULTI_TEAM_NT ? (int)(threadIdx.x % NT) : (int)threadIdx.x;
}
template <int NT>
__device__ __forceinline__ void qr_team_sync() {
    if (NT <= 32) __syncwarp();
    else if (NT == QR_MULTI_TEAM_NT) asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / NT)), "n"(NT) : "memory");
    // (immediate barrier numbers behind a switch on the team were measured: 6.8 -> 4.6 M robots/s, the four copies of
    // every barrier site inflate a kernel that is bound by instruction fetch)
    else __syncthreads();
}
// (rolled: a team loop runs one to three iterations per thread, and an unrolled copy of its body is only more code for
// kernels that are bound by instruction fetch)
#ifndef QR_FOR_UNROLL
#define QR_FOR_UNROLL _Pragma("unroll 1")
#endif
#define QR_FOR(i, n) QR_FOR_UNROLL for (int i = qr_tid<NT>(); i < (n); i += NT)
// strided loop over the entries (i, j) of an m x n row-major array: the row / column of an entry follow from the previous
// one by additions (one integer division per loop instead of one per entry)
#define QR_FOR_2D(idx, i, j, m, n)                                                                                    \
    QR_FOR_UNROLL for (int idx = qr_tid<NT>(), _qn = (n), _qdi = NT / _qn, _qdj = NT - _qdi * _qn, i = idx / _qn, j = idx - i * _qn; \
         idx < (m) * _qn; idx += NT, i += _qdi, j += _qdj, i += (j >= _qn), j -= (j >= _qn) ? _qn : 0)
#define QR_THREADS(t) for (int t = qr_tid<NT>(), _qr_once = 1; _qr_once; _qr_once = 0)
#define QR_SYNC() qr_team_sync<NT>()
// barrier + "does any thread of the team hold a non-zero flag"
template <int NT>
__device__ __forceinline__ int qr_team_any(int v) {
    if (NT <= 32) { int r = __any_sync(0xffffffffu, v); __syncwarp(); return r; }
    if (NT == QR_MULTI_TEAM_NT) {
        int r;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
                     : "=r"(r) : "r"(v), "r"(1 + (int)(threadIdx.x / NT)), "n"(NT) : "memory");
        return r;
    }
    return __syncthreads_or(v);
}
#define QR_ANY(v) qr_team_any<NT>(v)
#define QR_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
// float32 arithmetic that must not be contracted into FMAs (bit-exact condensing)
#define QR_FMUL(a, b) __fmul_rn((a), (b))
#define QR_FADD(a, b) __fadd_rn((a), (b))
#define QR_FSUB(a, b) __fsub_rn((a), (b))
#define QR_FDIV(a, b) __fdiv_rn((a), (b))
#define QR_DMUL(a, b) __dmul_rn((a), (b))
#define QR_DADD(a, b) __dadd_rn((a), (b))
#else
// ---- host emulation (tests only) ---------------------------------------------------------
#define QR_FOR(i, n) for (int i = 0; i < (n); ++i)
#define QR_FOR_2D(idx, i, j, m, n) \
    for (int idx = 0, _qn = (n), i = 0, j = 0; idx < (m) * _qn; ++idx, ++j, i += (j >= _qn), j -= (j >= _qn) ? _qn : 0)
#define QR_THREADS(t) for (int t = 0; t < NT; ++t)
#define QR_SYNC() ((void)0)
#define QR_ANY(v) (v)
#define QR_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#define QR_FMUL(a, b) ((a) * (b))
#define QR_FADD(a, b) ((a) + (b))
#define QR_FSUB(a, b) ((a) - (b))
#define QR_FDIV(a, b) ((a) / (b))
#define QR_DMUL(a, b) ((a) * (b))
#define QR_DADD(a, b) ((a) + (b))
#endif

// Optional per-phase cycle profile (debug builds only: -DQR_PROFILE).  Thread 0 of every team adds the
// cycles since the previous mark to a global table; marks sit right after barriers.
#if defined(QR_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long qr_prof_table[64];
#endif
#if defined(QR_PROFILE) && defined(__CUDA_ARCH__)
__device__ __forceinline__ void qr_prof_mark(int tag, long long& last) {
    if (threadIdx.x == 0) {
        const long long now = clock64();
        atomicAdd(&qr_prof_table[tag], (unsigned long long)(now - last));
        last = now;
    }
}
#define QR_PROF_DECL long long qr_prof_last = clock64()
#define QR_PROF(tag) qr_prof_mark(tag, qr_prof_last)
#define QR_PROF_ARG , long long& qr_prof_last
#define QR_PROF_PASS , qr_prof_last
#else
#define QR_PROF_DECL
#define QR_PROF(tag) ((void)0)
#define QR_PROF_ARG
#define QR_PROF_PASS
#endif

QR_DEV double qr_min(double a, double b) { return a < b ? a : b; }
QR_DEV double qr_max(double a, double b) { return a > b ? a : b; }
QR_DEV int qr_imin(int a, int b) { return a < b ? a : b; }
QR_DEV int qr_imax(int a, int b) { return a > b ? a : b; }
