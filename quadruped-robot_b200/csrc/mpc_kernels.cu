// mpc_kernels.cu -- CUDA kernels (sm_100a) and the extern "C" layer of libqr_gpu.so (include/qr_gpu.h).
//
// One CTA ("team") per MPC instance, persistent over the batch: grid = resident CTAs of the device,
// each CTA walks the batch with a stride.  Shared memory holds the matrix under factorisation and
// all solver vectors; the symmetric Hessian of the current instance sits in a per-CTA slice of a
// global scratch buffer that never leaves L2.  There is no CPU fallback in this library: every entry
// point fails with QR_ECUDA when no device is usable.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "mpc_problem.h"

#ifndef QR_NT
#define QR_NT 128
#endif

namespace {

// Size classes: workspace capacities in stance foot-steps.  8 .. 72 in steps of 8 keep the matrix under factorisation
// in shared memory; 96 and 128 (long horizons with most legs in stance) keep it in the L2-resident scratch.
constexpr int QR_CLASS_STEP = 8;
constexpr int QR_NCLASS_MAX = 11;
__host__ __device__ inline int qr_class_cap(int c) { return c < 9 ? QR_CLASS_STEP * (c + 1) : (c == 9 ? 96 : 128); }
__host__ __device__ inline int qr_class_of(int nf) {
    if (nf <= 72) { const int c = (nf + QR_CLASS_STEP - 1) / QR_CLASS_STEP - 1; return c < 0 ? 0 : c; }
    return nf <= 96 ? 9 : 10;
}
#ifndef QR_KG_FROM
#define QR_KG_FROM 96
#endif
constexpr int QR_KG_FROM_CAP = QR_KG_FROM;   // smallest capacity whose matrix under factorisation lives in the global scratch

// Size classification: one thread per instance counts its stance foot-steps (gait * f_max > 0, the
// rows SolveMPC would give a non-zero upper bound, qr_mpc_interface.cpp:387) and appends the instance
// to the work list of the smallest workspace class that holds it.
__global__ void qr_mpc_classify_kernel(const QrMpcArgs A, int nclass, int* counts, int* lists) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= A.batch) return;
    const int h4 = 4 * A.P.horizon;
    const float fmax = A.fmax_i ? A.fmax_i[prob] : A.P.f_max;
    const float* g = A.gait + (size_t)prob * h4;
    int nf = 0;
    for (int k = 0; k < h4; ++k) nf += (__fmul_rn(g[k], fmax) > 0.f) ? 1 : 0;
    int c = qr_class_of(nf);
    c = c >= nclass ? nclass - 1 : c;
    const int slot = atomicAdd(&counts[c], 1);
    lists[(size_t)c * A.batch + slot] = prob;
}

// Fused SolveMPCKernel + GetMPCSolution.  Persistent CTAs; instances are handed out through an
// atomic ticket so that the varying number of active-set rounds per instance balances out.
// The workspace capacity (size class) is a template parameter: all shared-memory pointers of the solver
// become compile-time offsets, which removes their re-computation from every inner loop.
// Occupancy.  With both block-packed matrices in shared memory (52.8 KB at capacity 24) and 128 registers the
// kernel stops at 4 CTAs per SM.  Measured on B200 (A1 trot, 65536 instances): Hessian in the L2-resident scratch
// and a 96-register cap -> 5 CTAs per SM, +5 % QP/s (3.11 -> 3.27 M); 80 registers / 6 CTAs: +2.5 % (spills);
// the Hessian in the scratch at unchanged occupancy costs nothing (its reads are three streaming passes per round).
// Team size and resident CTAs per size class.  The classes up to 24 foot-steps (the trot / walk gaits at h = 10) are
// bound by the instruction count of their busiest warp and by the SM's register file: 96 threads x 6 CTAs per SM
// measured +2.9 % QP/s over 128 x 5 (64 x 7: -3 %; 96 x 7 at 80 registers: spills, +-0).  From 32 foot-steps on,
// shared memory alone limits the SM to 4, 3, 2 and then 1 CTA: no register cap is needed there, and the classes left
// with one CTA per SM get a 256-thread team (the early steps of their factorisations have hundreds of tiles).
#ifndef QR_NT_SMALL
#define QR_NT_SMALL 96
#endif
#ifndef QR_CTAS_SMALL
#define QR_CTAS_SMALL 6
#endif
#ifndef QR_NT_LARGE
#define QR_NT_LARGE 256
#endif
#ifndef QR_CTAS_LARGE
#define QR_CTAS_LARGE 1   // resident CTAs per SM asked of the compiler for the 56..72 foot-step classes
#endif
#ifndef QR_HSG_FROM
#define QR_HSG_FROM 8     // smallest capacity that keeps the Hessian in the global scratch (8: every class)
#endif
__host__ __device__ constexpr int qr_fused_nt(int cap) {
    return cap <= 24 ? QR_NT_SMALL : (cap >= 56 && cap <= 72 ? QR_NT_LARGE : QR_NT);
}
__host__ __device__ constexpr int qr_fused_min_ctas(int cap) {
    return cap <= 24 ? QR_CTAS_SMALL : (cap == 32 ? 4 : (cap == 40 ? 3 : (cap == 48 ? 2 : (cap <= 72 ? QR_CTAS_LARGE : 5))));
}
template <int CAP, bool HSG, bool KG>
__global__ void __launch_bounds__(qr_fused_nt(CAP), qr_fused_min_ctas(CAP)) qr_mpc_fused_kernel(const QrMpcArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ticket;
    constexpr int NT = qr_fused_nt(CAP);
    QrMpcSmem S;
    qr_mpc_carve(S, smem, CAP, A.P.horizon, A.scratch + (size_t)blockIdx.x * qr_fallback_doubles(CAP),
                 HSG ? A.hs_global + (size_t)blockIdx.x * 9 * qr_ntri(CAP) : nullptr,
                 KG ? A.k_global + (size_t)blockIdx.x * 9 * qr_ntri(CAP) : nullptr);
    S.Hc = A.hc_global ? A.hc_global + (size_t)blockIdx.x * 9 * qr_ntri(CAP) : nullptr;
    S.Hc2 = (S.Hc && qr_coarse2_cap(CAP) > 0) ? S.Hc + 9 * qr_ntri(qr_coarse_cap(CAP)) : nullptr;   // ntri(3c/4) + ntri(9c/16) < ntri(c)
    qr_mpc_init_tables<NT>(S, CAP);
    // A.next == null (small batches): one launch, instances strided over the grid, no work lists.
    const int total = A.count ? *A.count : A.batch;
    int strided = blockIdx.x;
    for (;;) {
        if (A.next && threadIdx.x == 0) s_ticket = atomicAdd(A.next, 1);
        __syncthreads();
        const int k = A.next ? s_ticket : strided;
        __syncthreads();
        if (k >= total) break;
        strided += gridDim.x;
        qr_mpc_solve_problem<NT, (CAP >= QR_COARSE2_MIN_CAP)>(A, A.list ? A.list[k] : k, S);
    }
}

// Latency path (batches that cannot fill the device): one instantiation with the capacity as a runtime value, both
// matrices in shared memory whenever they fit and no register cap -- per-instance latency matters here, not occupancy.
#ifndef QR_COARSE_MAX_ROUNDS_LAT
#define QR_COARSE_MAX_ROUNDS_LAT 4   // coarse-round cap of the latency path
#endif
#ifndef QR_LAT_NT
#define QR_LAT_NT 256   // a wider team shortens the phases with plenty of parallel work (condensing, early LDL' steps)
#endif
template <bool L2>
__global__ void __launch_bounds__(QR_LAT_NT, 1) qr_mpc_fused_latency_kernel(const QrMpcArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = QR_LAT_NT;
    QrMpcSmem S;
    qr_mpc_carve(S, smem, A.nfcap, A.P.horizon, A.scratch + (size_t)blockIdx.x * qr_fallback_doubles(A.nfcap),
                 A.hs_global ? A.hs_global + (size_t)blockIdx.x * 9 * qr_ntri(A.nfcap) : nullptr,
                 A.k_global ? A.k_global + (size_t)blockIdx.x * 9 * qr_ntri(A.nfcap) : nullptr);
    S.Hc = A.hc_global ? A.hc_global + (size_t)blockIdx.x * 9 * qr_ntri(A.nfcap) : nullptr;
    S.Hc2 = (S.Hc && qr_coarse2_cap(A.nfcap) > 0) ? S.Hc + 9 * qr_ntri(qr_coarse_cap(A.nfcap)) : nullptr;
    qr_mpc_init_tables<NT>(S, A.nfcap);
    for (int prob = blockIdx.x; prob < A.batch; prob += gridDim.x) qr_mpc_solve_problem<NT, L2>(A, prob, S);
}

typedef void (*QrFusedKernel)(const QrMpcArgs);
// Instantiated size classes (capacities in stance foot-steps); the Hessian moves to the L2-resident scratch
// for the classes that do not fit two block-packed matrices in 227 KB of shared memory.
constexpr int QR_HSG_FROM_CAP = QR_HSG_FROM;
QrFusedKernel fused_kernel_for(int cap) {
    switch (cap) {
#define QR_CASE(C) case C: return qr_mpc_fused_kernel<C, (C >= QR_HSG_FROM), (C >= QR_KG_FROM_CAP)>;
#ifdef QR_EXP_ONLY_CAP   // experiment builds (tools/exp_bench.py): one size class, seconds to compile
        QR_CASE(QR_EXP_ONLY_CAP)
#else
        QR_CASE(8) QR_CASE(16) QR_CASE(24) QR_CASE(32) QR_CASE(40) QR_CASE(48) QR_CASE(56) QR_CASE(64) QR_CASE(72)
        QR_CASE(96) QR_CASE(128)
#endif
#undef QR_CASE
        default: return nullptr;
    }
}

__global__ void __launch_bounds__(QR_NT) qr_mpc_condense_kernel(const QrMpcArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = QR_NT;
    QrMpcSmem S;
    qr_mpc_carve(S, smem, A.nfcap, A.P.horizon, nullptr);
    for (int prob = blockIdx.x; prob < A.batch; prob += gridDim.x)
        qr_mpc_condense_problem<NT>(A, prob, S);
}

__global__ void __launch_bounds__(QR_NT) qr_qp_solve_kernel(const QrMpcArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = QR_NT;
    QrMpcSmem S;
    qr_mpc_carve(S, smem, A.nfcap, A.P.horizon, A.scratch + (size_t)blockIdx.x * qr_fallback_doubles(A.nfcap),
                 A.hs_global ? A.hs_global + (size_t)blockIdx.x * 9 * qr_ntri(A.nfcap) : nullptr,
                 A.k_global ? A.k_global + (size_t)blockIdx.x * 9 * qr_ntri(A.nfcap) : nullptr);
    S.Hc = A.hc_global ? A.hc_global + (size_t)blockIdx.x * 9 * qr_ntri(A.nfcap) : nullptr;
    S.Hc2 = (S.Hc && qr_coarse2_cap(A.nfcap) > 0) ? S.Hc + 9 * qr_ntri(qr_coarse_cap(A.nfcap)) : nullptr;
    qr_mpc_init_tables<NT>(S, A.nfcap);
    for (int prob = blockIdx.x; prob < A.batch; prob += gridDim.x)
        qr_qp_solve_problem<NT>(A, prob, S);
}

// ------------------------------------------------------------------------------------------
// host-side contexts: one per initialised device
// ------------------------------------------------------------------------------------------
// Threading and stream contract (include/qr_gpu.h): every entry point runs on the calling thread's CURRENT device,
// which must have been initialised by qr_gpu_init; device-pointer calls are asynchronous on the caller's stream and
// may be issued concurrently from several host threads and on several streams.  What makes that safe:
//  * workspaces ("lanes": the L2-resident scratch slices, the size-class counters, tickets and work lists of one
//    launch sequence) come from a per-device pool keyed by stream.  A lane is handed to the stream that used it last
//    (stream order protects it), else an idle lane is taken (its completion event has fired), else a new one is
//    created, and only when the pool is full does the stream wait on the least recently used lane's event;
//  * the *_host entry points take a private staging slot (device staging buffer, pinned mirror, two streams) for
//    the whole call, so two host threads never share a buffer; error paths drain the slot's streams before it is
//    released, so no copy from caller memory is left pending;
//  * a per-device mutex serialises only the bookkeeping and the launch calls themselves, never device execution.
constexpr int QR_MAX_DEVICES = 16;
constexpr int QR_MAX_LANES = 8;
constexpr int QR_MAX_SLOTS = 4;
constexpr int QR_MAX_WBC_MODELS = 4;

struct Plan {
    int grid = 0, occ = 0;
    size_t smem = 0;
    bool hsg = false, kg = false;
    size_t scratch_doubles(int nfcap) const {   // per launch: [grid][fallback] [grid][Hc] [grid][Hs] [grid][K]
        size_t d = (size_t)grid * qr_fallback_doubles(nfcap);
        d += (size_t)grid * 9 * qr_ntri(nfcap);   // coarse Hessian of the active-set prediction
        if (hsg) d += (size_t)grid * 9 * qr_ntri(nfcap);
        if (kg) d += (size_t)grid * 9 * qr_ntri(nfcap);
        return d;
    }
};
struct GeomEntry { const void* kern; int nfcap, horizon; bool want_hsg, want_kg; Plan plan; };

struct Lane {
    double* scratch = nullptr;
    size_t scratch_bytes = 0;
    int* work = nullptr;            // [2*nclass] counters (counts, tickets) followed by [nclass][batch] lists
    size_t work_bytes = 0;
    cudaEvent_t done = nullptr;     // recorded behind the last launch that used the lane
    cudaStream_t last = nullptr;    // stream of that launch
    bool used = false;              // `done` has been recorded at least once
    unsigned long long stamp = 0;   // least-recently-used order
};

struct HostSlot {
    bool created = false, busy = false;
    unsigned char* stage = nullptr;   // device staging buffer
    size_t stage_bytes = 0;
    unsigned char* pin = nullptr;     // pinned host mirror (small batches: one copy each way)
    size_t pin_bytes = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
};

struct WbcModelSlot {
    bool valid = false;
    qr_wbc_model key;
    void* dev = nullptr;              // QrWbcModelDev on this device
    unsigned long long stamp = 0;
};

struct Ctx {
    bool ready = false;
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    Lane lanes[QR_MAX_LANES];
    int nlanes = 0;
    unsigned long long clock = 0;
    HostSlot slots[QR_MAX_SLOTS];
    GeomEntry geom[96];
    int ngeom = 0;
    WbcModelSlot wbc_models[QR_MAX_WBC_MODELS];
    int wbc_occ = 0;                  // resident CTAs per SM of the WBC kernel (0: not queried yet)
};
Ctx g_dev[QR_MAX_DEVICES];
std::mutex g_mu;                                   // init / shutdown (the table of contexts)
std::mutex g_ctx_mu[QR_MAX_DEVICES];               // one per context: its lanes, plans, slots, WBC models -- calls on
std::condition_variable g_slot_cv[QR_MAX_DEVICES]; // different devices never contend (the multi-GPU call's threads)
std::mutex& ctx_mu(const Ctx& cx) { return g_ctx_mu[cx.device]; }
thread_local char t_err[256] = {0};

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(t_err, sizeof(t_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(t_err, sizeof(t_err), "%s", what);
    return code;
}

// Context of the calling thread's current device (null + error when that device was not initialised).
Ctx* current_ctx() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= QR_MAX_DEVICES || !g_dev[dev].ready) {
        fail(QR_ECUDA, "qr_gpu_init was not called for the current device (or no CUDA device)");
        return nullptr;
    }
    return &g_dev[dev];
}

qr_qp_options default_options() {
    qr_qp_options o;
    o.max_as_rounds = 32;
    o.max_ipm_iter = 40;
    o.max_polish_rounds = 12; o.flags = 0;
    o.ipm_tol = 1e-7;
    o.act_kappa = 1e3;
    o.feas_tol = 1e-9;
    o.mult_tol = 1e-11;
    return o;
}

// Launch plan of one kernel for one workspace capacity: where the two block-packed matrices live, dynamic shared
// memory, resident CTAs per SM.  want_hsg / want_kg are requests (true: keep that matrix in the global scratch) and
// are forced to true when the workspace would not fit in shared memory otherwise.
template <typename Kern>
int launch_geometry(Ctx& cx, Kern kern, int nfcap, int horizon, int batch, Plan* out, bool want_hsg = false,
                    bool want_kg = false, int nt = QR_NT) {
    if (!kern) return fail(QR_EINVAL, "no kernel instantiated for this size class");
    Plan pl;
    bool found = false;
    // attribute + occupancy queries cost microseconds each: remember them per (kernel, class, horizon, request)
    for (int i = 0; i < cx.ngeom && !found; ++i) {
        const GeomEntry& ge = cx.geom[i];
        if (ge.kern == (const void*)kern && ge.nfcap == nfcap && ge.horizon == horizon && ge.want_hsg == want_hsg &&
            ge.want_kg == want_kg) {
            pl = ge.plan;
            found = true;
        }
    }
    if (!found) {
        pl.hsg = want_hsg;
        pl.kg = want_kg;
        size_t bytes = qr_mpc_smem_bytes(nfcap, horizon, !pl.hsg, !pl.kg);
        if (bytes + 1024 > cx.smem_optin && !pl.hsg) { pl.hsg = true; bytes = qr_mpc_smem_bytes(nfcap, horizon, false, !pl.kg); }
        if (bytes + 1024 > cx.smem_optin && !pl.kg) { pl.kg = true; bytes = qr_mpc_smem_bytes(nfcap, horizon, false, false); }
        if (bytes + 1024 > cx.smem_optin) return fail(QR_EINVAL, "horizon needs more shared memory than one CTA may use");
        // opt in to the device maximum once per kernel (the launch's own byte count decides the occupancy)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cx.smem_optin - 256);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaFuncSetAttribute", e);
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, bytes);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaOccupancyMaxActiveBlocksPerMultiprocessor", e);
        pl.occ = occ < 1 ? 1 : occ;
        pl.smem = bytes;
        if (cx.ngeom < 96) cx.geom[cx.ngeom++] = GeomEntry{(const void*)kern, nfcap, horizon, want_hsg, want_kg, pl};
    }
    int g = cx.sm_count * pl.occ;
    if (g > batch) g = batch;
    if (g < 1) g = 1;
    pl.grid = g;
    *out = pl;
    return QR_OK;
}

// Workspace for a launch sequence on stream `st` (the context mutex held).  See the contract at the top of this section.
int acquire_lane(Ctx& cx, cudaStream_t st, Lane** out) {
    Lane* pick = nullptr;
    for (int i = 0; i < cx.nlanes && !pick; ++i)
        if (cx.lanes[i].used && cx.lanes[i].last == st) pick = &cx.lanes[i];
    for (int i = 0; i < cx.nlanes && !pick; ++i) {
        Lane& l = cx.lanes[i];
        if (!l.used) { pick = &l; break; }
        cudaError_t q = cudaEventQuery(l.done);
        if (q == cudaSuccess) pick = &l;
        else if (q != cudaErrorNotReady) return fail(QR_ECUDA, "cudaEventQuery(lane)", q);
    }
    if (!pick && cx.nlanes < QR_MAX_LANES) {
        Lane& l = cx.lanes[cx.nlanes];
        cudaError_t e = cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaEventCreate(lane)", e);
        ++cx.nlanes;
        pick = &l;
    }
    if (!pick) {
        pick = &cx.lanes[0];
        for (int i = 1; i < cx.nlanes; ++i)
            if (cx.lanes[i].stamp < pick->stamp) pick = &cx.lanes[i];
        cudaError_t e = cudaStreamWaitEvent(st, pick->done, 0);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaStreamWaitEvent(lane)", e);
    }
    pick->last = st;
    pick->stamp = ++cx.clock;
    *out = pick;
    return QR_OK;
}

// Mark the end of the lane's use on the stream (also on launch errors: whatever was enqueued still owns the lane).
void release_lane(Lane& l, cudaStream_t st) {
    if (cudaEventRecord(l.done, st) == cudaSuccess) l.used = true;
}

int ensure_scratch_doubles(Lane& l, size_t doubles) {
    const size_t need = doubles * sizeof(double);
    if (need <= l.scratch_bytes) return QR_OK;
    if (l.scratch) cudaFree(l.scratch);   // cudaFree waits for the device: nothing still reads the old buffer
    l.scratch = nullptr;
    l.scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&l.scratch, need);
    if (e != cudaSuccess) return fail(QR_ENOMEM, "cudaMalloc(scratch)", e);
    l.scratch_bytes = need;
    return QR_OK;
}

int ensure_work(Lane& l, int nclass, int batch) {
    const size_t need = ((size_t)2 * nclass + (size_t)nclass * batch) * sizeof(int);
    if (need <= l.work_bytes) return QR_OK;
    if (l.work) cudaFree(l.work);
    l.work = nullptr;
    l.work_bytes = 0;
    cudaError_t e = cudaMalloc(&l.work, need);
    if (e != cudaSuccess) return fail(QR_ENOMEM, "cudaMalloc(work lists)", e);
    l.work_bytes = need;
    return QR_OK;
}

// Point the kernel arguments at this launch's slices of the lane's scratch.
void bind_scratch(QrMpcArgs& A, const Plan& pl, int nfcap, Lane& l) {
    double* s = l.scratch;
    A.scratch = s;
    s += (size_t)pl.grid * qr_fallback_doubles(nfcap);
    A.hc_global = s;
    s += (size_t)pl.grid * 9 * qr_ntri(nfcap);
    A.hs_global = nullptr;
    A.k_global = nullptr;
    if (pl.hsg) { A.hs_global = s; s += (size_t)pl.grid * 9 * qr_ntri(nfcap); }
    if (pl.kg) A.k_global = s;
}

int check_params(const qr_mpc_params* P, int batch) {
    if (!P || batch < 0) return fail(QR_EINVAL, "null params or negative batch");
    if (P->horizon < 1 || P->horizon > QR_MAX_HORIZON) return fail(QR_EINVAL, "horizon out of range");
    if (!(P->dt > 0.f) || !(P->mass > 0.f) || !(P->mu > 0.f)) return fail(QR_EINVAL, "dt, mass and mu must be positive");
    return QR_OK;
}

// A staging slot of the *_host entry points, held for one call.  The destructor drains the slot's streams (so that
// no copy from or to caller memory is pending when the call returns, error paths included) and hands the slot back.
struct SlotHold {
    Ctx* cx = nullptr;
    HostSlot* s = nullptr;
    ~SlotHold() {
        if (!s) return;
        cudaStreamSynchronize(s->stream[0]);
        cudaStreamSynchronize(s->stream[1]);
        {
            std::lock_guard<std::mutex> lk(ctx_mu(*cx));
            s->busy = false;
        }
        g_slot_cv[cx->device].notify_one();
    }
};

// Take a free staging slot of the device (waits when all QR_MAX_SLOTS are in use) with at least `bytes` of device
// staging and, when pinned_bytes > 0, a pinned mirror of that size.
int acquire_slot(Ctx& cx, size_t bytes, size_t pinned_bytes, SlotHold& hold) {
    std::unique_lock<std::mutex> lk(ctx_mu(cx));
    HostSlot* s = nullptr;
    for (;;) {
        for (int i = 0; i < QR_MAX_SLOTS && !s; ++i)
            if (cx.slots[i].created && !cx.slots[i].busy) s = &cx.slots[i];
        for (int i = 0; i < QR_MAX_SLOTS && !s; ++i)
            if (!cx.slots[i].created) s = &cx.slots[i];
        if (s) break;
        g_slot_cv[cx.device].wait(lk);
    }
    if (!s->created) {
        for (int k = 0; k < 2; ++k) {
            cudaError_t e = cudaStreamCreateWithFlags(&s->stream[k], cudaStreamNonBlocking);
            if (e != cudaSuccess) return fail(QR_ECUDA, "cudaStreamCreate", e);
        }
        s->created = true;
    }
    s->busy = true;
    hold.cx = &cx;
    hold.s = s;
    if (bytes > s->stage_bytes) {
        if (s->stage) cudaFree(s->stage);
        s->stage = nullptr;
        s->stage_bytes = 0;
        cudaError_t e = cudaMalloc(&s->stage, bytes);
        if (e != cudaSuccess) return fail(QR_ENOMEM, "cudaMalloc(stage)", e);
        s->stage_bytes = bytes;
    }
    if (pinned_bytes > s->pin_bytes) {
        if (s->pin) cudaFreeHost(s->pin);
        s->pin = nullptr;
        s->pin_bytes = 0;
        cudaError_t e = cudaMallocHost(&s->pin, pinned_bytes);
        if (e != cudaSuccess) return fail(QR_ENOMEM, "cudaMallocHost(pinned stage)", e);
        s->pin_bytes = pinned_bytes;
    }
    return QR_OK;
}

void free_ctx(Ctx& cx) {
    if (!cx.ready) return;
    cudaSetDevice(cx.device);
    cudaDeviceSynchronize();
    for (int l = 0; l < cx.nlanes; ++l) {
        if (cx.lanes[l].scratch) cudaFree(cx.lanes[l].scratch);
        if (cx.lanes[l].work) cudaFree(cx.lanes[l].work);
        if (cx.lanes[l].done) cudaEventDestroy(cx.lanes[l].done);
    }
    for (HostSlot& s : cx.slots) {
        if (s.stage) cudaFree(s.stage);
        if (s.pin) cudaFreeHost(s.pin);
        for (int k = 0; k < 2; ++k)
            if (s.stream[k]) cudaStreamDestroy(s.stream[k]);
    }
    for (WbcModelSlot& m : cx.wbc_models)
        if (m.dev) cudaFree(m.dev);
    cx = Ctx();
}

}  // namespace

extern "C" const char* qr_gpu_last_error(void) { return t_err; }

extern "C" int qr_gpu_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(QR_ECUDA, "no CUDA device", e);
    if (device >= 0) {
        e = cudaSetDevice(device);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaSetDevice", e);
    }
    int dev = -1;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaGetDevice", e);
    if (dev < 0 || dev >= QR_MAX_DEVICES) return fail(QR_EINVAL, "device index out of range");
    Ctx& cx = g_dev[dev];
    if (!cx.ready) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaGetDeviceProperties", e);
        if (prop.major < 10) return fail(QR_ECUDA, "libqr_gpu.so is built for sm_100a only");
        cx.device = dev;
        cx.sm_count = prop.multiProcessorCount;
        cx.smem_optin = prop.sharedMemPerBlockOptin;
        cx.ready = true;
    }
    t_err[0] = 0;
    return QR_OK;
}

namespace { void workers_stop(); }
extern "C" void qr_gpu_shutdown(void) {
    workers_stop();
    std::lock_guard<std::mutex> lk(g_mu);
    int prev = -1;
    cudaGetDevice(&prev);
    for (Ctx& cx : g_dev) free_ctx(cx);
    if (prev >= 0) cudaSetDevice(prev);
}

namespace {
int num_classes(int horizon) { return qr_class_of(4 * horizon) + 1; }
int class_cap(int c, int /*horizon*/) { return qr_class_cap(c); }   // always one of the instantiated capacities
bool class_hsg(int cap) { return cap >= QR_HSG_FROM_CAP; }
bool class_kg(int cap) { return cap >= QR_KG_FROM_CAP; }
}  // namespace

extern "C" int qr_gpu_mpc_occupancy(int horizon, int stance_footsteps, int* sm_count, int* ctas_per_sm,
                                    int* threads_per_cta, int* smem_bytes) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (horizon < 1 || horizon > QR_MAX_HORIZON) return fail(QR_EINVAL, "horizon out of range");
    if (stance_footsteps < 0 || stance_footsteps > 4 * horizon) return fail(QR_EINVAL, "stance count out of range");
    const int cap = class_cap(qr_class_of(stance_footsteps), horizon);
    Plan pl;
    int rc = launch_geometry(*cx, fused_kernel_for(cap), cap, horizon, 1 << 30, &pl, class_hsg(cap), class_kg(cap), qr_fused_nt(cap));
    if (rc) return rc;
    if (sm_count) *sm_count = cx->sm_count;
    if (ctas_per_sm) *ctas_per_sm = pl.occ;
    if (threads_per_cta) *threads_per_cta = qr_fused_nt(cap);
    if (smem_bytes) *smem_bytes = (int)pl.smem;
    return QR_OK;
}

namespace {
// Enqueue the fused MPC solve of one batch on `st` (the context mutex held by the caller).
int mpc_enqueue(Ctx& cx, const qr_mpc_params* P, const qr_qp_options* opt, int batch, const float* p, const float* v,
                const float* quat, const float* w, const float* r_feet, const float* rpy, const float* traj,
                const float* gait, const float* mu_i, const float* fmax_i, float* grf_out, float* u_out,
                int32_t* status_out, int32_t* iters_out, cudaStream_t st, const qr_mpc_epilogue* ep = nullptr) {
    int rc = check_params(P, batch);
    if (rc) return rc;
    if (batch == 0) return QR_OK;
    if (!p || !v || !quat || !w || !r_feet || !rpy || !traj || !gait || !grf_out)
        return fail(QR_EINVAL, "null input/output pointer");
    const int h = P->horizon, nclass = num_classes(h);
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P;
    A.opt = opt ? *opt : default_options();
    A.batch = batch;
    A.p = p; A.v = v; A.quat = quat; A.w = w; A.r_feet = r_feet; A.rpy = rpy; A.traj = traj; A.gait = gait;
    A.mu_i = mu_i; A.fmax_i = fmax_i;
    A.grf_out = grf_out; A.u_out = u_out; A.status_out = status_out; A.iters_out = iters_out;
    if (ep) {
        if ((ep->f_ff_out || ep->tau_out) && !ep->q) return fail(QR_EINVAL, "epilogue: leg forces / torques need the motor angles");
        A.ep_q = ep->q; A.ep_ff = ep->f_ff_out; A.ep_tau = ep->tau_out; A.ep_cmd = ep->wbc_cmd_io;
        A.ep_hip = ep->hip_len; A.ep_upper = ep->upper_len; A.ep_lower = ep->lower_len;
    }
    Lane* lane = nullptr;
    if (batch <= cx.sm_count) {
        // Latency path: the grid cannot fill the device anyway, so skip the classification and launch the
        // largest size class once (its workspace holds any instance of this horizon).
        const int cap = class_cap(nclass - 1, h);
        Plan pl;
        const bool two_levels = qr_coarse2_cap(cap) > 0;
        rc = two_levels ? launch_geometry(cx, qr_mpc_fused_latency_kernel<true>, cap, h, batch, &pl, false, false, QR_LAT_NT)
                        : launch_geometry(cx, qr_mpc_fused_latency_kernel<false>, cap, h, batch, &pl, false, false, QR_LAT_NT);
        if (rc) return rc;
        rc = acquire_lane(cx, st, &lane);
        if (rc) return rc;
        rc = ensure_scratch_doubles(*lane, pl.scratch_doubles(cap));
        if (rc) return rc;
        A.nfcap = cap;
        A.coarse_rounds = two_levels ? 0 : QR_COARSE_MAX_ROUNDS_LAT;   // two-level classes: their own per-level caps
        bind_scratch(A, pl, cap, *lane);
        if (two_levels) qr_mpc_fused_latency_kernel<true><<<pl.grid, QR_LAT_NT, pl.smem, st>>>(A);
        else qr_mpc_fused_latency_kernel<false><<<pl.grid, QR_LAT_NT, pl.smem, st>>>(A);
        cudaError_t e1 = cudaGetLastError();
        release_lane(*lane, st);
        if (e1 != cudaSuccess) return fail(QR_ECUDA, "launch qr_mpc_fused_latency_kernel", e1);
        return QR_OK;
    }
    // plan of every class first (so that the scratch is sized once, before any launch)
    Plan plan[QR_NCLASS_MAX];
    size_t scratch_need = 0;
    for (int c = 0; c < nclass; ++c) {
        const int cap = class_cap(c, h);
#ifdef QR_EXP_ONLY_CAP
        if (cap != QR_EXP_ONLY_CAP) continue;
#endif
        rc = launch_geometry(cx, fused_kernel_for(cap), cap, h, batch, &plan[c], class_hsg(cap), class_kg(cap), qr_fused_nt(cap));
        if (rc) return rc;
        if (plan[c].hsg != class_hsg(cap) || plan[c].kg != class_kg(cap))
            return fail(QR_EINVAL, "unexpected shared-memory capacity for this size class");
        const size_t need = plan[c].scratch_doubles(cap);
        if (need > scratch_need) scratch_need = need;
    }
    rc = acquire_lane(cx, st, &lane);
    if (rc) return rc;
    rc = ensure_work(*lane, nclass, batch);
    if (rc) return rc;
    rc = ensure_scratch_doubles(*lane, scratch_need);
    if (rc) return rc;
    int* counts = lane->work;
    int* tickets = lane->work + nclass;
    int* lists = lane->work + 2 * nclass;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)2 * nclass * sizeof(int), st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemsetAsync(work counters)", e);
    qr_mpc_classify_kernel<<<(batch + 255) / 256, 256, 0, st>>>(A, nclass, counts, lists);
    e = cudaGetLastError();
    // largest workspaces first: their instances take longest.  The classes share the lane's scratch; they run one
    // after the other on the stream.
    for (int c = nclass - 1; c >= 0 && e == cudaSuccess; --c) {
        A.nfcap = class_cap(c, h);
#ifdef QR_EXP_ONLY_CAP
        if (A.nfcap != QR_EXP_ONLY_CAP) continue;
#endif
        bind_scratch(A, plan[c], A.nfcap, *lane);
        A.list = lists + (size_t)c * batch;
        A.count = counts + c;
        A.next = tickets + c;
        fused_kernel_for(A.nfcap)<<<plan[c].grid, qr_fused_nt(A.nfcap), plan[c].smem, st>>>(A);
        e = cudaGetLastError();
    }
    release_lane(*lane, st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_mpc_classify_kernel / qr_mpc_fused_kernel", e);
    return QR_OK;
}
}  // namespace

extern "C" int qr_gpu_mpc_solve_batch(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                                      const float* p, const float* v, const float* quat,
                                      const float* w, const float* r_feet, const float* rpy,
                                      const float* traj, const float* gait, const float* mu_i,
                                      const float* fmax_i, float* grf_out, float* u_out,
                                      int32_t* status_out, int32_t* iters_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    return mpc_enqueue(*cx, P, opt, batch, p, v, quat, w, r_feet, rpy, traj, gait, mu_i, fmax_i, grf_out, u_out,
                       status_out, iters_out, (cudaStream_t)cuda_stream);
}

extern "C" int qr_gpu_mpc_solve_batch_ex(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                                         const float* p, const float* v, const float* quat,
                                         const float* w, const float* r_feet, const float* rpy,
                                         const float* traj, const float* gait, const float* mu_i,
                                         const float* fmax_i, float* grf_out, float* u_out,
                                         int32_t* status_out, int32_t* iters_out, const qr_mpc_epilogue* epilogue,
                                         void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    return mpc_enqueue(*cx, P, opt, batch, p, v, quat, w, r_feet, rpy, traj, gait, mu_i, fmax_i, grf_out, u_out,
                       status_out, iters_out, (cudaStream_t)cuda_stream, epilogue);
}

extern "C" int qr_gpu_mpc_condense_batch(const qr_mpc_params* P, int batch, const float* p,
                                         const float* v, const float* quat, const float* w,
                                         const float* r_feet, const float* rpy, const float* traj,
                                         const float* gait, const float* fmax_i, float* H_out,
                                         float* g_out, float* ub_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    int rc = check_params(P, batch);
    if (rc) return rc;
    if (batch == 0) return QR_OK;
    if (!p || !v || !quat || !w || !r_feet || !rpy || !traj || !gait || !H_out || !g_out || !ub_out)
        return fail(QR_EINVAL, "null input/output pointer");
    // the condense-only kernel needs the tables and the staged rows, not the QP workspace (no lane)
    Plan pl;
    rc = launch_geometry(*cx, qr_mpc_condense_kernel, QR_CLASS_STEP, P->horizon, batch, &pl);
    if (rc) return rc;
    if (pl.hsg || pl.kg) return fail(QR_EINVAL, "workspace does not fit in shared memory");
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P;
    A.opt = default_options();
    A.batch = batch;
    A.nfcap = QR_CLASS_STEP;
    A.p = p; A.v = v; A.quat = quat; A.w = w; A.r_feet = r_feet; A.rpy = rpy; A.traj = traj; A.gait = gait;
    A.fmax_i = fmax_i;
    A.H_out = H_out; A.g_out = g_out; A.ub_out = ub_out;
    qr_mpc_condense_kernel<<<pl.grid, QR_NT, pl.smem, (cudaStream_t)cuda_stream>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_mpc_condense_kernel", e);
    return QR_OK;
}

extern "C" int qr_gpu_qp_solve_batch(int horizon, float mu, const qr_qp_options* opt, int batch,
                                     const float* H, const float* g, const float* ub,
                                     const float* mu_i, float* x_out, double* x_out_f64,
                                     int32_t* status_out, int32_t* iters_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    qr_mpc_params P;
    memset(&P, 0, sizeof(P));
    P.horizon = horizon; P.mu = mu; P.dt = 1.f; P.mass = 1.f;
    int rc = check_params(&P, batch);
    if (rc) return rc;
    if (batch == 0) return QR_OK;
    if (!H || !g || !ub || (!x_out && !x_out_f64)) return fail(QR_EINVAL, "null input/output pointer");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Plan pl;
    rc = launch_geometry(*cx, qr_qp_solve_kernel, 4 * horizon, horizon, batch, &pl);
    if (rc) return rc;
    Lane* lane = nullptr;
    rc = acquire_lane(*cx, st, &lane);
    if (rc) return rc;
    rc = ensure_scratch_doubles(*lane, pl.scratch_doubles(4 * horizon));
    if (rc) return rc;
    QrMpcArgs A;
    memset(&A, 0, sizeof(A));
    A.P = P;
    A.opt = opt ? *opt : default_options();
    A.batch = batch;
    A.nfcap = 4 * horizon;
    A.mu_i = mu_i;
    A.H_in = H; A.g_in = g; A.ub_in = ub;
    A.x_out = x_out; A.x_out_f64 = x_out_f64; A.status_out = status_out; A.iters_out = iters_out;
    bind_scratch(A, pl, A.nfcap, *lane);
    qr_qp_solve_kernel<<<pl.grid, QR_NT, pl.smem, st>>>(A);
    cudaError_t e = cudaGetLastError();
    release_lane(*lane, st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_qp_solve_kernel", e);
    return QR_OK;
}

extern "C" int qr_gpu_mpc_solve_batch_host(const qr_mpc_params* P, const qr_qp_options* opt, int batch,
                                           const float* p, const float* v, const float* quat,
                                           const float* w, const float* r_feet, const float* rpy,
                                           const float* traj, const float* gait, const float* mu_i,
                                           const float* fmax_i, float* grf_out, float* u_out,
                                           int32_t* status_out, int32_t* iters_out) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    {
        int rc = check_params(P, batch);
        if (rc) return rc;
    }
    if (batch == 0) return QR_OK;
    if (!p || !v || !quat || !w || !r_feet || !rpy || !traj || !gait || !grf_out)
        return fail(QR_EINVAL, "null input/output pointer");
    const int h = P->horizon;
    const size_t B = (size_t)batch;
    // device staging layout (floats unless noted)
    const size_t n_in = B * (3 + 3 + 4 + 3 + 12 + 3 + 12 * h + 4 * h + 2);
    const size_t n_out = B * (12 + (u_out ? 12 * h : 0));
    const size_t bytes = (n_in + n_out) * sizeof(float) + B * 3 * sizeof(int32_t) + 256;
    const bool packed = bytes <= (size_t)256 * 1024;
    SlotHold hold;   // private to this call; drained and released on every return path
    int rc = acquire_slot(*cx, bytes, packed ? (size_t)256 * 1024 : 0, hold);
    if (rc) return rc;
    HostSlot& S = *hold.s;
    float* d = reinterpret_cast<float*>(S.stage);
    cudaError_t e = cudaSuccess;
    if (!packed) {
        // Large batches: device row arrays [B][K] in the staging buffer; the batch is cut into (at most) two chunks of
        // whole rows, each chunk uploaded, solved and downloaded on its own stream with its own workspace lane.  The
        // first chunk is small, so the kernels start after a fraction of the upload and the rest of the upload (and
        // the first chunk's download) runs under them; the second chunk's CTAs move onto the SMs as the first
        // chunk's CTAs retire, so the cut costs no tail.
        struct Row { const float* src; float* dev; size_t k; };
        Row rows[10] = {{p, nullptr, 3}, {v, nullptr, 3}, {quat, nullptr, 4}, {w, nullptr, 3}, {r_feet, nullptr, 12},
                        {rpy, nullptr, 3}, {traj, nullptr, (size_t)12 * h}, {gait, nullptr, (size_t)4 * h},
                        {mu_i, nullptr, 1}, {fmax_i, nullptr, 1}};
        for (Row& r : rows) { r.dev = d; d += B * r.k; }
        float* dgrf = d; d += B * 12;
        float* du = nullptr;
        if (u_out) { du = d; d += B * 12 * h; }
        int32_t* dstat = reinterpret_cast<int32_t*>(d);
        int32_t* dit = dstat + B;
        const size_t first = B >= 8192 ? B / 8 : B;
        const size_t cut[3] = {0, first, B};
        const int nchunk = first < B ? 2 : 1;
        for (int c = 0; c < nchunk; ++c) {
            cudaStream_t cs = S.stream[c];
            const size_t b0 = cut[c], nb = cut[c + 1] - cut[c];
            for (const Row& r : rows)
                if (r.src && e == cudaSuccess)
                    e = cudaMemcpyAsync(r.dev + b0 * r.k, r.src + b0 * r.k, nb * r.k * sizeof(float), cudaMemcpyHostToDevice, cs);
            if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpyAsync H2D", e);
            auto at = [&](const Row& r) -> const float* { return r.src ? r.dev + b0 * r.k : nullptr; };
            {
                std::lock_guard<std::mutex> lk(ctx_mu(*cx));
                rc = mpc_enqueue(*cx, P, opt, (int)nb, at(rows[0]), at(rows[1]), at(rows[2]), at(rows[3]), at(rows[4]),
                                 at(rows[5]), at(rows[6]), at(rows[7]), at(rows[8]), at(rows[9]), dgrf + b0 * 12,
                                 du ? du + b0 * 12 * h : nullptr, dstat + b0, dit + 2 * b0, cs);
            }
            if (rc) return rc;
            e = cudaMemcpyAsync(grf_out + b0 * 12, dgrf + b0 * 12, nb * 12 * sizeof(float), cudaMemcpyDeviceToHost, cs);
            if (e == cudaSuccess && u_out)
                e = cudaMemcpyAsync(u_out + b0 * 12 * h, du + b0 * 12 * h, nb * 12 * h * sizeof(float), cudaMemcpyDeviceToHost, cs);
            if (e == cudaSuccess && status_out)
                e = cudaMemcpyAsync(status_out + b0, dstat + b0, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, cs);
            if (e == cudaSuccess && iters_out)
                e = cudaMemcpyAsync(iters_out + 2 * b0, dit + 2 * b0, 2 * nb * sizeof(int32_t), cudaMemcpyDeviceToHost, cs);
            if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpyAsync D2H", e);
        }
        e = cudaStreamSynchronize(S.stream[0]);
        if (e == cudaSuccess && nchunk > 1) e = cudaStreamSynchronize(S.stream[1]);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaStreamSynchronize", e);
        return QR_OK;
    }
    // Small batches (latency path): gather the rows into the slot's pinned mirror of the staging buffer so that
    // the whole call is one host->device and one device->host copy.
    cudaStream_t st = S.stream[0];
    float* hp = reinterpret_cast<float*>(S.pin);
    auto up = [&](const float* src, size_t cnt) -> float* {
        float* dst = d;
        d += cnt;
        if (src) memcpy(hp + (dst - reinterpret_cast<float*>(S.stage)), src, cnt * sizeof(float));
        return src ? dst : nullptr;
    };
    float* dp = up(p, B * 3);
    float* dv = up(v, B * 3);
    float* dq = up(quat, B * 4);
    float* dw = up(w, B * 3);
    float* dr = up(r_feet, B * 12);
    float* drpy = up(rpy, B * 3);
    float* dtraj = up(traj, B * 12 * h);
    float* dgait = up(gait, B * 4 * h);
    float* dmu = up(mu_i, B);
    float* dfm = up(fmax_i, B);
    e = cudaMemcpyAsync(S.stage, S.pin, n_in * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpyAsync H2D", e);
    float* dgrf = d; d += B * 12;
    float* du = nullptr;
    if (u_out) { du = d; d += B * 12 * h; }
    int32_t* dstat = reinterpret_cast<int32_t*>(d);
    int32_t* dit = dstat + B;
    {
        std::lock_guard<std::mutex> lk(ctx_mu(*cx));
        rc = mpc_enqueue(*cx, P, opt, batch, dp, dv, dq, dw, dr, drpy, dtraj, dgait, dmu, dfm, dgrf, du, dstat, dit, st);
    }
    if (rc) return rc;
    const size_t off = n_in * sizeof(float), nout = n_out * sizeof(float) + 3 * B * sizeof(int32_t);
    e = cudaMemcpyAsync(S.pin + off, S.stage + off, nout, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpyAsync D2H", e);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaStreamSynchronize", e);
    const unsigned char* hb = S.pin;
    memcpy(grf_out, hb + ((unsigned char*)dgrf - S.stage), B * 12 * sizeof(float));
    if (u_out) memcpy(u_out, hb + ((unsigned char*)du - S.stage), B * 12 * h * sizeof(float));
    if (status_out) memcpy(status_out, hb + ((unsigned char*)dstat - S.stage), B * sizeof(int32_t));
    if (iters_out) memcpy(iters_out, hb + ((unsigned char*)dit - S.stage), 2 * B * sizeof(int32_t));
    return QR_OK;
}

// One batch of host rows sharded over several devices of this process (the reference's seam is the
// single SolveDenseMPC call of qr_mpc_stance_leg_controller.cpp:385-410): contiguous shards [g*B/G, (g+1)*B/G), one
// host thread per device running the host entry point above on its shard, results written straight into the
// caller's arrays (the "final gather" is each shard's own device->host copy).  No collective.
// The per-device host threads are persistent (created on first use, bound to their device once): at 8192 instances
// per GPU a solve takes 2 ms, and creating and binding eight threads per call was a tenth of that.
namespace {
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = false, quit = false;
};
Worker* g_workers[QR_MAX_DEVICES] = {};
std::mutex g_multi_mu;   // one multi-device call at a time (the workers are shared)

void worker_main(Worker* w, int device) {
    cudaSetDevice(device);
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<void()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        job();
        lk.lock();
        w->done = true;
        w->cv.notify_all();
    }
}
void worker_submit(int device, std::function<void()> job) {
    Worker*& w = g_workers[device];
    if (!w) {
        w = new Worker();
        w->th = std::thread(worker_main, w, device);
    }
    std::lock_guard<std::mutex> lk(w->m);
    w->job = std::move(job);
    w->has_job = true;
    w->done = false;
    w->cv.notify_all();
}
void worker_wait(int device) {
    Worker* w = g_workers[device];
    std::unique_lock<std::mutex> lk(w->m);
    w->cv.wait(lk, [&] { return w->done; });
}
void workers_stop() {
    for (Worker*& w : g_workers) {
        if (!w) continue;
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
        w = nullptr;
    }
}
}  // namespace

extern "C" int qr_gpu_mpc_solve_batch_host_multi(int n_devices, const int* devices, const qr_mpc_params* P,
                                                 const qr_qp_options* opt, int batch, const float* p, const float* v,
                                                 const float* quat, const float* w, const float* r_feet,
                                                 const float* rpy, const float* traj, const float* gait,
                                                 const float* mu_i, const float* fmax_i, float* grf_out, float* u_out,
                                                 int32_t* status_out, int32_t* iters_out) {
    if (n_devices < 1 || n_devices > QR_MAX_DEVICES || !devices) return fail(QR_EINVAL, "bad device list");
    int rc = check_params(P, batch);
    if (rc) return rc;
    for (int g = 0; g < n_devices; ++g) {
        if (devices[g] < 0 || devices[g] >= QR_MAX_DEVICES) return fail(QR_EINVAL, "device index out of range");
        for (int k = 0; k < g; ++k)
            if (devices[k] == devices[g]) return fail(QR_EINVAL, "device listed twice");
    }
    if (batch == 0) return QR_OK;
    const int h = P->horizon;
    int codes[QR_MAX_DEVICES];
    char errs[QR_MAX_DEVICES][320];
    auto shard = [&](int g) {
        errs[g][0] = 0;
        const size_t b0 = (size_t)batch * g / n_devices, b1 = (size_t)batch * (g + 1) / n_devices, nb = b1 - b0;
        int r = qr_gpu_init(devices[g]);   // (re)binds the worker to its device; creates the context on first use
        if (r == QR_OK && nb > 0) {
            auto at = [&](const float* a, size_t k) -> const float* { return a ? a + b0 * k : nullptr; };
            r = qr_gpu_mpc_solve_batch_host(P, opt, (int)nb, at(p, 3), at(v, 3), at(quat, 4), at(w, 3), at(r_feet, 12),
                                            at(rpy, 3), at(traj, (size_t)12 * h), at(gait, (size_t)4 * h), at(mu_i, 1),
                                            at(fmax_i, 1), grf_out + b0 * 12, u_out ? u_out + b0 * 12 * h : nullptr,
                                            status_out ? status_out + b0 : nullptr, iters_out ? iters_out + 2 * b0 : nullptr);
        }
        codes[g] = r;
        if (r != QR_OK) snprintf(errs[g], sizeof(errs[g]), "device %d: %s", devices[g], t_err);
    };
    {
        std::lock_guard<std::mutex> lk(g_multi_mu);
        for (int g = 0; g < n_devices; ++g) worker_submit(devices[g], [&shard, g] { shard(g); });
        for (int g = 0; g < n_devices; ++g) worker_wait(devices[g]);
    }
    for (int g = 0; g < n_devices; ++g)
        if (codes[g] != QR_OK) { snprintf(t_err, sizeof(t_err), "%s", errs[g]); return codes[g]; }
    return QR_OK;
}

#ifdef QR_PROFILE
// Debug builds only (-DQR_PROFILE): read and clear the per-phase cycle table.
extern "C" int qr_gpu_debug_profile(unsigned long long* out64) {
    cudaError_t e = cudaMemcpyFromSymbol(out64, qr_prof_table, 64 * sizeof(unsigned long long));
    if (e != cudaSuccess) return QR_ECUDA;
    unsigned long long zero[64] = {0};
    cudaMemcpyToSymbol(qr_prof_table, zero, sizeof(zero));
    return QR_OK;
}
#endif

// ==================================================================================================
// Whole-body control and swing-foot kernels
// ==================================================================================================
#include "wbc_problem.h"

namespace {

// Threads per robot, measured on B200 with one robot per CTA and six robots per SM (Lite3, 65536 robots, outputs
// bit-identical): 32 / 64 / 96 / 128 threads -> 4.63 / 5.90 / 5.23 / 5.06 M robots/s (batch 1024: 4.47 / 5.63 / 6.29 / 6.01).
// (The workspace is 32.2 KB per robot now; the kernel turned out to be bound by instruction fetch, not by residency.)
#ifndef QR_WBC_NT_DEF
#define QR_WBC_NT_DEF 64
#endif
constexpr int QR_WBC_NT = QR_WBC_NT_DEF;

// QR_WBC_TEAMS robots per CTA, each with its own team of QR_WBC_NT threads, workspace and named barrier (qr_team.h): the
// teams run the same code at nearly the same time, so an instruction line is fetched once for all of them.  Measured
// (65536 robots, every variant on the named barrier): 1 / 2 / 3 / 4 robots per CTA -> 6.76 / 8.55 / 8.33 / 7.16 M robots/s.
#ifndef QR_WBC_TEAMS
#define QR_WBC_TEAMS 2
#endif
static_assert(QR_WBC_TEAMS == 1 || QR_WBC_NT == QR_MULTI_TEAM_NT, "several teams per CTA need the multi-team thread index");
__global__ void __launch_bounds__(QR_WBC_NT * QR_WBC_TEAMS) qr_wbc_kernel(const QrWbcArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = QR_WBC_NT;
    const int team = threadIdx.x / NT;
    QrWbcWork W;
    qr_wbc_carve(W, smem + (size_t)team * qr_wbc_smem_bytes());
    qr_wbc_init_tables<NT>(W);
    for (int prob = blockIdx.x * QR_WBC_TEAMS + team; prob < A.batch; prob += gridDim.x * QR_WBC_TEAMS) qr_wbc_problem<NT>(A, prob, W);
}

__global__ void qr_swing_parabola_kernel(int batch, const float* start, const float* end, const float* height,
                                         const float* phase, int phase_module, float* pos, int32_t* valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    float o[3] = {0.f, 0.f, 0.f};
    const int ok = qr_swing_parabola(start + 3 * (size_t)i, end + 3 * (size_t)i, height[i], phase[i], phase_module, o);
    pos[3 * (size_t)i] = o[0]; pos[3 * (size_t)i + 1] = o[1]; pos[3 * (size_t)i + 2] = o[2];
    if (valid) valid[i] = ok;
}

// Device copy of the robot constants for `model` on this context (the context mutex held).  Every distinct model keeps its own
// buffer (a few robots per process at most), so a batch in flight never sees its constants overwritten; when the
// table is full the least recently used entry is recycled after the device has drained.
int wbc_model_on_device(Ctx& cx, const qr_wbc_model* model, const QrWbcModelDev** out) {
    WbcModelSlot* pick = nullptr;
    for (WbcModelSlot& m : cx.wbc_models)
        if (m.valid && memcmp(&m.key, model, sizeof(*model)) == 0) pick = &m;
    if (!pick) {
        for (WbcModelSlot& m : cx.wbc_models)
            if (!m.valid && !pick) pick = &m;
        if (!pick) {
            pick = &cx.wbc_models[0];
            for (WbcModelSlot& m : cx.wbc_models)
                if (m.stamp < pick->stamp) pick = &m;
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return fail(QR_ECUDA, "cudaDeviceSynchronize(wbc model)", e);
            pick->valid = false;
        }
        if (!pick->dev) {
            cudaError_t e = cudaMalloc(&pick->dev, sizeof(QrWbcModelDev));
            if (e != cudaSuccess) return fail(QR_ENOMEM, "cudaMalloc(wbc model)", e);
        }
        QrWbcModelDev host;
        qr_wbc_host::build(model, &host);
        // synchronous upload into a buffer no launch references yet
        cudaError_t e = cudaMemcpy(pick->dev, &host, sizeof(host), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpy(wbc model)", e);
        pick->key = *model;
        pick->valid = true;
    }
    pick->stamp = ++cx.clock;
    *out = static_cast<const QrWbcModelDev*>(pick->dev);
    return QR_OK;
}

// The context mutex is held by the caller.
int wbc_launch(Ctx& cx, const qr_wbc_model* model, int batch, const float* state, const float* cmd, const int32_t* contact,
               QrWbcArgs& A, void* cuda_stream) {
    if (!model || batch < 0) return fail(QR_EINVAL, "null model or negative batch");
    if (batch == 0) return QR_OK;
    if (!state || !cmd || !contact) return fail(QR_EINVAL, "null input pointer");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const QrWbcModelDev* dev_model = nullptr;
    int rc = wbc_model_on_device(cx, model, &dev_model);
    if (rc) return rc;
    const size_t smem = qr_wbc_smem_bytes() * QR_WBC_TEAMS;
    if (!cx.wbc_occ) {
        cudaError_t e = cudaFuncSetAttribute(qr_wbc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaFuncSetAttribute(wbc)", e);
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qr_wbc_kernel, QR_WBC_NT * QR_WBC_TEAMS, smem);
        if (e != cudaSuccess) return fail(QR_ECUDA, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(wbc)", e);
        cx.wbc_occ = occ < 1 ? 1 : occ;
    }
    int grid = cx.sm_count * cx.wbc_occ;
    if (grid > (batch + QR_WBC_TEAMS - 1) / QR_WBC_TEAMS) grid = (batch + QR_WBC_TEAMS - 1) / QR_WBC_TEAMS;
    A.model = dev_model;
    A.opt = default_options();
    A.batch = batch;
    A.state = state; A.cmd = cmd; A.contact = contact;
    qr_wbc_kernel<<<grid, QR_WBC_NT * QR_WBC_TEAMS, smem, st>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_wbc_kernel", e);
    return QR_OK;
}

}  // namespace

extern "C" int qr_gpu_wbc_solve_batch(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                                      const int32_t* contact, float* tau_out, float* fr_out, float* qdes_out,
                                      float* qddes_out, int32_t* status_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch > 0 && !tau_out) return fail(QR_EINVAL, "null output pointer");
    QrWbcArgs A;
    memset(&A, 0, sizeof(A));
    A.tau32 = tau_out; A.fr32 = fr_out; A.qdes32 = qdes_out; A.qddes32 = qddes_out; A.status = status_out;
    return wbc_launch(*cx, model, batch, state, cmd, contact, A, cuda_stream);
}

extern "C" int qr_gpu_wbc_solve_batch_host(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                                           const int32_t* contact, float* tau_out, float* fr_out, float* qdes_out,
                                           float* qddes_out, int32_t* status_out) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    if (!model || batch < 0) return fail(QR_EINVAL, "null model or negative batch");
    if (batch == 0) return QR_OK;
    if (!state || !cmd || !contact || !tau_out) return fail(QR_EINVAL, "null pointer");
    const size_t B = (size_t)batch;
    const size_t bytes = B * ((37 + 66 + 4 * 12) * sizeof(float) + 5 * sizeof(int32_t)) + 256;
    SlotHold hold;   // private staging slot; drained and released on every return path
    int rc = acquire_slot(*cx, bytes, 0, hold);
    if (rc) return rc;
    HostSlot& S = *hold.s;
    cudaStream_t st = S.stream[0];
    float* d_state = reinterpret_cast<float*>(S.stage);
    float* d_cmd = d_state + B * 37;
    float* d_tau = d_cmd + B * 66;
    float* d_fr = d_tau + B * 12;
    float* d_qdes = d_fr + B * 12;
    float* d_qddes = d_qdes + B * 12;
    int32_t* d_contact = reinterpret_cast<int32_t*>(d_qddes + B * 12);
    int32_t* d_status = d_contact + B * 4;
    cudaError_t e = cudaMemcpyAsync(d_state, state, B * 37 * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cmd, cmd, B * 66 * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_contact, contact, B * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpyAsync H2D", e);
    rc = qr_gpu_wbc_solve_batch(model, batch, d_state, d_cmd, d_contact, d_tau, d_fr, d_qdes, d_qddes, d_status, st);
    if (rc) return rc;
    e = cudaMemcpyAsync(tau_out, d_tau, B * 12 * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && fr_out) e = cudaMemcpyAsync(fr_out, d_fr, B * 12 * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && qdes_out) e = cudaMemcpyAsync(qdes_out, d_qdes, B * 12 * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && qddes_out) e = cudaMemcpyAsync(qddes_out, d_qddes, B * 12 * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && status_out) e = cudaMemcpyAsync(status_out, d_status, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaMemcpyAsync D2H", e);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(QR_ECUDA, "cudaStreamSynchronize", e);
    return QR_OK;
}

extern "C" int qr_gpu_wbc_solve_batch_f64(const qr_wbc_model* model, int batch, const float* state, const float* cmd,
                                          const int32_t* contact, double* tau_out, double* fr_out, double* qdes_out,
                                          double* qddes_out, double* dbg_out, int32_t* status_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch > 0 && !tau_out) return fail(QR_EINVAL, "null output pointer");
    QrWbcArgs A;
    memset(&A, 0, sizeof(A));
    A.tau64 = tau_out; A.fr64 = fr_out; A.qdes64 = qdes_out; A.qddes64 = qddes_out; A.dbg = dbg_out; A.status = status_out;
    return wbc_launch(*cx, model, batch, state, cmd, contact, A, cuda_stream);
}

extern "C" int qr_gpu_swing_parabola_batch(int batch, const float* start, const float* end, const float* height,
                                           const float* phase, int phase_module, float* pos_out, int32_t* valid_out,
                                           void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch < 0) return fail(QR_EINVAL, "negative batch");
    if (batch == 0) return QR_OK;
    if (!start || !end || !height || !phase || !pos_out) return fail(QR_EINVAL, "null pointer");
    qr_swing_parabola_kernel<<<(batch + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(batch, start, end, height, phase,
                                                                                       phase_module, pos_out, valid_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_swing_parabola_kernel", e);
    return QR_OK;
}

// ==================================================================================================
// MPC pre- and post-processing (contact table, reference trajectory, GRF -> joint torques)
// ==================================================================================================
#include "mpc_io.h"

namespace {

__global__ void qr_mpc_inputs_kernel(int h, int num_horizon_l, float dt_mpc, int batch, const float* progress,
                                     const float* duty, const int32_t* early, const int32_t* contacts,
                                     const float* traj_init, const float* pos_xy, float* gait_out, float* traj_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    if (gait_out)
        qr_mpc_contact_table(h, num_horizon_l, progress + 4 * (size_t)i, duty + 4 * (size_t)i,
                             early ? early + 4 * (size_t)i : nullptr, contacts ? contacts + 4 * (size_t)i : nullptr,
                             gait_out + (size_t)4 * h * i);
    if (traj_out)
        qr_mpc_reference_traj(h, dt_mpc, traj_init + 12 * (size_t)i, pos_xy + 2 * (size_t)i, traj_out + (size_t)12 * h * i);
}

__global__ void qr_mpc_leg_torque_kernel(float hip_len, float upper_len, float lower_len, int batch, const float* quat,
                                         const float* q, const float* grf, float* f_ff_out, float* tau_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    qr_mpc_grf_to_torque(hip_len, upper_len, lower_len, quat + 4 * (size_t)i, q + 12 * (size_t)i, grf + 12 * (size_t)i,
                         f_ff_out ? f_ff_out + 12 * (size_t)i : nullptr, tau_out + 12 * (size_t)i);
}

}  // namespace

extern "C" int qr_gpu_mpc_inputs_batch(int horizon, int num_horizon_l, float dt_mpc, int batch, const float* progress,
                                       const float* duty, const int32_t* early_contact, const int32_t* contacts,
                                       const float* traj_init, const float* pos_xy, float* gait_out, float* traj_out,
                                       void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (horizon < 1 || horizon > QR_MAX_HORIZON || num_horizon_l < 1 || batch < 0) return fail(QR_EINVAL, "bad size argument");
    if (batch == 0) return QR_OK;
    if (gait_out && (!progress || !duty)) return fail(QR_EINVAL, "contact table needs progress and duty");
    if (traj_out && (!traj_init || !pos_xy)) return fail(QR_EINVAL, "trajectory needs traj_init and pos_xy");
    qr_mpc_inputs_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(
        horizon, num_horizon_l, dt_mpc, batch, progress, duty, early_contact, contacts, traj_init, pos_xy, gait_out, traj_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_mpc_inputs_kernel", e);
    return QR_OK;
}

extern "C" int qr_gpu_mpc_leg_torque_batch(float hip_len, float upper_len, float lower_len, int batch, const float* quat,
                                           const float* q, const float* grf, float* f_ff_out, float* tau_out,
                                           void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch < 0) return fail(QR_EINVAL, "negative batch");
    if (batch == 0) return QR_OK;
    if (!quat || !q || !grf || !tau_out) return fail(QR_EINVAL, "null pointer");
    qr_mpc_leg_torque_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(hip_len, upper_len, lower_len, batch,
                                                                                       quat, q, grf, f_ff_out, tau_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_mpc_leg_torque_kernel", e);
    return QR_OK;
}

// ==================================================================================================
// Force-balance stance QP (one thread per robot)
// ==================================================================================================
#include "fb_problem.h"

namespace {
__global__ void __launch_bounds__(64) qr_force_balance_kernel(const QrFbArgs A) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob < A.batch) qr_fb_problem(A, prob);
}
}  // namespace

extern "C" int qr_gpu_force_balance_batch(const qr_fb_params* P, int batch, const float* inertia, const float* foot,
                                          const float* acc, const int32_t* contact, const float* gravity,
                                          const float* frame, float* force_out, int32_t* status_out, int32_t* iters_out,
                                          void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (!P || batch < 0) return fail(QR_EINVAL, "null params or negative batch");
    if (!(P->mass > 0.f) || !(P->mu > 0.f)) return fail(QR_EINVAL, "mass and mu must be positive");
    if (batch == 0) return QR_OK;
    if (!foot || !acc || !contact || !force_out) return fail(QR_EINVAL, "null pointer");
    QrFbArgs A;
    memset(&A, 0, sizeof(A));
    A.P = *P;
    A.batch = batch;
    A.inertia = inertia; A.foot = foot; A.acc = acc; A.contact = contact; A.gravity = gravity; A.frame = frame;
    A.force_out = force_out; A.status_out = status_out; A.iters_out = iters_out;
    qr_force_balance_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)cuda_stream>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_force_balance_kernel", e);
    return QR_OK;
}

// ==================================================================================================
// Swing-leg targets: cubic B-spline trajectory and heuristic foothold (one thread per foot)
// ==================================================================================================
#include "swing_extra.h"

namespace {
__global__ void qr_swing_bspline_kernel(int batch, const float* ip, const float* tp, const float* height, const float* duration,
                                        const float* t0, const float* t, float* pos, float* vel, int32_t* valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    float p[3], v[3];
    const int ok = qr_swing_bspline(ip + 3 * (size_t)i, tp + 3 * (size_t)i, height[i], duration[i], t0[i], t[i], p, v);
    if (ok) {
        for (int k = 0; k < 3; ++k) { pos[3 * (size_t)i + k] = p[k]; vel[3 * (size_t)i + k] = v[k]; }
    }
    if (valid) valid[i] = ok;
}

struct QrFootholdRows {
    const float *com_vel, *w, *dR, *base_R, *rpy, *foot_base, *des_speed, *des_twist, *des_height, *swing_remain, *norm_phase;
    const int32_t *allow_switch, *swing_mask;
    float *foothold, *phase;
};
__global__ void qr_foothold_kernel(const QrFootholdParams P, int batch, const QrFootholdRows R) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 2, leg = idx & 3;
    if (i >= batch || !R.swing_mask[4 * (size_t)i + leg]) return;
    qr_foothold_heuristic(P, leg, R.com_vel + 3 * (size_t)i, R.w + 3 * (size_t)i, R.dR + 9 * (size_t)i, R.base_R + 9 * (size_t)i,
                          R.rpy + 3 * (size_t)i, R.foot_base + 12 * (size_t)i, R.des_speed + 3 * (size_t)i, R.des_twist[i],
                          R.des_height[i], R.swing_remain[4 * (size_t)i + leg], R.allow_switch[4 * (size_t)i + leg],
                          R.norm_phase[4 * (size_t)i + leg], R.foothold + 12 * (size_t)i, R.phase + 4 * (size_t)i + leg);
}
}  // namespace

extern "C" int qr_gpu_swing_bspline_batch(int batch, const float* initial_pos, const float* target_pos, const float* height,
                                          const float* duration, const float* initial_time, const float* time,
                                          float* pos_out, float* vel_out, int32_t* valid_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch < 0) return fail(QR_EINVAL, "negative batch");
    if (batch == 0) return QR_OK;
    if (!initial_pos || !target_pos || !height || !duration || !initial_time || !time || !pos_out || !vel_out)
        return fail(QR_EINVAL, "null pointer");
    qr_swing_bspline_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(batch, initial_pos, target_pos, height, duration,
                                                                                       initial_time, time, pos_out, vel_out, valid_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_swing_bspline_kernel", e);
    return QR_OK;
}

extern "C" int qr_gpu_foothold_heuristic_batch(const qr_foothold_params* P, int batch, const float* com_vel, const float* rpy_rate,
                                               const float* dR, const float* base_R, const float* rpy, const float* foot_base,
                                               const float* des_speed, const float* des_twist, const float* des_height,
                                               const float* swing_remain, const float* norm_phase, const int32_t* allow_switch,
                                               const int32_t* swing_mask, float* foothold_io, float* phase_io, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (!P || batch < 0) return fail(QR_EINVAL, "null params or negative batch");
    if (batch == 0) return QR_OK;
    if (!com_vel || !rpy_rate || !dR || !base_R || !rpy || !foot_base || !des_speed || !des_twist || !des_height ||
        !swing_remain || !norm_phase || !allow_switch || !swing_mask || !foothold_io || !phase_io)
        return fail(QR_EINVAL, "null pointer");
    QrFootholdParams Q;
    static_assert(sizeof(QrFootholdParams) == sizeof(qr_foothold_params), "layout");
    memcpy(&Q, P, sizeof(Q));
    QrFootholdRows R{com_vel, rpy_rate, dR, base_R, rpy, foot_base, des_speed, des_twist, des_height, swing_remain, norm_phase,
                     allow_switch, swing_mask, foothold_io, phase_io};
    qr_foothold_kernel<<<(4 * batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(Q, batch, R);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, "launch qr_foothold_kernel", e);
    return QR_OK;
}

// ==================================================================================================
// Controller arithmetic around the two solvers: lever arms, leg kinematics, MPC-mode swing targets, gait phase
// ==================================================================================================
#include "ctl_extra.h"

namespace {

__global__ void qr_mpc_lever_arms_kernel(int batch, const float* quat, const float* foot_base, float cx, float cy, float cz,
                                         float* r_feet) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const float com[3] = {cx, cy, cz};
    qr_mpc_lever_arms(quat + 4 * (size_t)i, foot_base + 12 * (size_t)i, com, r_feet + 12 * (size_t)i);
}

// one thread per leg
__global__ void qr_leg_kinematics_kernel(const QrLegGeom G, int batch, const float* q, const float* qd, float* foot_base,
                                         float* jac, float* foot_vel) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 2, leg = idx & 3;
    if (i >= batch) return;
    const float* t = q + 12 * (size_t)i + 3 * leg;
    if (foot_base) qr_leg_fk(G, leg, t, foot_base + 12 * (size_t)i + 3 * leg);
    if (jac || foot_vel) {
        float J[9];
        qr_leg_jacobian(G, leg, t, J);
        if (jac)
            for (int e = 0; e < 9; ++e) jac[36 * (size_t)i + 9 * leg + e] = J[e];
        if (foot_vel && qd) qr_mat3_vec(J, qd + 12 * (size_t)i + 3 * leg, foot_vel + 12 * (size_t)i + 3 * leg);
    }
}

__global__ void qr_leg_ik_kernel(const QrLegGeom G, int batch, const float* foot_base, const float* foot_vel,
                                 const int32_t* leg_mask, float* q_out, float* qd_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 2, leg = idx & 3;
    if (i >= batch || (leg_mask && !leg_mask[4 * (size_t)i + leg])) return;
    float t[3];
    qr_leg_ik(G, leg, foot_base + 12 * (size_t)i + 3 * leg, t);
    for (int a = 0; a < 3; ++a) q_out[12 * (size_t)i + 3 * leg + a] = t[a];
    if (qd_out && foot_vel) qr_leg_ik_velocity(G, leg, t, foot_vel + 12 * (size_t)i + 3 * leg, qd_out + 12 * (size_t)i + 3 * leg);
}

struct QrSwingRows {
    const float *base_pos, *quat, *v_world, *foothold, *planner_phase, *switch_pos, *swing_duration;
    const int32_t* swing_mask;
    float *cmd, *foot_base_des, *q_des, *qd_des;
    int32_t* valid;
};
__global__ void qr_swing_targets_kernel(const QrLegGeom G, int batch, int horizontal_terrain, const QrSwingRows R) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 2, leg = idx & 3;
    if (i >= batch) return;
    const size_t l = 4 * (size_t)i + leg;
    if (!R.swing_mask[l]) { if (R.valid) R.valid[l] = 0; return; }
    float* cmd = R.cmd + 66 * (size_t)i;
    const int ok = qr_swing_targets_leg(G, leg, R.base_pos + 3 * (size_t)i, R.quat + 4 * (size_t)i, R.v_world + 3 * (size_t)i,
                                        R.foothold + 3 * l, R.planner_phase[l], R.switch_pos + 3 * l, R.swing_duration[l],
                                        horizontal_terrain, cmd + 15 + 3 * leg, cmd + 27 + 3 * leg, cmd + 39 + 3 * leg,
                                        R.foot_base_des ? R.foot_base_des + 3 * l : nullptr, R.q_des ? R.q_des + 3 * l : nullptr,
                                        R.qd_des ? R.qd_des + 3 * l : nullptr);
    if (R.valid) R.valid[l] = ok;
}

struct QrGaitRows {
    const float *time, *cfg;
    const int32_t *contacts, *stop;
    int32_t* istate;
    float *fstate, *phase_full, *norm_phase, *swing_remain;
    int32_t *allow, *early, *swing_mask, *stance_mask;
};
__global__ void qr_gait_update_kernel(int batch, float contact_threshold, int advanced_trot, const QrGaitRows R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const size_t r4 = 4 * (size_t)i;
    int32_t al[4];
    float out[12];
    for (int l = 0; l < 4; ++l) {
        out[l] = R.phase_full[r4 + l];
        out[4 + l] = R.norm_phase[r4 + l];
        out[8 + l] = R.swing_remain[r4 + l];
    }
    int32_t* ist = R.istate + 20 * (size_t)i;
    qr_gait_update(R.time[i], R.cfg + 20 * (size_t)i, contact_threshold, R.contacts + r4, R.stop ? R.stop[i] : 0, advanced_trot,
                   ist, R.fstate + r4, out, al);
    for (int l = 0; l < 4; ++l) {
        R.phase_full[r4 + l] = out[l];
        R.norm_phase[r4 + l] = out[4 + l];
        R.swing_remain[r4 + l] = out[8 + l];
        const int ls = ist[12 + l];
        if (R.allow) R.allow[r4 + l] = al[l];
        if (R.early) R.early[r4 + l] = ls == 2;
        // the legs the swing controller moves (qr_swing_leg_controller.cpp:218-228)
        const int swing = !((ls == 1 && al[l]) || ls == 2);
        if (R.swing_mask) R.swing_mask[r4 + l] = swing;
        if (R.stance_mask) R.stance_mask[r4 + l] = !swing;
    }
}

QrLegGeom leg_geom_of(const qr_leg_geometry* g) {
    QrLegGeom G;
    G.hip_len = g->hip_len; G.upper_len = g->upper_len; G.lower_len = g->lower_len;
    for (int k = 0; k < 12; ++k) G.hip_offset[k] = g->hip_offset[k];
    return G;
}

int launch_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(QR_ECUDA, what, e);
    return QR_OK;
}

}  // namespace

extern "C" int qr_gpu_mpc_lever_arms_batch(int batch, const float* quat, const float* foot_base, const float* com_offset,
                                           float* r_feet_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch < 0) return fail(QR_EINVAL, "negative batch");
    if (batch == 0) return QR_OK;
    if (!quat || !foot_base || !com_offset || !r_feet_out) return fail(QR_EINVAL, "null pointer");
    qr_mpc_lever_arms_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(batch, quat, foot_base, com_offset[0],
                                                                                       com_offset[1], com_offset[2], r_feet_out);
    return launch_check("launch qr_mpc_lever_arms_kernel");
}

extern "C" int qr_gpu_leg_kinematics_batch(const qr_leg_geometry* geom, int batch, const float* q, const float* qd,
                                           float* foot_base_out, float* jac_out, float* foot_vel_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (!geom || batch < 0) return fail(QR_EINVAL, "null geometry or negative batch");
    if (batch == 0) return QR_OK;
    if (!q || (foot_vel_out && !qd)) return fail(QR_EINVAL, "null pointer");
    qr_leg_kinematics_kernel<<<(4 * batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(leg_geom_of(geom), batch, q, qd,
                                                                                           foot_base_out, jac_out, foot_vel_out);
    return launch_check("launch qr_leg_kinematics_kernel");
}

extern "C" int qr_gpu_leg_ik_batch(const qr_leg_geometry* geom, int batch, const float* foot_base, const float* foot_vel,
                                   const int32_t* leg_mask, float* q_out, float* qd_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (!geom || batch < 0) return fail(QR_EINVAL, "null geometry or negative batch");
    if (batch == 0) return QR_OK;
    if (!foot_base || !q_out || (qd_out && !foot_vel)) return fail(QR_EINVAL, "null pointer");
    qr_leg_ik_kernel<<<(4 * batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(leg_geom_of(geom), batch, foot_base, foot_vel,
                                                                                   leg_mask, q_out, qd_out);
    return launch_check("launch qr_leg_ik_kernel");
}

extern "C" int qr_gpu_swing_targets_batch(const qr_leg_geometry* geom, int batch, const float* base_pos, const float* quat,
                                          const float* v_world, const float* foothold, const float* planner_phase,
                                          const float* switch_pos, const float* swing_duration, const int32_t* swing_mask,
                                          int horizontal_terrain, float* wbc_cmd_io, float* foot_base_des_out, float* q_des_out,
                                          float* qd_des_out, int32_t* valid_out, void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (!geom || batch < 0) return fail(QR_EINVAL, "null geometry or negative batch");
    if (batch == 0) return QR_OK;
    if (!base_pos || !quat || !v_world || !foothold || !planner_phase || !switch_pos || !swing_duration || !swing_mask || !wbc_cmd_io)
        return fail(QR_EINVAL, "null pointer");
    if (qd_des_out && !q_des_out) return fail(QR_EINVAL, "joint velocities need the joint angles output");
    QrSwingRows R{base_pos, quat, v_world, foothold, planner_phase, switch_pos, swing_duration, swing_mask,
                  wbc_cmd_io, foot_base_des_out, q_des_out, qd_des_out, valid_out};
    qr_swing_targets_kernel<<<(4 * batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(leg_geom_of(geom), batch,
                                                                                          horizontal_terrain, R);
    return launch_check("launch qr_swing_targets_kernel");
}

extern "C" int qr_gpu_gait_update_batch(int batch, const float* time, const float* cfg, float contact_threshold,
                                        const int32_t* contacts, const int32_t* stop, int advanced_trot, int32_t* istate_io,
                                        float* fstate_io, float* phase_full_io, float* norm_phase_io, float* swing_remain_io,
                                        int32_t* allow_out, int32_t* early_out, int32_t* swing_mask_out, int32_t* stance_mask_out,
                                        void* cuda_stream) {
    Ctx* cx = current_ctx();
    if (!cx) return QR_ECUDA;
    std::lock_guard<std::mutex> lk(ctx_mu(*cx));
    if (batch < 0) return fail(QR_EINVAL, "negative batch");
    if (batch == 0) return QR_OK;
    if (!time || !cfg || !contacts || !istate_io || !fstate_io || !phase_full_io || !norm_phase_io || !swing_remain_io)
        return fail(QR_EINVAL, "null pointer");
    QrGaitRows R{time, cfg, contacts, stop, istate_io, fstate_io, phase_full_io, norm_phase_io, swing_remain_io,
                 allow_out, early_out, swing_mask_out, stance_mask_out};
    qr_gait_update_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(batch, contact_threshold, advanced_trot, R);
    return launch_check("launch qr_gait_update_kernel");
}
