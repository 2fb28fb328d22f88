// C ABI (include/ddb200.h): context management, kernel-plan selection, launch plumbing, host-buffer flavours.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/ddb200.h"
#include "philox.cuh"

namespace ddb {
size_t generic_smem_bytes(int m, int n, bool smem_tab);
int generic_block_threads(int m, int n, bool smem_tab);
cudaError_t launch_simplex_generic(const SolveArgs& a, bool smem_tab, int grid, int block, cudaStream_t st);
cudaError_t launch_generate(uint64_t key, long long first, long long B, int m, int n, double density, double* A,
                            double* b, double* c, double* x0, int sm_count, cudaStream_t st, int* launches);
bool regtile_supported(int m, int n);
cudaError_t launch_simplex_regtile(const SolveArgs& a, int sm_count, cudaStream_t st);
size_t regtile_scratch_bytes(int m, int n, int sm_count);
bool quadcol_supported(int m, int n);
int quadcol_grid(int sm_count);
int quadcol_gen_grid(int sm_count);
bool quadcol_gen_supported(int m, int n);
cudaError_t launch_simplex_quadcol_gen(const SolveArgs& a, int sm_count, cudaStream_t st);
cudaError_t launch_simplex_quadcol(const SolveArgs& a, int sm_count, cudaStream_t st);
#ifdef DDB_EXPERIMENTS   // `make experiments`: the measured negative results of round 1, not in the shipped library
bool tile2d_supported(int m, int n);
cudaError_t launch_simplex_tile2d(const SolveArgs& a, int sm_count, cudaStream_t st);
bool rowpipe_supported(int m, int n);
cudaError_t launch_simplex_rowpipe(const SolveArgs& a, int sm_count, cudaStream_t st);
#else
inline bool tile2d_supported(int, int) { return false; }
inline cudaError_t launch_simplex_tile2d(const SolveArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
inline bool rowpipe_supported(int, int) { return false; }
inline cudaError_t launch_simplex_rowpipe(const SolveArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
#endif
bool rowreg_supported(int m, int n);
int rowreg_grid(int m, int n, int sm_count);
size_t rowreg_rows_scratch_bytes(int m, int n, int grid);
size_t rowreg_d_scratch_bytes(int m, int n, int grid);
cudaError_t launch_simplex_rowreg(const SolveArgs& a, int sm_count, cudaStream_t st);
bool rowreg_gen_supported(int m, int n);
int rowreg_gen_grid(int m, int n, int sm_count);
cudaError_t launch_simplex_rowreg_gen(const SolveArgs& a, int sm_count, cudaStream_t st);
bool cluster_supported(int m, int n);
size_t cluster_scratch_bytes(int m, int n, int sm_count, long long B);
cudaError_t launch_simplex_cluster(const SolveArgs& a, int sm_count, cudaStream_t st);
bool s2v_gram_tc_supported(int m, int n);
size_t s2v_gram_out_floats(int m);
cudaError_t launch_s2v_gram_tc(long long B, int m, int n, const double* A, const double* b, const double* c, float* out,
                               int sm_count, cudaStream_t st);
cudaError_t launch_s2v_forward(const S2vArgs& a, int sm_count, long long smem_optin, cudaStream_t st, const char** why);
bool s2v_bipartite_dense_supported(int m, int n, int p, const void* A, long long smem_optin);
cudaError_t launch_s2v_bipartite_dense(const S2vArgs& a, int sm_count, cudaStream_t st);
struct S2vGradArgs {
    long long B;
    int m, n, p, T;
    const double* A;
    const double* b;
    const double* c;
    const float* params;
    const uint8_t* labels;
    float w0, w1;
    float* grad;
    double* loss;
    int* error_flag;
    int* inst_flag;
};
struct S2vGGradArgs {
    long long B;
    int m, n, p, T;
    const double* A;
    const double* b;
    const double* c;
    const float* params;
    const uint8_t* labels;
    float w0, w1;
    float* grad;
    double* loss;
    const uint8_t* row_ineq;
    const uint8_t* row_bound;
    const int* inst_flag;
    const int* flag_count;
    float* scratch;
};
size_t s2v_general_grad_smem_bytes(int m, int n, int p);
size_t s2v_general_grad_scratch_floats(int m, int n, int p, int T);
int s2v_general_grad_grid(long long B, int sm_count);
cudaError_t launch_s2v_bipartite_general_grad(const S2vGGradArgs& a, int grid, long long smem_optin, cudaStream_t st, const char** why);
cudaError_t launch_s2v_metrics(long long N, const float* logp, const float* probs, const uint8_t* labels, float thresh,
                               float w0, float w1, double* out, unsigned int* minbits, int sm_count, cudaStream_t st);
cudaError_t launch_s2v_bipartite_grad(const S2vGradArgs& a, int npar, int sm_count, long long smem_optin, cudaStream_t st,
                                      const char** why);
struct S2vCGradArgs {
    long long B;
    int m, p, T;
    const float* gram;
    int gram_pitch;
    const float* params;
    const uint8_t* labels;
    const uint8_t* row_ineq;
    float w0, w1;
    float* grad;
    double* loss;
};
size_t s2v_complete_grad_smem_bytes(int m, int p, int T);
cudaError_t launch_s2v_complete_grad(const S2vCGradArgs& a, int sm_count, long long smem_optin, cudaStream_t st, const char** why);
}  // namespace ddb

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return fail(DDB_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

namespace {
constexpr int kCounters = 64;
constexpr int kSlots = 3;   // chunks in flight in the host-buffer flavours (stage-in / H2D / solve + D2H overlap)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct HostBuf {             // pinned staging memory (cudaHostAlloc)
    void* p = nullptr;
    size_t cap = 0;
};

struct Retire {              // staged output: copy `bytes` from pinned `src` to the caller's pageable `dst` once the slot is done
    void* dst;
    const void* src;
    size_t bytes;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf A, b, c, mask, status, x, obj, labels, nact, piv, ties, viol;
    HostBuf in, out;         // pinned staging for pageable caller buffers
    std::vector<Retire> retire;
};
}  // namespace

struct ddb_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    int64_t smem_optin = 0;
    std::mutex mu;             // entry points that touch per-context state are serialised (one context per device per process)
    unsigned long long* counters = nullptr;   // kCounters work-queue counters followed by kCounters flag counters
    int next_counter = 0;
    DevBuf scratch;            // global tableau slabs / parked rows of the register kernels
    DevBuf dscr;               // row-per-thread kernel, hybrid rows: per-CTA global home of the crash inverse
    DevBuf slab;               // fused generate -> solve: per-CTA instance slabs (L2-resident)
    cudaEvent_t scratch_free = nullptr;   // recorded after every launch that uses scratch / dscr / slab: the next one waits
    bool scratch_in_use = false;
    DevBuf genA, genb, genc;   // unfused fallback of the fused call (odd n, shapes outside plan 0): chunk buffers
    cudaEvent_t gen_free = nullptr;       // last solve that read the chunk buffers
    bool gen_in_use = false;
    cudaStream_t gen_stream = nullptr;
    cudaEvent_t gen_done[2] = {nullptr, nullptr}, solve_done[2] = {nullptr, nullptr}, gen_fork = nullptr;
    cudaEvent_t chunk_ev[2] = {nullptr, nullptr};
    DevBuf gram;               // classifier: per-instance Gram row sums from the tensor-core kernel
    DevBuf s2vflag;            // classifier: per-instance "has a zero coefficient" flags of the dense bipartite kernel
    DevBuf s2vgscr;            // classifier: per-CTA scratch of the general-adjacency loss + gradient kernel
    Slot slots[kSlots];
    int forced_plan = -1;
    int fused_mode = 0;        // 0 automatic (the measured-faster path), 1 in-kernel generation, 2 generator kernel + solver kernel
    int64_t launches = 0;
};

static int ensure(DevBuf& buf, size_t bytes) {
    if (bytes <= buf.cap) return DDB_OK;
    if (buf.p) cudaFree(buf.p);
    buf.p = nullptr;
    buf.cap = 0;
    cudaError_t e = cudaMalloc(&buf.p, bytes);
    if (e != cudaSuccess) return fail(DDB_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    buf.cap = bytes;
    return DDB_OK;
}

extern "C" int ddb_abi_version(void) { return DDB_ABI_VERSION; }
extern "C" const char* ddb_last_error(void) { return g_err; }

static int create_resources(ddb_ctx* ctx) {
    CUDA_TRY(cudaMalloc(&ctx->counters, 3 * kCounters * sizeof(unsigned long long)));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->scratch_free, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->gen_free, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->gen_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->gen_fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->gen_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->solve_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kSlots; ++i) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->slots[i].stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->slots[i].done, cudaEventDisableTiming));
    }
    return DDB_OK;
}

extern "C" int ddb_destroy(ddb_ctx* ctx);

extern "C" int ddb_create(int device, ddb_ctx** out) {
    if (!out) return fail(DDB_EINVAL, "ddb_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(DDB_ECUDA, "ddb_create: no CUDA device (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(DDB_EINVAL, "ddb_create: device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DDB_EUNSUPPORTED, "ddb_create: device is sm_%d%d; this library is built for sm_100a only",
                    prop.major, prop.minor);
    ddb_ctx* ctx = new (std::nothrow) ddb_ctx();
    if (!ctx) return fail(DDB_ENOMEM, "ddb_create: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->smem_optin = (int64_t)prop.sharedMemPerBlockOptin;
    const int rc = create_resources(ctx);
    if (rc != DDB_OK) {            // nothing leaks on a half-built context
        ddb_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return DDB_OK;
}

static void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" int ddb_destroy(ddb_ctx* ctx) {
    if (!ctx) return DDB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kSlots; ++i) {
        Slot& s = ctx->slots[i];
        DevBuf* all[] = {&s.A, &s.b, &s.c, &s.mask, &s.status, &s.x, &s.obj, &s.labels, &s.nact, &s.piv, &s.ties, &s.viol};
        for (DevBuf* d : all) release(*d);
        if (s.in.p) cudaFreeHost(s.in.p);
        if (s.out.p) cudaFreeHost(s.out.p);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
    }
    release(ctx->scratch);
    release(ctx->dscr);
    release(ctx->slab);
    release(ctx->genA);
    release(ctx->genb);
    release(ctx->genc);
    release(ctx->gram);
    release(ctx->s2vgscr);
    release(ctx->s2vflag);
    if (ctx->scratch_free) cudaEventDestroy(ctx->scratch_free);
    if (ctx->gen_free) cudaEventDestroy(ctx->gen_free);
    if (ctx->gen_stream) cudaStreamDestroy(ctx->gen_stream);
    if (ctx->gen_fork) cudaEventDestroy(ctx->gen_fork);
    for (int i = 0; i < 2; ++i) {
        if (ctx->gen_done[i]) cudaEventDestroy(ctx->gen_done[i]);
        if (ctx->solve_done[i]) cudaEventDestroy(ctx->solve_done[i]);
        if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
    }
    if (ctx->counters) cudaFree(ctx->counters);
    delete ctx;
    return DDB_OK;
}

extern "C" int ddb_device_info(ddb_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin) {
    if (!ctx) return fail(DDB_EINVAL, "ddb_device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (smem_optin) *smem_optin = ctx->smem_optin;
    return DDB_OK;
}

extern "C" int64_t ddb_launch_count(ddb_ctx* ctx) { return ctx ? ctx->launches : 0; }

static int auto_plan(const ddb_ctx* ctx, int m, int n) {
    // column-block-per-warp register kernel: measured faster than the row-per-thread kernel wherever that one needs its
    // 101-column variants (n >= 72): 1.17x at (144,72) ... 1.27x at (200,100), 1.30x at (228,100); slower below (0.96x at (120,60))
    static const bool no_quadcol = [] { const char* e = getenv("DDB_NO_QUADCOL"); return e && e[0] == '1'; }();
    if (!no_quadcol && n >= 72 && ddb::quadcol_supported(m, n)) return 7;
    if (ddb::rowreg_supported(m, n) || ddb::regtile_supported(m, n)) return 0;
    if ((int64_t)ddb::generic_smem_bytes(m, n, true) <= ctx->smem_optin) return 1;
    static const bool no_cluster = [] { const char* e = getenv("DDB_NO_CLUSTER"); return e && e[0] == '1'; }();
    if (!no_cluster && ddb::cluster_supported(m, n)) return 6;
    return 2;
}

extern "C" int ddb_solve_plan(ddb_ctx* ctx, int m, int n) {
    if (!ctx) return fail(DDB_EINVAL, "ddb_solve_plan: ctx is NULL");
    if (m < 1 || n < 1) return fail(DDB_EINVAL, "ddb_solve_plan: m=%d n=%d", m, n);
    if (n > 512) return fail(DDB_EUNSUPPORTED, "ddb_solve_plan: n=%d > 512 is not supported yet", n);
    int plan = ctx->forced_plan >= 0 ? ctx->forced_plan : auto_plan(ctx, m, n);
    if (plan == 0 && !ddb::rowreg_supported(m, n) && !ddb::regtile_supported(m, n))
        return fail(DDB_EUNSUPPORTED, "register-tiled kernel does not cover m=%d n=%d", m, n);
    if (plan == 1 && (int64_t)ddb::generic_smem_bytes(m, n, true) > ctx->smem_optin)
        return fail(DDB_EUNSUPPORTED, "shared-memory tableau does not fit for m=%d n=%d", m, n);
    if (plan == 3 && !ddb::tile2d_supported(m, n))
        return fail(DDB_EUNSUPPORTED, "2-D register-tile kernel does not cover m=%d n=%d", m, n);
    if (plan == 4 && !ddb::regtile_supported(m, n))
        return fail(DDB_EUNSUPPORTED, "warp-tiled register kernel does not cover m=%d n=%d", m, n);
    if (plan == 5 && !ddb::rowpipe_supported(m, n))
        return fail(DDB_EUNSUPPORTED, "software-pipelined row-per-thread kernel does not cover m=%d n=%d", m, n);
    if (plan == 6 && !ddb::cluster_supported(m, n))
        return fail(DDB_EUNSUPPORTED, "thread-block-cluster kernel does not cover m=%d n=%d", m, n);
    if (plan == 7 && !ddb::quadcol_supported(m, n))
        return fail(DDB_EUNSUPPORTED, "column-block-per-warp register kernel does not cover m=%d n=%d", m, n);
    return plan;
}

extern "C" int ddb_set_fused_mode(ddb_ctx* ctx, int mode) {
    if (!ctx) return fail(DDB_EINVAL, "ddb_set_fused_mode: ctx is NULL");
    if (mode < 0 || mode > 2) return fail(DDB_EINVAL, "ddb_set_fused_mode: mode %d", mode);
    ctx->fused_mode = mode;
    return DDB_OK;
}

extern "C" int ddb_set_solve_plan(ddb_ctx* ctx, int plan) {
    if (!ctx) return fail(DDB_EINVAL, "ddb_set_solve_plan: ctx is NULL");
    if (plan < -1 || plan > 7) return fail(DDB_EINVAL, "ddb_set_solve_plan: plan %d", plan);
    ctx->forced_plan = plan;
    return DDB_OK;
}

extern "C" int ddb_generate_dev(ddb_ctx* ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n,
                                double density, double* A, double* b, double* c, double* x0, void* stream) {
    if (!ctx || !A || !b || !c) return fail(DDB_EINVAL, "ddb_generate_dev: NULL argument");
    if (B < 0 || m < 1 || n < 1 || !(density > 0.0 && density <= 1.0))
        return fail(DDB_EINVAL, "ddb_generate_dev: B=%lld m=%d n=%d density=%g", (long long)B, m, n, density);
    if (B == 0) return DDB_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    int launches = 0;
    CUDA_TRY(ddb::launch_generate(key, first_instance, B, m, n, density, A, b, c, x0, ctx->sm_count,
                                  (cudaStream_t)stream, &launches));
    ctx->launches += launches;
    return DDB_OK;
}


// ---------------------------------------------------------------------------------------------------------
// per-context scratch (parked rows / global tableau slabs, crash inverses, fused-mode instance slabs): one launch at a
// time uses it, whatever stream it runs on -- every launch waits for the previous one's event
// ---------------------------------------------------------------------------------------------------------
static int scratch_acquire(ddb_ctx* ctx, cudaStream_t st, size_t need_scratch, size_t need_dscr, size_t need_slab) {
    if (ctx->scratch_in_use) CUDA_TRY(cudaStreamWaitEvent(st, ctx->scratch_free, 0));
    if (need_scratch > ctx->scratch.cap || need_dscr > ctx->dscr.cap || need_slab > ctx->slab.cap)
        CUDA_TRY(cudaDeviceSynchronize());          // growing: nothing may still be running on the old block
    int rc;
    if ((rc = ensure(ctx->scratch, need_scratch))) return rc;
    if ((rc = ensure(ctx->dscr, need_dscr))) return rc;
    if ((rc = ensure(ctx->slab, need_slab))) return rc;
    return DDB_OK;
}
static int scratch_release(ddb_ctx* ctx, cudaStream_t st) {
    CUDA_TRY(cudaEventRecord(ctx->scratch_free, st));
    ctx->scratch_in_use = true;
    return DDB_OK;
}

static long long generic_grid(const ddb_ctx* ctx, int m, int n, int plan, long long B) {
    const bool smem_tab = (plan == 1);
    const int block = ddb::generic_block_threads(m, n, smem_tab);
    const size_t smem = ddb::generic_smem_bytes(m, n, smem_tab);
    // persistent grid: as many CTAs as are co-resident (bounded by shared memory and threads), never more than B
    int per_sm = (int)((size_t)ctx->smem_optin / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int by_threads = 2048 / block;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm > 16) per_sm = 16;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (!smem_tab) {
        // global-memory tableau: every pivot streams the live part of a CTA's slab, so the slabs of all co-resident CTAs
        // should stay in the 126 MB L2 -- cap the grid by an L2 budget (DDB_PLAN2_L2_MB overrides, 0 = no cap)
        static const long long l2_mb = [] { const char* e = getenv("DDB_PLAN2_L2_MB"); return e ? atoll(e) : 96ll; }();
        if (l2_mb > 0) {
            long long cap = (l2_mb << 20) / ((long long)m * n * (long long)sizeof(double));
            if (cap < 8) cap = 8;
            if (grid > cap) grid = cap;
        }
    }
    if (grid > B) grid = B;
    return grid;
}

// scratch must have been acquired by the caller (generic_scratch_bytes / slab for a.gen)
static size_t generic_scratch_bytes(const ddb_ctx* ctx, int m, int n, int plan, long long B) {
    return plan == 1 ? 0 : (size_t)generic_grid(ctx, m, n, plan, B) * m * n * sizeof(double);
}

static int launch_generic(ddb_ctx* ctx, ddb::SolveArgs& a, int plan, cudaStream_t st) {
    const bool smem_tab = (plan == 1);
    const int block = ddb::generic_block_threads(a.m, a.n, smem_tab);
    const long long grid = generic_grid(ctx, a.m, a.n, plan, a.B);
    if (!smem_tab) a.gtab = (double*)ctx->scratch.p;
    CUDA_TRY(ddb::launch_simplex_generic(a, smem_tab, (int)grid, block, st));
    ctx->launches += 1;
    return DDB_OK;
}

struct GenSpec {              // fused mode: draw the instances inside the solver kernels
    uint64_t key;
    int64_t first;
    double density;
};

// Common body of ddb_solve_label_dev and the fused call.  gen == nullptr: A, b, c are the caller's instances.
// gen != nullptr: the solver CTAs draw instance gen->first + i themselves; A / b / c (all or none) receive them.
static int solve_launch(ddb_ctx* ctx, int64_t B, int m, int n, const double* A, const double* b, const double* c,
                        double threshold, const uint8_t* row_mask, int32_t* status, double* x, double* obj, uint8_t* labels,
                        int32_t* n_active, int32_t* pivots, int32_t* ties, int32_t* violations, const GenSpec* gen,
                        cudaStream_t st) {
    const int plan = ddb_solve_plan(ctx, m, n);
    if (plan < 0) return plan;

    ddb::SolveArgs a;
    a.m = m; a.n = n; a.B = B;
    a.A = A; a.b = b; a.c = c; a.row_mask = row_mask; a.thr = threshold;
    a.status = status; a.x = x; a.obj = obj; a.labels = labels; a.n_active = n_active;
    a.pivots = pivots; a.ties = ties; a.violations = violations;
    a.max_iter = 50 * (m + n);
    a.gtab = nullptr; a.dscr = nullptr; a.slab = nullptr;
    a.gen = gen ? 1 : 0;
    a.gen_key = gen ? gen->key : 0; a.gen_first = gen ? gen->first : 0; a.gen_density = gen ? gen->density : 1.0;
    const int slot = ctx->next_counter;
    ctx->next_counter = (ctx->next_counter + 1) % kCounters;
    a.counter = ctx->counters + 3 * slot;               // [0] plan-0 queue, [1] fix-up queue, [2] flag count
    a.flag_count = reinterpret_cast<int*>(ctx->counters + 3 * slot + 2);
    a.only_flagged = 0;
    CUDA_TRY(cudaMemsetAsync(a.counter, 0, 3 * sizeof(unsigned long long), st));
    const size_t per_lp = ((size_t)m * n + m + n) * sizeof(double);
    int rc;

    if (plan == 0 || plan == 3 || plan == 4 || plan == 5 || plan == 7) {
        // Register-resident kernels: row-per-thread (plan 0 default; hybrid register + shared-memory rows at (200,100)) and
        // the warp-tiled kernel (plan 4, the fallback of plan 0 for shapes the row kernel does not cover).  Plans 3 / 5
        // (2-D tile, software-pipelined rows) exist only in the `make experiments` build.
        // (shapes of the warp-tiled kernel, 100 < n <= 111 with few live rows, stay on it)
        int which = (ddb::rowreg_supported(m, n) && !(n > 100 && ddb::regtile_supported(m, n))) ? 1 : 2;
        if (plan == 3) which = 0;
        if (plan == 4) which = 2;
        if (plan == 5) which = 3;
        if (plan == 7) which = 4;
        // in-solver generator (the caller checked *_gen_supported): in the column-block kernel at its shapes, else in the
        // row-per-thread kernel
        const bool qgen = gen && plan == 7 && ddb::quadcol_gen_supported(m, n);
        if (gen) which = qgen ? 4 : 1;
        const int grid = qgen ? ddb::quadcol_gen_grid(ctx->sm_count)
                              : (which == 4 ? ddb::quadcol_grid(ctx->sm_count)
                                            : (gen ? ddb::rowreg_gen_grid(m, n, ctx->sm_count) : ddb::rowreg_grid(m, n, ctx->sm_count)));
        const int fplan = ((int64_t)ddb::generic_smem_bytes(m, n, true) <= ctx->smem_optin) ? 1 : 2;
        const long long fgrid = generic_grid(ctx, m, n, fplan, B);
        size_t need = (which == 2) ? ddb::regtile_scratch_bytes(m, n, ctx->sm_count)
                                   : (which == 1 ? ddb::rowreg_rows_scratch_bytes(m, n, grid) : 0);
        const size_t fneed = generic_scratch_bytes(ctx, m, n, fplan, B);
        if (fneed > need) need = fneed;
        const size_t need_d = (which == 1) ? ddb::rowreg_d_scratch_bytes(m, n, grid) : 0;
        const size_t need_slab = (gen && !A) ? (size_t)(grid > fgrid ? grid : fgrid) * ddb::slab_doubles(m, n) * sizeof(double) : 0;
        if ((rc = scratch_acquire(ctx, st, need, need_d, need_slab))) return rc;
        a.gtab = (double*)ctx->scratch.p;
        a.dscr = (double*)ctx->dscr.p;
        a.slab = (double*)ctx->slab.p;
        if (qgen)
            CUDA_TRY(ddb::launch_simplex_quadcol_gen(a, ctx->sm_count, st));
        else if (gen)
            CUDA_TRY(ddb::launch_simplex_rowreg_gen(a, ctx->sm_count, st));
        else if (which == 0)
            CUDA_TRY(ddb::launch_simplex_tile2d(a, ctx->sm_count, st));
        else if (which == 1)
            CUDA_TRY(ddb::launch_simplex_rowreg(a, ctx->sm_count, st));
        else if (which == 3)
            CUDA_TRY(ddb::launch_simplex_rowpipe(a, ctx->sm_count, st));
        else if (which == 4)
            CUDA_TRY(ddb::launch_simplex_quadcol(a, ctx->sm_count, st));
        else
            CUDA_TRY(ddb::launch_simplex_regtile(a, ctx->sm_count, st));
        ctx->launches += 1;
        // fix-up pass: instances the tile could not hold / whose static crash basis was singular were flagged
        // status = -1; the generic kernel re-solves exactly those (it returns at once when none were flagged).
        a.only_flagged = 1;
        a.counter = ctx->counters + 3 * slot + 1;
        rc = launch_generic(ctx, a, fplan, st);
        if (rc) return rc;
        return scratch_release(ctx, st);
    }

    if (plan == 6) {
        // Thread-block-cluster kernel: the live tableau in the distributed shared memory of 1-8 CTAs, one LP per cluster.
        // Instances it flags (singular static crash basis, ill-conditioned vertex) go to the generic kernel afterwards.
        const int fplan = ((int64_t)ddb::generic_smem_bytes(m, n, true) <= ctx->smem_optin) ? 1 : 2;
        size_t need = ddb::cluster_scratch_bytes(m, n, ctx->sm_count, B);
        const size_t fneed = generic_scratch_bytes(ctx, m, n, fplan, B);
        if (fneed > need) need = fneed;
        if ((rc = scratch_acquire(ctx, st, need, 0, 0))) return rc;
        a.gtab = (double*)ctx->scratch.p;
        CUDA_TRY(ddb::launch_simplex_cluster(a, ctx->sm_count, st));
        ctx->launches += 1;
        a.only_flagged = 1;
        a.counter = ctx->counters + 3 * slot + 1;
        rc = launch_generic(ctx, a, fplan, st);
        if (rc) return rc;
        return scratch_release(ctx, st);
    }

    if ((rc = scratch_acquire(ctx, st, generic_scratch_bytes(ctx, m, n, plan, B), 0, 0))) return rc;
    rc = launch_generic(ctx, a, plan, st);
    if (rc) return rc;
    return scratch_release(ctx, st);
}

extern "C" int ddb_solve_label_dev(ddb_ctx* ctx, int64_t B, int m, int n, const double* A, const double* b,
                                   const double* c, double threshold, const uint8_t* row_mask, int32_t* status,
                                   double* x, double* obj, uint8_t* labels, int32_t* n_active, int32_t* pivots,
                                   int32_t* ties, int32_t* violations, void* stream) {
    if (!ctx || !A || !b || !c || !status || !labels) return fail(DDB_EINVAL, "ddb_solve_label_dev: NULL argument");
    if (B < 0 || m < 1 || n < 1 || !(threshold >= 0.0))
        return fail(DDB_EINVAL, "ddb_solve_label_dev: B=%lld m=%d n=%d threshold=%g", (long long)B, m, n, threshold);
    if (B == 0) return DDB_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return solve_launch(ctx, B, m, n, A, b, c, threshold, row_mask, status, x, obj, labels, n_active, pivots, ties, violations,
                        nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer plumbing: pinned caller buffers are DMA'd directly; pageable ones go through the context's pinned staging
// ring (parallel memcpy on a few host threads into the slot's staging block, one cudaMemcpyAsync per array from there)
// ---------------------------------------------------------------------------------------------------------
static bool is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static int ensure_pinned(HostBuf& hb, size_t bytes) {
    if (bytes <= hb.cap) return DDB_OK;
    if (hb.p) cudaFreeHost(hb.p);
    hb.p = nullptr;
    hb.cap = 0;
    cudaError_t e = cudaHostAlloc(&hb.p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(DDB_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    hb.cap = bytes;
    return DDB_OK;
}

static int copy_threads() {
    static const int v = [] {
        const char* e = getenv("DDB_COPY_THREADS");
        int t = e ? atoi(e) : 0;
        if (t <= 0) {
            const unsigned hc = std::thread::hardware_concurrency();
            t = hc >= 4 ? (int)hc : 1;
            if (t > 16) t = 16;
        }
        return t > 32 ? 32 : t;
    }();
    return v;
}

// One thread's slice of a staging copy.  The destination is the pinned ring, which the CPU never reads back: non-temporal
// (streaming) stores skip the read-for-ownership of every destination line, i.e. a third of the host memory traffic of a
// plain memcpy (glibc only switches to them above a per-thread threshold of ~3/4 of the last-level cache).
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_copy_avx2(char* dst, const char* src, size_t bytes) {
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > bytes) head = bytes;
    memcpy(dst, src, head);
    dst += head; src += head; bytes -= head;
    const size_t blocks = bytes / 128;
    for (size_t i = 0; i < blocks; ++i) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src) + 0);
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src) + 1);
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src) + 2);
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src) + 3);
        _mm_prefetch(src + 1024, _MM_HINT_NTA);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst) + 0, a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst) + 1, b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst) + 2, c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst) + 3, d);
        src += 128; dst += 128;
    }
    _mm_sfence();
    memcpy(dst, src, bytes - blocks * 128);
}
#endif
static void slice_copy(char* dst, const char* src, size_t bytes) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("DDB_NO_STREAM_COPY");
    if (avx2 && bytes >= (size_t)(1u << 20)) {
        stream_copy_avx2(dst, src, bytes);
        return;
    }
#endif
    memcpy(dst, src, bytes);
}

// memcpy split over the host cores (at most 16): one core moves ~5-10 GB/s, PCIe 5 x16 wants ~55
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const int nt = copy_threads();
    if (nt <= 1 || bytes < (size_t)(8u << 20)) {
        slice_copy((char*)dst, (const char*)src, bytes);
        return;
    }
    const size_t slice = ((bytes / nt) + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        const size_t off = (size_t)t * slice;
        if (off >= bytes) break;
        const size_t nb = (bytes - off < slice) ? (bytes - off) : slice;
        th.emplace_back([=] { slice_copy((char*)dst + off, (const char*)src + off, nb); });
    }
    for (auto& t : th) t.join();
}

static int slot_wait_and_retire(Slot& s) {
    CUDA_TRY(cudaEventSynchronize(s.done));
    for (const Retire& r : s.retire) memcpy(r.dst, r.src, r.bytes);
    s.retire.clear();
    return DDB_OK;
}

// H2D of one array of a chunk: direct when the caller's memory is pinned, else through `stage` (advances stage_off)
static int h2d(Slot& s, void* dev, const void* host, size_t bytes, bool pinned, size_t& stage_off) {
    if (bytes == 0) return DDB_OK;
    if (pinned) {
        CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s.stream));
    } else {
        char* st = (char*)s.in.p + stage_off;
        parallel_memcpy(st, host, bytes);
        CUDA_TRY(cudaMemcpyAsync(dev, st, bytes, cudaMemcpyHostToDevice, s.stream));
        stage_off += (bytes + 255) & ~(size_t)255;
    }
    return DDB_OK;
}
// D2H of one output array of a chunk: direct when pinned, else into the slot's pinned out-block + a retire entry
static int d2h(Slot& s, cudaStream_t st, void* host, const void* dev, size_t bytes, bool pinned, size_t& stage_off) {
    if (!host || bytes == 0) return DDB_OK;
    if (pinned) {
        CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
    } else {
        char* stg = (char*)s.out.p + stage_off;
        CUDA_TRY(cudaMemcpyAsync(stg, dev, bytes, cudaMemcpyDeviceToHost, st));
        s.retire.push_back({host, stg, bytes});
        stage_off += (bytes + 255) & ~(size_t)255;
    }
    return DDB_OK;
}

struct HostOut {              // the caller's host output arrays and whether each is page-locked
    int32_t* status; double* x; double* obj; uint8_t* labels; int32_t* n_active; int32_t* pivots; int32_t* ties;
    int32_t* violations;
    bool pinned;              // all of them (one pageable array sends every output through staging: they are small)
};

static size_t out_bytes_per_lp(int m, int n) { return 4 + (size_t)n * 8 + 8 + m + 4 + 16 + 4 + 4; }   // status x obj labels n_active pivots ties violations

static int ensure_slot_outputs(Slot& s, int64_t nb, int m, int n, const HostOut& o) {
    int rc;
    if ((rc = ensure(s.status, (size_t)nb * 4))) return rc;
    if ((rc = ensure(s.labels, (size_t)nb * m))) return rc;
    if (o.x && (rc = ensure(s.x, (size_t)nb * n * 8))) return rc;
    if (o.obj && (rc = ensure(s.obj, (size_t)nb * 8))) return rc;
    if (o.n_active && (rc = ensure(s.nact, (size_t)nb * 4))) return rc;
    if (o.pivots && (rc = ensure(s.piv, (size_t)nb * 16))) return rc;
    if (o.ties && (rc = ensure(s.ties, (size_t)nb * 4))) return rc;
    if (o.violations && (rc = ensure(s.viol, (size_t)nb * 4))) return rc;
    if (!o.pinned && (rc = ensure_pinned(s.out, (size_t)nb * (out_bytes_per_lp(m, n) + 8) + 8 * 256))) return rc;
    return DDB_OK;
}

static int slot_d2h_outputs(Slot& s, int64_t off, int64_t nb, int m, int n, const HostOut& o) {
    size_t so = 0;
    int rc;
    cudaStream_t st = s.stream;
    if ((rc = d2h(s, st, o.status + off, s.status.p, (size_t)nb * 4, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.labels + (size_t)off * m, s.labels.p, (size_t)nb * m, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.x ? o.x + (size_t)off * n : nullptr, s.x.p, (size_t)nb * n * 8, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.obj ? o.obj + off : nullptr, s.obj.p, (size_t)nb * 8, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.n_active ? o.n_active + off : nullptr, s.nact.p, (size_t)nb * 4, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.pivots ? o.pivots + (size_t)off * 4 : nullptr, s.piv.p, (size_t)nb * 16, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.ties ? o.ties + off : nullptr, s.ties.p, (size_t)nb * 4, o.pinned, so))) return rc;
    if ((rc = d2h(s, st, o.violations ? o.violations + off : nullptr, s.viol.p, (size_t)nb * 4, o.pinned, so))) return rc;
    return DDB_OK;
}

static HostOut make_host_out(int32_t* status, double* x, double* obj, uint8_t* labels, int32_t* n_active, int32_t* pivots,
                             int32_t* ties, int32_t* violations) {
    HostOut o{status, x, obj, labels, n_active, pivots, ties, violations, true};
    o.pinned = is_pinned(status) && is_pinned(x) && is_pinned(obj) && is_pinned(labels) && is_pinned(n_active) &&
               is_pinned(pivots) && is_pinned(ties) && is_pinned(violations);
    return o;
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer flavour: chunked, kSlots chunks in flight (stage-in and H2D of chunk k+1 overlap the solve of chunk k)
// ---------------------------------------------------------------------------------------------------------
extern "C" int ddb_solve_label_host(ddb_ctx* ctx, int64_t B, int m, int n, const double* A, const double* b,
                                    const double* c, double threshold, const uint8_t* row_mask, int32_t* status,
                                    double* x, double* obj, uint8_t* labels, int32_t* n_active, int32_t* pivots,
                                    int32_t* ties, int32_t* violations) {
    if (!ctx || !A || !b || !c || !status || !labels) return fail(DDB_EINVAL, "ddb_solve_label_host: NULL argument");
    if (B < 0 || m < 1 || n < 1) return fail(DDB_EINVAL, "ddb_solve_label_host: B=%lld m=%d n=%d", (long long)B, m, n);
    if (B == 0) return DDB_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t per_lp = ((size_t)m * n + m + n) * sizeof(double);
    // chunks of ~192 MB (DDB_HOST_CHUNK_MB): the copy engine is the bottleneck (PCIe ~52 GB/s vs 8(mn+m+n) bytes per LP), so
    // the pipeline fill (first H2D) and drain (last solve) are what chunking can shrink; at least 8 LPs per SM keep the
    // solver's persistent CTAs balanced
    static const long long chunk_mb = [] { const char* e = getenv("DDB_HOST_CHUNK_MB"); return e ? atoll(e) : 192ll; }();
    int64_t chunk = (int64_t)((size_t)(chunk_mb << 20) / per_lp);
    const int64_t min_chunk = (int64_t)ctx->sm_count * 8;
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > B) chunk = B;
    const bool pinA = is_pinned(A), pinb = is_pinned(b), pinc = is_pinned(c), pinm = is_pinned(row_mask);
    const HostOut o = make_host_out(status, x, obj, labels, n_active, pivots, ties, violations);
    int rc = DDB_OK;
    int k = 0;
    for (int64_t off = 0; off < B; off += chunk, ++k) {
        const int64_t nb = (B - off < chunk) ? (B - off) : chunk;
        Slot& s = ctx->slots[k % kSlots];
        if ((rc = slot_wait_and_retire(s))) return rc;   // previous use of this slot has drained
        if ((rc = ensure(s.A, (size_t)nb * m * n * 8))) return rc;
        if ((rc = ensure(s.b, (size_t)nb * m * 8))) return rc;
        if ((rc = ensure(s.c, (size_t)nb * n * 8))) return rc;
        if (row_mask && (rc = ensure(s.mask, (size_t)nb * m))) return rc;
        if ((rc = ensure_slot_outputs(s, nb, m, n, o))) return rc;
        size_t stage_need = 1024;
        if (!pinA) stage_need += (size_t)nb * m * n * 8 + 256;
        if (!pinb) stage_need += (size_t)nb * m * 8 + 256;
        if (!pinc) stage_need += (size_t)nb * n * 8 + 256;
        if (row_mask && !pinm) stage_need += (size_t)nb * m + 256;
        if (stage_need > 1024 && (rc = ensure_pinned(s.in, stage_need))) return rc;
        size_t so = 0;
        if ((rc = h2d(s, s.A.p, A + (size_t)off * m * n, (size_t)nb * m * n * 8, pinA, so))) return rc;
        if ((rc = h2d(s, s.b.p, b + (size_t)off * m, (size_t)nb * m * 8, pinb, so))) return rc;
        if ((rc = h2d(s, s.c.p, c + (size_t)off * n, (size_t)nb * n * 8, pinc, so))) return rc;
        if (row_mask && (rc = h2d(s, s.mask.p, row_mask + (size_t)off * m, (size_t)nb * m, pinm, so))) return rc;
        rc = solve_launch(ctx, nb, m, n, (const double*)s.A.p, (const double*)s.b.p, (const double*)s.c.p, threshold,
                          row_mask ? (const uint8_t*)s.mask.p : nullptr, (int32_t*)s.status.p,
                          x ? (double*)s.x.p : nullptr, obj ? (double*)s.obj.p : nullptr, (uint8_t*)s.labels.p,
                          n_active ? (int32_t*)s.nact.p : nullptr, pivots ? (int32_t*)s.piv.p : nullptr,
                          ties ? (int32_t*)s.ties.p : nullptr, violations ? (int32_t*)s.viol.p : nullptr, nullptr, s.stream);
        if (rc) return rc;
        if ((rc = slot_d2h_outputs(s, off, nb, m, n, o))) return rc;
        CUDA_TRY(cudaEventRecord(s.done, s.stream));
    }
    for (int i = 0; i < kSlots; ++i)
        if ((rc = slot_wait_and_retire(ctx->slots[i]))) return rc;
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fused generate -> solve -> label
// ---------------------------------------------------------------------------------------------------------
static int fused_launch(ddb_ctx* ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n, double density,
                        double threshold, int32_t* status, double* x, double* obj, uint8_t* labels, int32_t* n_active,
                        int32_t* pivots, int32_t* ties, int32_t* violations, double* A_out, double* b_out, double* c_out,
                        cudaStream_t st) {
    const bool keep = A_out && b_out && c_out;
    // IN-KERNEL GENERATION (even n, shapes of the row-per-thread kernel): ONE launch; every solver CTA draws its instance
    // itself (Philox, counter = global instance index) into a per-CTA slab that lives in L2 -- A never travels through HBM
    // and no generator kernel runs.  Measured on B200 at (200,100), 262 144 instances: 403 k LP/s, against 430 k LP/s for
    // the generator kernel + solver kernel pipeline below (512 MB chunks): the solver kernel is latency-bound at ~11
    // warp-cycles per issued instruction, so the generator's instructions cost more inside it (+14 %) than in a kernel of
    // their own at full occupancy (+8.5 %), and the HBM round trip of A (2 x 163 KB per LP = 2 % of the HBM bandwidth) is
    // cheap.  So automatic mode takes the two-kernel pipeline; ddb_set_fused_mode(ctx, 1) or DDB_FUSED_INKERNEL=1 selects the
    // in-kernel generator (same bits, same results -- tests/test_gpu_solve.py checks both).
    static const int env_mode = [] { const char* e = getenv("DDB_FUSED_INKERNEL"); return e ? (e[0] == '0' ? 2 : 1) : 0; }();
    const int mode = ctx->fused_mode ? ctx->fused_mode : env_mode;
    const bool inkernel = (mode == 1);
    const int plan = ddb_solve_plan(ctx, m, n);
    if (plan < 0) return plan;
    if (inkernel && (plan == 0 || plan == 7) && ctx->forced_plan < 0 &&
        ((plan == 7 && ddb::quadcol_gen_supported(m, n)) || ddb::rowreg_gen_supported(m, n)) &&
        (!keep || ((reinterpret_cast<uintptr_t>(A_out) | reinterpret_cast<uintptr_t>(c_out)) & 15) == 0)) {
        GenSpec g{key, first_instance, density};
        return solve_launch(ctx, B, m, n, keep ? A_out : nullptr, keep ? b_out : nullptr, keep ? c_out : nullptr, threshold,
                            nullptr, status, x, obj, labels, n_active, pivots, ties, violations, &g, st);
    }
    // Fallback (odd n, shapes outside plan 0): generator kernel -> chunk buffers in context scratch -> solver, chunk i + 1
    // generated on a side stream while chunk i is solved.
    const size_t per_lp = ((size_t)m * n + m + n) * sizeof(double);
    // (measured on plan 7, 262 144 instances of (200,100): 512 MB 548 k LP/s, 1 GB 554 k, 2 GB 557 k, 4 GB 559 k, 8 GB 561 k)
    static const long long chunk_mb = [] { const char* e = getenv("DDB_FUSED_CHUNK_MB"); return e ? atoll(e) : 2048ll; }();
    int64_t chunk = (int64_t)((size_t)(chunk_mb << 20) / per_lp);
    const int64_t min_chunk = (int64_t)ctx->sm_count * 4;
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > B) chunk = B;
    int rc;
    if (!keep) {
        // the chunk buffers are per-context: wait for the last solve that read them, whatever stream it ran on
        if (ctx->gen_in_use) CUDA_TRY(cudaEventSynchronize(ctx->gen_free));
        if ((rc = ensure(ctx->genA, (size_t)2 * chunk * m * n * 8))) return rc;
        if ((rc = ensure(ctx->genb, (size_t)2 * chunk * m * 8))) return rc;
        if ((rc = ensure(ctx->genc, (size_t)2 * chunk * n * 8))) return rc;
    }
    CUDA_TRY(cudaEventRecord(ctx->gen_fork, st));   // the side stream starts after whatever the caller has queued so far
    CUDA_TRY(cudaStreamWaitEvent(ctx->gen_stream, ctx->gen_fork, 0));
    int64_t idx = 0;
    for (int64_t off = 0; off < B; off += chunk, ++idx) {
        const int64_t nb = (B - off < chunk) ? (B - off) : chunk;
        const int h = (int)(idx & 1);
        double* Ap = keep ? A_out + (size_t)off * m * n : (double*)ctx->genA.p + (size_t)h * chunk * m * n;
        double* bp = keep ? b_out + (size_t)off * m : (double*)ctx->genb.p + (size_t)h * chunk * m;
        double* cp = keep ? c_out + (size_t)off * n : (double*)ctx->genc.p + (size_t)h * chunk * n;
        if (!keep && idx >= 2) CUDA_TRY(cudaStreamWaitEvent(ctx->gen_stream, ctx->solve_done[h], 0));   // half h is free again
        int launches = 0;
        CUDA_TRY(ddb::launch_generate(key, first_instance + off, nb, m, n, density, Ap, bp, cp, nullptr, ctx->sm_count,
                                      ctx->gen_stream, &launches));
        ctx->launches += launches;
        CUDA_TRY(cudaEventRecord(ctx->gen_done[h], ctx->gen_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, ctx->gen_done[h], 0));
        rc = solve_launch(ctx, nb, m, n, Ap, bp, cp, threshold, nullptr, status + off, x ? x + (size_t)off * n : nullptr,
                          obj ? obj + off : nullptr, labels + (size_t)off * m, n_active ? n_active + off : nullptr,
                          pivots ? pivots + (size_t)off * 4 : nullptr, ties ? ties + off : nullptr,
                          violations ? violations + off : nullptr, nullptr, st);
        if (rc) return rc;
        if (!keep) CUDA_TRY(cudaEventRecord(ctx->solve_done[h], st));
    }
    if (!keep) {
        CUDA_TRY(cudaEventRecord(ctx->gen_free, st));
        ctx->gen_in_use = true;
    }
    return DDB_OK;
}

extern "C" int ddb_generate_solve_label_dev(ddb_ctx* ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n,
                                            double density, double threshold, int32_t* status, double* x, double* obj,
                                            uint8_t* labels, int32_t* n_active, int32_t* pivots, int32_t* ties,
                                            int32_t* violations, double* A_out, double* b_out, double* c_out, void* stream) {
    if (!ctx || !status || !labels) return fail(DDB_EINVAL, "ddb_generate_solve_label_dev: NULL argument");
    if (B < 0 || m < 1 || n < 1 || !(density > 0.0 && density <= 1.0) || !(threshold >= 0.0))
        return fail(DDB_EINVAL, "ddb_generate_solve_label_dev: B=%lld m=%d n=%d density=%g threshold=%g", (long long)B, m, n,
                    density, threshold);
    if ((A_out || b_out || c_out) && !(A_out && b_out && c_out))
        return fail(DDB_EINVAL, "ddb_generate_solve_label_dev: A_out, b_out, c_out must be given together");
    if (B == 0) return DDB_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return fused_launch(ctx, key, first_instance, B, m, n, density, threshold, status, x, obj, labels, n_active, pivots, ties,
                        violations, A_out, b_out, c_out, (cudaStream_t)stream);
}

// Fused call with HOST outputs: what the reference's generate -> solve -> label loop hands to its caller (labels, status,
// objective, x, ...) lands in host arrays; nothing travels host -> device.  Chunks of DDB_FUSED_HOST_CHUNK instances:
// the D2H of chunk k overlaps the solve of chunk k + 1.
extern "C" int ddb_generate_solve_label_host(ddb_ctx* ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n,
                                             double density, double threshold, int32_t* status, double* x, double* obj,
                                             uint8_t* labels, int32_t* n_active, int32_t* pivots, int32_t* ties,
                                             int32_t* violations) {
    if (!ctx || !status || !labels) return fail(DDB_EINVAL, "ddb_generate_solve_label_host: NULL argument");
    if (B < 0 || m < 1 || n < 1 || !(density > 0.0 && density <= 1.0) || !(threshold >= 0.0))
        return fail(DDB_EINVAL, "ddb_generate_solve_label_host: B=%lld m=%d n=%d density=%g threshold=%g", (long long)B, m, n,
                    density, threshold);
    if (B == 0) return DDB_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    static const long long chunk_lps = [] { const char* e = getenv("DDB_FUSED_HOST_CHUNK"); return e ? atoll(e) : 16384ll; }();
    int64_t chunk = chunk_lps;
    const int64_t min_chunk = (int64_t)ctx->sm_count * 8;
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > B) chunk = B;
    const HostOut o = make_host_out(status, x, obj, labels, n_active, pivots, ties, violations);
    int rc = DDB_OK;
    int k = 0;
    for (int64_t off = 0; off < B; off += chunk, ++k) {
        const int64_t nb = (B - off < chunk) ? (B - off) : chunk;
        Slot& s = ctx->slots[k % kSlots];
        if ((rc = slot_wait_and_retire(s))) return rc;
        if ((rc = ensure_slot_outputs(s, nb, m, n, o))) return rc;
        rc = fused_launch(ctx, key, first_instance + off, nb, m, n, density, threshold, (int32_t*)s.status.p,
                          x ? (double*)s.x.p : nullptr, obj ? (double*)s.obj.p : nullptr, (uint8_t*)s.labels.p,
                          n_active ? (int32_t*)s.nact.p : nullptr, pivots ? (int32_t*)s.piv.p : nullptr,
                          ties ? (int32_t*)s.ties.p : nullptr, violations ? (int32_t*)s.viol.p : nullptr, nullptr, nullptr,
                          nullptr, s.stream);
        if (rc) return rc;
        if ((rc = slot_d2h_outputs(s, off, nb, m, n, o))) return rc;
        CUDA_TRY(cudaEventRecord(s.done, s.stream));
    }
    for (int i = 0; i < kSlots; ++i)
        if ((rc = slot_wait_and_retire(ctx->slots[i]))) return rc;
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// classifier forward
// ---------------------------------------------------------------------------------------------------------
extern "C" int ddb_s2v_param_count(int graph, int p) {
    if (p < 1) return DDB_EINVAL;
    if (graph == 0) return 2 * p + 6 * p * p + 3 * p + 3 * p * p + 2 * 2 * p;
    if (graph == 1) return p + 4 * p + p + 4 * p * p + 2 * p + 3 * p * p + 2 * (2 * p + 4);
    return DDB_EINVAL;
}

static int s2v_forward_impl(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A, const double* b,
                            const double* c, const float* params, const uint8_t* row_ineq, const uint8_t* row_bound, float* logp,
                            float* probs, void* stream);

extern "C" int ddb_s2v_forward_dev(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A,
                                   const double* b, const double* c, const float* params, float* logp, float* probs,
                                   void* stream) {
    return s2v_forward_impl(ctx, graph, B, m, n, p, T, A, b, c, params, nullptr, nullptr, logp, probs, stream);
}

extern "C" int ddb_s2v_forward_flags_dev(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A,
                                         const double* b, const double* c, const float* params, const uint8_t* row_ineq,
                                         const uint8_t* row_bound, float* logp, float* probs, void* stream) {
    if (graph == 1 && ((row_ineq == nullptr) != (row_bound == nullptr)))
        return fail(DDB_EINVAL, "ddb_s2v_forward_flags_dev: bipartite items carry both row flags or none");
    return s2v_forward_impl(ctx, graph, B, m, n, p, T, A, b, c, params, row_ineq, graph == 1 ? row_bound : nullptr, logp, probs,
                            stream);
}

static int s2v_forward_impl(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A, const double* b,
                            const double* c, const float* params, const uint8_t* row_ineq, const uint8_t* row_bound, float* logp,
                            float* probs, void* stream) {
    if (!ctx || !A || !b || !c || !params || !logp) return fail(DDB_EINVAL, "ddb_s2v_forward_dev: NULL argument");
    if (graph != 0 && graph != 1) return fail(DDB_EINVAL, "ddb_s2v_forward_dev: Graph not recognised (%d)", graph);
    if (B < 0 || m < 1 || n < 1 || p < 1 || T < 0)
        return fail(DDB_EINVAL, "ddb_s2v_forward_dev: B=%lld m=%d n=%d p=%d T=%d", (long long)B, m, n, p, T);
    if (B == 0) return DDB_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int slot = ctx->next_counter;
    ctx->next_counter = (ctx->next_counter + 1) % kCounters;
    int* err = reinterpret_cast<int*>(ctx->counters + 3 * slot);
    CUDA_TRY(cudaMemsetAsync(err, 0, sizeof(int), st));
    ddb::S2vArgs a;
    a.graph = graph; a.B = B; a.m = m; a.n = n; a.p = p; a.T = T;
    a.A = A; a.b = b; a.c = c; a.params = params; a.logp = logp; a.probs = probs;
    a.error_flag = err; a.store_A = 0;
    a.gram = nullptr; a.gram_pitch = 0;
    a.inst_flag = nullptr; a.flag_count = nullptr; a.only_flagged = 0;
    a.row_ineq = row_ineq; a.row_bound = row_bound;
    // bipartite variant: dense instances (the reference's distribution) go through the HBM-streaming kernel; it flags
    // instances with zero coefficients and the general kernel below then processes exactly those.  Items with row flags
    // (MPS / PLNN: equality rows, bound rows -- sparse by nature) go to the general kernel directly.
    static const bool no_dense = [] { const char* e = getenv("DDB_S2V_NO_DENSE"); return e && e[0] == '1'; }();
    if (graph == 1 && !no_dense && !row_ineq && ddb::s2v_bipartite_dense_supported(m, n, p, A, ctx->smem_optin)) {
        const size_t need = (size_t)B * sizeof(int);
        if (need > ctx->s2vflag.cap) CUDA_TRY(cudaStreamSynchronize(st));
        int rc = ensure(ctx->s2vflag, need);
        if (rc) return rc;
        a.inst_flag = (int*)ctx->s2vflag.p;
        a.flag_count = err + 1;
        CUDA_TRY(cudaMemsetAsync(a.inst_flag, 0, need, st));
        CUDA_TRY(cudaMemsetAsync(a.flag_count, 0, sizeof(int), st));
        CUDA_TRY(ddb::launch_s2v_bipartite_dense(a, ctx->sm_count, st));
        ctx->launches += 1;
        a.only_flagged = 1;
    }
    // complete variant: the Gram product W = G G^T (the only dense contraction of the forward) runs on the tensor cores
    // (tcgen05 kind::tf32, 3xTF32) when m + 1 <= 256; larger shapes keep the fused CUDA-core Gram inside the forward
    static const bool no_tc = [] { const char* e = getenv("DDB_S2V_NO_TC"); return e && e[0] == '1'; }();
    if (graph == 0 && !no_tc && ddb::s2v_gram_tc_supported(m, n)) {
        const size_t need = (size_t)B * ddb::s2v_gram_out_floats(m) * sizeof(float);
        if (need > ctx->gram.cap) CUDA_TRY(cudaStreamSynchronize(st));
        int rc = ensure(ctx->gram, need);
        if (rc) return rc;
        CUDA_TRY(ddb::launch_s2v_gram_tc(B, m, n, A, b, c, (float*)ctx->gram.p, ctx->sm_count, st));
        ctx->launches += 1;
        a.gram = (const float*)ctx->gram.p;
        a.gram_pitch = (int)(ddb::s2v_gram_out_floats(m) / 3);
    }
    const char* why = "";
    cudaError_t e = ddb::launch_s2v_forward(a, ctx->sm_count, ctx->smem_optin, st, &why);
    if (e != cudaSuccess) {
        if (why[0]) return fail(DDB_EUNSUPPORTED, "%s (m=%d n=%d p=%d)", why, m, n, p);
        return fail(DDB_ECUDA, "s2v forward launch: %s", cudaGetErrorString(e));
    }
    ctx->launches += 1;
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// classifier loss + gradient (training step of the reference: src/ml/train.py:59-66, criterion src/benchmark.py:70-75)
// ---------------------------------------------------------------------------------------------------------
static int s2v_loss_grad_impl(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A, const double* b,
                              const double* c, const float* params, const uint8_t* labels, const uint8_t* row_ineq,
                              const uint8_t* row_bound, float w0, float w1, float* grad, double* loss, int32_t* not_dense,
                              void* stream);

extern "C" int ddb_s2v_loss_grad_dev(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A,
                                     const double* b, const double* c, const float* params, const uint8_t* labels,
                                     float w0, float w1, float* grad, double* loss, int32_t* not_dense, void* stream) {
    return s2v_loss_grad_impl(ctx, graph, B, m, n, p, T, A, b, c, params, labels, nullptr, nullptr, w0, w1, grad, loss, not_dense,
                              stream);
}

extern "C" int ddb_s2v_loss_grad_flags_dev(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A,
                                           const double* b, const double* c, const float* params, const uint8_t* labels,
                                           const uint8_t* row_ineq, const uint8_t* row_bound, float w0, float w1, float* grad,
                                           double* loss, int32_t* not_dense, void* stream) {
    if (graph == 1 && ((row_ineq == nullptr) != (row_bound == nullptr)))
        return fail(DDB_EINVAL, "ddb_s2v_loss_grad_flags_dev: bipartite items carry both row flags or none");
    return s2v_loss_grad_impl(ctx, graph, B, m, n, p, T, A, b, c, params, labels, row_ineq, graph == 1 ? row_bound : nullptr, w0, w1,
                              grad, loss, not_dense, stream);
}

static int s2v_loss_grad_impl(ddb_ctx* ctx, int graph, int64_t B, int m, int n, int p, int T, const double* A, const double* b,
                              const double* c, const float* params, const uint8_t* labels, const uint8_t* row_ineq,
                              const uint8_t* row_bound, float w0, float w1, float* grad, double* loss, int32_t* not_dense,
                              void* stream) {
    if (!ctx || !A || !b || !c || !params || !labels || !grad || !loss || !not_dense)
        return fail(DDB_EINVAL, "ddb_s2v_loss_grad_dev: NULL argument");
    if (graph != 0 && graph != 1) return fail(DDB_EINVAL, "ddb_s2v_loss_grad_dev: Graph not recognised (%d)", graph);
    if (B < 0 || m < 1 || n < 1 || p < 1 || T < 0)
        return fail(DDB_EINVAL, "ddb_s2v_loss_grad_dev: B=%lld m=%d n=%d p=%d T=%d", (long long)B, m, n, p, T);
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int npar = ddb_s2v_param_count(graph, p);
    CUDA_TRY(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)npar, st));
    CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(not_dense, 0, sizeof(int32_t), st));
    if (B == 0) return DDB_OK;
    if (graph == 0) {
        // complete variant: relu row sums of W = G G^T from the tcgen05 Gram kernel (W depends on the data only), then
        // forward + hand-written backward on them (csrc/s2v_complete_backward.cu)
        if (!ddb::s2v_gram_tc_supported(m, n) || (long long)ddb::s2v_complete_grad_smem_bytes(m, p, T) > ctx->smem_optin || p > 64)
            return fail(DDB_EUNSUPPORTED, "classifier backward (complete): embeddings do not fit in shared memory (m=%d n=%d p=%d T=%d)", m, n, p, T);
        const size_t need = (size_t)B * ddb::s2v_gram_out_floats(m) * sizeof(float);
        if (need > ctx->gram.cap) CUDA_TRY(cudaStreamSynchronize(st));
        int rc = ensure(ctx->gram, need);
        if (rc) return rc;
        CUDA_TRY(ddb::launch_s2v_gram_tc(B, m, n, A, b, c, (float*)ctx->gram.p, ctx->sm_count, st));
        ctx->launches += 1;
        ddb::S2vCGradArgs g;
        g.B = B; g.m = m; g.p = p; g.T = T;
        g.gram = (const float*)ctx->gram.p; g.gram_pitch = (int)(ddb::s2v_gram_out_floats(m) / 3);
        g.params = params; g.labels = labels; g.row_ineq = row_ineq; g.w0 = w0; g.w1 = w1; g.grad = grad; g.loss = loss;
        const char* why = "";
        cudaError_t e = ddb::launch_s2v_complete_grad(g, ctx->sm_count, ctx->smem_optin, st, &why);
        if (e != cudaSuccess) {
            if (why[0]) return fail(DDB_EUNSUPPORTED, "%s (m=%d n=%d p=%d T=%d)", why, m, n, p, T);
            return fail(DDB_ECUDA, "s2v backward launch: %s", cudaGetErrorString(e));
        }
        ctx->launches += 1;
        return DDB_OK;
    }
    ddb::S2vGradArgs a;
    a.B = B; a.m = m; a.n = n; a.p = p; a.T = T;
    a.A = A; a.b = b; a.c = c; a.params = params; a.labels = labels; a.w0 = w0; a.w1 = w1;
    a.grad = grad; a.loss = loss; a.error_flag = not_dense;
    // dense instances (the reference's distribution) go through the streaming kernel, which flags every instance with a zero
    // coefficient; the general-adjacency kernel then adds exactly those (3 us when there are none)
    const int ggrid = ddb::s2v_general_grad_grid(B, ctx->sm_count);
    const size_t need_flag = (size_t)(B + 1) * sizeof(int);
    const size_t need_scr = (size_t)ggrid * ddb::s2v_general_grad_scratch_floats(m, n, p, T) * sizeof(float);
    if (need_flag > ctx->s2vflag.cap || need_scr > ctx->s2vgscr.cap) CUDA_TRY(cudaStreamSynchronize(st));
    int rc = ensure(ctx->s2vflag, need_flag);
    if (rc) return rc;
    if ((rc = ensure(ctx->s2vgscr, need_scr))) return rc;
    a.inst_flag = (int*)ctx->s2vflag.p;
    CUDA_TRY(cudaMemsetAsync(a.inst_flag, 0, need_flag, st));
    const char* why = "";
    static const bool no_dense_grad = [] { const char* e = getenv("DDB_S2V_NO_DENSE"); return e && e[0] == '1'; }();
    bool all_general = no_dense_grad || row_ineq != nullptr;      // items with row flags: the general-adjacency kernel takes them all
    if (!all_general) {
        cudaError_t e = ddb::launch_s2v_bipartite_grad(a, npar, ctx->sm_count, ctx->smem_optin, st, &why);
        if (e == cudaErrorInvalidValue && why[0]) {
            all_general = true;                  // the dense kernel does not fit this shape: everything through the general kernel
        } else if (e != cudaSuccess) {
            return fail(DDB_ECUDA, "s2v backward launch: %s", cudaGetErrorString(e));
        } else {
            ctx->launches += 1;
        }
    }
    if (all_general) CUDA_TRY(cudaMemsetAsync(not_dense, 1, 1, st));      // little-endian int32 1: every instance takes the general kernel
    ddb::S2vGGradArgs g;
    g.B = B; g.m = m; g.n = n; g.p = p; g.T = T;
    g.A = A; g.b = b; g.c = c; g.params = params; g.labels = labels; g.w0 = w0; g.w1 = w1;
    g.grad = grad; g.loss = loss;
    g.row_ineq = row_ineq; g.row_bound = row_bound;
    g.inst_flag = all_general ? nullptr : a.inst_flag;
    g.flag_count = all_general ? nullptr : a.inst_flag + B;
    g.scratch = (float*)ctx->s2vgscr.p;
    cudaError_t e = ddb::launch_s2v_bipartite_general_grad(g, ggrid, ctx->smem_optin, st, &why);
    if (e != cudaSuccess) {
        if (why[0]) return fail(DDB_EUNSUPPORTED, "%s (m=%d n=%d p=%d T=%d)", why, m, n, p, T);
        return fail(DDB_ECUDA, "s2v backward launch: %s", cudaGetErrorString(e));
    }
    ctx->launches += 1;
    return DDB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// classifier evaluation metrics (src/ml/train.py:118-150, 174-246; src/ml/test.py:10-54)
// ---------------------------------------------------------------------------------------------------------
extern "C" int ddb_s2v_metrics_dev(ddb_ctx* ctx, int64_t N, const float* logp, const float* probs, const uint8_t* labels,
                                   float thresh, float w0, float w1, double* out, void* stream) {
    if (!ctx || !probs || !labels || !out) return fail(DDB_EINVAL, "ddb_s2v_metrics_dev: NULL argument");
    if (N < 0) return fail(DDB_EINVAL, "ddb_s2v_metrics_dev: N=%lld", (long long)N);
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int slot = ctx->next_counter;
    ctx->next_counter = (ctx->next_counter + 1) % kCounters;
    unsigned int* minbits = reinterpret_cast<unsigned int*>(ctx->counters + 3 * slot);
    CUDA_TRY(ddb::launch_s2v_metrics(N, logp, probs, labels, thresh, w0, w1, out, minbits, ctx->sm_count, (cudaStream_t)stream));
    ctx->launches += 2;
    return DDB_OK;
}
