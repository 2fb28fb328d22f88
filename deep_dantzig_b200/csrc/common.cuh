// Shared declarations for the ddb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddb {

// Tolerances of the device simplex (absolute; instance data is O(1), see DESIGN.md "Numerics").
constexpr double kTolPivot = 1e-9;   // smallest |entry| accepted as a ratio-test pivot
constexpr double kTolFeas  = 1e-9;   // primal / dual feasibility tolerance
constexpr double kTolCrash = 1e-7;   // smallest |entry| accepted as a crash pivot

// Row states of the condensed tableau.
enum : uint8_t { ROW_LIVE = 0, ROW_CRASHED = 1, ROW_EXCLUDED = 2 };

// Status codes == Gurobi's (reference src/data/gurobi_lp.py:447-461).
enum : int { ST_OPTIMAL = 2, ST_INFEASIBLE = 3, ST_UNBOUNDED = 5, ST_ITERATION_LIMIT = 7, ST_NUMERIC = 12 };

struct SolveArgs {
    int m, n;
    long long B;
    const double* A;
    const double* b;
    const double* c;
    const uint8_t* row_mask;   // nullable
    double thr;
    int* status;
    double* x;                 // nullable
    double* obj;               // nullable
    uint8_t* labels;
    int* n_active;             // nullable
    int* pivots;               // nullable, [B,4]
    int* ties;                 // nullable
    int* violations;           // nullable
    unsigned long long* counter;   // work queue
    double* gtab;              // per-CTA global tableau slabs (plan 2) or per-CTA scratch (plan 0)
    int max_iter;
    int only_flagged;          // generic kernel: solve only instances whose status is -1 (flagged by plan 0)
    int* flag_count;           // number of instances plan 0 flagged for the generic kernel
    double* dscr;              // plan 0, hybrid rows: per-CTA global home of the crash inverse D (L2-resident, read back by bulk TMA)
    // fused generate -> solve -> label (gen != 0): instance lp is drawn inside the solver CTA from (gen_key, gen_first + lp);
    // A / b / c then point at the caller's output arrays (nullable: the instance lives only in a per-CTA slab in `slab`)
    int gen;
    unsigned long long gen_key;
    long long gen_first;
    double gen_density;
    double* slab;              // [grid][m n + m + n] per-CTA instance slabs (used when A is NULL)
};

// Arguments of the classifier forward kernels (s2v_forward.cu, s2v_bipartite_dense.cu).
struct S2vArgs {
    int graph;                 // 0 complete, 1 bipartite
    long long B;
    int m, n, p, T;
    const double* A;
    const double* b;
    const double* c;
    const float* params;       // flat, reference state_dict order (oracle/classifier.py: *_PARAMS)
    float* logp;               // [B, m, 2]
    float* probs;              // [B, m, 2] (nullable)
    int* error_flag;           // set to 1 if a sparse instance did not fit the shared-memory plan
    int store_A;               // bipartite general kernel: 2 = normalised A + aggregation buffer in shared memory,
                               // 1 = aggregation buffer only (adjacency from global A), 0 = dense instances only
    const float* gram;         // complete: [B][3][gram_pitch] Wp, Wn, wc from the tensor-core Gram kernel (nullable)
    int gram_pitch;
    int* inst_flag;            // bipartite: [B] set to 1 by the dense kernel for instances with a zero coefficient
    int* flag_count;           // number of flagged instances
    int only_flagged;          // general bipartite kernel: process only instances whose inst_flag is 1
    // node flags of items that are not plain random LPs (MPS / PLNN items: equality rows, bound rows), [B, m] 0 / 1, nullable:
    //   bipartite: c_feats[:, 0] = is_inequality (default 1), c_feats[:, 2] = is_bound (default 0)   (gurobi_lp.py:157-158)
    //   complete : node_features of the row nodes = is_inequality (default 1)                        (gurobi_lp.py:329, 360)
    const uint8_t* row_ineq;
    const uint8_t* row_bound;
};

// ---------------------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// all state spaces: orders this thread's generic-proxy accesses (global and shared) before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async_all() {
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const long long t0 = clock64();
    do {
        if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a lost TMA transaction becomes an error, not a hang
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// same with an L2 cache policy (createpolicy): per-CTA scratch that is rewritten by every LP should stay in L2
__device__ __forceinline__ void tma_load_1d_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_global_v2_hint(double* p, double x, double y, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(x), "d"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_global_hint(double* p, double x, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(x), "l"(pol) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// warp argmin helpers with Bland (lowest variable index) tie-breaking
// ---------------------------------------------------------------------------------------------------------
struct Cand {
    double val;   // key to minimise
    int var;      // tie-break: lowest variable index wins
    int idx;      // payload (row or column position)
};
__device__ __forceinline__ bool cand_better(double v1, int var1, double v2, int var2) {
    return (v1 < v2) || (v1 == v2 && var1 < var2);
}
__device__ __forceinline__ Cand warp_argmin(Cand c) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, c.val, off);
        int ovar = __shfl_xor_sync(0xffffffffu, c.var, off);
        int oidx = __shfl_xor_sync(0xffffffffu, c.idx, off);
        if (cand_better(ov, ovar, c.val, c.var)) {
            c.val = ov;
            c.var = ovar;
            c.idx = oidx;
        }
    }
    return c;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}


// ---------------------------------------------------------------------------------------------------------
// order-preserving 64-bit keys + redux-based warp argmin, fast reciprocal (register-resident kernels)
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned long long dkey(double v) {   // order-preserving map double -> uint64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// lane holding the minimum key (lowest lane on ties); kmin = that key.  Two redux + one ballot.
__device__ __forceinline__ int warp_argmin_key(unsigned long long key, unsigned long long& kmin) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(FULL, hi);
    const unsigned lo2 = (hi == mhi) ? lo : 0xffffffffu;
    const unsigned mlo = __reduce_min_sync(FULL, lo2);
    const unsigned ball = __ballot_sync(FULL, hi == mhi && lo2 == mlo);
    kmin = ((unsigned long long)mhi << 32) | mlo;
    return __ffs(ball) - 1;
}
constexpr unsigned long long KEY_INF = 0xfff0000000000000ull;   // dkey(+inf)

// 1/p to ~1 ulp: 20-bit hardware seed + two Newton steps (5 instructions instead of the ~35 of an IEEE division).
__device__ __forceinline__ double fast_rcp(double p) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(p));
    double e = fma(-p, x, 1.0);
    x = fma(x, e, x);
    e = fma(-p, x, 1.0);
    x = fma(x, e, x);
    return x;
}

constexpr double kInf = 1e300;
constexpr int kBigVar = 0x7fffffff;

}  // namespace ddb
