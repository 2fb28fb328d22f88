// Tensor-core Gram stage of the 'complete' classifier variant: W = G G^T fused with its relu row sums, on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), one LP instance per CTA pass.
//
// Replaces the only dense contraction of the reference forward (src/ml/models/s2v.py:155-162: W = G.mm(G.t()),
// diagonal zeroed) together with the bmm -> relu -> sum reductions that consume W (s2v.py:112-116): the kernel never
// stores W, it emits per node i
//     Wp[i] = sum_{j<m, j!=i} relu(W_ij),   Wn[i] = sum_{j<m, j!=i} relu(-W_ij),   wc[i] = W_im (cost-node column)
// which is all the s2v rounds need (see s2v_forward.cu).
//
// G = [normalize([A | b]) ; [c, 0]] is (m+1) x (n+1), fp32 in the reference.  tcgen05 has no fp32 kind, so the
// product is computed as 3xTF32: G = Ghi + Glo (both tf32), W ~= Ghi Ghi^T + Ghi Glo^T + Glo Ghi^T, accumulated in
// fp32 in TMEM; the dropped Glo Glo^T term is 2^-22 relative, i.e. fp32-level accuracy (the parity tests keep their
// 5e-5 tolerance on log-probabilities).  The product runs on the unnormalised rows: the diagonal of the accumulator
// is the squared row norm, so the normalisation is an epilogue scaling and A is streamed from HBM exactly once.
//
// Shapes: M = m+1 <= 256 rows (one or two 128-row accumulator tiles), N = M rounded up to 16, K = n+1 streamed in
// chunks of 32 columns.  Operands are written by the CTA's threads straight into the canonical K-major no-swizzle
// UMMA layout (8-row x 16-byte core matrices; element (row, k) of a chunk at
//     (k / 4) * PLANE + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4     bytes)
// so the shared-memory descriptors are  start | LBO = PLANE | SBO = 128.  PLANE is padded by 16 bytes so that the 8
// planes a warp writes at once fall into different banks.
#include "common.cuh"

namespace ddb {

struct S2vGramArgs {
    long long B;
    int m, n;
    const double* A;
    const double* b;
    const double* c;
    float* out;        // [B][3][MP]: Wp, Wn, wc
    int MP;            // pitch of one output vector (floats)
};

namespace {

constexpr int KC = 32;                     // K columns per chunk
constexpr int NPL = KC / 4;                // 16-byte planes per chunk

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 accumulator columns of my TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no swizzle: start address, leading (K) byte offset, stride (8-row group) byte offset, sm_100 version bit
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

struct GramLayout {
    int RG;            // 8-row groups per plane
    int PLB;           // plane pitch in bytes
    size_t hi, lo, stage, inv, rowacc, bar, slot, total;   // hi / lo of stage 0; stage = byte offset between the two stages
};
__host__ __device__ inline GramLayout gram_layout(int m) {
    GramLayout L;
    const int M1 = m + 1;
    const int MT = (M1 + 127) / 128;
    L.RG = 16 * MT;
    L.PLB = L.RG * 128 + 16;
    size_t off = 0;
    L.hi = off;   off += (size_t)NPL * L.PLB;
    L.lo = off;   off += (size_t)NPL * L.PLB;
    off = (off + 127) / 128 * 128;
    L.stage = off;
    off *= 2;                                  // second stage: produce chunk k+1 while the tensor core reads chunk k
    off = (off + 15) / 16 * 16;
    L.inv = off;  off += (size_t)(M1 + 16) * 4;
    off = (off + 15) / 16 * 16;
    L.rowacc = off; off += (size_t)4 * 3 * 256 * 4;      // epilogue: per-row partial sums of the four column groups
    off = (off + 15) / 16 * 16;
    L.bar = off;  off += 32;
    L.slot = off; off += 16;
    L.total = off;
    return L;
}

constexpr int GT = 512;                    // threads per CTA: 16 warps stream A, warps 0-3 own the accumulator rows

template <int TMEM_COLS>
__global__ void __launch_bounds__(GT, 1) s2v_gram_tc_kernel(S2vGramArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n, M1 = m + 1, K = n + 1;
    const int MT = (M1 + 127) / 128;
    const int NN = (M1 + 15) / 16 * 16;            // MMA N
    const GramLayout L = gram_layout(m);
    unsigned char* hi = smem_raw + L.hi;
    unsigned char* lo = smem_raw + L.lo;
    float* rowacc = reinterpret_cast<float*>(smem_raw + L.rowacc);   // [4 column groups][3: Wp, Wn, wc][256 rows]
    float* inv = reinterpret_cast<float*>(smem_raw + L.inv);       // 1 / |[a_i, b_i]| per row (1 for the cost node)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + L.slot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows_pl = L.RG * 8;
    const int nkc = (K + KC - 1) / KC;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    uint32_t par0 = 0, par1 = 0;          // phase parity of the two stage barriers
    bool pend0 = false, pend1 = false;    // stage has MMAs in flight that have not been waited for

    // instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = NN
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi_s = smem_u32(hi), lo_s = smem_u32(lo);

    // software-pipelined producer: the global loads of chunk k + 1 are issued before chunk k is converted, stored and
    // multiplied (RPW rows per warp and chunk, one column per lane)
    constexpr int RPW = 256 / (GT / 32);
    double vnext[RPW];
    auto load_chunk = [&](long long lpq, int kc) {
        const double* Ag = a.A + (size_t)lpq * m * n;
        const double* bg = a.b + (size_t)lpq * m;
        const double* cg = a.c + (size_t)lpq * n;
        const int kabs = kc * KC + lane;
#pragma unroll
        for (int q = 0; q < RPW; ++q) {
            const int r = warp + q * (GT / 32);
            double v = 0.0;
            if (r < m) {
                if (kabs < n) v = __ldg(Ag + (size_t)r * n + kabs);
                else if (kabs == n) v = __ldg(bg + r);
            } else if (r == m) {
                if (kabs < n) v = __ldg(cg + kabs);
            }
            vnext[q] = v;
        }
    };
    if ((long long)blockIdx.x < a.B) load_chunk(blockIdx.x, 0);

    for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {

        // The product runs on the UNNORMALISED rows G' = [[A | b] ; [c, 0]]: W'_ii is the squared row norm, so the
        // normalisation of the reference (s2v.py:145) becomes a row / column scaling of the accumulators in the
        // epilogue and A is read exactly once.
        for (int kc = 0; kc < nkc; ++kc) {
            const int stg = kc & 1;
            // this chunk's values were loaded one chunk ago (vnext); start the loads of the NEXT chunk (the next
            // instance's first one after the last) now, so that they fly during the convert / store / MMA issue below
            double vcur[RPW];
#pragma unroll
            for (int q = 0; q < RPW; ++q) vcur[q] = vnext[q];
            {
                const bool wrap = (kc + 1 == nkc);
                const long long lpn = wrap ? lp + gridDim.x : lp;
                if (lpn < a.B) load_chunk(lpn, wrap ? 0 : kc + 1);
            }
            // the MMAs that read this stage two chunks ago must have finished before it is overwritten
            if (stg == 0 && pend0) { mbar_wait(bar, par0); par0 ^= 1; pend0 = false; }
            if (stg == 1 && pend1) { mbar_wait(bar + 1, par1); par1 ^= 1; pend1 = false; }
            unsigned char* hs = hi + (size_t)stg * L.stage;
            unsigned char* ls = lo + (size_t)stg * L.stage;
            // ---- produce the chunk: rows of G' -> (hi, lo) tf32 in the canonical UMMA layout -------------------------------
            const uint32_t coff = (uint32_t)(lane >> 2) * L.PLB + (lane & 3) * 4;
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
                const int r = warp + q * (GT / 32);
                if (r < rows_pl) {
                    const float g = (float)vcur[q];
                    const uint32_t h = to_tf32(g);
                    const uint32_t l = to_tf32(g - __uint_as_float(h));
                    const uint32_t off = coff + (uint32_t)(r >> 3) * 128 + (r & 7) * 16;
                    *reinterpret_cast<uint32_t*>(hs + off) = h;
                    *reinterpret_cast<uint32_t*>(ls + off) = l;
                }
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t hb = hi_s + (uint32_t)(stg * L.stage), lb = lo_s + (uint32_t)(stg * L.stage);
#pragma unroll 1
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t d = tmem_base + (uint32_t)(mt * NN);
#pragma unroll 1
                    for (int s = 0; s < KC / 8; ++s) {
                        const uint32_t ko = (uint32_t)(2 * s) * L.PLB;
                        const uint64_t a_hi = smem_desc(hb + ko + mt * 2048, L.PLB, 128);
                        const uint64_t a_lo = smem_desc(lb + ko + mt * 2048, L.PLB, 128);
                        const uint64_t b_hi = smem_desc(hb + ko, L.PLB, 128);
                        const uint64_t b_lo = smem_desc(lb + ko, L.PLB, 128);
                        umma_tf32(d, a_hi, b_hi, idesc, (kc | s) != 0);
                        umma_tf32(d, a_hi, b_lo, idesc, 1);
                        umma_tf32(d, a_lo, b_hi, idesc, 1);
                    }
                }
                umma_commit(bar + stg);
            }
            if (stg == 0) pend0 = true; else pend1 = true;
        }
        // all MMAs of this instance have to be complete before the accumulators are read
        if (pend0) { mbar_wait(bar, par0); par0 ^= 1; pend0 = false; }
        if (pend1) { mbar_wait(bar + 1, par1); par1 ^= 1; pend1 = false; }
        tc_fence_after();

        // ---- epilogue (warps 0-3: thread t owns accumulator row t of each tile) ------------------------------------------
        // pass 1: the diagonal gives the row norms
        if (warp < 4) {
            for (int mt = 0; mt < MT; ++mt) {
                const int i = mt * 128 + tid;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * NN);
                // the diagonal entries of my warp's 32 rows sit in two 16-column chunks (warp-uniform addresses:
                // tcgen05.ld is a warp-collective)
                const int cb = mt * 128 + warp * 32;
                float dg = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (cb + 16 * h < NN) {
                        float v[16];
                        tmem_ld16(taddr + cb + 16 * h, v);
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (lane == 16 * h + q) dg = v[q];
                    }
                }
                if (i < M1) inv[i] = (i < m) ? 1.0f / fmaxf(sqrtf(dg), 1e-12f) : 1.0f;
            }
        }
        __syncthreads();
        // pass 2: W_ij = W'_ij inv_i inv_j, relu row sums.  All 16 warps: warp w reads the TMEM lanes of its quadrant
        // (w % 4, the hardware's lane restriction) and every fourth 16-column chunk (w / 4); the four partial sums of a
        // row meet in shared memory.
        {
            const int wq = warp & 3, wg = warp >> 2;
            for (int mt = 0; mt < MT; ++mt) {
                const int i = mt * 128 + wq * 32 + lane;
                const float ii = (i < M1) ? inv[i] : 0.f;
                float sp = 0.f, sn = 0.f, wl = 0.f;
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * NN);
                for (int c0 = 16 * wg; c0 < NN; c0 += 64) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int j = c0 + q;
                        if (j != i && j <= m) {
                            const float w = v[q] * ii * inv[j];
                            if (j < m) { sp += fmaxf(w, 0.f); sn += fmaxf(-w, 0.f); }
                            else wl = w;
                        }
                    }
                }
                float* ra = rowacc + (size_t)wg * 3 * 256 + (mt * 128 + wq * 32 + lane);
                ra[0] = sp;
                ra[256] = sn;
                ra[512] = wl;
            }
        }
        tc_fence_before();
        __syncthreads();
        for (int i = tid; i < M1; i += GT) {
            float* o = a.out + (size_t)lp * 3 * a.MP;
            // fixed summation order over the four column groups: the result does not depend on warp timing
            o[i] = (rowacc[i] + rowacc[768 + i]) + (rowacc[1536 + i] + rowacc[2304 + i]);
            o[a.MP + i] = (rowacc[256 + i] + rowacc[1024 + i]) + (rowacc[1792 + i] + rowacc[2560 + i]);
            o[2 * a.MP + i] = (rowacc[512 + i] + rowacc[1280 + i]) + (rowacc[2048 + i] + rowacc[2816 + i]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// shapes the tensor-core Gram stage covers: m + 1 <= 256 accumulator rows
bool s2v_gram_tc_supported(int m, int n) { return m >= 1 && n >= 1 && m + 1 <= 256; }
size_t s2v_gram_out_floats(int m) { return (size_t)3 * ((m + 1 + 3) & ~3); }

cudaError_t launch_s2v_gram_tc(long long B, int m, int n, const double* A, const double* b, const double* c, float* out,
                               int sm_count, cudaStream_t st) {
    S2vGramArgs a;
    a.B = B; a.m = m; a.n = n; a.A = A; a.b = b; a.c = c; a.out = out;
    a.MP = (m + 1 + 3) & ~3;
    const size_t smem = gram_layout(m).total;
    const int M1 = m + 1, MT = (M1 + 127) / 128, NN = (M1 + 15) / 16 * 16;
    const int cols = MT * NN;
    long long grid = sm_count;
    if (grid > B) grid = B;
    cudaError_t e;
#define DDB_GRAM_LAUNCH(TC)                                                                                         \
    do {                                                                                                            \
        e = cudaFuncSetAttribute(s2v_gram_tc_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e != cudaSuccess) return e;                                                                             \
        s2v_gram_tc_kernel<TC><<<(int)grid, GT, smem, st>>>(a);                                                    \
    } while (0)
    if (cols <= 32) DDB_GRAM_LAUNCH(32);
    else if (cols <= 64) DDB_GRAM_LAUNCH(64);
    else if (cols <= 128) DDB_GRAM_LAUNCH(128);
    else if (cols <= 256) DDB_GRAM_LAUNCH(256);
    else DDB_GRAM_LAUNCH(512);
#undef DDB_GRAM_LAUNCH
    return cudaGetLastError();
}

}  // namespace ddb
