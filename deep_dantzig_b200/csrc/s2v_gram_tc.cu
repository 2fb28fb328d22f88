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
    size_t hi, lo, stage, inv, invs, diag, rowacc, bar, slot, total;   // hi / lo of stage 0; stage = byte offset between the two stages
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
    L.inv = off;  off += (size_t)288 * 4;      // 1 / |[a_i, b_i]| per row, 1 for the cost node, 0 beyond
    L.invs = off; off += (size_t)288 * 4;      // the same with 0 from the cost node on: weights of the relu row sums
    L.diag = off; off += (size_t)256 * 4;
    L.rowacc = off; off += (size_t)4 * 3 * 256 * 4;      // epilogue: per-row partial sums of the four column groups
    off = (off + 15) / 16 * 16;
    L.bar = off;  off += 64;
    L.slot = off; off += 16;
    L.total = off;
    return L;
}

constexpr int GPW = 16;                    // producer / epilogue warps
constexpr int GPT = GPW * 32;              // ... threads
constexpr int GT = GPT + 32;               // + one MMA-issue warp

__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GPT) : "memory"); }

// two 16-column loads of my TMEM lane in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t t0, uint32_t t1, bool second, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(t0));
    if (second) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(t1));
    } else {
#pragma unroll
        for (int i = 16; i < 32; ++i) r[i] = 0u;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// WARP-SPECIALISED: warps 0 .. 15 stream A (fp64 -> tf32 hi / lo in the canonical UMMA layout, two stages) and run the
// epilogue; warp 16 only issues the MMAs.  full[stage] (16 producer-warp arrivals) / empty[stage] (tcgen05.commit) hand
// the stages back and forth, tfree (16 arrivals) tells the MMA warp that the accumulators of the previous instance
// have been read.  The producers never wait for the issue of an MMA and the issuing thread never converts a number.
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
    float* inv = reinterpret_cast<float*>(smem_raw + L.inv);
    float* invs = reinterpret_cast<float*>(smem_raw + L.invs);
    float* diag = reinterpret_cast<float*>(smem_raw + L.diag);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L.bar);   // [2]
    uint64_t* empty = full + 2;                                       // [2]
    uint64_t* tfree = full + 4;
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + L.slot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows_pl = L.RG * 8;
    const int nkc = (K + KC - 1) / KC;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(full, GPW);
        mbar_init(full + 1, GPW);
        mbar_init(empty, 1);
        mbar_init(empty + 1, 1);
        mbar_init(tfree, GPW);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;

    if (warp == GPW) {
        // ================= MMA warp: lane 0 issues, the warp stays converged on the waits =================
        // instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = NN
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dbase = ((uint64_t)((uint32_t)L.PLB >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint32_t hi_s = smem_u32(hi), lo_s = smem_u32(lo);
        long long g = 0;
        int inst = 0;
        for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x, ++inst) {
            if (inst >= 1) mbar_wait(tfree, (uint32_t)((inst - 1) & 1));     // the previous accumulators have been read
            for (int kc = 0; kc < nkc; ++kc, ++g) {
                const int stg = (int)(g & 1);
                mbar_wait(full + stg, (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t hb = hi_s + (uint32_t)(stg * L.stage), lb = lo_s + (uint32_t)(stg * L.stage);
#pragma unroll 1
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d = tmem_base + (uint32_t)(mt * NN);
#pragma unroll
                        for (int s = 0; s < KC / 8; ++s) {
                            const uint32_t ko = (uint32_t)(2 * s) * L.PLB;
                            const uint64_t a_hi = dbase | (uint64_t)(((hb + ko + mt * 2048) & 0x3FFFF) >> 4);
                            const uint64_t a_lo = dbase | (uint64_t)(((lb + ko + mt * 2048) & 0x3FFFF) >> 4);
                            const uint64_t b_hi = dbase | (uint64_t)(((hb + ko) & 0x3FFFF) >> 4);
                            const uint64_t b_lo = dbase | (uint64_t)(((lb + ko) & 0x3FFFF) >> 4);
                            umma_tf32(d, a_hi, b_hi, idesc, (kc | s) != 0);
                            umma_tf32(d, a_hi, b_lo, idesc, 1);
                            umma_tf32(d, a_lo, b_hi, idesc, 1);
                        }
                    }
                    umma_commit(empty + stg);      // arrives when every MMA issued so far has finished reading / writing
                }
                __syncwarp();
            }
        }
    } else {
        // ================= producer / epilogue warps =================
        // software-pipelined producer: the global loads of chunk k + 1 are issued before chunk k is converted and stored
        // (RPW rows per warp and chunk, one column per lane)
        constexpr int RPW = 256 / GPW;
        double vnext[RPW];
        // my rows are r = warp + q GPW: the first qlim of them are rows of [A | b], slot qc (if any) is the cost row [c, 0]
        const int qlim = (m > warp) ? (m - warp + GPW - 1) / GPW : 0;
        const int qc = (m >= warp && ((m - warp) % GPW) == 0) ? (m - warp) / GPW : -1;
        auto load_chunk = [&](long long lpq, int kc) {
            // a lane owns ONE column of the chunk: a column of A (stride n), the column b (stride 1) or padding
            const int kabs = kc * KC + lane;
            const bool colA = kabs < n, colB = kabs == n;
            const size_t stride = colA ? (size_t)n : 1;
            const double* p = (colA ? a.A + (size_t)lpq * m * n + kabs : a.b + (size_t)lpq * m) + (size_t)warp * stride;
            const double* pc = a.c + (size_t)lpq * n + (colA ? kabs : 0);
            const bool any = colA || colB;
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
                double v = 0.0;
                if (q < qlim) {
                    if (any) v = __ldg(p + (size_t)q * GPW * stride);
                } else if (q == qc) {
                    if (colA) v = __ldg(pc);
                }
                vnext[q] = v;
            }
        };
        if ((long long)blockIdx.x < a.B) load_chunk(blockIdx.x, 0);
        long long g = 0;

        for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {
            // The product runs on the UNNORMALISED rows G' = [[A | b] ; [c, 0]]: W'_ii is the squared row norm, so the
            // normalisation of the reference (s2v.py:145) becomes a row / column scaling of the accumulators in the
            // epilogue and A is read exactly once.
            for (int kc = 0; kc < nkc; ++kc, ++g) {
                const int stg = (int)(g & 1);
                double vcur[RPW];
#pragma unroll
                for (int q = 0; q < RPW; ++q) vcur[q] = vnext[q];
                {
                    const bool wrap = (kc + 1 == nkc);
                    const long long lpn = wrap ? lp + gridDim.x : lp;
                    if (lpn < a.B) load_chunk(lpn, wrap ? 0 : kc + 1);
                }
                // the MMAs that read this stage two chunks ago must have finished before it is overwritten
                if (g >= 2) mbar_wait(empty + stg, (uint32_t)(((g >> 1) - 1) & 1));
                // element (row r, column lane) of the chunk: plane lane / 4, 8-row group r / 8, row r % 8, word lane % 4; with
                // r = warp + q GPW the address is a per-thread base + q * (GPW / 8) * 128 (an immediate).  hi = tf32(g) rounded,
                // lo = g - hi as it is: kind::tf32 reads the upper 19 bits of an fp32 container, which drops 2^-22 |g| at most --
                // the size of the Glo Glo^T term the 3xTF32 product leaves out anyway
                const uint32_t coff = (uint32_t)(lane >> 2) * L.PLB + (lane & 3) * 4 + (uint32_t)(warp >> 3) * 128 + (warp & 7) * 16;
                const uint32_t hsa = smem_u32(hi) + (uint32_t)(stg * L.stage) + coff;
                const uint32_t lsa = smem_u32(lo) + (uint32_t)(stg * L.stage) + coff;
                const int qrows = (rows_pl - warp + GPW - 1) / GPW;          // slots whose row lies inside the planes
#pragma unroll
                for (int q = 0; q < RPW; ++q) {
                    if (q < qrows) {
                        const float gq = (float)vcur[q];
                        const uint32_t h = to_tf32(gq);
                        const float l = gq - __uint_as_float(h);
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(hsa + (uint32_t)q * (GPW / 8) * 128), "r"(h) : "memory");
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(lsa + (uint32_t)q * (GPW / 8) * 128), "f"(l) : "memory");
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full + stg);
            }
            // all MMAs of this instance are complete when the last chunk's commit has arrived
            {
                const long long gl = g - 1;
                mbar_wait(empty + (int)(gl & 1), (uint32_t)((gl >> 1) & 1));
            }
            tc_fence_after();

            // ---- epilogue ------------------------------------------------------------------------------------------------
            // pass 1 (warps 0-3: thread t owns accumulator row t of each tile): the diagonal gives the row norms
            if (warp < 4) {
                for (int mt = 0; mt < MT; ++mt) {
                    const int i = mt * 128 + tid;
                    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * NN);
                    const int cb = mt * 128 + warp * 32;       // my warp's 32 diagonal entries sit in two 16-column chunks
                    float dg = 0.f;
                    if (cb < NN) {
                        float v[32];
                        tmem_ld16x2(taddr + cb, taddr + cb + 16, cb + 16 < NN, v);
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (lane == q) dg = v[q];
                    }
                    if (i < 256) {
                        const float iv = (i < m) ? 1.0f / fmaxf(sqrtf(dg), 1e-12f) : (i == m ? 1.0f : 0.f);
                        inv[i] = iv;
                        invs[i] = (i < m) ? iv : 0.f;
                        diag[i] = dg;
                    }
                }
            }
            for (int i = 256 + tid; i < 288; i += GPT) { inv[i] = 0.f; invs[i] = 0.f; }
            producers_sync();
            // pass 2: W_ij = W'_ij inv_i inv_j; relu row sums  Wp_i = inv_i sum_j inv_j relu(W'_ij),  Wn_i = Wp_i - inv_i sum_j inv_j W'_ij
            // (inv > 0).  Warp w reads the TMEM lanes of its quadrant (w % 4, the hardware's lane restriction) and every fourth
            // 16-column chunk (w / 4), two chunks in flight; chunks that hold my rows' diagonal or the cost column take the
            // guarded path; the four partial sums of a row meet in shared memory.
            {
                const int wq = warp & 3, wg = warp >> 2;
                for (int mt = 0; mt < MT; ++mt) {
                    const int i0 = mt * 128 + wq * 32, i = i0 + lane;
                    const float ii = inv[i];
                    float sp = 0.f, sa = 0.f, wl = 0.f;
                    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * NN);
                    for (int c0 = 16 * wg; c0 < NN; c0 += 128) {
                        float v[32];
                        const bool second = c0 + 64 < NN;
                        tmem_ld16x2(taddr + c0, taddr + c0 + 64, second, v);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int cc = c0 + 64 * h;
                            if (h == 1 && !second) break;
                            const bool special = (cc < i0 + 32 && cc + 16 > i0) || (cc <= m && cc + 16 > m);   // warp-uniform
                            if (!special) {
#pragma unroll
                                for (int q4 = 0; q4 < 4; ++q4) {
                                    const float4 w4 = *reinterpret_cast<const float4*>(invs + cc + 4 * q4);
                                    const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const float x = v[16 * h + 4 * q4 + u];
                                        sp = fmaf(ws[u], fmaxf(x, 0.f), sp);
                                        sa = fmaf(ws[u], x, sa);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    const int j = cc + q;
                                    const float x = v[16 * h + q];
                                    if (j != i && j < m) {
                                        const float wj = invs[j];
                                        sp = fmaf(wj, fmaxf(x, 0.f), sp);
                                        sa = fmaf(wj, x, sa);
                                    }
                                    if (j == m && j != i) wl = x * ii;
                                }
                            }
                        }
                    }
                    float* ra = rowacc + (size_t)wg * 3 * 256 + i;
                    ra[0] = sp * ii;
                    ra[256] = (sp - sa) * ii;
                    ra[512] = wl;
                }
            }
            // the accumulators have been read: the MMA warp may start the next instance
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tfree);
            producers_sync();
            for (int i = tid; i < M1; i += GPT) {
                float* o = a.out + (size_t)lp * 3 * a.MP;
                // fixed summation order over the four column groups: the result does not depend on warp timing
                o[i] = (rowacc[i] + rowacc[768 + i]) + (rowacc[1536 + i] + rowacc[2304 + i]);
                o[a.MP + i] = (rowacc[256 + i] + rowacc[1024 + i]) + (rowacc[1792 + i] + rowacc[2560 + i]);
                o[2 * a.MP + i] = (rowacc[512 + i] + rowacc[1280 + i]) + (rowacc[2048 + i] + rowacc[2816 + i]);
            }
            producers_sync();
        }
    }

    tc_fence_before();
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
