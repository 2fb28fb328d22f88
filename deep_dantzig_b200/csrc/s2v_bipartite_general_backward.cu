// Batched loss + gradient of the reference classifier, bipartite variant, GENERAL ADJACENCY (instances that contain zero
// coefficients): one CTA per instance, fp32, backward by hand, gradients summed over the batch.
//
// Second half of ddb_s2v_loss_grad_dev for graph 1: the dense kernel (s2v_backward.cu) handles the reference's dense random
// LPs and FLAGS every instance with a zero in A; this kernel processes exactly the flagged ones, so that nothing on the
// training path (reference src/ml/train.py:59-66, criterion src/benchmark.py:70-75) goes through autograd.
//
// Forward = s2v_bipartite_kernel's general path (s2v_forward.cu; s2v.py:253-323, 218-251, quirk B9 kept):
//   agg_v[:, j] = mean of mu_c over the rows adjacent to column j,  agg_c[:, i] = mean of mu_v over the columns adjacent to i
//   term2 position q:  q < n -> t2c agg_v[:, q],  q >= n -> t2v agg_c[:, q - n]      (variables first, B9)
//   mu position q   :  q < m constraint q, q >= m variable q - m;  mu' = relu(base + term2)  position by position
// With a general adjacency the gradient of a round is per node, so the kernel keeps the embeddings and aggregates of every
// round.  They live in a per-CTA global scratch (L2-resident, rewritten by every instance) in NODE-MAJOR layout [node][PP]:
// an embedding vector is one coalesced 4 PP-byte read, an adjacency product is "for every set bit: add one vector"
// (cost proportional to the non-zeros), and a p x p product per node reads its input vector once.
// The adjacency is a bit mask in shared memory in both orientations (rows of A and columns of A).
#include "common.cuh"

namespace ddb {

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
    const uint8_t* row_ineq;   // [B, m] node flags of MPS / PLNN items (c_feats[:, 0] / c_feats[:, 2]), nullable = 1 / 0; a label
    const uint8_t* row_bound;  // value of 2 marks a row outside the item's in_loss set (no loss, no gradient from it)
    const int* inst_flag;      // [B]: process only instances whose flag is set (nullable: all)
    const int* flag_count;     // number of flagged instances (nullable)
    float* scratch;            // [grid][s2v_general_grad_scratch_floats]
};

namespace {

constexpr int kGT = 256;
constexpr int kPartSlots = 12;       // per-warp partial-sum slots of the parameter reductions

__host__ __device__ inline int gpad4(int v) { return (v + 3) & ~3; }

__device__ __forceinline__ float gwsum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ void gmatvec(const float* __restrict__ W, int p, const float* x, float* y, int warp, int lane, int nw) {
    for (int k = warp; k < p; k += nw) {
        float acc = 0.f;
        for (int l = lane; l < p; l += 32) acc = fmaf(__ldg(W + k * p + l), x[l], acc);
        acc = gwsum(acc);
        if (lane == 0) y[k] = acc;
    }
}
__device__ __forceinline__ void gmatvecT(const float* __restrict__ W, int p, const float* x, float* y, int warp, int lane, int nw) {
    for (int l = warp; l < p; l += nw) {
        float acc = 0.f;
        for (int k = lane; k < p; k += 32) acc = fmaf(__ldg(W + k * p + l), x[k], acc);
        acc = gwsum(acc);
        if (lane == 0) y[l] = acc;
    }
}

struct GGradLayout {           // shared memory, offsets in floats
    size_t w2cT, w2vT, w2cN, w2vN, t7T, t7N, adj, adjT, vecs, total;
};
__host__ __device__ inline GGradLayout ggrad_layout(int m, int n, int p) {
    const int PP = gpad4(p);
    GGradLayout L;
    size_t off = 0;
    L.w2cT = off; off += (size_t)p * PP;
    L.w2vT = off; off += (size_t)p * PP;
    L.w2cN = off; off += (size_t)p * PP;
    L.w2vN = off; off += (size_t)p * PP;
    L.t7T = off;  off += (size_t)p * PP;
    L.t7N = off;  off += (size_t)p * PP;
    L.adj = off;  off += (size_t)m * ((n + 31) / 32);
    L.adjT = off; off += (size_t)n * ((m + 31) / 32);
    L.vecs = off;
    off += (size_t)8 * m + 5 * n + (size_t)40 * PP + (size_t)(kGT / 32) * kPartSlots * PP + 64;
    L.total = off;
    return L;
}
// per-CTA global scratch (floats): base, T embeddings, T aggregates, D1, D2, DP, Z
__host__ __device__ inline size_t ggrad_scratch_floats(int m, int n, int p, int T) {
    const size_t PP = gpad4(p), NP = (size_t)m + n;
    return (size_t)(2 * (T > 0 ? T : 1) + 4) * NP * PP + (size_t)m * PP;
}

// S = slots of the element-wise p x p register accumulators: ceil(p * p / kGT)
template <int S>
__global__ void __launch_bounds__(kGT, 2) s2v_bipartite_general_grad_kernel(S2vGGradArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int m = a.m, n = a.n, p = a.p, T = a.T, PP = gpad4(p), NP = m + n;
    const int NW32 = (n + 31) / 32, MW32 = (m + 31) / 32;
    const GGradLayout L = ggrad_layout(m, n, p);
    float* w2cT = sm + L.w2cT;   // [l][k] = t2c[k][l]
    float* w2vT = sm + L.w2vT;
    float* w2cN = sm + L.w2cN;   // [k][l] = t2c[k][l]
    float* w2vN = sm + L.w2vN;
    float* t7T = sm + L.t7T;
    float* t7N = sm + L.t7N;
    unsigned* adj = reinterpret_cast<unsigned*>(sm + L.adj);     // [m][NW32]: bit j of row i
    unsigned* adjT = reinterpret_cast<unsigned*>(sm + L.adjT);   // [n][MW32]: bit i of column j
    float* v = sm + L.vecs;
    float* rb = v;     v += m;   float* cosv = v;  v += m;   float* Sp = v;    v += m;   float* Sn = v;    v += m;
    float* rinv = v;   v += m;   float* ds0 = v;   v += m;   float* ds1 = v;   v += m;   float* lossn = v; v += m;
    float* cj = v;     v += n;   float* Cp = v;    v += n;   float* Cn = v;    v += n;   float* ccnt = v;  v += n;
    float* cinv = v;   v += n;
    float* w3cp = v;   v += PP;  float* w3cn = v;  v += PP;  float* w3vp = v;  v += PP;  float* w3vn = v;  v += PP;
    float* r4 = v;     v += 4 * PP;
    float* meanc = v;  v += PP;  float* meanv = v; v += PP;  float* u6pre = v; v += PP;  float* u6r = v;   v += PP;
    float* du6 = v;    v += PP;  float* tmp1 = v;  v += PP;  float* tmp2 = v;  v += PP;
    float* gw3 = v;    v += 4 * PP;             // d w3cp, d w3cn, d w3vp, d w3vn
    float* g_t0 = v;   v += PP;  float* g_t1c = v; v += 4 * PP;   float* g_t1v = v; v += PP;
    float* g_t8 = v;   v += 8 * PP;             // [c][u6 block | z block | 4 features (at 2 PP)] , c = 0 at 0, c = 1 at 4 PP
    float* part = v;   v += (size_t)(kGT / 32) * kPartSlots * PP;        // [warp][10][PP] per-warp partial sums
    float* scal = v;

    const int tid = threadIdx.x, nt = kGT, lane = tid & 31, warp = tid >> 5, nw = kGT / 32;
    const float* P = a.params;
    const float* t0 = P;   P += p;        const int o_t0 = 0;
    const float* t1c = P;  P += 4 * p;    const int o_t1c = o_t0 + p;
    const float* t1v = P;  P += p;        const int o_t1v = o_t1c + 4 * p;
    const float* t2c = P;  P += p * p;    const int o_t2c = o_t1v + p;
    const float* t2v = P;  P += p * p;    const int o_t2v = o_t2c + p * p;
    const float* t3c = P;  P += p * p;    const int o_t3c = o_t2v + p * p;
    const float* t3v = P;  P += p * p;    const int o_t3v = o_t3c + p * p;
    const float* t4c = P;  P += p;        const int o_t4c = o_t3v + p * p;
    const float* t4v = P;  P += p;        const int o_t4v = o_t4c + p;
    const float* t6c = P;  P += p * p;    const int o_t6c = o_t4v + p;
    const float* t6v = P;  P += p * p;    const int o_t6v = o_t6c + p * p;
    const float* t7 = P;   P += p * p;    const int o_t7 = o_t6v + p * p;
    const float* t8 = P;                  const int o_t8 = o_t7 + p * p;
    const int W8 = 2 * p + 4;

    if (a.flag_count && *a.flag_count == 0) return;       // the dense kernel handled every instance (uniform)

    for (int e = tid; e < p * PP; e += nt) {
        const int r = e / PP, q = e - r * PP;
        w2cT[e] = (q < p) ? __ldg(t2c + q * p + r) : 0.f;
        w2vT[e] = (q < p) ? __ldg(t2v + q * p + r) : 0.f;
        w2cN[e] = (q < p) ? __ldg(t2c + r * p + q) : 0.f;
        w2vN[e] = (q < p) ? __ldg(t2v + r * p + q) : 0.f;
        t7T[e] = (q < p) ? __ldg(t7 + q * p + r) : 0.f;
        t7N[e] = (q < p) ? __ldg(t7 + r * p + q) : 0.f;
    }
    for (int l = tid; l < PP; l += nt) {
        const float c4 = (l < p) ? __ldg(t4c + l) : 0.f, v4 = (l < p) ? __ldg(t4v + l) : 0.f;
        r4[l] = fmaxf(c4, 0.f); r4[PP + l] = fmaxf(-c4, 0.f); r4[2 * PP + l] = fmaxf(v4, 0.f); r4[3 * PP + l] = fmaxf(-v4, 0.f);
        g_t0[l] = 0.f; g_t1v[l] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { gw3[q * PP + l] = 0.f; g_t1c[q * PP + l] = 0.f; }
#pragma unroll
        for (int q = 0; q < 8; ++q) g_t8[q * PP + l] = 0.f;
    }
    __syncthreads();
    gmatvec(t3c, p, r4, w3cp, warp, lane, nw);
    gmatvec(t3c, p, r4 + PP, w3cn, warp, lane, nw);
    gmatvec(t3v, p, r4 + 2 * PP, w3vp, warp, lane, nw);
    gmatvec(t3v, p, r4 + 3 * PP, w3vn, warp, lane, nw);
    __syncthreads();

    // global scratch of this CTA, node-major [node][PP]
    const size_t NV = (size_t)NP * PP;
    float* base = a.scratch + (size_t)blockIdx.x * ggrad_scratch_floats(m, n, p, T);
    float* MU = base + NV;                         // [T]: embeddings after round 0 .. T-1 (mu positions)
    float* AG = MU + (size_t)(T > 0 ? T : 1) * NV; // [T]: aggregates that entered round t (term2 positions); index 0 unused
    float* D1 = AG + (size_t)(T > 0 ? T : 1) * NV;
    float* D2 = D1 + NV;
    float* DP = D2 + NV;                           // sum over the rounds of d pre (what the base parameters see)
    float* Z = DP + NV;                            // [m][PP]: relu(t7 mu_i), then d z

    // p x p gradient accumulators in registers, element e = tid + s kGT -> (k, l) = (e / p, e % p)
    float a_t2c[S], a_t2v[S], a_t6c[S], a_t6v[S], a_t7[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { a_t2c[s] = 0.f; a_t2v[s] = 0.f; a_t6c[s] = 0.f; a_t6v[s] = 0.f; a_t7[s] = 0.f; }
    double loss_cta = 0.0;

    // out[node q][k] = epi(q, k, sum_l Wq[l][k] * in[q][l]) for q in [q0, q1): one thread per (node, 4 outputs)
    auto node_product = [&](const float* Wq, const float* in, int q0, int q1, auto epi) {
        const int KG = PP / 4;
        for (int w = tid; w < (q1 - q0) * KG; w += nt) {
            const int q = q0 + w / KG, kb = (w % KG) * 4;
            const float* x = in + (size_t)q * PP;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int l = 0; l < p; ++l) {
                const float xv = x[l];
                const float4 wv = *reinterpret_cast<const float4*>(Wq + l * PP + kb);
                a0 = fmaf(wv.x, xv, a0); a1 = fmaf(wv.y, xv, a1); a2 = fmaf(wv.z, xv, a2); a3 = fmaf(wv.w, xv, a3);
            }
            epi(q, kb, a0, a1, a2, a3);
        }
    };
    // out[o][:] = oscale[o] * sum over the set bits r of bits[o][*] of iscale[r] * in[r][:]   (one warp per output node)
    auto bit_aggregate = [&](float* out, const float* in, const unsigned* bits, int words, int nout, const float* iscale,
                             const float* oscale) {
        for (int o = warp; o < nout; o += nw) {
            float acc0 = 0.f, acc1 = 0.f;
            for (int w = 0; w < words; ++w) {
                unsigned bw = bits[o * words + w];
                while (bw) {
                    const int r = w * 32 + __ffs(bw) - 1;
                    bw &= bw - 1;
                    const float sc = iscale ? iscale[r] : 1.f;
                    const float* x = in + (size_t)r * PP;
                    if (lane < PP) acc0 = fmaf(sc, x[lane], acc0);
                    if (lane + 32 < PP) acc1 = fmaf(sc, x[lane + 32], acc1);
                }
            }
            const float os = oscale ? oscale[o] : 1.f;
            if (lane < PP) out[(size_t)o * PP + lane] = acc0 * os;
            if (lane + 32 < PP) out[(size_t)o * PP + lane + 32] = acc1 * os;
        }
    };
    // acc[k][l] += sum_{q in [q0, q1)} Am[q][k] * Bm[q][l]
    auto reduce_nodes = [&](float (&acc)[S], const float* Am, const float* Bm, int q0, int q1) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int e = tid + s * kGT;
            if (e < p * p) {
                const int k = e / p, l = e - k * p;
                float r0 = 0.f, r1 = 0.f;
                int q = q0;
                for (; q + 1 < q1; q += 2) {
                    r0 = fmaf(Am[(size_t)q * PP + k], Bm[(size_t)q * PP + l], r0);
                    r1 = fmaf(Am[(size_t)(q + 1) * PP + k], Bm[(size_t)(q + 1) * PP + l], r1);
                }
                if (q < q1) r0 = fmaf(Am[(size_t)q * PP + k], Bm[(size_t)q * PP + l], r0);
                acc[s] += r0 + r1;
            }
        }
    };
    auto rank1_acc = [&](float (&acc)[S], const float* u, const float* w) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int e = tid + s * kGT;
            if (e < p * p) acc[s] = fmaf(u[e / p], w[e % p], acc[s]);
        }
    };
    // part[warp][slot][l] partial sums -> dst[l] += sum over the warps
    auto fold_part = [&](int slot, float* dst) {
        for (int l = tid; l < p; l += nt) {
            float s = 0.f;
            for (int w = 0; w < nw; ++w) s += part[((size_t)w * kPartSlots + slot) * PP + l];
            dst[l] += s;
        }
    };

    for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {
        if (a.inst_flag && a.inst_flag[lp] == 0) continue;
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg = a.c + (size_t)lp * n;
        const uint8_t* yl = a.labels + (size_t)lp * m;
        // row flags of MPS / PLNN items (random LPs: every row a non-bound inequality)
        const uint8_t* fil = a.row_ineq ? a.row_ineq + (size_t)lp * m : nullptr;
        const uint8_t* fbl = a.row_bound ? a.row_bound + (size_t)lp * m : nullptr;
        auto fiq = [&](int i) -> float { return fil ? (float)fil[i] : 1.f; };
        auto fbq = [&](int i) -> float { return fbl ? (float)fbl[i] : 0.f; };
        for (int j = tid; j < n; j += nt) { cj[j] = (float)cg[j]; Cp[j] = 0.f; Cn[j] = 0.f; ccnt[j] = 0.f; }
        for (int e = tid; e < n * MW32; e += nt) adjT[e] = 0u;
        __syncthreads();

        // ---- pass over A: row normalisation, cosines, relu row / column sums, adjacency (same arithmetic as s2v_forward.cu) ----
        for (int i = warp; i < m; i += nw) {
            const float bi = (float)bg[i];
            float ss = 0.f;
            for (int j = lane; j < n; j += 32) {
                const float x = (float)Ag[(size_t)i * n + j];
                ss = fmaf(x, x, ss);
            }
            ss = gwsum(ss) + bi * bi;
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            float cs = 0.f, sp = 0.f, sn = 0.f, cnt = 0.f;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                float x = 0.f;
                if (j < n) x = (float)Ag[(size_t)i * n + j] * inv;
                const bool nz = (j < n) && (x != 0.f);
                const unsigned word = __ballot_sync(0xffffffffu, nz);
                if (lane == 0) adj[i * NW32 + (j0 >> 5)] = word;
                if (j < n) {
                    cs = fmaf(x, cj[j], cs);
                    sp += fmaxf(x, 0.f);
                    sn += fmaxf(-x, 0.f);
                    cnt += nz ? 1.f : 0.f;
                    atomicAdd(&Cp[j], fmaxf(x, 0.f));
                    atomicAdd(&Cn[j], fmaxf(-x, 0.f));
                    if (nz) {
                        atomicAdd(&ccnt[j], 1.f);
                        atomicOr(&adjT[j * MW32 + (i >> 5)], 1u << (i & 31));
                    }
                }
            }
            cs = gwsum(cs); sp = gwsum(sp); sn = gwsum(sn); cnt = gwsum(cnt);
            if (lane == 0) {
                rb[i] = bi * inv; cosv[i] = cs; Sp[i] = sp; Sn[i] = sn;
                rinv[i] = 1.f / fmaxf(cnt, 1e-12f);
            }
        }
        __syncthreads();
        for (int j = tid; j < n; j += nt) cinv[j] = 1.f / fmaxf(ccnt[j], 1e-12f);

        // ---- base and round 0 (the embeddings start at zero) ---------------------------------------------------------------
        for (int e = tid; e < NP * PP; e += nt) {
            const int q = e / PP, l = e - q * PP;
            float val = 0.f;
            if (l < p) {
                val = __ldg(t0 + l);
                if (q < m) {
                    val += __ldg(t1c + 4 * l) * fiq(q) + __ldg(t1c + 4 * l + 1) * rb[q] + __ldg(t1c + 4 * l + 2) * fbq(q) + __ldg(t1c + 4 * l + 3) * cosv[q];
                    val += w3cp[l] * Sp[q] + w3cn[l] * Sn[q];
                } else {
                    const int j = q - m;
                    val += __ldg(t1v + l) * cj[j] + w3vp[l] * Cp[j] + w3vn[l] * Cn[j];
                }
            }
            base[e] = val;
            if (T >= 1) MU[e] = (l < p) ? fmaxf(val, 0.f) : 0.f;
            DP[e] = 0.f;
        }
        __syncthreads();

        // ---- forward rounds 1 .. T-1 --------------------------------------------------------------------------------------
        for (int t = 1; t < T; ++t) {
            const float* mu_in = MU + (size_t)(t - 1) * NV;
            float* mu_out = MU + (size_t)t * NV;
            float* ag = AG + (size_t)t * NV;
            // term2 positions: q < n <- mean of mu_c over the rows adjacent to column q; q >= n <- mean of mu_v over row q - n
            bit_aggregate(ag, mu_in, adjT, MW32, n, nullptr, cinv);
            bit_aggregate(ag + (size_t)n * PP, mu_in + (size_t)m * PP, adj, NW32, m, nullptr, rinv);
            __syncthreads();
            auto epi = [&](int q, int kb, float v0, float v1, float v2, float v3) {
                const float vv[4] = {v0, v1, v2, v3};
                float o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) o[u] = (kb + u < p) ? fmaxf(base[(size_t)q * PP + kb + u] + vv[u], 0.f) : 0.f;
                *reinterpret_cast<float4*>(mu_out + (size_t)q * PP + kb) = make_float4(o[0], o[1], o[2], o[3]);
            };
            node_product(w2cT, ag, 0, n, epi);
            node_product(w2vT, ag, n, NP, epi);
            __syncthreads();
        }
        const float* muT = MU + (size_t)(T >= 1 ? T - 1 : 0) * NV;      // T == 0: treated as zeros below

        // ---- head forward -------------------------------------------------------------------------------------------------
        {   // group means: warps stride over the nodes, lanes over l
            float c0 = 0.f, c1 = 0.f, v0 = 0.f, v1 = 0.f;
            if (T >= 1) {
                for (int q = warp; q < NP; q += nw) {
                    const float* x = muT + (size_t)q * PP;
                    const float x0 = (lane < PP) ? x[lane] : 0.f, x1 = (lane + 32 < PP) ? x[lane + 32] : 0.f;
                    if (q < m) { c0 += x0; c1 += x1; } else { v0 += x0; v1 += x1; }
                }
            }
            if (lane < PP) { part[((size_t)warp * kPartSlots + 0) * PP + lane] = c0; part[((size_t)warp * kPartSlots + 1) * PP + lane] = v0; }
            if (lane + 32 < PP) { part[((size_t)warp * kPartSlots + 0) * PP + lane + 32] = c1; part[((size_t)warp * kPartSlots + 1) * PP + lane + 32] = v1; }
        }
        for (int l = tid; l < PP; l += nt) { meanc[l] = 0.f; meanv[l] = 0.f; }
        __syncthreads();
        fold_part(0, meanc);
        fold_part(1, meanv);
        __syncthreads();
        for (int l = tid; l < p; l += nt) { meanc[l] /= (float)m; meanv[l] /= (float)n; }
        __syncthreads();
        gmatvec(t6c, p, meanc, tmp1, warp, lane, nw);
        gmatvec(t6v, p, meanv, tmp2, warp, lane, nw);
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            u6pre[l] = tmp1[l] + tmp2[l];
            u6r[l] = fmaxf(u6pre[l], 0.f);
        }
        if (T >= 1) {
            node_product(t7T, muT, 0, m, [&](int q, int kb, float v0, float v1, float v2, float v3) {
                *reinterpret_cast<float4*>(Z + (size_t)q * PP + kb) =
                    make_float4(fmaxf(v0, 0.f), fmaxf(v1, 0.f), fmaxf(v2, 0.f), fmaxf(v3, 0.f));
            });
        } else {
            for (int e = tid; e < m * PP; e += nt) Z[e] = 0.f;
        }
        __syncthreads();
        for (int i = tid; i < m; i += nt) {
            float s0 = 0.f, s1 = 0.f;
            for (int l = 0; l < p; ++l) {
                s0 = fmaf(__ldg(t8 + l), u6r[l], s0);
                s1 = fmaf(__ldg(t8 + W8 + l), u6r[l], s1);
            }
            for (int k = 0; k < p; ++k) {
                const float z = Z[(size_t)i * PP + k];
                s0 = fmaf(__ldg(t8 + p + k), z, s0);
                s1 = fmaf(__ldg(t8 + W8 + p + k), z, s1);
            }
            const float f0 = fiq(i), f1 = rb[i], f2 = fbq(i), f3 = cosv[i];
            s0 += __ldg(t8 + 2 * p) * f0 + __ldg(t8 + 2 * p + 1) * f1 + __ldg(t8 + 2 * p + 2) * f2 + __ldg(t8 + 2 * p + 3) * f3;
            s1 += __ldg(t8 + W8 + 2 * p) * f0 + __ldg(t8 + W8 + 2 * p + 1) * f1 + __ldg(t8 + W8 + 2 * p + 2) * f2 + __ldg(t8 + W8 + 2 * p + 3) * f3;
            const float mx = fmaxf(s0, s1);
            const float lse = mx + logf(expf(s0 - mx) + expf(s1 - mx));
            const float p0 = expf(s0 - lse), p1 = expf(s1 - lse);
            const int y = (yl[i] == 1) ? 1 : 0;
            const float w = (yl[i] >= 2) ? 0.f : (y ? a.w1 : a.w0);      // label 2: row outside in_loss
            lossn[i] = -w * (y ? (s1 - lse) : (s0 - lse));
            ds0[i] = w * (p0 - (y == 0 ? 1.f : 0.f));
            ds1[i] = w * (p1 - (y == 1 ? 1.f : 0.f));
        }
        __syncthreads();

        // ---- head backward ------------------------------------------------------------------------------------------------
        if (warp == 0) {
            float s0 = 0.f, s1 = 0.f, ls = 0.f, f0a = 0.f, f0b = 0.f, f1a = 0.f, f1b = 0.f, f2a = 0.f, f2b = 0.f, f3a = 0.f, f3b = 0.f;
            for (int i = lane; i < m; i += 32) {
                s0 += ds0[i]; s1 += ds1[i]; ls += lossn[i];
                f0a += ds0[i] * fiq(i); f0b += ds1[i] * fiq(i); f2a += ds0[i] * fbq(i); f2b += ds1[i] * fbq(i);
                f1a += ds0[i] * rb[i]; f1b += ds1[i] * rb[i]; f3a += ds0[i] * cosv[i]; f3b += ds1[i] * cosv[i];
            }
            s0 = gwsum(s0); s1 = gwsum(s1); ls = gwsum(ls);
            f0a = gwsum(f0a); f0b = gwsum(f0b); f2a = gwsum(f2a); f2b = gwsum(f2b);
            f1a = gwsum(f1a); f1b = gwsum(f1b); f3a = gwsum(f3a); f3b = gwsum(f3b);
            if (lane == 0) {
                scal[0] = s0; scal[1] = s1; loss_cta += (double)ls;
                g_t8[2 * PP + 0] += f0a;  g_t8[4 * PP + 2 * PP + 0] += f0b;
                g_t8[2 * PP + 1] += f1a;  g_t8[4 * PP + 2 * PP + 1] += f1b;
                g_t8[2 * PP + 2] += f2a;  g_t8[4 * PP + 2 * PP + 2] += f2b;
                g_t8[2 * PP + 3] += f3a;  g_t8[4 * PP + 2 * PP + 3] += f3b;
            }
        }
        {   // d t8 (z block): per-warp partial sums over the nodes (lanes over k); d z in place of z
            float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
            for (int i = warp; i < m; i += nw) {
                const float d0 = ds0[i], d1 = ds1[i];
                float* zrow = Z + (size_t)i * PP;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = lane + 32 * h;
                    if (k < p) {
                        const float z = zrow[k];
                        if (h == 0) { g00 = fmaf(d0, z, g00); g10 = fmaf(d1, z, g10); } else { g01 = fmaf(d0, z, g01); g11 = fmaf(d1, z, g11); }
                        zrow[k] = (z > 0.f) ? (__ldg(t8 + p + k) * d0 + __ldg(t8 + W8 + p + k) * d1) : 0.f;
                    }
                }
            }
            if (lane < PP) { part[((size_t)warp * kPartSlots + 2) * PP + lane] = g00; part[((size_t)warp * kPartSlots + 3) * PP + lane] = g10; }
            if (lane + 32 < PP) { part[((size_t)warp * kPartSlots + 2) * PP + lane + 32] = g01; part[((size_t)warp * kPartSlots + 3) * PP + lane + 32] = g11; }
        }
        __syncthreads();
        fold_part(2, g_t8 + PP);
        fold_part(3, g_t8 + 4 * PP + PP);
        for (int l = tid; l < p; l += nt) {
            const float S0 = scal[0], S1 = scal[1];
            g_t8[l] += S0 * u6r[l];
            g_t8[4 * PP + l] += S1 * u6r[l];
            const float g = __ldg(t8 + l) * S0 + __ldg(t8 + W8 + l) * S1;
            du6[l] = (u6pre[l] > 0.f) ? g : 0.f;
        }
        __syncthreads();
        rank1_acc(a_t6c, du6, meanc);
        rank1_acc(a_t6v, du6, meanv);
        gmatvecT(t6c, p, du6, tmp1, warp, lane, nw);
        gmatvecT(t6v, p, du6, tmp2, warp, lane, nw);
        if (T >= 1) reduce_nodes(a_t7, Z, muT, 0, m);            // d t7[k][l] += sum_i dz[i][k] mu[i][l]
        __syncthreads();
        if (T >= 1) {
            // d mu (last round): constraints t7^T dz + t6c^T du6 / m, variables t6v^T du6 / n
            node_product(t7N, Z, 0, m, [&](int q, int lb, float v0, float v1, float v2, float v3) {
                const float im = 1.f / (float)m;
                *reinterpret_cast<float4*>(D1 + (size_t)q * PP + lb) =
                    make_float4(lb + 0 < p ? v0 + tmp1[lb + 0] * im : 0.f, lb + 1 < p ? v1 + tmp1[lb + 1] * im : 0.f,
                                lb + 2 < p ? v2 + tmp1[lb + 2] * im : 0.f, lb + 3 < p ? v3 + tmp1[lb + 3] * im : 0.f);
            });
            for (int e = tid; e < n * PP; e += nt) {
                const int l = e % PP;
                D1[(size_t)m * PP + e] = (l < p) ? tmp2[l] / (float)n : 0.f;
            }
        }
        __syncthreads();

        // ---- rounds backward ----------------------------------------------------------------------------------------------
        for (int t = T - 1; t >= 0; --t) {
            const float* mu_out = MU + (size_t)t * NV;
            for (int e = tid; e < NP * PP; e += nt) {
                const float d = (mu_out[e] > 0.f) ? D1[e] : 0.f;
                D1[e] = d;
                DP[e] += d;
            }
            __syncthreads();
            if (t >= 1) {
                const float* ag = AG + (size_t)t * NV;
                reduce_nodes(a_t2c, D1, ag, 0, n);
                reduce_nodes(a_t2v, D1, ag, n, NP);
                auto epi = [&](int q, int lb, float v0, float v1, float v2, float v3) {
                    *reinterpret_cast<float4*>(D2 + (size_t)q * PP + lb) = make_float4(v0, v1, v2, v3);
                };
                node_product(w2cN, D1, 0, n, epi);               // d agg_v[:, q] = t2c^T d pre[:, q]
                node_product(w2vN, D1, n, NP, epi);              // d agg_c[:, q - n] = t2v^T d pre[:, q]
                __syncthreads();
                // d mu_c[i] = sum_{j adjacent} d agg_v[j] / ccnt_j ;  d mu_v[j] = sum_{i adjacent} d agg_c[i] / rcnt_i
                bit_aggregate(D1, D2, adj, NW32, m, cinv, nullptr);
                bit_aggregate(D1 + (size_t)m * PP, D2 + (size_t)n * PP, adjT, MW32, n, rinv, nullptr);
                __syncthreads();
            }
        }

        // ---- base parameters: reductions of DP over the nodes (lanes over l, warps stride over the nodes) --------------------
        {
            float r[11][2];                 // 9, 10: d base weighted by the row flags (t1c columns 0 and 2)
#pragma unroll
            for (int q = 0; q < 11; ++q) { r[q][0] = 0.f; r[q][1] = 0.f; }
            for (int q = warp; q < NP; q += nw) {
                const float* d = DP + (size_t)q * PP;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int l = lane + 32 * h;
                    if (l < p) {
                        const float dv = d[l];
                        if (q < m) {
                            r[0][h] += dv;
                            r[9][h] = fmaf(dv, fiq(q), r[9][h]); r[10][h] = fmaf(dv, fbq(q), r[10][h]);
                            r[1][h] = fmaf(dv, rb[q], r[1][h]); r[2][h] = fmaf(dv, cosv[q], r[2][h]);
                            r[3][h] = fmaf(dv, Sp[q], r[3][h]); r[4][h] = fmaf(dv, Sn[q], r[4][h]);
                        } else {
                            const int j = q - m;
                            r[5][h] += dv;
                            r[6][h] = fmaf(dv, cj[j], r[6][h]); r[7][h] = fmaf(dv, Cp[j], r[7][h]); r[8][h] = fmaf(dv, Cn[j], r[8][h]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 11; ++q) {
                if (lane < PP) part[((size_t)warp * kPartSlots + q) * PP + lane] = r[q][0];
                if (lane + 32 < PP) part[((size_t)warp * kPartSlots + q) * PP + lane + 32] = r[q][1];
            }
        }
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            float s[11];
#pragma unroll
            for (int q = 0; q < 11; ++q) {
                s[q] = 0.f;
                for (int w = 0; w < nw; ++w) s[q] += part[((size_t)w * kPartSlots + q) * PP + l];
            }
            g_t0[l] += s[0] + s[5];
            g_t1c[0 * PP + l] += s[9]; g_t1c[1 * PP + l] += s[1]; g_t1c[2 * PP + l] += s[10]; g_t1c[3 * PP + l] += s[2];
            gw3[l] += s[3]; gw3[PP + l] += s[4];
            g_t1v[l] += s[6]; gw3[2 * PP + l] += s[7]; gw3[3 * PP + l] += s[8];
        }
        __syncthreads();
    }

    // ---- per CTA: push d w3 through t3 / t4, then everything to the global gradient ------------------------------------------
    auto add = [&](int idx, float g) {
        if (g != 0.f) atomicAdd(a.grad + idx, g);
    };
    gmatvecT(t3c, p, gw3, tmp1, warp, lane, nw);
    gmatvecT(t3c, p, gw3 + PP, tmp2, warp, lane, nw);
    __syncthreads();
    for (int l = tid; l < p; l += nt) {
        const float tv = __ldg(t4c + l);
        add(o_t4c + l, (tv > 0.f ? tmp1[l] : 0.f) - (tv < 0.f ? tmp2[l] : 0.f));
    }
    __syncthreads();
    gmatvecT(t3v, p, gw3 + 2 * PP, tmp1, warp, lane, nw);
    gmatvecT(t3v, p, gw3 + 3 * PP, tmp2, warp, lane, nw);
    __syncthreads();
    for (int l = tid; l < p; l += nt) {
        const float tv = __ldg(t4v + l);
        add(o_t4v + l, (tv > 0.f ? tmp1[l] : 0.f) - (tv < 0.f ? tmp2[l] : 0.f));
        add(o_t0 + l, g_t0[l]);
        add(o_t1v + l, g_t1v[l]);
#pragma unroll
        for (int q = 0; q < 4; ++q) add(o_t1c + 4 * l + q, g_t1c[q * PP + l]);
        add(o_t8 + l, g_t8[l]);
        add(o_t8 + p + l, g_t8[PP + l]);
        add(o_t8 + W8 + l, g_t8[4 * PP + l]);
        add(o_t8 + W8 + p + l, g_t8[4 * PP + PP + l]);
    }
    if (tid < 4) {
        add(o_t8 + 2 * p + tid, g_t8[2 * PP + tid]);
        add(o_t8 + W8 + 2 * p + tid, g_t8[4 * PP + 2 * PP + tid]);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int e = tid + s * kGT;
        if (e < p * p) {
            const int k = e / p, l = e - k * p;
            add(o_t2c + e, a_t2c[s]);
            add(o_t2v + e, a_t2v[s]);
            add(o_t6c + e, a_t6c[s]);
            add(o_t6v + e, a_t6v[s]);
            add(o_t7 + e, a_t7[s]);
            add(o_t3c + e, gw3[k] * r4[l] + gw3[PP + k] * r4[PP + l]);
            add(o_t3v + e, gw3[2 * PP + k] * r4[2 * PP + l] + gw3[3 * PP + k] * r4[3 * PP + l]);
        }
    }
    if (tid == 0) atomicAdd(a.loss, loss_cta);
}

}  // namespace

size_t s2v_general_grad_smem_bytes(int m, int n, int p) { return ggrad_layout(m, n, p).total * 4; }
size_t s2v_general_grad_scratch_floats(int m, int n, int p, int T) { return ggrad_scratch_floats(m, n, p, T); }
// two CTAs per SM (latency-bound: the embeddings of every round live in an L2-resident scratch)
int s2v_general_grad_grid(long long B, int sm_count) { return (int)((B < 2 * sm_count) ? B : 2 * sm_count); }

cudaError_t launch_s2v_bipartite_general_grad(const S2vGGradArgs& a, int grid, long long smem_optin, cudaStream_t st, const char** why) {
    *why = "";
    if (a.p > 64) { *why = "classifier backward (general adjacency): p > 64 is not supported"; return cudaErrorInvalidValue; }
    const size_t smem = s2v_general_grad_smem_bytes(a.m, a.n, a.p);
    if ((long long)smem > smem_optin) { *why = "classifier backward (general adjacency): weights and adjacency masks do not fit in shared memory"; return cudaErrorInvalidValue; }
    const int slots = (a.p * a.p + kGT - 1) / kGT;
    auto kern = slots <= 2 ? s2v_bipartite_general_grad_kernel<2> : slots <= 4 ? s2v_bipartite_general_grad_kernel<4>
                : slots <= 8 ? s2v_bipartite_general_grad_kernel<8> : s2v_bipartite_general_grad_kernel<16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kGT, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ddb
