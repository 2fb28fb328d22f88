// Batched loss + gradient of the reference classifier, COMPLETE variant: one CTA per LP instance, fp32, forward kept in
// shared memory, backward by hand, gradients summed over the batch.
//
// Replaces, for graph == 'complete', the inner loop of train_net (reference src/ml/train.py:59-66) with its criterion
// NLLLoss(weight=[w0, w1], size_average=False) (src/benchmark.py:70-75): the summed loss of a batch and the accumulated
// gradient of every parameter in the flat state_dict order of include/ddb200.h (4).
//
// Forward = s2v_complete_kernel (s2v_forward.cu; s2v.py:124-187, 91-122, quirk B10 kept) on the relu row sums of
// W = G G^T that the tcgen05 Gram kernel (s2v_gram_tc.cu) produces -- W depends on the data only, so nothing flows back
// through it:
//   base_i   = t0 + t1 + w3p Wp_i + w3n Wn_i + s,      w3p = t3rr relu(t4rr), w3n = t3rr relu(-t4rr),
//              s = t4rc . (relu(t4rc) S+ + relu(-t4rc) S-)   (B10: a scalar),  S+- = sum_j relu(+-W[m][j])
//   round t  : mu_i' = relu(base_i + t2rr mu_i + t2rc mu_c),  mu_c' = relu(t0 + t2cr mean_i(mu_i) + t3cr relu_cr)
//   head     : scores_i = t8 [relu(t6r mean_i(mu_i) + t6c mu_c) ; relu(t7 mu_i)]
// Unlike the bipartite variant on dense LPs, every round needs PER-NODE gradients (t2rr mu_i is per node): the kernel
// keeps the embeddings of every round in shared memory (T p m floats; activity = embedding > 0), runs the p x p x m
// products of forward and backward as 4 x 4 register tiles and accumulates the p x p parameter gradients in REGISTERS
// across the instances of a CTA (one atomic add per parameter per CTA at the end).
#include "common.cuh"

namespace ddb {

struct S2vCGradArgs {
    long long B;
    int m, p, T;
    const float* gram;         // [B][3][gram_pitch]: Wp, Wn, wc from the Gram kernel
    int gram_pitch;
    const float* params;
    const uint8_t* labels;     // [B, m] 0 / 1; 2 = row outside the item's in_loss set (no loss, no gradient from it)
    const uint8_t* row_ineq;   // [B, m] node features of the row nodes of MPS / PLNN items (1 = inequality row), nullable = all 1
    float w0, w1;
    float* grad;               // [param_count], accumulated with atomics (zeroed by the caller)
    double* loss;              // scalar, accumulated with atomics (zeroed by the caller)
};

namespace {

constexpr int kCT = 512;       // threads per CTA (one CTA per SM: the stored embeddings are the footprint)

__host__ __device__ inline int cpad4(int v) { return (v + 3) & ~3; }
// pitch of a [p][m] array: a multiple of 4 floats with pitch / 4 odd, so that lanes reading float4 of consecutive rows
// hit different bank groups
__host__ __device__ inline int cpitch(int m) {
    int v = cpad4(m);
    if (((v / 4) & 1) == 0) v += 4;
    return v;
}

__device__ __forceinline__ float cwsum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// y[k] = sum_l W[k][l] x[l] (row-major p x p in global memory), one warp per output
__device__ __forceinline__ void cmatvec(const float* __restrict__ W, int p, const float* x, float* y, int warp, int lane, int nw) {
    for (int k = warp; k < p; k += nw) {
        float acc = 0.f;
        for (int l = lane; l < p; l += 32) acc = fmaf(__ldg(W + k * p + l), x[l], acc);
        acc = cwsum(acc);
        if (lane == 0) y[k] = acc;
    }
}
// y[l] = sum_k W[k][l] x[k]
__device__ __forceinline__ void cmatvecT(const float* __restrict__ W, int p, const float* x, float* y, int warp, int lane, int nw) {
    for (int l = warp; l < p; l += nw) {
        float acc = 0.f;
        for (int k = lane; k < p; k += 32) acc = fmaf(__ldg(W + k * p + l), x[k], acc);
        acc = cwsum(acc);
        if (lane == 0) y[l] = acc;
    }
}

// out(r, i) = sum_q Wq[q][r] * X[q][i]  for r < p, i < MP: register tile, thread = 4 nodes x 4 outputs (two LDS.128 per
// 16 FMAs); Wq has pitch PP, X pitch MP.  The epilogue receives (r, i0, four accumulators of nodes i0 .. i0+3).
template <class Epi>
__device__ __forceinline__ void tile_product(const float* __restrict__ Wq, int PP, int p, const float* X, int MP, int tid, int nt,
                                             Epi epi) {
    const int NG = MP / 4, KG = PP / 4;
    for (int w = tid; w < NG * KG; w += nt) {
        const int ng = w % NG, kg = w / NG;
        float acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
        for (int l = 0; l < p; ++l) {
            const float4 xv = *reinterpret_cast<const float4*>(X + l * MP + 4 * ng);
            const float4 wv = *reinterpret_cast<const float4*>(Wq + l * PP + 4 * kg);
            const float xq[4] = {xv.x, xv.y, xv.z, xv.w};
            const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(wr[r], xq[q], acc[q][r]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (4 * kg + r < p) epi(4 * kg + r, 4 * ng, acc[0][r], acc[1][r], acc[2][r], acc[3][r]);
    }
}

struct CGradLayout {           // offsets in floats
    size_t t2N, t2T, t7N, t7T, mus, X, Y, vecs, total;
};
__host__ __device__ inline CGradLayout cgrad_layout(int m, int p, int T) {
    const int PP = cpad4(p), MP = cpitch(m);
    CGradLayout L;
    size_t off = 0;
    L.t2N = off; off += (size_t)p * PP;
    L.t2T = off; off += (size_t)p * PP;
    L.t7N = off; off += (size_t)p * PP;
    L.t7T = off; off += (size_t)p * PP;
    L.mus = off; off += (size_t)(T > 0 ? T : 1) * p * MP;
    L.X = off;   off += (size_t)p * MP;
    L.Y = off;   off += (size_t)p * MP;
    L.vecs = off;
    off += (size_t)7 * MP + (size_t)(33 + 2 * (T + 1)) * PP + 64;
    L.total = off;
    return L;
}

// S = slots of the element-wise p x p register accumulators: ceil(p * p / kCT)
template <int S>
__global__ void __launch_bounds__(kCT, 1) s2v_complete_grad_kernel(S2vCGradArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int m = a.m, p = a.p, T = a.T, PP = cpad4(p), MP = cpitch(m);
    const CGradLayout L = cgrad_layout(m, p, T);
    float* t2N = sm + L.t2N;     // t2N[k][l] = t2rr[k][l]
    float* t2T = sm + L.t2T;     // t2T[l][k] = t2rr[k][l]
    float* t7N = sm + L.t7N;
    float* t7T = sm + L.t7T;
    float* mus = sm + L.mus;     // [T][p][MP]: mu_r after round 0 .. T-1 (padding nodes hold 0)
    float* X = sm + L.X;         // z / dz, then ping-pong with Y
    float* Y = sm + L.Y;         // d mu_r / d pre
    float* v = sm + L.vecs;
    float* Wp = v;     v += MP;   float* Wn = v;     v += MP;   float* wc = v;    v += MP;
    float* ds0 = v;    v += MP;   float* ds1 = v;    v += MP;   float* lossn = v; v += MP;
    float* nfs = v;    v += MP;   // node feature of every row node (1 on random LPs), 0 on the padding nodes
    float* w3p = v;    v += PP;   float* w3n = v;    v += PP;   float* r4p = v;   v += PP;   float* r4n = v;   v += PP;
    float* k0 = v;     v += PP;   float* u3c = v;    v += PP;   float* relucr = v; v += PP;  float* y1 = v;    v += PP;
    float* y2 = v;     v += PP;   float* u6pre = v;  v += PP;   float* u6r = v;   v += PP;   float* du6 = v;   v += PP;
    float* tmpA = v;   v += PP;   float* tmpB = v;   v += PP;   float* srow = v;  v += PP;   float* swp = v;   v += PP;
    float* swn = v;    v += PP;   float* dpc = v;    v += PP;   float* dmuc = v;  v += PP;   float* gu3i = v;  v += PP;
    float* g_t0 = v;   v += PP;   float* g_t1 = v;   v += PP;   float* g_w3p = v; v += PP;   float* g_w3n = v; v += PP;
    float* g_t4rc = v; v += PP;   float* g_t4cr = v; v += PP;   float* g_t8 = v;  v += 4 * PP;
    float* sfl = v;    v += PP;   // per-coordinate sums of d pre weighted by the node features (gradient of t1)
    float* mucs = v;   v += (size_t)(T + 1) * PP;     // mu_c before round 0 (zero), after round 0, ...
    float* means = v;  v += (size_t)(T + 1) * PP;     // mean_i mu_r, same indexing
    float* scal = v;                                   // [0] s (B10), [1] S+, [2] S-, [3] S0, [4] S1, [5] d s of this instance

    const int tid = threadIdx.x, nt = kCT, lane = tid & 31, warp = tid >> 5, nw = kCT / 32;
    const float* P = a.params;
    const float* t0 = P;     P += p;        const int o_t0 = 0;
    const float* t1 = P;     P += p;        const int o_t1 = o_t0 + p;
    const float* t2rr = P;   P += p * p;    const int o_t2rr = o_t1 + p;
    const float* t2rc = P;   P += p * p;    const int o_t2rc = o_t2rr + p * p;
    const float* t2cr = P;   P += p * p;    const int o_t2cr = o_t2rc + p * p;
    const float* t3rr = P;   P += p * p;    const int o_t3rr = o_t2cr + p * p;
    P += p * p;                             const int o_t3rc = o_t3rr + p * p;   // unused by the forward (B10): zero gradient
    const float* t3cr = P;   P += p * p;    const int o_t3cr = o_t3rc + p * p;
    const float* t4rr = P;   P += p;        const int o_t4rr = o_t3cr + p * p;
    const float* t4rc = P;   P += p;        const int o_t4rc = o_t4rr + p;
    const float* t4cr = P;   P += p;        const int o_t4cr = o_t4rc + p;
    const float* t6r = P;    P += p * p;    const int o_t6r = o_t4cr + p;
    const float* t6c = P;    P += p * p;    const int o_t6c = o_t6r + p * p;
    const float* t7 = P;     P += p * p;    const int o_t7 = o_t6c + p * p;
    const float* t8 = P;                    const int o_t8 = o_t7 + p * p;
    const int W8 = 2 * p;

    for (int e = tid; e < p * PP; e += nt) {
        const int r = e / PP, q = e - r * PP;
        t2N[e] = (q < p) ? __ldg(t2rr + r * p + q) : 0.f;
        t2T[e] = (q < p) ? __ldg(t2rr + q * p + r) : 0.f;
        t7N[e] = (q < p) ? __ldg(t7 + r * p + q) : 0.f;
        t7T[e] = (q < p) ? __ldg(t7 + q * p + r) : 0.f;
    }
    for (int l = tid; l < PP; l += nt) {
        const float tv = (l < p) ? __ldg(t4rr + l) : 0.f;
        r4p[l] = fmaxf(tv, 0.f);
        r4n[l] = fmaxf(-tv, 0.f);
        k0[l] = (l < p) ? __ldg(t0 + l) + __ldg(t1 + l) : 0.f;
        g_t0[l] = 0.f; g_t1[l] = 0.f; g_w3p[l] = 0.f; g_w3n[l] = 0.f; g_t4rc[l] = 0.f; g_t4cr[l] = 0.f;
        g_t8[l] = 0.f; g_t8[PP + l] = 0.f; g_t8[2 * PP + l] = 0.f; g_t8[3 * PP + l] = 0.f;
        mucs[l] = 0.f; means[l] = 0.f;
    }
    __syncthreads();
    cmatvec(t3rr, p, r4p, w3p, warp, lane, nw);
    cmatvec(t3rr, p, r4n, w3n, warp, lane, nw);
    __syncthreads();

    // register accumulators of the p x p parameter gradients
    //   element-wise (e = tid + s kCT -> (k, l) = (e / p, e % p)): rank-1 terms of t2rc, t2cr, t3cr, t6r, t6c
    //   reductions over the nodes (w = tid + s kCT -> l = w % p, k0 = 4 (w / p)): t2rr, t7
    float a_t2rc[S], a_t2cr[S], a_t3cr[S], a_t6r[S], a_t6c[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { a_t2rc[s] = 0.f; a_t2cr[s] = 0.f; a_t3cr[s] = 0.f; a_t6r[s] = 0.f; a_t6c[s] = 0.f; }
    constexpr int SW = (S >= 4) ? 2 : 1;      // (PP / 4) p work items: <= 2 kCT up to p = 64, <= kCT up to p = 32
    float a_t2rr[SW][4], a_t7[SW][4];
#pragma unroll
    for (int s = 0; s < SW; ++s)
#pragma unroll
        for (int r = 0; r < 4; ++r) { a_t2rr[s][r] = 0.f; a_t7[s][r] = 0.f; }
    double loss_cta = 0.0;
    const float inv_m = 1.f / (float)m;

    // acc[k][l] += sum_i Am[k][i] * Bm[l][i]   (thread = 4 outputs k x one l; Am rows are broadcast reads)
    auto reduce_product = [&](float (&acc)[SW][4], const float* Am, const float* Bm) {
#pragma unroll
        for (int s = 0; s < SW; ++s) {
            const int w = tid + s * kCT;
            if (w < (PP / 4) * p) {
                const int l = w % p, kq = (w / p) * 4;
                float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
                const float* b = Bm + l * MP;
                const float* a0 = Am + (kq + 0 < p ? kq + 0 : 0) * MP;
                const float* a1 = Am + (kq + 1 < p ? kq + 1 : 0) * MP;
                const float* a2 = Am + (kq + 2 < p ? kq + 2 : 0) * MP;
                const float* a3 = Am + (kq + 3 < p ? kq + 3 : 0) * MP;
                for (int i = 0; i < MP; i += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(b + i);
                    const float4 z0 = *reinterpret_cast<const float4*>(a0 + i);
                    const float4 z1 = *reinterpret_cast<const float4*>(a1 + i);
                    const float4 z2 = *reinterpret_cast<const float4*>(a2 + i);
                    const float4 z3 = *reinterpret_cast<const float4*>(a3 + i);
                    r0 = fmaf(z0.x, x.x, fmaf(z0.y, x.y, fmaf(z0.z, x.z, fmaf(z0.w, x.w, r0))));
                    r1 = fmaf(z1.x, x.x, fmaf(z1.y, x.y, fmaf(z1.z, x.z, fmaf(z1.w, x.w, r1))));
                    r2 = fmaf(z2.x, x.x, fmaf(z2.y, x.y, fmaf(z2.z, x.z, fmaf(z2.w, x.w, r2))));
                    r3 = fmaf(z3.x, x.x, fmaf(z3.y, x.y, fmaf(z3.z, x.z, fmaf(z3.w, x.w, r3))));
                }
                acc[s][0] += r0; acc[s][1] += r1; acc[s][2] += r2; acc[s][3] += r3;
            }
        }
    };
    // acc[k][l] += u[k] * w[l]
    auto rank1_acc = [&](float (&acc)[S], const float* u, const float* w) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int e = tid + s * kCT;
            if (e < p * p) acc[s] = fmaf(u[e / p], w[e % p], acc[s]);
        }
    };
    // mean over the nodes of a [p][MP] array (padding nodes hold 0)
    auto node_means = [&](const float* Mu, float* out) {
        for (int l = warp; l < p; l += nw) {
            float sr = 0.f;
            for (int i = lane; i < MP; i += 32) sr += Mu[l * MP + i];
            sr = cwsum(sr);
            if (lane == 0) out[l] = sr * inv_m;
        }
    };

    for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {
        const float* gr = a.gram + (size_t)lp * 3 * a.gram_pitch;
        const uint8_t* yl = a.labels + (size_t)lp * m;
        for (int i = tid; i < MP; i += nt) {
            Wp[i] = (i < m) ? __ldg(gr + i) : 0.f;
            Wn[i] = (i < m) ? __ldg(gr + a.gram_pitch + i) : 0.f;
            wc[i] = (i < m) ? __ldg(gr + 2 * a.gram_pitch + i) : 0.f;
            nfs[i] = (i < m) ? (a.row_ineq ? (float)a.row_ineq[(size_t)lp * m + i] : 1.f) : 0.f;
        }
        for (int l = tid; l < PP; l += nt) gu3i[l] = 0.f;
        __syncthreads();
        if (warp == 0) {
            float sp = 0.f, sn = 0.f;
            for (int j = lane; j < m; j += 32) { sp += fmaxf(wc[j], 0.f); sn += fmaxf(-wc[j], 0.f); }
            sp = cwsum(sp); sn = cwsum(sn);
            float acc = 0.f;
            for (int l = lane; l < p; l += 32) {
                const float tv = __ldg(t4rc + l);
                acc += tv * (fmaxf(tv, 0.f) * sp + fmaxf(-tv, 0.f) * sn);
            }
            acc = cwsum(acc);
            if (lane == 0) { scal[0] = acc; scal[1] = sp; scal[2] = sn; scal[5] = 0.f; }
        }
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            const float tv = __ldg(t4cr + l);
            relucr[l] = fmaxf(tv, 0.f) * scal[1] + fmaxf(-tv, 0.f) * scal[2];
        }
        __syncthreads();
        cmatvec(t3cr, p, relucr, u3c, warp, lane, nw);
        __syncthreads();
        const float sB = scal[0];

        // ---- forward rounds -------------------------------------------------------------------------------------------------
        // round 0: the embeddings start at zero, so it is the base alone
        if (T >= 1) {
            float* mu1 = mus;
            for (int e = tid; e < p * MP; e += nt) {
                const int l = e / MP, i = e - l * MP;
                const float kq = a.row_ineq ? __ldg(t0 + l) + __ldg(t1 + l) * nfs[i] : k0[l];
                mu1[e] = (i < m) ? fmaxf(kq + w3p[l] * Wp[i] + w3n[l] * Wn[i] + sB, 0.f) : 0.f;
            }
            for (int l = tid; l < p; l += nt) mucs[PP + l] = fmaxf(__ldg(t0 + l) + u3c[l], 0.f);
            __syncthreads();
            node_means(mu1, means + PP);
            __syncthreads();
        }
        for (int t = 1; t < T; ++t) {
            const float* mu_in = mus + (size_t)(t - 1) * p * MP;
            float* mu_out = mus + (size_t)t * p * MP;
            cmatvec(t2rc, p, mucs + t * PP, y1, warp, lane, nw);
            cmatvec(t2cr, p, means + t * PP, y2, warp, lane, nw);
            __syncthreads();
            tile_product(t2T, PP, p, mu_in, MP, tid, nt, [&](int k, int i0, float v0, float v1, float v2, float v3) {
                const float add = k0[k] + y1[k] + sB, wp = w3p[k], wn = w3n[k];
                const float t0k = __ldg(t0 + k), t1k = __ldg(t1 + k), y1k = y1[k];
                const float vv[4] = {v0, v1, v2, v3};
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + q;
                    const float addq = a.row_ineq ? (t0k + t1k * nfs[i < m ? i : 0]) + y1k + sB : add;
                    o[q] = (i < m) ? fmaxf(vv[q] + addq + wp * Wp[i] + wn * Wn[i], 0.f) : 0.f;
                }
                *reinterpret_cast<float4*>(mu_out + k * MP + i0) = make_float4(o[0], o[1], o[2], o[3]);
            });
            for (int l = tid; l < p; l += nt) mucs[(t + 1) * PP + l] = fmaxf(__ldg(t0 + l) + u3c[l] + y2[l], 0.f);
            __syncthreads();
            node_means(mu_out, means + (t + 1) * PP);
            __syncthreads();
        }
        const float* muT = (T >= 1) ? mus + (size_t)(T - 1) * p * MP : nullptr;
        const float* mucT = mucs + T * PP;
        const float* meanT = means + T * PP;

        // ---- head forward ---------------------------------------------------------------------------------------------------
        cmatvec(t6r, p, meanT, tmpA, warp, lane, nw);
        cmatvec(t6c, p, mucT, tmpB, warp, lane, nw);
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            u6pre[l] = tmpA[l] + tmpB[l];
            u6r[l] = fmaxf(u6pre[l], 0.f);
        }
        if (T >= 1) {
            tile_product(t7T, PP, p, muT, MP, tid, nt, [&](int k, int i0, float v0, float v1, float v2, float v3) {
                *reinterpret_cast<float4*>(X + k * MP + i0) =
                    make_float4(fmaxf(v0, 0.f), fmaxf(v1, 0.f), fmaxf(v2, 0.f), fmaxf(v3, 0.f));
            });
        } else {
            for (int e = tid; e < p * MP; e += nt) X[e] = 0.f;
        }
        __syncthreads();
        for (int i = tid; i < MP; i += nt) {
            float d0 = 0.f, d1 = 0.f, ln = 0.f;
            if (i < m) {
                float s0 = 0.f, s1 = 0.f;
                for (int l = 0; l < p; ++l) {
                    s0 = fmaf(__ldg(t8 + l), u6r[l], s0);
                    s1 = fmaf(__ldg(t8 + W8 + l), u6r[l], s1);
                }
                for (int k = 0; k < p; ++k) {
                    const float z = X[k * MP + i];
                    s0 = fmaf(__ldg(t8 + p + k), z, s0);
                    s1 = fmaf(__ldg(t8 + W8 + p + k), z, s1);
                }
                const float mx = fmaxf(s0, s1);
                const float lse = mx + logf(expf(s0 - mx) + expf(s1 - mx));
                const float p0 = expf(s0 - lse), p1 = expf(s1 - lse);
                const int y = (yl[i] == 1) ? 1 : 0;
                const float w = (yl[i] >= 2) ? 0.f : (y ? a.w1 : a.w0);      // label 2: row outside in_loss
                ln = -w * (y ? (s1 - lse) : (s0 - lse));
                d0 = w * (p0 - (y == 0 ? 1.f : 0.f));
                d1 = w * (p1 - (y == 1 ? 1.f : 0.f));
            }
            lossn[i] = ln; ds0[i] = d0; ds1[i] = d1;
        }
        __syncthreads();

        // ---- head backward --------------------------------------------------------------------------------------------------
        if (warp == 0) {
            float s0 = 0.f, s1 = 0.f, ls = 0.f;
            for (int i = lane; i < MP; i += 32) { s0 += ds0[i]; s1 += ds1[i]; ls += lossn[i]; }
            s0 = cwsum(s0); s1 = cwsum(s1); ls = cwsum(ls);
            if (lane == 0) { scal[3] = s0; scal[4] = s1; loss_cta += (double)ls; }
        }
        // d t8 (z block) and d z in place of z
        for (int k = warp; k < p; k += nw) {
            float g0 = 0.f, g1 = 0.f;
            const float w0k = __ldg(t8 + p + k), w1k = __ldg(t8 + W8 + p + k);
            for (int i = lane; i < MP; i += 32) {
                const float z = X[k * MP + i];
                g0 = fmaf(ds0[i], z, g0);
                g1 = fmaf(ds1[i], z, g1);
                X[k * MP + i] = (z > 0.f) ? (w0k * ds0[i] + w1k * ds1[i]) : 0.f;
            }
            g0 = cwsum(g0); g1 = cwsum(g1);
            if (lane == 0) { g_t8[PP + k] += g0; g_t8[3 * PP + k] += g1; }
        }
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            const float S0 = scal[3], S1 = scal[4];
            g_t8[l] += S0 * u6r[l];
            g_t8[2 * PP + l] += S1 * u6r[l];
            const float g = __ldg(t8 + l) * S0 + __ldg(t8 + W8 + l) * S1;
            du6[l] = (u6pre[l] > 0.f) ? g : 0.f;
        }
        __syncthreads();
        rank1_acc(a_t6r, du6, meanT);
        rank1_acc(a_t6c, du6, mucT);
        cmatvecT(t6r, p, du6, tmpA, warp, lane, nw);      // gradient of mean_i mu_r (last round)
        cmatvecT(t6c, p, du6, dmuc, warp, lane, nw);      // gradient of mu_c (last round)
        if (T >= 1) reduce_product(a_t7, X, muT);         // d t7[k][l] += sum_i dz[k][i] mu[l][i]
        __syncthreads();
        if (T >= 1) {
            // d mu_r (last round) = t7^T dz + t6r^T du6 / m
            tile_product(t7N, PP, p, X, MP, tid, nt, [&](int l, int i0, float v0, float v1, float v2, float v3) {
                const float add = tmpA[l] * inv_m;
                *reinterpret_cast<float4*>(Y + l * MP + i0) =
                    make_float4(i0 + 0 < m ? v0 + add : 0.f, i0 + 1 < m ? v1 + add : 0.f, i0 + 2 < m ? v2 + add : 0.f,
                                i0 + 3 < m ? v3 + add : 0.f);
            });
        }
        __syncthreads();

        // ---- rounds backward ------------------------------------------------------------------------------------------------
        float* dcur = Y;       // gradient of mu_r after round t
        float* dnext = X;
        for (int t = T - 1; t >= 0; --t) {
            const float* mu_out = mus + (size_t)t * p * MP;
            // d pre = d mu . [mu > 0], in place; per-coordinate sums over the nodes
            for (int l = warp; l < p; l += nw) {
                float s_all = 0.f, s_p = 0.f, s_n = 0.f, s_f = 0.f;
                for (int i = lane; i < MP; i += 32) {
                    const float d = (mu_out[l * MP + i] > 0.f) ? dcur[l * MP + i] : 0.f;
                    dcur[l * MP + i] = d;
                    s_all += d;
                    s_f = fmaf(d, nfs[i], s_f);
                    s_p = fmaf(d, Wp[i], s_p);
                    s_n = fmaf(d, Wn[i], s_n);
                }
                s_all = cwsum(s_all); s_p = cwsum(s_p); s_n = cwsum(s_n); s_f = cwsum(s_f);
                if (lane == 0) {
                    srow[l] = s_all; swp[l] = s_p; swn[l] = s_n; sfl[l] = s_f;
                    dpc[l] = (mucs[(t + 1) * PP + l] > 0.f) ? dmuc[l] : 0.f;
                }
            }
            __syncthreads();
            if (warp == 0) {
                float tot = 0.f;
                for (int l = lane; l < p; l += 32) tot += srow[l];
                tot = cwsum(tot);
                if (lane == 0) scal[5] += tot;
            }
            for (int l = tid; l < p; l += nt) {
                g_t0[l] += srow[l] + dpc[l];
                g_t1[l] += a.row_ineq ? sfl[l] : srow[l];
                g_w3p[l] += swp[l];
                g_w3n[l] += swn[l];
                gu3i[l] += dpc[l];
            }
            if (t >= 1) {
                const float* mu_in = mus + (size_t)(t - 1) * p * MP;
                rank1_acc(a_t2rc, srow, mucs + t * PP);
                rank1_acc(a_t2cr, dpc, means + t * PP);
                reduce_product(a_t2rr, dcur, mu_in);
                cmatvecT(t2cr, p, dpc, tmpB, warp, lane, nw);     // gradient of mean_i mu_r (input of round t)
                cmatvecT(t2rc, p, srow, tmpA, warp, lane, nw);    // gradient of mu_c (input of round t)
                __syncthreads();
                tile_product(t2N, PP, p, dcur, MP, tid, nt, [&](int l, int i0, float v0, float v1, float v2, float v3) {
                    const float add = tmpB[l] * inv_m;
                    *reinterpret_cast<float4*>(dnext + l * MP + i0) =
                        make_float4(i0 + 0 < m ? v0 + add : 0.f, i0 + 1 < m ? v1 + add : 0.f, i0 + 2 < m ? v2 + add : 0.f,
                                    i0 + 3 < m ? v3 + add : 0.f);
                });
                for (int l = tid; l < p; l += nt) dmuc[l] = tmpA[l];
                float* sw = dcur; dcur = dnext; dnext = sw;
            }
            __syncthreads();
        }

        // ---- per instance: the data-dependent scalars S+-, through s (B10) and relu_cr -------------------------------------------
        rank1_acc(a_t3cr, gu3i, relucr);
        cmatvecT(t3cr, p, gu3i, tmpA, warp, lane, nw);            // gradient of relu_cr
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            const float sp = scal[1], sn = scal[2];
            const float trc = __ldg(t4rc + l), tcr = __ldg(t4cr + l);
            g_t4rc[l] += scal[5] * 2.f * (fmaxf(trc, 0.f) * sp + fmaxf(-trc, 0.f) * sn);
            g_t4cr[l] += tmpA[l] * ((tcr > 0.f ? sp : 0.f) - (tcr < 0.f ? sn : 0.f));
        }
        __syncthreads();
    }

    // ---- per CTA: d w3 through t3rr / t4rr, then everything to the global gradient ------------------------------------------
    cmatvecT(t3rr, p, g_w3p, tmpA, warp, lane, nw);
    cmatvecT(t3rr, p, g_w3n, tmpB, warp, lane, nw);
    __syncthreads();
    auto add = [&](int idx, float g) {
        if (g != 0.f) atomicAdd(a.grad + idx, g);
    };
    for (int l = tid; l < p; l += nt) {
        const float tv = __ldg(t4rr + l);
        add(o_t0 + l, g_t0[l]);
        add(o_t1 + l, g_t1[l]);
        add(o_t4rr + l, (tv > 0.f ? tmpA[l] : 0.f) - (tv < 0.f ? tmpB[l] : 0.f));
        add(o_t4rc + l, g_t4rc[l]);
        add(o_t4cr + l, g_t4cr[l]);
        add(o_t8 + l, g_t8[l]);
        add(o_t8 + p + l, g_t8[PP + l]);
        add(o_t8 + W8 + l, g_t8[2 * PP + l]);
        add(o_t8 + W8 + p + l, g_t8[3 * PP + l]);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int e = tid + s * kCT;
        if (e < p * p) {
            const int k = e / p, l = e - k * p;
            add(o_t2rc + e, a_t2rc[s]);
            add(o_t2cr + e, a_t2cr[s]);
            add(o_t3cr + e, a_t3cr[s]);
            add(o_t6r + e, a_t6r[s]);
            add(o_t6c + e, a_t6c[s]);
            add(o_t3rr + e, g_w3p[k] * r4p[l] + g_w3n[k] * r4n[l]);
        }
    }
#pragma unroll
    for (int s = 0; s < SW; ++s) {
        const int w = tid + s * kCT;
        if (w < (PP / 4) * p) {
            const int l = w % p, kq = (w / p) * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (kq + r < p) {
                    add(o_t2rr + (kq + r) * p + l, a_t2rr[s][r]);
                    add(o_t7 + (kq + r) * p + l, a_t7[s][r]);
                }
            }
        }
    }
    if (tid == 0) atomicAdd(a.loss, loss_cta);
}

}  // namespace

size_t s2v_complete_grad_smem_bytes(int m, int p, int T) { return cgrad_layout(m, p, T).total * 4; }

cudaError_t launch_s2v_complete_grad(const S2vCGradArgs& a, int sm_count, long long smem_optin, cudaStream_t st, const char** why) {
    *why = "";
    if (a.p > 64) { *why = "classifier backward (complete): p > 64 is not supported"; return cudaErrorInvalidValue; }
    const size_t smem = s2v_complete_grad_smem_bytes(a.m, a.p, a.T);
    if ((long long)smem > smem_optin) { *why = "classifier backward (complete): embeddings do not fit in shared memory"; return cudaErrorInvalidValue; }
    const int slots = (a.p * a.p + kCT - 1) / kCT;
    auto kern = slots <= 1 ? s2v_complete_grad_kernel<1> : slots <= 2 ? s2v_complete_grad_kernel<2>
                : slots <= 4 ? s2v_complete_grad_kernel<4> : s2v_complete_grad_kernel<8>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = sm_count;
    if (grid > a.B) grid = a.B;
    kern<<<(int)grid, kCT, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ddb
