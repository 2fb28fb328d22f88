// Batched loss + gradient of the reference classifier, bipartite variant on dense instances: one CTA per LP instance,
// fp32, forward recomputed in shared memory, backward by hand, gradients summed over the batch.
//
// Replaces the inner loop of train_net (reference src/ml/train.py:59-66: zero_grad; for every instance
// loss = criterion(model(x), y); loss.backward(); optimizer.step()) with its criterion
// NLLLoss(weight=[w0, w1], size_average=False) (src/benchmark.py:70-75) for a whole batch: the summed loss and the
// accumulated gradient of every parameter, in the flat state_dict order of include/ddb200.h (4).
//
// Forward = s2v_bipartite_kernel (s2v_forward.cu; s2v.py:253-323, 218-251, quirk B9 kept).  What makes the backward
// cheap on the reference's dense random LPs (adj all ones): the messages a node receives are group means, so the
// gradient that flows back through a round is the SAME vector for every constraint node (d mean_c / m) and for every
// variable node (d mean_v / n); only the last round needs a per-node gradient (through t7).  Per round the kernel keeps
// one p-bit activity mask per node and the two input means, nothing else.
//   head:   scores_i = t8 . [relu(u6) ; relu(t7 mu_i) ; c_feats_i],  u6 = t6c mean_c + t6v mean_v
//   round:  mu_q = relu(base_q + (q < n ? t2c mean_c : t2v mean_v)),  base_q = t0 + t1 . feats_q + w3 . relu-sums_q
//   w3cp = t3c relu(t4c), w3cn = t3c relu(-t4c) (same for v): their gradients are accumulated as p-vectors and pushed
//   through t3 / t4 once per CTA.
#include "common.cuh"

namespace ddb {

struct S2vGradArgs {
    long long B;
    int m, n, p, T;
    const double* A;
    const double* b;
    const double* c;
    const float* params;
    const uint8_t* labels;     // [B, m] 0 / 1
    float w0, w1;              // class weights of the NLL loss
    float* grad;               // [param_count], accumulated with atomics (zeroed by the caller)
    double* loss;              // scalar, accumulated with atomics (zeroed by the caller)
    int* error_flag;           // set to 1 when an instance is not dense (general adjacency is not handled here)
    int* inst_flag;            // [B + 1]: inst_flag[lp] = 1 for every such instance, inst_flag[B] counts them -- the general
                               // kernel (s2v_bipartite_general_backward.cu) then processes exactly those
};

namespace {

__host__ __device__ inline int bpad4(int v) { return (v + 3) & ~3; }

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// y[k] = sum_l W[k][l] x[l], W row-major p x p in global memory; one warp per output
__device__ __forceinline__ void matvec_g(const float* __restrict__ W, int p, const float* x, float* y, int warp, int lane, int nw) {
    for (int k = warp; k < p; k += nw) {
        float acc = 0.f;
        for (int l = lane; l < p; l += 32) acc = fmaf(__ldg(W + k * p + l), x[l], acc);
        acc = wsum(acc);
        if (lane == 0) y[k] = acc;
    }
}
// y[l] = sum_k W[k][l] x[k]  (transposed product), one warp per output
__device__ __forceinline__ void matvec_gT(const float* __restrict__ W, int p, const float* x, float* y, int warp, int lane, int nw) {
    for (int l = warp; l < p; l += nw) {
        float acc = 0.f;
        for (int k = lane; k < p; k += 32) acc = fmaf(__ldg(W + k * p + l), x[k], acc);
        acc = wsum(acc);
        if (lane == 0) y[l] = acc;
    }
}

constexpr int kGradThreads = 512;   // 16 warps per CTA (one CTA per SM: the shared-memory footprint is the limit)

struct GradLayout {          // offsets in floats
    size_t t7T, t7N, mu, zr, dmu, gacc, mask, part, vecs, total;
};
__host__ __device__ inline GradLayout grad_layout(int m, int n, int p, int T, int npar) {
    const int PP = bpad4(p), NP = m + n;
    GradLayout L;
    size_t off = 0;
    L.t7T = off;  off += (size_t)p * PP;
    L.t7N = off;  off += (size_t)p * PP;
    L.mu = off;   off += (size_t)p * NP;
    L.zr = off;   off += (size_t)p * m;
    L.dmu = off;  off += (size_t)p * m;
    L.gacc = off; off += (size_t)((npar + 3) & ~3);
    off = (off + 1) & ~(size_t)1;
    L.mask = off; off += (size_t)2 * T * NP;                       // one 64-bit mask per (round, node)
    L.part = off; off += (size_t)12 * (kGradThreads / (p < 1 ? 1 : p) + 1) * PP;   // [group][12 sums][l]: register-path round reductions
    L.vecs = off;
    off += (size_t)8 * m + 5 * n + (size_t)(24 + 4 * T) * PP + 64;
    L.total = off;
    return L;
}

constexpr int AU = 16;   // columns per thread of the streaming A pass (n <= 128)

// NT = 512 (one CTA per SM, big shapes whose shared-memory footprint allows only one anyway) or 256 (two CTAs per SM)
template <int NT>
__global__ void __launch_bounds__(NT, (NT == 256) ? 2 : 1) s2v_bipartite_grad_kernel(S2vGradArgs a, int npar) {
    extern __shared__ __align__(16) float sm[];
    const int m = a.m, n = a.n, p = a.p, T = a.T, PP = bpad4(p), NP = m + n;
    const GradLayout L = grad_layout(m, n, p, T, npar);
    float* t7T = sm + L.t7T;                  // t7T[l][k] = t7[k][l]
    float* t7N = sm + L.t7N;                  // t7N[k][l] = t7[k][l]
    float* mu = sm + L.mu;                    // [p][NP] embeddings of the current round
    float* zr = sm + L.zr;                    // [p][m] relu(t7 mu_i), later d z
    float* dmu = sm + L.dmu;                  // [p][m] gradient wrt the last round's constraint embeddings
    float* gacc = sm + L.gacc;                // per-CTA gradient accumulator, flat parameter order
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(sm + L.mask);   // [T][NP]
    float* v = sm + L.vecs;
    float* rb = v;     v += m;    float* cosv = v;  v += m;    float* Sp = v;   v += m;    float* Sn = v;  v += m;
    float* ds0 = v;    v += m;    float* ds1 = v;   v += m;    float* lossn = v; v += m;   float* spare = v; v += m;
    float* cj = v;     v += n;    float* Cp = v;    v += n;    float* Cn = v;   v += n;    float* ccnt = v; v += 2 * n;
    float* w3cp = v;   v += PP;   float* w3cn = v;  v += PP;   float* w3vp = v; v += PP;   float* w3vn = v; v += PP;
    float* meanc = v;  v += PP;   float* meanv = v; v += PP;   float* yv = v;   v += PP;   float* yc = v;   v += PP;
    float* u6pre = v;  v += PP;   float* u6r = v;   v += PP;   float* du6 = v;  v += PP;   float* tmp1 = v; v += PP;
    float* tmp2 = v;   v += PP;   float* dmc = v;   v += PP;   float* dmv = v;  v += PP;   float* da = v;   v += PP;
    float* db2 = v;    v += PP;   float* r4 = v;    v += 4 * PP;
    float* gw3 = v;    v += 4 * PP;            // per-CTA accumulators of d w3cp, d w3cn, d w3vp, d w3vn
    float* mcs = v;    v += (size_t)T * PP;    // input mean_c of every round
    float* mvs = v;    v += (size_t)T * PP;    // input mean_v of every round
    float* yvs = v;    v += (size_t)T * PP;    // t2c . mean_c added in every round (node positions < n)
    float* ycs = v;    v += (size_t)T * PP;    // t2v . mean_v added in every round (node positions >= n)
    float* part = sm + L.part;
    int* iflag = reinterpret_cast<int*>(v);
    (void)spare;

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
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

    for (int e = tid; e < p * PP; e += nt) {
        const int r = e / PP, q = e - r * PP;
        t7T[e] = (q < p) ? __ldg(t7 + q * p + r) : 0.f;
        t7N[e] = (q < p) ? __ldg(t7 + r * p + q) : 0.f;
    }
    for (int e = tid; e < npar; e += nt) gacc[e] = 0.f;
    for (int e = tid; e < 4 * PP; e += nt) gw3[e] = 0.f;
    for (int l = tid; l < p; l += nt) {
        r4[l] = fmaxf(__ldg(t4c + l), 0.f);
        r4[PP + l] = fmaxf(-__ldg(t4c + l), 0.f);
        r4[2 * PP + l] = fmaxf(__ldg(t4v + l), 0.f);
        r4[3 * PP + l] = fmaxf(-__ldg(t4v + l), 0.f);
    }
    __syncthreads();
    matvec_g(t3c, p, r4, w3cp, warp, lane, nw);
    matvec_g(t3c, p, r4 + PP, w3cn, warp, lane, nw);
    matvec_g(t3v, p, r4 + 2 * PP, w3vp, warp, lane, nw);
    matvec_g(t3v, p, r4 + 3 * PP, w3vn, warp, lane, nw);
    __syncthreads();
    double loss_cta = 0.0;
    // register path of the rounds: thread (l, group); RC / RV node slots per thread
    constexpr int RC = 20, RV = 12;
    const int G = nt / p;
    const int rl = tid % p, rgrp = tid / p;
    const bool ractive = rgrp < G;
    const bool regfit = (m <= RC * G) && (n <= RV * G) && T >= 1;      // uniform
    const float kc0 = __ldg(t0 + rl) + __ldg(t1c + 4 * rl), kc1 = __ldg(t1c + 4 * rl + 1), kc3 = __ldg(t1c + 4 * rl + 3);
    const float kcp = w3cp[rl], kcn = w3cn[rl];
    const float kv0 = __ldg(t0 + rl), kv1 = __ldg(t1v + rl), kvp = w3vp[rl], kvn = w3vn[rl];

    for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg = a.c + (size_t)lp * n;
        const uint8_t* yl = a.labels + (size_t)lp * m;
        for (int j = tid; j < n; j += nt) {
            cj[j] = (float)cg[j];
            Cp[j] = 0.f; Cn[j] = 0.f; ccnt[j] = 0.f;
        }
        if (tid == 0) *iflag = 0;
        __syncthreads();
        // ---- pass over A: row normalisation, cosines, relu row / column sums ---------------------------------------------------
        if (n <= 8 * AU) {
            // streaming layout of s2v_bipartite_dense.cu: 32-row chunks, thread = (row of the chunk, column phase s, half g)
            // owning columns g * SL + s + 4u; the 8 threads of a row are lanes of one warp (xor 1, 2, 16), the column sums
            // accumulate in registers over all rows; the next chunk's loads are in flight while this one is reduced
            const int SL = ((n + 1) / 2 + 3) & ~3;
            const int as = lane & 3, ar = (lane >> 2) & 3, ag = lane >> 4;
            const int arow = warp * 4 + ar, col0 = ag * SL + as;
            float cmk[AU], ccol[AU], accp[AU], accs[AU];
            int npad = 0;
#pragma unroll
            for (int u = 0; u < AU; ++u) {
                const bool ok = (4 * u + as < SL) && (col0 + 4 * u < n);
                cmk[u] = ok ? 1.f : 0.f;
                npad += ok ? 0 : 1;
                ccol[u] = ok ? cj[col0 + 4 * u] : 0.f;
                accp[u] = 0.f;
                accs[u] = 0.f;
            }
            const int crow = 4 * nw;                  // rows per chunk: four per warp
            const int nchunk = (m + crow - 1) / crow;
            double xn[AU];
            auto load_chunk = [&](int ci) {
                const int i = ci * crow + arow;
                const double* rp = Ag + (size_t)(i < m ? i : 0) * n;
#pragma unroll
                for (int u = 0; u < AU; ++u) xn[u] = __ldg(rp + ((cmk[u] != 0.f) ? col0 + 4 * u : 0));
            };
            load_chunk(0);
            int sparse = 0;
            for (int ci = 0; ci < nchunk; ++ci) {
                float x[AU];
#pragma unroll
                for (int u = 0; u < AU; ++u) x[u] = (float)xn[u] * cmk[u];
                if (ci + 1 < nchunk) load_chunk(ci + 1);
                const int i = ci * crow + arow;
                const bool rowok = i < m;
                float ss = 0.f, cs = 0.f, sp = 0.f, sx = 0.f;
                int nzero = 0;
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    const float v = x[u];
                    ss = fmaf(v, v, ss);
                    cs = fmaf(v, ccol[u], cs);
                    sp += fmaxf(v, 0.f);
                    sx += v;
                    nzero += (v == 0.f) ? 1 : 0;
                }
                sparse |= (rowok && nzero != npad);
#pragma unroll
                for (int off = 1; off <= 16; off = (off == 2) ? 16 : off * 2) {
                    ss += __shfl_xor_sync(0xffffffffu, ss, off);
                    cs += __shfl_xor_sync(0xffffffffu, cs, off);
                    sp += __shfl_xor_sync(0xffffffffu, sp, off);
                    sx += __shfl_xor_sync(0xffffffffu, sx, off);
                }
                const float bi = (float)bg[rowok ? i : 0];
                ss = fmaf(bi, bi, ss);
                const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
                const float invm = rowok ? inv : 0.f;
#pragma unroll
                for (int u = 0; u < AU; ++u) {
                    const float xs = x[u] * invm;
                    accp[u] += fmaxf(xs, 0.f);
                    accs[u] += xs;
                }
                if (rowok && as == 0 && ag == 0) {
                    rb[i] = bi * inv; cosv[i] = cs * inv; Sp[i] = sp * inv; Sn[i] = (sp - sx) * inv;
                }
            }
            if (sparse) *iflag = 1;
            // column sums: over the 4 rows of a warp by shuffles, over the 8 warps by shared-memory atomics
#pragma unroll
            for (int u = 0; u < AU; ++u) {
                float cp = accp[u], cn = accp[u] - accs[u];
                cp += __shfl_xor_sync(0xffffffffu, cp, 4); cn += __shfl_xor_sync(0xffffffffu, cn, 4);
                cp += __shfl_xor_sync(0xffffffffu, cp, 8); cn += __shfl_xor_sync(0xffffffffu, cn, 8);
                if (ar == 0 && cmk[u] != 0.f) {
                    atomicAdd(&Cp[col0 + 4 * u], cp);
                    atomicAdd(&Cn[col0 + 4 * u], cn);
                }
            }
        } else {
            for (int i = warp; i < m; i += nw) {
                const float bi = (float)bg[i];
                float ss = 0.f;
                for (int j = lane; j < n; j += 32) {
                    const float x = (float)Ag[(size_t)i * n + j];
                    ss = fmaf(x, x, ss);
                }
                ss = wsum(ss) + bi * bi;
                const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
                float cs = 0.f, sp = 0.f, sn = 0.f, cnt = 0.f;
                for (int j = lane; j < n; j += 32) {
                    const float x = (float)Ag[(size_t)i * n + j] * inv;
                    cs = fmaf(x, cj[j], cs);
                    sp += fmaxf(x, 0.f);
                    sn += fmaxf(-x, 0.f);
                    cnt += (x != 0.f) ? 1.f : 0.f;
                    atomicAdd(&Cp[j], fmaxf(x, 0.f));
                    atomicAdd(&Cn[j], fmaxf(-x, 0.f));
                }
                cs = wsum(cs); sp = wsum(sp); sn = wsum(sn); cnt = wsum(cnt);
                if (lane == 0) {
                    rb[i] = bi * inv; cosv[i] = cs; Sp[i] = sp; Sn[i] = sn;
                    if (cnt != (float)n) *iflag = 1;
                }
            }
        }
        if (!regfit)
            for (int e = tid; e < p * NP; e += nt) mu[e] = 0.f;
        __syncthreads();
        if (*iflag) {                     // sparse instance: not handled by this kernel (uniform decision)
            if (tid == 0) {
                *a.error_flag = 1;
                a.inst_flag[lp] = 1;
                atomicAdd(a.inst_flag + a.B, 1);
            }
            __syncthreads();
            continue;
        }

        // ---- forward rounds ------------------------------------------------------------------------------------------------------
        // Register path (shapes whose per-thread share of the nodes fits): thread (l, group) keeps base_l(q) of its nodes in
        // registers -- five FMAs on the node statistics, the same for every round -- so a round is add / relu / sum per node,
        // and the backward recomputes the activity of a (node, l) pair from base + y_t instead of storing masks.
        float bc[RC], bv[RV];
        int qc = 0, qv = 0;
        if (regfit) {
            const float ninf = __int_as_float(0xff800000);
#pragma unroll
            for (int q = 0; q < RC; ++q) {
                const int i = rgrp + q * G;
                float val = ninf;
                if (ractive && i < m) {
                    val = kc0;
                    val = fmaf(kc1, rb[i], val); val = fmaf(kc3, cosv[i], val);
                    val = fmaf(kcp, Sp[i], val); val = fmaf(kcn, Sn[i], val);
                }
                bc[q] = val;
            }
#pragma unroll
            for (int q = 0; q < RV; ++q) {
                const int j = rgrp + q * G;
                float val = ninf;
                if (ractive && j < n) {
                    val = kv0;
                    val = fmaf(kv1, cj[j], val); val = fmaf(kvp, Cp[j], val); val = fmaf(kvn, Cn[j], val);
                }
                bv[q] = val;
            }
            // node positions < n get yv, positions >= n get yc (quirk B9): my first qc constraint / qv variable nodes
            qc = (n > rgrp) ? (n - rgrp + G - 1) / G : 0;
            qv = (n - m > rgrp) ? (n - m - rgrp + G - 1) / G : 0;
            for (int l = tid; l < p; l += nt) { meanc[l] = 0.f; meanv[l] = 0.f; }
            __syncthreads();
            for (int t = 0; t < T; ++t) {
                const bool last = (t == T - 1);
                // input means of this round and what they add: y = t2 . mean (one thread per output)
                if (tid < 2 * p) {
                    const int kk = (tid < p) ? tid : tid - p;
                    const float* Wr = ((tid < p) ? t2c : t2v) + kk * p;
                    const float* xin = (tid < p) ? meanc : meanv;
                    float a0 = 0.f, a1 = 0.f;
                    int q = 0;
                    for (; q + 1 < p; q += 2) { a0 = fmaf(__ldg(Wr + q), xin[q], a0); a1 = fmaf(__ldg(Wr + q + 1), xin[q + 1], a1); }
                    if (q < p) a0 = fmaf(__ldg(Wr + q), xin[q], a0);
                    ((tid < p) ? yvs : ycs)[t * PP + kk] = a0 + a1;
                    ((tid < p) ? mcs : mvs)[t * PP + kk] = xin[kk];
                }
                __syncthreads();
                float sc = 0.f, sv = 0.f;
                if (ractive) {
                    const float ya = yvs[t * PP + rl], yb = ycs[t * PP + rl];
#pragma unroll
                    for (int q = 0; q < RC; ++q) {
                        const float val = fmaxf(bc[q] + (q < qc ? ya : yb), 0.f);
                        sc += val;
                        if (last) {
                            const int i = rgrp + q * G;
                            if (i < m) mu[rl * NP + i] = val;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < RV; ++q) sv += fmaxf(bv[q] + (q < qv ? ya : yb), 0.f);
                    part[(rgrp * 12 + 0) * PP + rl] = sc;
                    part[(rgrp * 12 + 1) * PP + rl] = sv;
                }
                __syncthreads();
                if (tid < p) {
                    float tc = 0.f, tv = 0.f;
                    for (int g2 = 0; g2 < G; ++g2) { tc += part[(g2 * 12 + 0) * PP + tid]; tv += part[(g2 * 12 + 1) * PP + tid]; }
                    meanc[tid] = tc / (float)m;          // means of this round's output: input of the next round / of the head
                    meanv[tid] = tv / (float)n;
                }
                __syncthreads();
            }
        } else {
            for (int t = 0; t < T; ++t) {
                for (int l = warp; l < p; l += nw) {
                    float sc = 0.f, svv = 0.f;
                    for (int i = lane; i < m; i += 32) sc += mu[l * NP + i];
                    for (int j = lane; j < n; j += 32) svv += mu[l * NP + m + j];
                    sc = wsum(sc); svv = wsum(svv);
                    if (lane == 0) {
                        meanc[l] = sc / (float)m; meanv[l] = svv / (float)n;
                        mcs[t * PP + l] = meanc[l]; mvs[t * PP + l] = meanv[l];
                    }
                }
                __syncthreads();
                matvec_g(t2c, p, meanc, yv, warp, lane, nw);
                matvec_g(t2v, p, meanv, yc, warp, lane, nw);
                __syncthreads();
                for (int q = tid; q < NP; q += nt) {
                    unsigned long long bits = 0ull;
                    for (int l = 0; l < p; ++l) {
                        float val = __ldg(t0 + l);
                        if (q < m) {
                            val += __ldg(t1c + 4 * l) + __ldg(t1c + 4 * l + 1) * rb[q] + __ldg(t1c + 4 * l + 3) * cosv[q];
                            val += w3cp[l] * Sp[q] + w3cn[l] * Sn[q];
                        } else {
                            const int j = q - m;
                            val += __ldg(t1v + l) * cj[j] + w3vp[l] * Cp[j] + w3vn[l] * Cn[j];
                        }
                        val += (q < n) ? yv[l] : yc[l];
                        if (val > 0.f) bits |= 1ull << l;
                        mu[l * NP + q] = fmaxf(val, 0.f);
                    }
                    mask[(size_t)t * NP + q] = bits;
                }
                __syncthreads();
            }

        }

        // ---- head forward: scores, loss, d scores ---------------------------------------------------------------------------
        if (!regfit) {
            for (int l = warp; l < p; l += nw) {
                float sc = 0.f, svv = 0.f;
                for (int i = lane; i < m; i += 32) sc += mu[l * NP + i];
                for (int j = lane; j < n; j += 32) svv += mu[l * NP + m + j];
                sc = wsum(sc); svv = wsum(svv);
                if (lane == 0) { meanc[l] = sc / (float)m; meanv[l] = svv / (float)n; }
            }
            __syncthreads();
        }
        matvec_g(t6c, p, meanc, tmp1, warp, lane, nw);
        matvec_g(t6v, p, meanv, tmp2, warp, lane, nw);
        __syncthreads();
        for (int l = tid; l < p; l += nt) {
            u6pre[l] = tmp1[l] + tmp2[l];
            u6r[l] = fmaxf(u6pre[l], 0.f);
        }
        __syncthreads();
        const bool tiled = ((m & 3) == 0) && ((NP & 3) == 0);     // float4 rows of mu / zr / dmu (uniform)
        if (tiled) {
            // z = relu(t7 mu_c): register tile, thread = 4 nodes x 4 outputs, two LDS.128 per 16 FMAs
            const int NG = m / 4, KG = PP / 4;
            for (int w = tid; w < NG * KG; w += nt) {
                const int ng = w % NG, kg = w / NG;
                float acc[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
                for (int l = 0; l < p; ++l) {
                    const float4 mv = *reinterpret_cast<const float4*>(mu + l * NP + 4 * ng);
                    const float4 wv = *reinterpret_cast<const float4*>(t7T + l * PP + 4 * kg);
                    const float mq[4] = {mv.x, mv.y, mv.z, mv.w};
                    const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(wr[r], mq[q], acc[q][r]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (4 * kg + r < p)
                        *reinterpret_cast<float4*>(zr + (4 * kg + r) * m + 4 * ng) =
                            make_float4(fmaxf(acc[0][r], 0.f), fmaxf(acc[1][r], 0.f), fmaxf(acc[2][r], 0.f), fmaxf(acc[3][r], 0.f));
            }
            __syncthreads();
        }
        for (int i = tid; i < m; i += nt) {
            float s0 = 0.f, s1 = 0.f;
            for (int l = 0; l < p; ++l) {
                s0 = fmaf(__ldg(t8 + l), u6r[l], s0);
                s1 = fmaf(__ldg(t8 + W8 + l), u6r[l], s1);
            }
            if (tiled) {
                for (int k = 0; k < p; ++k) {
                    const float z = zr[k * m + i];
                    s0 = fmaf(__ldg(t8 + p + k), z, s0);
                    s1 = fmaf(__ldg(t8 + W8 + p + k), z, s1);
                }
            } else {
                for (int kb = 0; kb < PP; kb += 4) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    for (int l = 0; l < p; ++l) {
                        const float x = mu[l * NP + i];
                        const float4 w = *reinterpret_cast<const float4*>(t7T + l * PP + kb);
                        a0 = fmaf(w.x, x, a0); a1 = fmaf(w.y, x, a1); a2 = fmaf(w.z, x, a2); a3 = fmaf(w.w, x, a3);
                    }
                    const float r[4] = {fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f)};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (kb + u < p) {
                            zr[(kb + u) * m + i] = r[u];
                            s0 = fmaf(__ldg(t8 + p + kb + u), r[u], s0);
                            s1 = fmaf(__ldg(t8 + W8 + p + kb + u), r[u], s1);
                        }
                    }
                }
            }
            const float f1 = rb[i], f3 = cosv[i];
            s0 += __ldg(t8 + 2 * p) + __ldg(t8 + 2 * p + 1) * f1 + __ldg(t8 + 2 * p + 3) * f3;
            s1 += __ldg(t8 + W8 + 2 * p) + __ldg(t8 + W8 + 2 * p + 1) * f1 + __ldg(t8 + W8 + 2 * p + 3) * f3;
            const float mx = fmaxf(s0, s1);
            const float lse = mx + logf(expf(s0 - mx) + expf(s1 - mx));
            const float p0 = expf(s0 - lse), p1 = expf(s1 - lse);
            const int y = yl[i] ? 1 : 0;
            const float w = y ? a.w1 : a.w0;
            lossn[i] = -w * (y ? (s1 - lse) : (s0 - lse));
            ds0[i] = w * (p0 - (y == 0 ? 1.f : 0.f));
            ds1[i] = w * (p1 - (y == 1 ? 1.f : 0.f));
        }
        __syncthreads();

        // ---- head backward ------------------------------------------------------------------------------------------------------
        // S_c = sum_i ds_c[i];  loss
        if (warp == 0) {
            float s0 = 0.f, s1 = 0.f, ls = 0.f;
            for (int i = lane; i < m; i += 32) { s0 += ds0[i]; s1 += ds1[i]; ls += lossn[i]; }
            s0 = wsum(s0); s1 = wsum(s1); ls = wsum(ls);
            if (lane == 0) { tmp1[0] = s0; tmp1[1] = s1; loss_cta += (double)ls; }
        }
        __syncthreads();
        {
            const float S0 = tmp1[0], S1 = tmp1[1];
            // d t8: u6 block, c_feats block
            for (int l = tid; l < p; l += nt) {
                gacc[o_t8 + l] += S0 * u6r[l];
                gacc[o_t8 + W8 + l] += S1 * u6r[l];
                const float g = __ldg(t8 + l) * S0 + __ldg(t8 + W8 + l) * S1;
                du6[l] = (u6pre[l] > 0.f) ? g : 0.f;
            }
            if (warp == 1) {
                float f1a = 0.f, f1b = 0.f, f3a = 0.f, f3b = 0.f;
                for (int i = lane; i < m; i += 32) {
                    f1a += ds0[i] * rb[i]; f1b += ds1[i] * rb[i];
                    f3a += ds0[i] * cosv[i]; f3b += ds1[i] * cosv[i];
                }
                f1a = wsum(f1a); f1b = wsum(f1b); f3a = wsum(f3a); f3b = wsum(f3b);
                if (lane == 0) {
                    gacc[o_t8 + 2 * p] += S0;          gacc[o_t8 + W8 + 2 * p] += S1;
                    gacc[o_t8 + 2 * p + 1] += f1a;     gacc[o_t8 + W8 + 2 * p + 1] += f1b;
                    gacc[o_t8 + 2 * p + 3] += f3a;     gacc[o_t8 + W8 + 2 * p + 3] += f3b;
                }
            }
        }
        // d t8 (z block) and d z in place of zr
        for (int k = warp; k < p; k += nw) {
            float g0 = 0.f, g1 = 0.f;
            const float w0k = __ldg(t8 + p + k), w1k = __ldg(t8 + W8 + p + k);
            for (int i = lane; i < m; i += 32) {
                const float z = zr[k * m + i];
                g0 = fmaf(ds0[i], z, g0);
                g1 = fmaf(ds1[i], z, g1);
                zr[k * m + i] = (z > 0.f) ? (w0k * ds0[i] + w1k * ds1[i]) : 0.f;
            }
            g0 = wsum(g0); g1 = wsum(g1);
            if (lane == 0) { gacc[o_t8 + p + k] += g0; gacc[o_t8 + W8 + p + k] += g1; }
        }
        __syncthreads();
        // d t6c, d t6v;  t6c^T du6, t6v^T du6
        for (int e = tid; e < p * p; e += nt) {
            const int k = e / p, l = e - k * p;
            gacc[o_t6c + e] += du6[k] * meanc[l];
            gacc[o_t6v + e] += du6[k] * meanv[l];
        }
        matvec_gT(t6c, p, du6, tmp1, warp, lane, nw);     // tmp1 = t6c^T du6
        matvec_gT(t6v, p, du6, tmp2, warp, lane, nw);     // tmp2 = t6v^T du6
        // d t7[k][l] = sum_i dz[k][i] mu[l][i]
        if (tiled) {
            // thread = (4 outputs k, one l): lanes run over l (mu rows, conflict-free LDS.128), the dz rows are broadcast
            for (int w = tid; w < (PP / 4) * p; w += nt) {
                const int l = w % p, k0 = (w / p) * 4;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < m; i += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(mu + l * NP + i);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (k0 + r < p) {
                            const float4 z = *reinterpret_cast<const float4*>(zr + (k0 + r) * m + i);
                            acc[r] = fmaf(z.x, x.x, fmaf(z.y, x.y, fmaf(z.z, x.z, fmaf(z.w, x.w, acc[r]))));
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (k0 + r < p) gacc[o_t7 + (k0 + r) * p + l] += acc[r];
            }
        } else {
            for (int e = tid; e < p * p; e += nt) {
                const int k = e / p, l = e - k * p;
                float acc = 0.f;
                for (int i = 0; i < m; ++i) acc = fmaf(zr[k * m + i], mu[l * NP + i], acc);
                gacc[o_t7 + e] += acc;
            }
        }
        __syncthreads();
        // d mu_c (last round) = t7^T dz + t6c^T du6 / m
        if (tiled) {
            // register tile: thread = 4 nodes x 4 coordinates l, two LDS.128 per 16 FMAs
            const int NG = m / 4, LG = PP / 4;
            for (int w = tid; w < NG * LG; w += nt) {
                const int ng = w % NG, lg = w / NG;
                float acc[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
                for (int k = 0; k < p; ++k) {
                    const float4 zv = *reinterpret_cast<const float4*>(zr + k * m + 4 * ng);
                    const float4 wv = *reinterpret_cast<const float4*>(t7N + k * PP + 4 * lg);
                    const float zq[4] = {zv.x, zv.y, zv.z, zv.w};
                    const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(wr[r], zq[q], acc[q][r]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (4 * lg + r < p) {
                        const float add = tmp1[4 * lg + r] / (float)m;
                        *reinterpret_cast<float4*>(dmu + (4 * lg + r) * m + 4 * ng) =
                            make_float4(acc[0][r] + add, acc[1][r] + add, acc[2][r] + add, acc[3][r] + add);
                    }
                }
            }
        } else {
            for (int i = tid; i < m; i += nt) {
                for (int lb = 0; lb < PP; lb += 4) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    for (int k = 0; k < p; ++k) {
                        const float x = zr[k * m + i];
                        const float4 w = *reinterpret_cast<const float4*>(t7N + k * PP + lb);
                        a0 = fmaf(w.x, x, a0); a1 = fmaf(w.y, x, a1); a2 = fmaf(w.z, x, a2); a3 = fmaf(w.w, x, a3);
                    }
                    const float r[4] = {a0, a1, a2, a3};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (lb + u < p) dmu[(lb + u) * m + i] = r[u] + tmp1[lb + u] / (float)m;
                }
            }
        }
        for (int l = tid; l < p; l += nt) dmv[l] = tmp2[l] / (float)n;      // same for every variable node
        __syncthreads();

        // ---- rounds backward ---------------------------------------------------------------------------------------------------
        if (regfit) {
            for (int t = T - 1; t >= 0; --t) {
                const bool last = (t == T - 1);
                if (ractive) {
                    const float ya = yvs[t * PP + rl], yb = ycs[t * PP + rl];
                    const float gc = last ? 0.f : dmc[rl], gv = dmv[rl];
                    float s_c = 0.f, s_rb = 0.f, s_cos = 0.f, s_sp = 0.f, s_sn = 0.f, s_a = 0.f, s_b = 0.f;
                    float s_v = 0.f, s_cj = 0.f, s_cp = 0.f, s_cn = 0.f;
#pragma unroll
                    for (int q = 0; q < RC; ++q) {
                        const int i = rgrp + q * G;
                        const bool on = (bc[q] + (q < qc ? ya : yb)) > 0.f;       // -inf base of padding slots: never active
                        float d = 0.f;
                        if (on) d = last ? dmu[rl * m + i] : gc;
                        const int ii = (i < m) ? i : 0;
                        s_c += d;
                        s_rb = fmaf(d, rb[ii], s_rb); s_cos = fmaf(d, cosv[ii], s_cos);
                        s_sp = fmaf(d, Sp[ii], s_sp); s_sn = fmaf(d, Sn[ii], s_sn);
                        if (q < qc) s_a += d; else s_b += d;
                    }
#pragma unroll
                    for (int q = 0; q < RV; ++q) {
                        const int j = rgrp + q * G;
                        const bool on = (bv[q] + (q < qv ? ya : yb)) > 0.f;
                        const float d = on ? gv : 0.f;
                        const int jj = (j < n) ? j : 0;
                        s_v += d;
                        s_cj = fmaf(d, cj[jj], s_cj); s_cp = fmaf(d, Cp[jj], s_cp); s_cn = fmaf(d, Cn[jj], s_cn);
                        if (q < qv) s_a += d; else s_b += d;
                    }
                    float* pp = part + (size_t)rgrp * 12 * PP + rl;
                    pp[0 * PP] = s_c; pp[1 * PP] = s_rb; pp[2 * PP] = s_cos; pp[3 * PP] = s_sp; pp[4 * PP] = s_sn;
                    pp[5 * PP] = s_a; pp[6 * PP] = s_b; pp[7 * PP] = s_v; pp[8 * PP] = s_cj; pp[9 * PP] = s_cp; pp[10 * PP] = s_cn;
                }
                __syncthreads();
                if (tid < p) {
                    float r[11];
#pragma unroll
                    for (int q = 0; q < 11; ++q) r[q] = 0.f;
                    for (int g2 = 0; g2 < G; ++g2)
#pragma unroll
                        for (int q = 0; q < 11; ++q) r[q] += part[((size_t)g2 * 12 + q) * PP + tid];
                    const int l = tid;
                    gacc[o_t0 + l] += r[0] + r[7];
                    gacc[o_t1c + 4 * l] += r[0];
                    gacc[o_t1c + 4 * l + 1] += r[1];
                    gacc[o_t1c + 4 * l + 3] += r[2];
                    gacc[o_t1v + l] += r[8];
                    gw3[l] += r[3]; gw3[PP + l] += r[4]; gw3[2 * PP + l] += r[9]; gw3[3 * PP + l] += r[10];
                    da[l] = r[5]; db2[l] = r[6];
                }
                __syncthreads();
                // d t2c += da (x) mean_c(t), d t2v += db2 (x) mean_v(t); gradient of the input means (one thread per output)
                for (int e = tid; e < p * p; e += nt) {
                    const int k = e / p, l = e - k * p;
                    gacc[o_t2c + e] += da[k] * mcs[t * PP + l];
                    gacc[o_t2v + e] += db2[k] * mvs[t * PP + l];
                }
                if (tid < 2 * p) {
                    const int l = (tid < p) ? tid : tid - p;
                    const float* W = (tid < p) ? t2c : t2v;
                    const float* xin = (tid < p) ? da : db2;
                    float a0 = 0.f, a1 = 0.f;
                    int k = 0;
                    for (; k + 1 < p; k += 2) { a0 = fmaf(__ldg(W + k * p + l), xin[k], a0); a1 = fmaf(__ldg(W + (k + 1) * p + l), xin[k + 1], a1); }
                    if (k < p) a0 = fmaf(__ldg(W + k * p + l), xin[k], a0);
                    ((tid < p) ? tmp1 : tmp2)[l] = a0 + a1;
                }
                __syncthreads();
                for (int l = tid; l < p; l += nt) {
                    dmc[l] = tmp1[l] / (float)m;
                    dmv[l] = tmp2[l] / (float)n;
                }
                __syncthreads();
            }
        } else {
            for (int t = T - 1; t >= 0; --t) {
                const bool last = (t == T - 1);
                // one warp per embedding coordinate l: the linear reductions of d pre[l][:] over the nodes
                for (int l = warp; l < p; l += nw) {
                    float s_all = 0.f, s_t1c1 = 0.f, s_t1c3 = 0.f, s_c = 0.f, s_sp = 0.f, s_sn = 0.f;
                    float s_t1v = 0.f, s_cp = 0.f, s_cn = 0.f, s_a = 0.f, s_b = 0.f;
                    const float gc = last ? 0.f : dmc[l], gv = dmv[l];
                    for (int q = lane; q < NP; q += 32) {
                        const bool on = (mask[(size_t)t * NP + q] >> l) & 1ull;
                        float d = 0.f;
                        if (on) d = (q < m) ? (last ? dmu[l * m + q] : gc) : gv;
                        s_all += d;
                        if (q < m) {
                            s_c += d; s_t1c1 = fmaf(d, rb[q], s_t1c1); s_t1c3 = fmaf(d, cosv[q], s_t1c3);
                            s_sp = fmaf(d, Sp[q], s_sp); s_sn = fmaf(d, Sn[q], s_sn);
                        } else {
                            const int j = q - m;
                            s_t1v = fmaf(d, cj[j], s_t1v); s_cp = fmaf(d, Cp[j], s_cp); s_cn = fmaf(d, Cn[j], s_cn);
                        }
                        if (q < n) s_a += d; else s_b += d;
                    }
                    s_all = wsum(s_all); s_c = wsum(s_c); s_t1c1 = wsum(s_t1c1); s_t1c3 = wsum(s_t1c3);
                    s_sp = wsum(s_sp); s_sn = wsum(s_sn); s_t1v = wsum(s_t1v); s_cp = wsum(s_cp); s_cn = wsum(s_cn);
                    s_a = wsum(s_a); s_b = wsum(s_b);
                    if (lane == 0) {
                        gacc[o_t0 + l] += s_all;
                        gacc[o_t1c + 4 * l] += s_c;
                        gacc[o_t1c + 4 * l + 1] += s_t1c1;
                        gacc[o_t1c + 4 * l + 3] += s_t1c3;
                        gacc[o_t1v + l] += s_t1v;
                        gw3[l] += s_sp; gw3[PP + l] += s_sn; gw3[2 * PP + l] += s_cp; gw3[3 * PP + l] += s_cn;
                        da[l] = s_a; db2[l] = s_b;
                    }
                }
                __syncthreads();
                // d t2c += da (x) mean_c(t), d t2v += db2 (x) mean_v(t); gradient of the input means
                for (int e = tid; e < p * p; e += nt) {
                    const int k = e / p, l = e - k * p;
                    gacc[o_t2c + e] += da[k] * mcs[t * PP + l];
                    gacc[o_t2v + e] += db2[k] * mvs[t * PP + l];
                }
                matvec_gT(t2c, p, da, tmp1, warp, lane, nw);
                matvec_gT(t2v, p, db2, tmp2, warp, lane, nw);
                __syncthreads();
                for (int l = tid; l < p; l += nt) {
                    dmc[l] = tmp1[l] / (float)m;
                    dmv[l] = tmp2[l] / (float)n;
                }
                __syncthreads();
            }
        }
    }

    // ---- per CTA: push d w3 through t3 / t4, then add the accumulator to the global gradient --------------------------------
    __syncthreads();
    for (int e = tid; e < p * p; e += nt) {
        const int k = e / p, l = e - k * p;
        gacc[o_t3c + e] += gw3[k] * r4[l] + gw3[PP + k] * r4[PP + l];
        gacc[o_t3v + e] += gw3[2 * PP + k] * r4[2 * PP + l] + gw3[3 * PP + k] * r4[3 * PP + l];
    }
    matvec_gT(t3c, p, gw3, tmp1, warp, lane, nw);
    matvec_gT(t3c, p, gw3 + PP, tmp2, warp, lane, nw);
    __syncthreads();
    for (int l = tid; l < p; l += nt) {
        const float tv = __ldg(t4c + l);
        gacc[o_t4c + l] += (tv > 0.f ? tmp1[l] : 0.f) - (tv < 0.f ? tmp2[l] : 0.f);
    }
    __syncthreads();
    matvec_gT(t3v, p, gw3 + 2 * PP, tmp1, warp, lane, nw);
    matvec_gT(t3v, p, gw3 + 3 * PP, tmp2, warp, lane, nw);
    __syncthreads();
    for (int l = tid; l < p; l += nt) {
        const float tv = __ldg(t4v + l);
        gacc[o_t4v + l] += (tv > 0.f ? tmp1[l] : 0.f) - (tv < 0.f ? tmp2[l] : 0.f);
    }
    __syncthreads();
    for (int e = tid; e < npar; e += nt) {
        const float g = gacc[e];
        if (g != 0.f) atomicAdd(a.grad + e, g);
    }
    if (tid == 0) atomicAdd(a.loss, loss_cta);
}

}  // namespace

size_t s2v_grad_smem_bytes(int m, int n, int p, int T, int npar) { return grad_layout(m, n, p, T, npar).total * 4; }

cudaError_t launch_s2v_bipartite_grad(const S2vGradArgs& a, int npar, int sm_count, long long smem_optin, cudaStream_t st,
                                      const char** why) {
    *why = "";
    if (a.p > 64) { *why = "classifier backward: p > 64 is not supported (one 64-bit activity mask per node)"; return cudaErrorInvalidValue; }
    const size_t smem = s2v_grad_smem_bytes(a.m, a.n, a.p, a.T, npar);
    if ((long long)smem > smem_optin) { *why = "classifier backward: embeddings do not fit in shared memory"; return cudaErrorInvalidValue; }
    // two 256-thread CTAs per SM when two footprints fit in shared memory, else one 512-thread CTA
    const bool two = 2 * (smem + 1024) <= (size_t)smem_optin;
    auto kern = two ? s2v_bipartite_grad_kernel<256> : s2v_bipartite_grad_kernel<512>;
    const int nthreads = two ? 256 : 512;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem);
    if (e != cudaSuccess) return e;
    long long grid = (long long)sm_count * (per_sm > 0 ? per_sm : 1);
    if (grid > a.B) grid = a.B;
    kern<<<(int)grid, nthreads, smem, st>>>(a, npar);
    return cudaGetLastError();
}

}  // namespace ddb
