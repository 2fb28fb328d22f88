// Batched forward pass of the reference classifier (structure2vec), one CTA per LP instance, fp32.
//
// Replaces Model.forward for a whole batch (reference src/ml/models/s2v.py:45-54): _forward_bipartite + _s2v_bipartite
// (:253-323, 218-251) and _forward_complete + _s2v_complete (:124-187, 91-122), plus the per-instance Python batch
// loop of ml/utils.py:3-25.  Inputs are the fp64 (A, b, c) the solver path produces; in_loss is every row (A1).
//
// What makes it cheap (SURVEY.md section 7 "classifier is thin"):
//   * the bmm -> relu -> sum terms (s2v.py:112, 236, 239) are linear in relu(+t4), relu(-t4):
//         sum_j relu(t4_k w_j) = relu(t4_k) sum_j relu(w_j) + relu(-t4_k) sum_j relu(-w_j),
//     so only two row sums / column sums of the normalised matrix are needed, and t3 . relu(+-t4) are p-vectors
//     computed once per launch;
//   * bipartite, dense instance: adj is all ones, so mu . normalize(adj) is a broadcast mean -> a round costs
//     O(p^2 + p (m+n)); the whole forward is one pass over A (HBM-bound);
//   * complete: only row sums of relu(+-W), W = G G^T, are needed, never W itself: the Gram product is fused with
//     its relu-row-sum epilogue and W is never stored.
// Quirks B9 (term2 laid out variables-first) and B10 (scalar t4rc . relu_rc added to every row) are reproduced.
#include "common.cuh"

namespace ddb {


__host__ __device__ inline int pad4(int v) { return (v + 3) & ~3; }

// out[k][q] = sum_l Wt[l][k] * X[l][q] for nodes q in [0, nq); Wt is the transposed weight (pitch PP, multiple of 4),
// X has pitch px, out has pitch po.  One thread per node, 4 outputs per inner loop (1 LDS.32 + 1 LDS.128 per 4 FMA).
__device__ __forceinline__ void node_matvec(const float* __restrict__ Wt, int PP, int p, const float* X, int px,
                                            float* out, int po, int nq, int tid, int nt) {
    for (int q = tid; q < nq; q += nt) {
        for (int kb = 0; kb < PP; kb += 4) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int l = 0; l < p; ++l) {
                const float x = X[l * px + q];
                const float4 w = *reinterpret_cast<const float4*>(Wt + l * PP + kb);
                a0 = fmaf(w.x, x, a0);
                a1 = fmaf(w.y, x, a1);
                a2 = fmaf(w.z, x, a2);
                a3 = fmaf(w.w, x, a3);
            }
            if (kb + 0 < p) out[(kb + 0) * po + q] = a0;
            if (kb + 1 < p) out[(kb + 1) * po + q] = a1;
            if (kb + 2 < p) out[(kb + 2) * po + q] = a2;
            if (kb + 3 < p) out[(kb + 3) * po + q] = a3;
        }
    }
}

// out[k][q] = sum_l Wt[l][k] * X[l][q]: register tile, thread = 4 nodes x 4 outputs (two LDS.128 per 16 FMAs).  The pitches
// px / po are multiples of 4 floats, nq4 = number of nodes rounded up to 4 (the padding nodes hold garbage: never read).
__device__ __forceinline__ void node_matmul_tiled(const float* __restrict__ Wt, int PP, int p, const float* X, int px,
                                                  float* out, int po, int nq4, int tid, int nt) {
    const int NG = nq4 / 4, KG = PP / 4;
    for (int w = tid; w < NG * KG; w += nt) {
        const int ng = w % NG, kg = w / NG;
        float acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
        for (int l = 0; l < p; ++l) {
            const float4 xv = *reinterpret_cast<const float4*>(X + l * px + 4 * ng);
            const float4 wv = *reinterpret_cast<const float4*>(Wt + l * PP + 4 * kg);
            const float xq[4] = {xv.x, xv.y, xv.z, xv.w};
            const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(wr[r], xq[q], acc[q][r]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (4 * kg + r < p)
                *reinterpret_cast<float4*>(out + (4 * kg + r) * po + 4 * ng) = make_float4(acc[0][r], acc[1][r], acc[2][r], acc[3][r]);
    }
}

// y[k] = sum_l W[k][l] x[l] (row-major W in global memory), one warp per output.
__device__ __forceinline__ void small_matvec(const float* __restrict__ W, int p, const float* x, float* y, int warp,
                                             int lane, int nw) {
    for (int k = warp; k < p; k += nw) {
        float acc = 0.f;
        for (int l = lane; l < p; l += 32) acc = fmaf(__ldg(W + k * p + l), x[l], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) y[k] = acc;
    }
}

__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// stage a p x p row-major matrix from global memory into shared memory transposed: Wt[l][k] = W[k][l]
__device__ __forceinline__ void stage_transposed(const float* __restrict__ W, int p, int PP, float* Wt, int tid, int nt) {
    for (int e = tid; e < p * PP; e += nt) {
        const int l = e / PP, k = e - l * PP;
        Wt[e] = (k < p) ? __ldg(W + k * p + l) : 0.f;
    }
}

// =====================================================================================================================
// bipartite
// =====================================================================================================================
struct BipLayout {
    size_t t2c, t2v, t7, mu, agg, An, vecs, total;   // offsets in floats
};
// mode 2: normalised A and the aggregation buffer in shared memory; 1: aggregation buffer only (adjacency re-read from
// global memory); 0: neither (dense instances only -- a sparse instance then gets NaN outputs and sets the error flag)
__host__ __device__ inline BipLayout bip_layout(int m, int n, int p, int mode) {
    const int PP = pad4(p);
    BipLayout L;
    size_t off = 0;
    L.t2c = off; off += (size_t)p * PP;
    L.t2v = off; off += (size_t)p * PP;
    L.t7 = off;  off += (size_t)p * PP;
    L.mu = off;  off += (size_t)p * (m + n);
    L.agg = off; off += mode >= 1 ? (size_t)p * (m + n) : 0;
    L.An = off;  off += mode >= 2 ? (size_t)m * n : 0;
    L.vecs = off;
    off += (size_t)6 * m + 5 * n + 16 * PP + 64;     // row/col statistics + small p-vectors
    L.total = off;
    return L;
}
size_t s2v_bipartite_smem_bytes(int m, int n, int p, int mode) { return bip_layout(m, n, p, mode).total * 4; }

__global__ void __launch_bounds__(256) s2v_bipartite_kernel(S2vArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int m = a.m, n = a.n, p = a.p, PP = pad4(p), NP = m + n;
    const BipLayout L = bip_layout(m, n, p, a.store_A);
    float* t2cT = sm + L.t2c;
    float* t2vT = sm + L.t2v;
    float* t7T = sm + L.t7;
    float* mu = sm + L.mu;
    float* agg = sm + L.agg;
    float* An = sm + L.An;
    float* v = sm + L.vecs;
    float* rb = v;            v += m;    // b_i / norm_i            (c_feats[:,1] after s2v.py:293)
    float* cosv = v;          v += m;    // <a_i', c>               (c_feats[:,3], s2v.py:297)
    float* Sp = v;            v += m;    // sum_j relu(a'_ij)
    float* Sn = v;            v += m;    // sum_j relu(-a'_ij)
    float* rcnt = v;          v += m;    // nonzeros in row i
    float* z7 = v;            v += m;    // scratch
    float* cj = v;            v += n;    // objective coefficients (v_feats)
    float* Cp = v;            v += n;    // sum_i relu(a'_ij)
    float* Cn = v;            v += n;    // sum_i relu(-a'_ij)
    float* ccnt = v;          v += n;    // nonzeros in column j
    float* ctmp = v;          v += n;
    float* w3cp = v;          v += PP;   // t3c . relu(t4c)
    float* w3cn = v;          v += PP;   // t3c . relu(-t4c)
    float* w3vp = v;          v += PP;
    float* w3vn = v;          v += PP;
    float* meanc = v;         v += PP;
    float* meanv = v;         v += PP;
    float* yv = v;            v += PP;   // t2c . mean_c   (goes to node positions < n : quirk B9)
    float* yc = v;            v += PP;   // t2v . mean_v   (goes to node positions >= n)
    float* u6 = v;            v += PP;
    float* tmp1 = v;          v += PP;
    float* tmp2 = v;          v += PP;
    float* r4 = v;            v += 4 * PP;
    int* iflag = reinterpret_cast<int*>(v);

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    // parameter block offsets (reference state_dict order)
    const float* P = a.params;
    const float* t0 = P;                 P += p;
    const float* t1c = P;                P += 4 * p;
    const float* t1v = P;                P += p;
    const float* t2c = P;                P += p * p;
    const float* t2v = P;                P += p * p;
    const float* t3c = P;                P += p * p;
    const float* t3v = P;                P += p * p;
    const float* t4c = P;                P += p;
    const float* t4v = P;                P += p;
    const float* t6c = P;                P += p * p;
    const float* t6v = P;                P += p * p;
    const float* t7 = P;                 P += p * p;
    const float* t8 = P;

    // ---- once per CTA: weights into shared memory, t3 . relu(+-t4) ---------------------------------------------------
    stage_transposed(t2c, p, PP, t2cT, tid, nt);
    stage_transposed(t2v, p, PP, t2vT, tid, nt);
    stage_transposed(t7, p, PP, t7T, tid, nt);
    for (int l = tid; l < p; l += nt) {
        r4[l] = fmaxf(__ldg(t4c + l), 0.f);
        r4[PP + l] = fmaxf(-__ldg(t4c + l), 0.f);
        r4[2 * PP + l] = fmaxf(__ldg(t4v + l), 0.f);
        r4[3 * PP + l] = fmaxf(-__ldg(t4v + l), 0.f);
    }
    __syncthreads();
    small_matvec(t3c, p, r4, w3cp, warp, lane, nw);
    small_matvec(t3c, p, r4 + PP, w3cn, warp, lane, nw);
    small_matvec(t3v, p, r4 + 2 * PP, w3vp, warp, lane, nw);
    small_matvec(t3v, p, r4 + 3 * PP, w3vn, warp, lane, nw);
    __syncthreads();

    if (a.only_flagged && *a.flag_count == 0) return;   // the dense kernel handled every instance
    for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {
        if (a.only_flagged && a.inst_flag[lp] == 0) continue;
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg = a.c + (size_t)lp * n;
        for (int j = tid; j < n; j += nt) {
            cj[j] = (float)cg[j];
            Cp[j] = 0.f; Cn[j] = 0.f; ccnt[j] = 0.f;
        }
        if (tid == 0) *iflag = 0;
        __syncthreads();

        // ---- pass over A: row normalisation, cosines, relu row/column sums, adjacency degrees -------------------------
        for (int i = warp; i < m; i += nw) {
            const float bi = (float)bg[i];
            float ss = 0.f;
            for (int j = lane; j < n; j += 32) {
                const float x = (float)Ag[(size_t)i * n + j];
                ss = fmaf(x, x, ss);
            }
            ss = warp_sumf(ss) + bi * bi;                     // ||[a_i | -b_i]||^2
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
            float cs = 0.f, sp = 0.f, sn = 0.f, cnt = 0.f;
            for (int j = lane; j < n; j += 32) {
                const float x = (float)Ag[(size_t)i * n + j] * inv;
                cs = fmaf(x, cj[j], cs);
                sp += fmaxf(x, 0.f);
                sn += fmaxf(-x, 0.f);
                const float nz = (x != 0.f) ? 1.f : 0.f;
                cnt += nz;
                // column statistics: every (warp, j) pair is touched by one lane only -> shared-memory atomics are cheap
                atomicAdd(&Cp[j], fmaxf(x, 0.f));
                atomicAdd(&Cn[j], fmaxf(-x, 0.f));
                atomicAdd(&ccnt[j], nz);
                if (a.store_A >= 2) An[(size_t)i * n + j] = x;
            }
            cs = warp_sumf(cs); sp = warp_sumf(sp); sn = warp_sumf(sn); cnt = warp_sumf(cnt);
            if (lane == 0) {
                rb[i] = bi * inv;
                cosv[i] = cs;
                Sp[i] = sp; Sn[i] = sn; rcnt[i] = cnt;
                if (cnt != (float)n) *iflag = 1;              // benign race: any writer sets the same value
            }
        }
        for (int e = tid; e < p * NP; e += nt) mu[e] = 0.f;
        __syncthreads();
        const bool dense = (*iflag == 0);
        if (!dense && a.store_A == 0) {
            // general adjacency does not fit in shared memory for this shape: fail loudly (NaN outputs + error flag)
            if (tid == 0) *a.error_flag = 1;
            const float qnan = __int_as_float(0x7fc00000);
            for (int e = tid; e < 2 * m; e += nt) {
                a.logp[(size_t)lp * m * 2 + e] = qnan;
                if (a.probs) a.probs[(size_t)lp * m * 2 + e] = qnan;
            }
            __syncthreads();
            continue;
        }
        // adjacency test of the general path: from the shared-memory copy, else from the caller's A (fp32 cast as the
        // reference's item tensors, s2v.py:275-283)
        auto adjacent = [&](int i, int j) -> bool {
            return (a.store_A >= 2) ? (An[(size_t)i * n + j] != 0.f) : ((float)Ag[(size_t)i * n + j] != 0.f);
        };

        // ---- T rounds of message passing ---------------------------------------------------------------------------------
        for (int t = 0; t < a.T; ++t) {
            if (dense) {
                for (int l = warp; l < p; l += nw) {
                    float sc = 0.f, svv = 0.f;
                    for (int i = lane; i < m; i += 32) sc += mu[l * NP + i];
                    for (int j = lane; j < n; j += 32) svv += mu[l * NP + m + j];
                    sc = warp_sumf(sc); svv = warp_sumf(svv);
                    if (lane == 0) { meanc[l] = sc / (float)m; meanv[l] = svv / (float)n; }
                }
                __syncthreads();
                small_matvec(t2c, p, meanc, yv, warp, lane, nw);
                small_matvec(t2v, p, meanv, yc, warp, lane, nw);
                __syncthreads();
            } else {
                // general adjacency: agg_v[:,j] = mean of mu_c over rows adjacent to j, agg_c[:,i] likewise over columns
                for (int e = tid; e < p * n; e += nt) {
                    const int l = e / n, j = e - l * n;
                    float acc = 0.f;
                    for (int i = 0; i < m; ++i)
                        if (adjacent(i, j)) acc += mu[l * NP + i];
                    agg[l * NP + j] = acc / fmaxf(ccnt[j], 1e-12f);
                }
                for (int e = tid; e < p * m; e += nt) {
                    const int l = e / m, i = e - l * m;
                    float acc = 0.f;
                    for (int j = 0; j < n; ++j)
                        if (adjacent(i, j)) acc += mu[l * NP + m + j];
                    agg[l * NP + n + i] = acc / fmaxf(rcnt[i], 1e-12f);
                }
                __syncthreads();
                // term2 in place: positions < n get t2c . agg_v, positions >= n get t2v . agg_c  (B9 layout)
                node_matvec(t2cT, PP, p, agg, NP, mu, NP, n, tid, nt);
                node_matvec(t2vT, PP, p, agg + n, NP, mu + n, NP, m, tid, nt);
                __syncthreads();
            }
            for (int e = tid; e < p * NP; e += nt) {
                const int l = e / NP, q = e - l * NP;
                float val = __ldg(t0 + l);
                if (q < m) {
                    const float fi = a.row_ineq ? (float)a.row_ineq[(size_t)lp * m + q] : 1.f;
                    const float fb = a.row_bound ? (float)a.row_bound[(size_t)lp * m + q] : 0.f;
                    val += __ldg(t1c + 4 * l) * fi + __ldg(t1c + 4 * l + 1) * rb[q] + __ldg(t1c + 4 * l + 2) * fb + __ldg(t1c + 4 * l + 3) * cosv[q];
                    val += w3cp[l] * Sp[q] + w3cn[l] * Sn[q];
                } else {
                    const int j = q - m;
                    val += __ldg(t1v + l) * cj[j] + w3vp[l] * Cp[j] + w3vn[l] * Cn[j];
                }
                if (dense) val += (q < n) ? yv[l] : yc[l];
                else val += mu[e];
                mu[e] = fmaxf(val, 0.f);
            }
            __syncthreads();
        }

        // ---- head: u6, t7 . mu[:, i], relu, t8, (log-)softmax ----------------------------------------------------------------
        for (int l = warp; l < p; l += nw) {
            float sc = 0.f, svv = 0.f;
            for (int i = lane; i < m; i += 32) sc += mu[l * NP + i];
            for (int j = lane; j < n; j += 32) svv += mu[l * NP + m + j];
            sc = warp_sumf(sc); svv = warp_sumf(svv);
            if (lane == 0) { meanc[l] = sc / (float)m; meanv[l] = svv / (float)n; }
        }
        __syncthreads();
        small_matvec(t6c, p, meanc, tmp1, warp, lane, nw);
        small_matvec(t6v, p, meanv, tmp2, warp, lane, nw);
        __syncthreads();
        for (int l = tid; l < p; l += nt) u6[l] = fmaxf(tmp1[l] + tmp2[l], 0.f);   // relu(u6) is what t8 sees
        __syncthreads();
        // scores_i = t8[:, :p] . relu(u6) + t8[:, p:2p] . relu(t7 mu_i) + t8[:, 2p:2p+4] . c_feats_i
        const int W8 = 2 * p + 4;
        for (int i = tid; i < m; i += nt) {
            float s0 = 0.f, s1 = 0.f;
            for (int l = 0; l < p; ++l) {
                s0 = fmaf(__ldg(t8 + l), u6[l], s0);
                s1 = fmaf(__ldg(t8 + W8 + l), u6[l], s1);
            }
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
                        s0 = fmaf(__ldg(t8 + p + kb + u), r[u], s0);
                        s1 = fmaf(__ldg(t8 + W8 + p + kb + u), r[u], s1);
                    }
                }
            }
            // c_feats = [is_inequality, rhs', is_bound, cosine]  (flags: 1 / 0 on random LPs, per row on MPS / PLNN items)
            const float f0 = a.row_ineq ? (float)a.row_ineq[(size_t)lp * m + i] : 1.f, f1 = rb[i];
            const float f2 = a.row_bound ? (float)a.row_bound[(size_t)lp * m + i] : 0.f, f3 = cosv[i];
            s0 += __ldg(t8 + 2 * p) * f0 + __ldg(t8 + 2 * p + 1) * f1 + __ldg(t8 + 2 * p + 2) * f2 + __ldg(t8 + 2 * p + 3) * f3;
            s1 += __ldg(t8 + W8 + 2 * p) * f0 + __ldg(t8 + W8 + 2 * p + 1) * f1 + __ldg(t8 + W8 + 2 * p + 2) * f2 + __ldg(t8 + W8 + 2 * p + 3) * f3;
            const float mx = fmaxf(s0, s1);
            const float lse = mx + logf(expf(s0 - mx) + expf(s1 - mx));
            float* lo = a.logp + ((size_t)lp * m + i) * 2;
            lo[0] = s0 - lse;
            lo[1] = s1 - lse;
            if (a.probs) {
                float* po = a.probs + ((size_t)lp * m + i) * 2;
                po[0] = expf(s0 - lse);
                po[1] = expf(s1 - lse);
            }
        }
        __syncthreads();
    }
    (void)z7; (void)ctmp;
}

// =====================================================================================================================
// complete
// =====================================================================================================================
struct CmpLayout {
    size_t t2rr, t7, G, mu, mu2, vecs, total;   // floats
};
// with_G = false when the Gram row sums come from the tensor-core kernel: G is not staged and two CTAs fit per SM
__host__ __device__ inline CmpLayout cmp_layout(int m, int n, int p, bool with_G) {
    const int PP = pad4(p);
    int PG = n + 1;
    if ((PG & 1) == 0) PG += 1;                 // odd pitch: conflict-free column walks
    CmpLayout L;
    size_t off = 0;
    L.t2rr = off; off += (size_t)p * PP;
    L.t7 = off;   off += (size_t)p * PP;
    L.G = off;    off += with_G ? (size_t)(m + 1) * PG : 0;
    off = (off + 3) & ~(size_t)3;
    L.mu = off;   off += (size_t)p * pad4(m + 1);      // pitch pad4(m + 1): float4 rows for the register-tiled products
    L.mu2 = off;  off += (size_t)p * pad4(m + 1);
    L.vecs = off; off += (size_t)3 * (m + 1) + 16 * PP + 64;
    L.total = off;
    return L;
}
size_t s2v_complete_smem_bytes(int m, int n, int p, bool with_G) { return cmp_layout(m, n, p, with_G).total * 4; }

__global__ void __launch_bounds__(256) s2v_complete_kernel(S2vArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int m = a.m, n = a.n, p = a.p, PP = pad4(p), M1 = m + 1, MP = pad4(m + 1);
    int PG = n + 1;
    if ((PG & 1) == 0) PG += 1;
    const CmpLayout L = cmp_layout(m, n, p, a.gram == nullptr);
    float* t2rrT = sm + L.t2rr;
    float* t7T = sm + L.t7;
    float* G = sm + L.G;
    float* mu = sm + L.mu;
    float* mu2 = sm + L.mu2;
    float* v = sm + L.vecs;
    float* Wp = v;     v += M1;    // sum_{j<m, j!=i} relu(W_ij)
    float* Wn = v;     v += M1;    // sum_{j<m, j!=i} relu(-W_ij)
    float* wc = v;     v += M1;    // W[m][j] = <c, a_j'> (cost node against row j)
    float* w3p = v;    v += PP;    // t3rr . relu(t4rr)
    float* w3n = v;    v += PP;    // t3rr . relu(-t4rr)
    float* meanr = v;  v += PP;
    float* muc = v;    v += PP;
    float* y1 = v;     v += PP;    // t2rc . mu_c
    float* y2 = v;     v += PP;    // t2cr . mean(mu_r)
    float* u3c = v;    v += PP;    // t3cr . relu_cr
    float* u6 = v;     v += PP;
    float* tmp1 = v;   v += PP;
    float* tmp2 = v;   v += PP;
    float* r4 = v;     v += 4 * PP;
    float* scal = v;               // [0] = scalar t4rc . relu_rc (B10), [1] = sum relu(wc), [2] = sum relu(-wc)

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const float* P = a.params;
    const float* t0 = P;     P += p;
    const float* t1 = P;     P += p;
    const float* t2rr = P;   P += p * p;
    const float* t2rc = P;   P += p * p;
    const float* t2cr = P;   P += p * p;
    const float* t3rr = P;   P += p * p;
    P += p * p;              // t3rc: unused by the reference forward (B10)
    const float* t3cr = P;   P += p * p;
    const float* t4rr = P;   P += p;
    const float* t4rc = P;   P += p;
    const float* t4cr = P;   P += p;
    const float* t6r = P;    P += p * p;
    const float* t6c = P;    P += p * p;
    const float* t7 = P;     P += p * p;
    const float* t8 = P;

    stage_transposed(t2rr, p, PP, t2rrT, tid, nt);
    stage_transposed(t7, p, PP, t7T, tid, nt);
    for (int l = tid; l < p; l += nt) {
        r4[l] = fmaxf(__ldg(t4rr + l), 0.f);
        r4[PP + l] = fmaxf(-__ldg(t4rr + l), 0.f);
    }
    __syncthreads();
    small_matvec(t3rr, p, r4, w3p, warp, lane, nw);
    small_matvec(t3rr, p, r4 + PP, w3n, warp, lane, nw);
    __syncthreads();

    for (long long lp = blockIdx.x; lp < a.B; lp += gridDim.x) {
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg = a.c + (size_t)lp * n;
        if (a.gram) {
            // Wp, Wn, wc were produced by the tcgen05 Gram kernel (s2v_gram_tc.cu): G never has to be built here
            const float* gr = a.gram + (size_t)lp * 3 * a.gram_pitch;
            for (int i = tid; i < M1; i += nt) {
                Wp[i] = __ldg(gr + i);
                Wn[i] = __ldg(gr + a.gram_pitch + i);
                wc[i] = __ldg(gr + 2 * a.gram_pitch + i);
            }
            for (int e = tid; e < p * MP; e += nt) mu[e] = 0.f;
            __syncthreads();
        } else {
            // ---- G = [normalize([A | b]) ; [c, 0]]  (normalisation in fp64 as the reference does, s2v.py:145) ----------------
            for (int i = warp; i < m; i += nw) {
                const double bi = bg[i];
                double ss = 0.0;
                for (int j = lane; j < n; j += 32) {
                    const double x = Ag[(size_t)i * n + j];
                    ss = fma(x, x, ss);
                }
                ss = warp_sum(ss) + bi * bi;
                const double inv = 1.0 / fmax(sqrt(ss), 1e-12);
                for (int j = lane; j < n; j += 32) G[i * PG + j] = (float)(Ag[(size_t)i * n + j] * inv);
                if (lane == 0) G[i * PG + n] = (float)(bi * inv);
            }
            for (int j = tid; j <= n; j += nt) G[m * PG + j] = (j < n) ? (float)cg[j] : 0.f;
            for (int e = tid; e < p * MP; e += nt) mu[e] = 0.f;
            __syncthreads();

            // ---- fused Gram + relu row sums: W_ij = <G_i, G_j>, never stored ----------------------------------------------------
            // each warp takes 4 rows at a time; lanes walk the columns j (conflict-free: odd pitch)
            for (int i0 = warp * 4; i0 < M1; i0 += nw * 4) {
                float sp[4] = {0.f, 0.f, 0.f, 0.f}, sn[4] = {0.f, 0.f, 0.f, 0.f}, wl[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = lane; j < M1; j += 32) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    const float* gj = G + j * PG;
                    for (int k = 0; k <= n; ++k) {
                        const float x = gj[k];
    #pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = (i0 + u < M1) ? i0 + u : m;
                            acc[u] = fmaf(G[i * PG + k], x, acc[u]);
                        }
                    }
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u;
                        if (i < M1 && j != i) {
                            if (j < m) { sp[u] += fmaxf(acc[u], 0.f); sn[u] += fmaxf(-acc[u], 0.f); }
                            else wl[u] = acc[u];                      // column m: against the cost node
                        }
                    }
                }
    #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float a0 = warp_sumf(sp[u]), a1 = warp_sumf(sn[u]), a2 = warp_sumf(wl[u]);
                    if (lane == 0 && i0 + u < M1) {
                        Wp[i0 + u] = a0; Wn[i0 + u] = a1; wc[i0 + u] = a2;
                    }
                }
            }
            __syncthreads();
        }
        // cost-node statistics: sums of relu(+-W[m][j]) over rows j < m  (W is symmetric: W[:m, m] == W[m, :m])
        if (warp == 0) {
            float sp = 0.f, sn = 0.f;
            for (int j = lane; j < m; j += 32) { sp += fmaxf(wc[j], 0.f); sn += fmaxf(-wc[j], 0.f); }
            sp = warp_sumf(sp); sn = warp_sumf(sn);
            if (lane == 0) { scal[1] = sp; scal[2] = sn; }
        }
        __syncthreads();
        if (warp == 0) {   // scalar of B10: t4rc . (relu(t4rc) S+ + relu(-t4rc) S-)
            float acc = 0.f;
            for (int l = lane; l < p; l += 32) {
                const float tv = __ldg(t4rc + l);
                acc += tv * (fmaxf(tv, 0.f) * scal[1] + fmaxf(-tv, 0.f) * scal[2]);
            }
            acc = warp_sumf(acc);
            if (lane == 0) scal[0] = acc;
        }
        for (int l = tid; l < p; l += nt) {
            const float tv = __ldg(t4cr + l);
            tmp1[l] = fmaxf(tv, 0.f) * scal[1] + fmaxf(-tv, 0.f) * scal[2];      // relu_cr
        }
        __syncthreads();
        small_matvec(t3cr, p, tmp1, u3c, warp, lane, nw);
        __syncthreads();

        // ---- T rounds ----------------------------------------------------------------------------------------------------------
        // mu ping-pongs between the two buffers; a round's update (base + t2rr mu_i + t2rc mu_c, relu) is the epilogue of the
        // register-tiled product, and round 0 -- the embeddings start at zero -- is the base alone.  Padding nodes hold 0.
        float* mu_cur = mu;
        float* mu_nxt = mu2;
        for (int l = tid; l < PP; l += nt) { muc[l] = 0.f; meanr[l] = 0.f; }
        __syncthreads();
        for (int t = 0; t < a.T; ++t) {
            if (t > 0) {
                small_matvec(t2rc, p, muc, y1, warp, lane, nw);
                small_matvec(t2cr, p, meanr, y2, warp, lane, nw);
                __syncthreads();
            }
            {
                const int NG = MP / 4, KG = PP / 4;
                for (int w = tid; w < NG * KG; w += nt) {
                    const int ng = w % NG, kg = w / NG;
                    float acc[4][4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
                    if (t > 0) {
                        for (int l = 0; l < p; ++l) {
                            const float4 xv = *reinterpret_cast<const float4*>(mu_cur + l * MP + 4 * ng);
                            const float4 wv = *reinterpret_cast<const float4*>(t2rrT + l * PP + 4 * kg);
                            const float xq[4] = {xv.x, xv.y, xv.z, xv.w};
                            const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q)
#pragma unroll
                                for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(wr[r], xq[q], acc[q][r]);
                        }
                    }
                    float wpq[4], wnq[4], nfq[4];      // nfq: node feature of the row node (1 = inequality row; s2v.py:100 u1 = t0 + t1 feat)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = 4 * ng + q;
                        wpq[q] = (i < m) ? Wp[i] : 0.f;
                        wnq[q] = (i < m) ? Wn[i] : 0.f;
                        nfq[q] = (a.row_ineq && i < m) ? (float)a.row_ineq[(size_t)lp * m + i] : 1.f;
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int l = 4 * kg + r;
                        if (l < p) {
                            const float t0l = __ldg(t0 + l), t1l = __ldg(t1 + l), y1l = (t > 0 ? y1[l] : 0.f);
                            const float add = t0l + t1l + scal[0] + y1l;          // every row an inequality (random LPs)
                            const float cp = w3p[l], cn = w3n[l];
                            float o[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int i = 4 * ng + q;
                                const float addq = a.row_ineq ? ((t0l + t1l * nfq[q]) + scal[0]) + y1l : add;
                                o[q] = (i < m) ? fmaxf(acc[q][r] + addq + cp * wpq[q] + cn * wnq[q], 0.f) : 0.f;
                            }
                            *reinterpret_cast<float4*>(mu_nxt + l * MP + 4 * ng) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            }
            __syncthreads();
            // the cost node (its node feature is 0) and the mean over the row nodes of the NEW embeddings
            for (int l = tid; l < p; l += nt) muc[l] = fmaxf(__ldg(t0 + l) + (t > 0 ? y2[l] : 0.f) + u3c[l], 0.f);
            for (int l = warp; l < p; l += nw) {
                float sr = 0.f;
                for (int i = lane; i < m; i += 32) sr += mu_nxt[l * MP + i];
                sr = warp_sumf(sr);
                if (lane == 0) meanr[l] = sr / (float)m;
            }
            float* sw = mu_cur; mu_cur = mu_nxt; mu_nxt = sw;
            __syncthreads();
        }
        if (a.T == 0) {
            for (int e = tid; e < p * MP; e += nt) mu_cur[e] = 0.f;
            __syncthreads();
        }
        mu = mu_cur;
        mu2 = mu_nxt;

        // ---- head (meanr / muc are those of the final embeddings) ------------------------------------------------------------------
        small_matvec(t6r, p, meanr, tmp1, warp, lane, nw);
        small_matvec(t6c, p, muc, tmp2, warp, lane, nw);
        __syncthreads();
        for (int l = tid; l < p; l += nt) u6[l] = fmaxf(tmp1[l] + tmp2[l], 0.f);
        __syncthreads();
        const int W8 = 2 * p;
        node_matmul_tiled(t7T, PP, p, mu, MP, mu2, MP, pad4(m), tid, nt);     // t7 . mu_r (the head's only p x p x m product)
        __syncthreads();
        for (int i = tid; i < m; i += nt) {
            float s0 = 0.f, s1 = 0.f;
            for (int l = 0; l < p; ++l) {
                s0 = fmaf(__ldg(t8 + l), u6[l], s0);
                s1 = fmaf(__ldg(t8 + W8 + l), u6[l], s1);
            }
            for (int kk = 0; kk < p; ++kk) {
                const float z = fmaxf(mu2[kk * MP + i], 0.f);          // relu(t7 mu_i), product computed above
                s0 = fmaf(__ldg(t8 + p + kk), z, s0);
                s1 = fmaf(__ldg(t8 + W8 + p + kk), z, s1);
            }
            const float mx = fmaxf(s0, s1);
            const float lse = mx + logf(expf(s0 - mx) + expf(s1 - mx));
            float* lo = a.logp + ((size_t)lp * m + i) * 2;
            lo[0] = s0 - lse;
            lo[1] = s1 - lse;
            if (a.probs) {
                float* po = a.probs + ((size_t)lp * m + i) * 2;
                po[0] = expf(s0 - lse);
                po[1] = expf(s1 - lse);
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_s2v_forward(const S2vArgs& a0, int sm_count, long long smem_optin, cudaStream_t st, const char** why) {
    S2vArgs a = a0;
    *why = "";
    if (a.graph == 1) {
        a.store_A = (long long)s2v_bipartite_smem_bytes(a.m, a.n, a.p, 2) <= smem_optin ? 2
                    : ((long long)s2v_bipartite_smem_bytes(a.m, a.n, a.p, 1) <= smem_optin ? 1 : 0);
        const size_t smem = s2v_bipartite_smem_bytes(a.m, a.n, a.p, a.store_A);
        if ((long long)smem > smem_optin) { *why = "bipartite forward: embeddings do not fit in shared memory"; return cudaErrorInvalidValue; }
        cudaError_t e = cudaFuncSetAttribute(s2v_bipartite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s2v_bipartite_kernel, 256, smem);
        if (e != cudaSuccess) return e;
        long long grid = (long long)sm_count * (per_sm > 0 ? per_sm : 1);
        if (grid > a.B) grid = a.B;
        s2v_bipartite_kernel<<<(int)grid, 256, smem, st>>>(a);
    } else {
        const size_t smem = s2v_complete_smem_bytes(a.m, a.n, a.p, a.gram == nullptr);
        if ((long long)smem > smem_optin) { *why = "complete forward: G and embeddings do not fit in shared memory"; return cudaErrorInvalidValue; }
        cudaError_t e = cudaFuncSetAttribute(s2v_complete_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s2v_complete_kernel, 256, smem);
        if (e != cudaSuccess) return e;
        long long grid = (long long)sm_count * (per_sm > 0 ? per_sm : 1);
        if (grid > a.B) grid = a.B;
        s2v_complete_kernel<<<(int)grid, 256, smem, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace ddb
