// Batched fp64 simplex for tableaus beyond one SM (plan 6): ONE LP PER THREAD-BLOCK CLUSTER, the live tableau spread over
// the shared memory of the cluster's CTAs (distributed shared memory), e.g. (m,n) = (500,250): 250 x 251 doubles = 502 KB
// in a cluster of four SMs.
//
// Same algorithm as the other kernels: crash as an explicit inverse (Gauss-Jordan on the n x n block
// of the rows most opposed to c), the remaining m - n rows enter through P_N = -A_N D, phase 1 (most negative slack
// leaves, ratio test along its row), phase 2 (Dantzig), x = xv - D sigma, labels from the caller's A.  What is new is where
// the data lives and how a pivot travels:
//   * CTA `rank` of the cluster owns the tableau rows [rank * LR, (rank + 1) * LR) -- crash rows first, live rows later --
//     row-major in its shared memory with an ODD pitch, so that a column read (one row per lane) is conflict-free;
//   * inside a CTA a thread owns tableau COLUMN(S) tid, tid + 256: the rank-1 update  T[i][j] -= f_i * prow[j]  keeps
//     prow[j] in a register, reads f_i as a shared-memory broadcast and streams its column of the local rows with
//     coalesced 8-byte accesses -- no tableau entry ever leaves the SM that owns it;
//   * per pivot two things cross the cluster, both through distributed shared memory followed by a cluster barrier:
//     the per-CTA candidates of the leaving row (16 bytes each, written into every CTA's candidate table) and the pivot
//     row (n + 1 doubles, written by its owner into every CTA's row buffer).  The column vectors (reduced costs g, ghat,
//     the column -> constraint map) are replicated: every CTA updates its own copy from the broadcast row with the same
//     arithmetic, so pricing needs no communication at all.
//   * the crash inverse D goes to a per-cluster global scratch (L2-resident); the product reads D[k][j] once per thread
//     and k from there and the coefficients A_N[i][k] as shared-memory broadcasts (the CTA's rows of A_N are staged in the
//     tableau's own space, 32 accumulators per thread in registers).
// The cluster's global scratch (D, x, the basic-slack table) is written by one CTA and read by the others after a cluster
// barrier; those reads are L2-only loads (ld.global.cg), so no stale line of an earlier LP in an SM's L1 can be seen.
// Instances with a singular static crash basis or an ill-conditioned vertex (an active row with a visible residual) are
// flagged status = -1 and re-solved by the generic kernel on the device (capi.cu), as the register kernels do.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

// dev-only section accounting (tools/cluster_timing.cu builds with -DDDB_CL_TIMING; never defined in the library build)
#ifdef DDB_CL_TIMING
#define CL_T(i) do { if (tid == 0 && rank == 0) { const long long t_ = clock64(); tacc[i] += (double)(t_ - tlast); tlast = t_; } } while (0)
#else
#define CL_T(i)
#endif

namespace ddb {

namespace {

constexpr int kClNT = 512;           // threads per CTA: (row group, column) = (tid / CW, tid % CW), CW = n + 1 rounded up to 32
constexpr int kClWarps = kClNT / 32;
constexpr int kClChunk = 32;         // rows of the product a thread accumulates in registers at a time
constexpr int kClBatch = 8;          // rows of the rank-1 update a thread keeps in flight

struct ClCand {                      // a CTA's candidate for the leaving row
    unsigned long long key;          // order-preserving key of the slack (phase 1) / ratio (phase 2); KEY_INF = none
    int row;                         // global live-row index
    int var;                         // constraint whose slack is basic in that row
};
struct ClPub {                       // what the pivot row's owner publishes beside the row
    double p;                        // pivot entry (crash only: the others read it off the row)
    int k;                           // crash: pivot column (-1 = singular)
    int var;                         // constraint whose slack was basic in the pivot row
};

struct ClLayout {
    int LR, PD, NCP;
    size_t T, prow, g, gh, ebuf, xs, sig, scores, order, cvmap, colvar0, pivcol, rowvar, cand, pub, wred, cnt, lp, total;
};
__host__ __device__ inline size_t cl_align(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline ClLayout cl_layout(int m, int n, int CL) {
    ClLayout L;
    const int ncol = n + 1;
    const int rows = (m - n > n) ? (m - n) : n;
    L.LR = (rows + CL - 1) / CL;
    L.PD = ncol | 1;                                  // odd pitch: a column read over consecutive rows hits distinct banks
    L.NCP = (ncol + 1) & ~1;
    size_t off = 0;
    L.T = off;       off += cl_align((size_t)L.LR * L.PD * 8);
    L.prow = off;    off += cl_align((size_t)2 * CL * L.NCP * 8);   // [parity][candidate CTA][NCP]
    L.g = off;       off += cl_align((size_t)L.NCP * 8);
    L.gh = off;      off += cl_align((size_t)L.NCP * 8);
    L.ebuf = off;    off += cl_align((size_t)L.LR * 8);
    L.xs = off;      off += cl_align((size_t)n * 8);
    L.sig = off;     off += cl_align((size_t)n * 8);
    L.scores = off;  off += cl_align((size_t)m * 8);
    L.order = off;   off += cl_align((size_t)m * 4);
    L.cvmap = off;   off += cl_align((size_t)L.NCP * 4);
    L.colvar0 = off; off += cl_align((size_t)n * 4);
    L.pivcol = off;  off += cl_align((size_t)n * 4);
    L.rowvar = off;  off += cl_align((size_t)L.LR * 4);
    L.cand = off;    off += cl_align((size_t)2 * 8 * sizeof(ClCand));
    L.pub = off;     off += cl_align((size_t)2 * sizeof(ClPub));
    L.wred = off;    off += cl_align((size_t)2 * kClWarps * 16);
    L.cnt = off;     off += cl_align((size_t)8 * 4 * 4);
    L.lp = off;      off += 16;
    L.total = off;
    return L;
}

// block-wide argmin of a 64-bit key with an int payload; ties: lowest payload wins.  Every thread gets the result.
// `wred` = 2 x kClWarps x {key, payload}: the staging area alternates (`flip`), so ONE barrier per call is enough.
__device__ __forceinline__ void block_argmin(unsigned long long key, int payload, unsigned long long* wred, int& flip, int tid,
                                             unsigned long long& kmin, int& pmin) {
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long k1;
    warp_argmin_key(key, k1);
    int pbest = (key == k1) ? payload : 0x7fffffff;   // among the lanes that hold the minimum key: the lowest payload
    pbest = __reduce_min_sync(FULL, pbest);
    unsigned long long* w = wred + (size_t)flip * 2 * kClWarps;
    flip ^= 1;
    if (lane == 0) {
        w[2 * warp] = k1;
        w[2 * warp + 1] = (unsigned long long)(unsigned)pbest;
    }
    __syncthreads();
    // second level: lane l < kClWarps takes warp l's result
    unsigned long long kw = (lane < kClWarps) ? w[2 * lane] : ~0ull;
    const int pw = (lane < kClWarps) ? (int)w[2 * lane + 1] : 0x7fffffff;
    unsigned long long k2;
    warp_argmin_key(kw, k2);
    int p2 = (kw == k2) ? pw : 0x7fffffff;
    p2 = __reduce_min_sync(FULL, p2);
    kmin = k2;
    pmin = p2;
}

}  // namespace

template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kClNT, 1) simplex_cluster_kernel(SolveArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n, ncol = n + 1;
    const ClLayout L = cl_layout(m, n, CL);
    const int LR = L.LR, PD = L.PD, NCP = L.NCP;
    double* Tl = reinterpret_cast<double*>(smem_raw + L.T);          // [LR][PD] my rows of the tableau; column n = right-hand side
    double* prow = reinterpret_cast<double*>(smem_raw + L.prow);     // [2][CL][NCP] candidate pivot rows as broadcast by their owners (raw)
    double* g = reinterpret_cast<double*>(smem_raw + L.g);           // true reduced costs (replicated)
    double* gh = reinterpret_cast<double*>(smem_raw + L.gh);         // artificial costs of phase 1 (replicated)
    double* ebuf = reinterpret_cast<double*>(smem_raw + L.ebuf);     // entering column of my rows
    double* xs = reinterpret_cast<double*>(smem_raw + L.xs);
    double* sig = reinterpret_cast<double*>(smem_raw + L.sig);
    double* scores = reinterpret_cast<double*>(smem_raw + L.scores);
    int* order = reinterpret_cast<int*>(smem_raw + L.order);
    int* cvmap = reinterpret_cast<int*>(smem_raw + L.cvmap);         // column -> constraint whose slack is nonbasic (-1 free x_j)
    int* colvar0 = reinterpret_cast<int*>(smem_raw + L.colvar0);
    int* pivcol = reinterpret_cast<int*>(smem_raw + L.pivcol);
    int* rowvar = reinterpret_cast<int*>(smem_raw + L.rowvar);
    ClCand* cand = reinterpret_cast<ClCand*>(smem_raw + L.cand);     // [2][8]
    ClPub* pub = reinterpret_cast<ClPub*>(smem_raw + L.pub);         // [2]
    unsigned long long* wred = reinterpret_cast<unsigned long long*>(smem_raw + L.wred);
    int* cnt = reinterpret_cast<int*>(smem_raw + L.cnt);             // [8][4] per-CTA label counts, gathered in rank 0
    long long* cur_lp = reinterpret_cast<long long*>(smem_raw + L.lp);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = rank * kClWarps + warp, gnw = CL * kClWarps;
    // thread -> (row group, column): the columns of a row group fill whole warps
    const int CW = (ncol + 31) & ~31;
    const int NG = (kClNT / CW) > 0 ? (kClNT / CW) : 1;
    const int rg = tid / CW, cj = tid - rg * CW;
    const bool colact = (rg < NG) && (cj < ncol);      // I own column cj of the rows rg, rg + NG, ...
    const bool col0 = (rg == 0) && (cj < ncol);        // one thread per column
    int flip = 0;
    const long long cluster_id = blockIdx.x / CL;
    // per-cluster global scratch: D [n][NCP], x [n], val [m], flag [m] (ints)
    const size_t per_cluster = (size_t)n * NCP + n + m + (size_t)((m + 1) / 2);
    double* Dg = a.gtab + (size_t)cluster_id * per_cluster;
    double* xg = Dg + (size_t)n * NCP;
    double* valg = xg + n;
    int* flagg = reinterpret_cast<int*>(valg + m);

    // remote views of the shared buffers every CTA writes into
    double* r_prow[CL];
    ClCand* r_cand[CL];
    ClPub* r_pub[CL];
    double* r_scores[CL];
#pragma unroll
    for (int c = 0; c < CL; ++c) {
        r_prow[c] = cluster.map_shared_rank(prow, c);
        r_cand[c] = cluster.map_shared_rank(cand, c);
        r_pub[c] = cluster.map_shared_rank(pub, c);
        r_scores[c] = cluster.map_shared_rank(scores, c);
    }

    // rank-1 update of my rows [0, nloc) from the broadcast pivot row pr (raw), pivot entry p = pr[k]:
    //   rows i != pivot : f = ebuf[i] / p;  T[i][j] -= f pr[j] (j != k);  T[i][k] = -f
    //   pivot row (mine if lr >= 0): T[lr][j] = pr[j] / p (j != k);  T[lr][k] = 1 / p
    // A thread streams ITS column of the rows of its row group, kClBatch rows in flight (loads first, then the FMAs and the
    // stores: written as one loop the store of a row and the load of the next serialise on possible aliasing).
    auto update_rows = [&](const double* pr, int k, double rp, int nloc, int lr) {
        if (colact) {
            const double pj = pr[cj];
            const bool isk = (cj == k);
            double* col = Tl + cj;
            // full batches carry no guards at all (a guarded store drags its load into the branch and the rows serialise);
            // the pivot row is updated like any other and rewritten below by the same thread
            int i0 = rg;
            for (; i0 + (kClBatch - 1) * NG < nloc; i0 += kClBatch * NG) {
                double f[kClBatch], v[kClBatch];
#pragma unroll
                for (int u = 0; u < kClBatch; ++u) {
                    f[u] = ebuf[i0 + u * NG];
                    v[u] = col[(size_t)(i0 + u * NG) * PD];
                }
#pragma unroll
                for (int u = 0; u < kClBatch; ++u) {
                    const double fu = f[u] * rp;
                    col[(size_t)(i0 + u * NG) * PD] = isk ? -fu : fma(-fu, pj, v[u]);
                }
            }
            for (; i0 < nloc; i0 += NG) {
                const double fu = ebuf[i0] * rp;
                const double vv = col[(size_t)i0 * PD];
                col[(size_t)i0 * PD] = isk ? -fu : fma(-fu, pj, vv);
            }
            if (lr >= 0 && rg == lr % NG) Tl[(size_t)lr * PD + cj] = isk ? rp : pj * rp;
        }
    };
    // the owner writes row lr of its tableau into every CTA's row buffer
    auto publish_row = [&](int lr, int buf, int slot) {
        if (col0) {
            const double v = Tl[(size_t)lr * PD + cj];
#pragma unroll
            for (int c = 0; c < CL; ++c) r_prow[c][(size_t)(buf * CL + slot) * NCP + cj] = v;
        }
    };

    for (;;) {
        if (rank == 0 && tid == 0) {
            const long long v = (long long)atomicAdd(a.counter, 1ull);
#pragma unroll
            for (int c = 0; c < CL; ++c) *cluster.map_shared_rank(cur_lp, c) = v;
        }
        cluster.sync();
        const long long lp = *cur_lp;
        if (lp >= a.B) break;
#ifdef DDB_CL_TIMING
        double tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = clock64();
#endif
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg_ = a.c + (size_t)lp * n;
        const uint8_t* mask = a.row_mask ? a.row_mask + (size_t)lp * m : nullptr;   // reduced LP: rows with mask 0 are left out

        // ---- stage 0: crash order (rows dealt to the warps of the whole cluster, scores written to every CTA) --------------
        for (int i = gwarp; i < m; i += gnw) {
            double dot = 0.0, nn = 0.0;
            for (int j = lane; j < n; j += 32) {
                const double v = __ldg(Ag + (size_t)i * n + j);
                dot = fma(v, __ldg(cg_ + j), dot);
                nn = fma(v, v, nn);
            }
            dot = warp_sum(dot);
            nn = warp_sum(nn);
            if (lane == 0) {
                const bool excl = mask && mask[i] == 0;
                const double sc = excl ? kInf : ((nn > 0.0) ? dot / sqrt(nn) : kInf * 0.5);
#pragma unroll
                for (int c = 0; c < CL; ++c) r_scores[c][i] = sc;
            }
        }
        cluster.sync();
        for (int i = tid; i < m; i += kClNT) {
            const double v = scores[i];
            int rk = 0;
            for (int i2 = 0; i2 < m; ++i2) {
                const double v2 = scores[i2];
                rk += (v2 < v) || (v2 == v && i2 < i);
            }
            order[rk] = i;
        }
        for (int j = tid; j < NCP; j += kClNT) {
            g[j] = (j < n) ? __ldg(cg_ + j) : 0.0;
            gh[j] = (j < n) ? 1.0 : 0.0;
            cvmap[j] = (j < n) ? -1 : -2;
        }
        __syncthreads();
        CL_T(0);
        int m_eff = m;
        if (mask) {                                    // rows of the reduced LP (uniform over the cluster: same scores everywhere)
            int cntm = 0;
            for (int i = tid; i < m; i += kClNT) cntm += (scores[i] < kInf);
            m_eff = __syncthreads_count(0);            // (barrier) ...
            __shared__ int meff_sm;
            if (tid == 0) meff_sm = 0;
            __syncthreads();
            cntm = __reduce_add_sync(FULL, cntm);
            if (lane == 0) atomicAdd(&meff_sm, cntm);
            __syncthreads();
            m_eff = meff_sm;
        }
        const int nN = m_eff - n;
        bool need_generic = (nN < 0) || (nN > CL * LR);
        int npiv_crash = 0, npiv_p1 = 0, npiv_p2 = 0;
        int status = ST_OPTIMAL;

        // ---- stage 1: Gauss-Jordan on the crash rows [A_B0 | b_B0], crash row t lives in CTA t / LR -------------------------
        int nloc = 0;
        if (!need_generic) {
            nloc = n - rank * LR;
            nloc = nloc < 0 ? 0 : (nloc > LR ? LR : nloc);
            for (int li = warp; li < nloc; li += kClWarps) {
                const int row = order[rank * LR + li];
                for (int j = lane; j < n; j += 32) Tl[(size_t)li * PD + j] = __ldg(Ag + (size_t)row * n + j);
                if (lane == 0) {
                    Tl[(size_t)li * PD + n] = __ldg(bg + row);
                    rowvar[li] = row;
                }
            }
            __syncthreads();
            for (int t = 0; t < n; ++t) {
                const int buf = t & 1, owner = t / LR, lr = t - owner * LR;
                if (rank == owner) {
                    // pivot column: largest |entry| of the row among the still-free columns
                    unsigned long long best = ~0ull;
                    int bj = 0x7fffffff;
                    if (col0 && cj < n && cvmap[cj] == -1) {
                        best = ~(unsigned long long)__double_as_longlong(fabs(Tl[(size_t)lr * PD + cj]));
                        bj = cj;
                    }
                    unsigned long long kmin;
                    int k;
                    block_argmin(best, bj, wred, flip, tid, kmin, k);
                    const double pabs = __longlong_as_double((long long)~kmin);
                    const bool ok = (kmin != ~0ull) && pabs >= kTolCrash;
                    publish_row(lr, buf, 0);
                    if (tid == 0) {
                        const double p = ok ? Tl[(size_t)lr * PD + k] : 0.0;
#pragma unroll
                        for (int c = 0; c < CL; ++c) {
                            r_pub[c][buf].p = p;
                            r_pub[c][buf].k = ok ? k : -1;
                        }
                    }
                }
                CL_T(1);
                cluster.sync();
                CL_T(2);
                const int k = pub[buf].k;
                if (k < 0) { need_generic = true; break; }
                const double p = pub[buf].p;
                const double rp = fast_rcp(p);
                const double* pr = prow + (size_t)(buf * CL) * NCP;
                const double fg = g[k] * rp;
                if (tid < nloc) ebuf[tid] = Tl[(size_t)tid * PD + k];
                __syncthreads();
                for (int j = tid; j < n; j += kClNT) g[j] = (j == k) ? -fg : fma(-fg, pr[j], g[j]);
                if (tid == 0) {
                    cvmap[k] = 0;                     // no longer free
                    pivcol[t] = k;
                }
                update_rows(pr, k, rp, nloc, (rank == owner) ? lr : -1);
                ++npiv_crash;
                __syncthreads();
                CL_T(3);
            }
        }

        if (!need_generic) {
            // D' (row of x_k stored at index k; column n holds the x-vertex) -> the cluster's global scratch
            for (int li = warp; li < nloc; li += kClWarps) {
                const int k = pivcol[rank * LR + li];
                for (int j = lane; j < ncol; j += 32) Dg[(size_t)k * NCP + j] = Tl[(size_t)li * PD + j];
            }
            for (int t = tid; t < n; t += kClNT) {
                const int k = pivcol[t];
                colvar0[k] = order[t];
                cvmap[k] = order[t];
            }
            cluster.sync();                            // D is complete and visible to the whole cluster

            // ---- stage 2: my rows of P_N = -A_N D, s_N = b_N - A_N xv ------------------------------------------------------
            nloc = nN - rank * LR;
            nloc = nloc < 0 ? 0 : (nloc > LR ? LR : nloc);
            for (int li = warp; li < nloc; li += kClWarps) {
                const int row = order[n + rank * LR + li];
                for (int j = lane; j < n; j += 32) Tl[(size_t)li * PD + j] = __ldg(Ag + (size_t)row * n + j);
                if (lane == 0) rowvar[li] = row;
            }
            __syncthreads();
            // thread (rg, cj) accumulates column cj of the rows c0 + rg + u NG (u < kClChunk) of a chunk of kClChunk * NG rows:
            // per k one D entry from L2 (requested four k ahead) and kClChunk broadcast reads of A_N from shared memory
            for (int c0 = 0; c0 < nloc; c0 += kClChunk * NG) {
                double acc[kClChunk];
#pragma unroll
                for (int i = 0; i < kClChunk; ++i) acc[i] = 0.0;
                constexpr int PF = 4;
                double dn[PF];
#pragma unroll
                for (int q = 0; q < PF; ++q) dn[q] = (colact && q < n) ? __ldcg(Dg + (size_t)q * NCP + cj) : 0.0;
                const double* arow = Tl + (size_t)(c0 + rg) * PD;
                for (int k0 = 0; k0 < n; k0 += PF) {
#pragma unroll
                    for (int q = 0; q < PF; ++q) {
                        const int k = k0 + q;
                        const double d = dn[q];
                        dn[q] = (colact && k + PF < n) ? __ldcg(Dg + (size_t)(k + PF) * NCP + cj) : 0.0;
                        if (k < n && colact) {
#pragma unroll
                            for (int i = 0; i < kClChunk; ++i) {
                                const int li = c0 + rg + i * NG;
                                const double av = (li < nloc) ? arow[(size_t)i * NG * PD + k] : 0.0;      // broadcast
                                acc[i] = fma(-av, d, acc[i]);
                            }
                        }
                    }
                }
                __syncthreads();                       // every thread has read the chunk's rows of A_N
                if (colact) {
#pragma unroll
                    for (int i = 0; i < kClChunk; ++i) {
                        const int li = c0 + rg + i * NG;
                        if (li < nloc) Tl[(size_t)li * PD + cj] = acc[i] + ((cj == n) ? __ldg(bg + rowvar[li]) : 0.0);
                    }
                }
            }
            __syncthreads();
            CL_T(4);

            // ---- stage 3a: phase 1 -------------------------------------------------------------------------------------------
            // ONE cluster barrier per pivot: every CTA sends its candidate (most negative slack of its rows) TOGETHER WITH that
            // candidate's row to all CTAs; after the barrier everybody picks the winner and already holds its row.
            int it = 0;                                // candidate / row buffers alternate
            for (;;) {
                const int buf = it & 1;
                {
                    unsigned long long key = KEY_INF;
                    if (tid < nloc) {
                        const double s = Tl[(size_t)tid * PD + n];
                        if (s < -kTolFeas) key = dkey(s);
                    }
                    unsigned long long kmin;
                    int li;
                    block_argmin(key, tid, wred, flip, tid, kmin, li);
                    if (kmin != KEY_INF) publish_row(li, buf, rank);
                    if (tid == 0) {
                        const int var = (kmin != KEY_INF) ? rowvar[li] : -1;
#pragma unroll
                        for (int c = 0; c < CL; ++c) {
                            r_cand[c][buf * 8 + rank].key = kmin;
                            r_cand[c][buf * 8 + rank].row = rank * LR + li;
                            r_cand[c][buf * 8 + rank].var = var;
                        }
                    }
                }
                cluster.sync();
                unsigned long long kmin = KEY_INF;
                int rq = -1, var_r = -1, owner = 0;
#pragma unroll
                for (int c = 0; c < CL; ++c) {
                    const unsigned long long kc = cand[buf * 8 + c].key;
                    if (kc < kmin) { kmin = kc; rq = cand[buf * 8 + c].row; var_r = cand[buf * 8 + c].var; owner = c; }
                }
                if (kmin == KEY_INF) break;            // s >= 0 everywhere: phase 1 finished
                if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int lr = rq - owner * LR;
                const double* pr = prow + (size_t)(buf * CL + owner) * NCP;
                // entering column: ratio test along the row, min ghat_j / (-e_j) over e_j < -tol (every CTA, same arithmetic)
                unsigned long long key = KEY_INF;
                int bj = 0x7fffffff;
                if (col0 && cj < n) {
                    const double e = -pr[cj];
                    if (e > kTolPivot) {
                        key = dkey(fmax(gh[cj], 0.0) / e);
                        bj = cj;
                    }
                }
                int k;
                block_argmin(key, bj, wred, flip, tid, kmin, k);
                if (kmin == KEY_INF) { status = ST_INFEASIBLE; break; }
                const double p = pr[k];
                const double rp = fast_rcp(p);
                const double fv = gh[k] * rp, fg = g[k] * rp;
                const int cv = cvmap[k];
                if (tid < nloc) ebuf[tid] = Tl[(size_t)tid * PD + k];
                __syncthreads();
                if (col0 && cj < n) {
                    gh[cj] = (cj == k) ? -fv : fma(-fv, pr[cj], gh[cj]);
                    g[cj] = (cj == k) ? -fg : fma(-fg, pr[cj], g[cj]);
                }
                if (tid == 0) {
                    cvmap[k] = var_r;                  // becomes nonbasic in column k
                    if (rank == owner) rowvar[lr] = cv;   // becomes basic in the pivot row
                }
                update_rows(pr, k, rp, nloc, (rank == owner) ? lr : -1);
                ++npiv_p1;
                ++it;
                __syncthreads();
            }
            __syncthreads();

            // ---- stage 3b: phase 2 (Dantzig) ---------------------------------------------------------------------------------
            auto price = [&]() -> int {               // most negative reduced cost (every CTA, same arithmetic)
                unsigned long long key = KEY_INF;
                int bj = 0x7fffffff;
                if (col0 && cj < n) {
                    key = dkey(g[cj]);
                    bj = cj;
                }
                unsigned long long kmin;
                int k;
                block_argmin(key, bj, wred, flip, tid, kmin, k);
                return (kmin < dkey(-kTolFeas)) ? k : -1;
            };
            CL_T(5);
            int k = (status == ST_OPTIMAL) ? price() : -1;
            while (status == ST_OPTIMAL && k >= 0) {
                const int buf = it & 1;
                {
                    unsigned long long key = KEY_INF;
                    if (tid < nloc) {
                        const double e = Tl[(size_t)tid * PD + k];
                        ebuf[tid] = e;
                        if (e > kTolPivot) key = dkey(fmax(Tl[(size_t)tid * PD + n], 0.0) / e);
                    }
                    unsigned long long kmin;
                    int li;
                    block_argmin(key, tid, wred, flip, tid, kmin, li);
                    if (kmin != KEY_INF) publish_row(li, buf, rank);       // my candidate's row travels with the candidate
                    if (tid == 0) {
                        const int var = (kmin != KEY_INF) ? rowvar[li] : -1;
#pragma unroll
                        for (int c = 0; c < CL; ++c) {
                            r_cand[c][buf * 8 + rank].key = kmin;
                            r_cand[c][buf * 8 + rank].row = rank * LR + li;
                            r_cand[c][buf * 8 + rank].var = var;
                        }
                    }
                }
                CL_T(6);
                cluster.sync();                        // the pivot's only cluster barrier
                CL_T(7);
                unsigned long long kmin = KEY_INF;
                int rq = -1, var_r = -1, owner = 0;
#pragma unroll
                for (int c = 0; c < CL; ++c) {
                    const unsigned long long kc = cand[buf * 8 + c].key;
                    if (kc < kmin) { kmin = kc; rq = cand[buf * 8 + c].row; var_r = cand[buf * 8 + c].var; owner = c; }
                }
                if (kmin == KEY_INF) { status = ST_UNBOUNDED; break; }
                if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int lr = rq - owner * LR;
                const double* pr = prow + (size_t)(buf * CL + owner) * NCP;
                const double p = pr[k];
                const double rp = fast_rcp(p);
                const double fg = g[k] * rp;
                const int cv = cvmap[k];
                __syncthreads();                       // everyone has read g[k], cvmap[k]
                CL_T(8);
                // reduced costs first, so that the next entering column is priced while the rows are still being updated
                unsigned long long pkey = KEY_INF;
                int pj_ = 0x7fffffff;
                if (col0 && cj < n) {
                    const double gn = (cj == k) ? -fg : fma(-fg, pr[cj], g[cj]);
                    g[cj] = gn;
                    pkey = dkey(gn);
                    pj_ = cj;
                }
                if (tid == 0) {
                    cvmap[k] = var_r;
                    if (rank == owner) rowvar[lr] = cv;
                }
                CL_T(9);
                update_rows(pr, k, rp, nloc, (rank == owner) ? lr : -1);
                CL_T(10);
                int kn;
                block_argmin(pkey, pj_, wred, flip, tid, kmin, kn);        // its barrier also closes the row update
                k = (kmin < dkey(-kTolFeas)) ? kn : -1;
                ++npiv_p2;
                ++it;
                CL_T(11);
            }
        }

        // ---- stage 4: x, objective, slacks, labels ------------------------------------------------------------------------------
        __syncthreads();
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0, nref = 0;
        if (!need_generic && status == ST_OPTIMAL) {
            // which constraints are basic, and at what slack: through the cluster's global scratch
            for (int i = rank * kClNT + tid; i < m; i += CL * kClNT) flagg[i] = 0;
            cluster.sync();
            if (tid < nloc) {
                const int q = rowvar[tid];
                if (q >= 0) {
                    valg[q] = Tl[(size_t)tid * PD + n];
                    flagg[q] = 1;
                }
            }
            cluster.sync();
            for (int j = tid; j < n; j += kClNT) {
                const int q0 = colvar0[j];
                sig[j] = __ldcg(flagg + q0) ? __ldcg(valg + q0) : 0.0;
            }
            __syncthreads();
            for (int kx = gwarp; kx < n; kx += gnw) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(__ldcg(Dg + (size_t)kx * NCP + j), sig[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) xg[kx] = __ldcg(Dg + (size_t)kx * NCP + n) - acc;
            }
            cluster.sync();
            for (int j = tid; j < n; j += kClNT) xs[j] = __ldcg(xg + j);
            __syncthreads();
            for (int i = gwarp; i < m; i += gnw) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(__ldg(Ag + (size_t)i * n + j), xs[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) {
                    const double slack = __ldg(bg + i) - acc;
                    const double as = fabs(slack);
                    const int active = as <= a.thr;
                    const int nonbasic = (__ldcg(flagg + i) == 0);
                    const bool excl = mask && mask[i] == 0;
                    lab[i] = (uint8_t)active;
                    nact += active;
                    nties += ((as >= a.thr * 0.1 && as <= a.thr * 10.0) || (!excl && active != nonbasic)) ? 1 : 0;
                    nviol += (slack < -a.thr);
                    nref += (!excl && nonbasic && as > a.thr * 0.01);
                }
            }
            if (rank == 0) {
                if (warp == 0) {
                    double acc = 0.0;
                    for (int j = lane; j < n; j += 32) acc = fma(__ldg(cg_ + j), xs[j], acc);
                    acc = warp_sum(acc);
                    if (lane == 0 && a.obj) a.obj[lp] = acc;
                }
                if (a.x)
                    for (int j = tid; j < n; j += kClNT) a.x[(size_t)lp * n + j] = xs[j];
            }
        } else if (!need_generic) {
            for (int i = rank * kClNT + tid; i < m; i += CL * kClNT) lab[i] = 0;
            if (rank == 0) {
                if (a.x)
                    for (int j = tid; j < n; j += kClNT) a.x[(size_t)lp * n + j] = 0.0;
                if (tid == 0 && a.obj) a.obj[lp] = __longlong_as_double(0x7ff8000000000000ll);
            }
        }
        // label counts: lanes 0 of all warps -> rank 0
        {
            nact = __reduce_add_sync(FULL, nact);
            nties = __reduce_add_sync(FULL, nties);
            nviol = __reduce_add_sync(FULL, nviol);
            nref = __reduce_add_sync(FULL, nref);
            __shared__ int wcnt[kClWarps][4];
            if (lane == 0) { wcnt[warp][0] = nact; wcnt[warp][1] = nties; wcnt[warp][2] = nviol; wcnt[warp][3] = nref; }
            __syncthreads();
            if (tid < 4) {
                int s = 0;
                for (int w = 0; w < kClWarps; ++w) s += wcnt[w][tid];
                cluster.map_shared_rank(cnt, 0)[rank * 4 + tid] = s;
            }
        }
        cluster.sync();
        if (rank == 0 && tid == 0) {
            int t0 = 0, t1 = 0, t2 = 0, t3 = 0;
            for (int c = 0; c < CL; ++c) { t0 += cnt[c * 4]; t1 += cnt[c * 4 + 1]; t2 += cnt[c * 4 + 2]; t3 += cnt[c * 4 + 3]; }
            if (need_generic || (status == ST_OPTIMAL && t3 > 0)) status = -1;   // singular crash basis / ill-conditioned vertex: generic kernel
            a.status[lp] = status;
            if (status == -1) {
                atomicAdd(a.flag_count, 1);
            } else {
                if (a.n_active) a.n_active[lp] = t0;
                if (a.ties) a.ties[lp] = t1;
                if (a.violations) a.violations[lp] = t2;
                if (a.pivots) {
                    int* pv = a.pivots + (size_t)lp * 4;
                    pv[0] = npiv_crash;
                    pv[1] = npiv_p1;
                    pv[2] = npiv_p2;
                    pv[3] = npiv_crash + npiv_p1 + npiv_p2;
                }
            }
        }
#ifdef DDB_CL_TIMING
        CL_T(13);
        if (tid == 0 && rank == 0) {
            for (int q = 0; q < 14; ++q) atomicAdd(a.dscr + q, tacc[q]);
            atomicAdd(a.dscr + 14, (double)npiv_crash); atomicAdd(a.dscr + 15, (double)npiv_p1); atomicAdd(a.dscr + 16, (double)npiv_p2);
        }
#endif
        // (the next cluster.sync -- the LP fetch -- separates this LP's reads of the shared buffers from the next LP's writes)
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kClSmemBudget = 232448 - 2048;

template <int CL>
cudaError_t launch_cluster_variant(const SolveArgs& a, int sm_count, cudaStream_t st) {
    auto kern = simplex_cluster_kernel<CL>;
    const size_t smem = cl_layout(a.m, a.n, CL).total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long clusters = sm_count / CL;
    if (clusters > a.B) clusters = a.B;
    if (clusters < 1) clusters = 1;
    kern<<<(int)(clusters * CL), kClNT, smem, st>>>(a);
    return cudaGetLastError();
}

// Can the device co-schedule clusters of CL CTAs of this kernel with `smem` bytes each?  (Non-power-of-two sizes are legal
// but depend on the SM count per GPC; asked once per size from the occupancy API.)
template <int CL>
bool cluster_size_schedulable(size_t smem) {
    auto kern = simplex_cluster_kernel<CL>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, 1, 1);
    cfg.blockDim = dim3(kClNT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    return nclusters > 0;
}

bool size_ok(int cl, size_t smem) {
    static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // 0 unknown, 1 yes, -1 no (probed with the largest footprint)
    if (cl == 1 || cl == 2 || cl == 4 || cl == 8) return true;
    if (cache[cl] == 0) {
        bool ok = false;
        switch (cl) {
            case 3: ok = cluster_size_schedulable<3>(kClSmemBudget); break;
            case 5: ok = cluster_size_schedulable<5>(kClSmemBudget); break;
            case 6: ok = cluster_size_schedulable<6>(kClSmemBudget); break;
            case 7: ok = cluster_size_schedulable<7>(kClSmemBudget); break;
        }
        cache[cl] = ok ? 1 : -1;
    }
    (void)smem;
    return cache[cl] > 0;
}

// smallest cluster whose CTAs hold their share of the tableau: throughput goes with 1 / (cluster size x time per LP), and the
// time per LP falls more slowly than 1 / size (the per-pivot barrier and candidate exchange do not shrink)
int cluster_size_for(int m, int n) {
    if (m < n || n + 1 > kClNT) return 0;
    static const bool pow2_only = [] { const char* e = getenv("DDB_CLUSTER_POW2"); return e && e[0] == '1'; }();
    for (int cl = 1; cl <= 8; ++cl) {
        if (pow2_only && (cl & (cl - 1))) continue;
        const ClLayout L = cl_layout(m, n, cl);
        if (L.total <= kClSmemBudget && L.LR <= kClNT && size_ok(cl, L.total)) return cl;
    }
    return 0;
}
}  // namespace

bool cluster_supported(int m, int n) { return cluster_size_for(m, n) > 0; }

int cluster_count(int m, int n, int sm_count, long long B) {
    const int cl = cluster_size_for(m, n);
    if (!cl) return 0;
    long long c = sm_count / cl;
    if (c > B) c = B;
    return (int)(c < 1 ? 1 : c);
}

size_t cluster_scratch_bytes(int m, int n, int sm_count, long long B) {
    const int cl = cluster_size_for(m, n);
    if (!cl) return 0;
    const int ncp = (n + 2) & ~1;
    const size_t per_cluster = (size_t)n * ncp + n + m + (size_t)((m + 1) / 2);
    return (size_t)cluster_count(m, n, sm_count, B) * per_cluster * sizeof(double);
}

cudaError_t launch_simplex_cluster(const SolveArgs& a, int sm_count, cudaStream_t st) {
    switch (cluster_size_for(a.m, a.n)) {
        case 1: return launch_cluster_variant<1>(a, sm_count, st);
        case 2: return launch_cluster_variant<2>(a, sm_count, st);
        case 3: return launch_cluster_variant<3>(a, sm_count, st);
        case 4: return launch_cluster_variant<4>(a, sm_count, st);
        case 5: return launch_cluster_variant<5>(a, sm_count, st);
        case 6: return launch_cluster_variant<6>(a, sm_count, st);
        case 7: return launch_cluster_variant<7>(a, sm_count, st);
        case 8: return launch_cluster_variant<8>(a, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

int cluster_size(int m, int n) { return cluster_size_for(m, n); }

}  // namespace ddb
