// 2-D register-tiled batched fp64 simplex (plan 0): one LP per CTA, the condensed tableau lives in the REGISTER
// FILE as an R x C block per thread.
//
// Why a 2-D block per thread.  A pivot is a rank-1 update  T[i][j] -= f_i * q_j.  On B200 an fp64 FMA issues every
// ~1.6 clk per SM sub-partition, but delivering ONE double from shared memory to the 32 lanes of a warp costs 4 clk
// of the load/store pipe even when all lanes read the same address (measured: tools/rank1_bench.cu -- the
// register write-back moves 64 B/clk per sub-partition).  A thread that owns a whole row (C = n + 1, R = 1) needs
// C operands for C FMAs and is 2.5x load-bound; a thread that owns an R x C block needs R + C operands for R * C
// FMAs (R = 5, C = 13: 18 loads for 65 FMAs), so the update runs at the speed of the fp64 pipe.
//
// Layout.  lane = (lr, lc) with lr = lane / 8 (4 row groups), lc = lane % 8 (8 column groups).
//   tile row    i = (warp * 4 + lr) * R + rs      rs < R  : register row slot
//   tile column j = lc * C + cs                   cs < C  : register column slot;  the last column is the rhs
// Rows and the cost vectors travel through shared memory in the same column-group format (8 groups of CP = C
// rounded up to even doubles, so every thread fetches its C operands with 16-byte loads, conflict-free).
//   * entering column k lives in the lanes with lc == k / C of every warp: a C-way uniform switch extracts their R
//     entries into shared memory (fcol); the ratio test is then one row per LANE (+ a redux-based warp argmin);
//   * the pivot row r lives in the 8 lanes (lr == ..) of one warp: an R-way uniform switch publishes it with 7
//     16-byte stores per lane; that warp also updates the cost vector(s) and prices the next entering column;
//   * everything else is R * C FMAs per thread with R + C operands from shared memory.
//
// Stages per LP (same algorithm and tolerances as simplex_generic.cu):
//   0. crash order by cosine score (A streamed once from HBM through the idle tile registers)
//   1. crash as an explicit inverse: Gauss-Jordan on the n x n block A_B0, 1 barrier / pivot
//   2. remaining rows enter through  P_N = -A_N D  (register-blocked: R + C operands per R * C FMAs)
//   3. phase 1 (most negative slack leaves, ratio test along the published row), phase 2 (Dantzig), 2 barriers / pivot
//   4. x = xv - D sigma, one step of iterative refinement on the final active set, slack = b - A x from the
//      caller's A, labels = |slack| <= threshold
// Instances the tile cannot hold or whose static crash basis is singular are flagged status = -1 and re-solved by
// the generic kernel on the device (capi.cu); nothing ever falls back to the CPU.
#include <cstdlib>
#include <type_traits>

#include "../common.cuh"

namespace ddb {

// Optional per-stage cycle accounting (debug builds only: -DDDB_TIMING; results in a.gtab, 16 doubles per LP).
#ifdef DDB_TIMING
#define TSTAGE(i)                                   \
    do {                                            \
        const long long _t = clock64();             \
        tacc[i] += _t - tlast;                      \
        tlast = _t;                                 \
    } while (0)
#else
#define TSTAGE(i) do { } while (0)
#endif

// Warp-uniform dynamic slot -> compile-time slot.  The bodies contain stores / asm volatile, so the compiler keeps
// real (uniform) branches instead of if-converting into selects over every register.
template <int N, class F>
__device__ __forceinline__ void slot_switch(int i, F&& f) {
    static_assert(N <= 16, "slot_switch covers 16 slots");
    switch (i) {
#define DDB_SLOT(I) \
    case I:         \
        if constexpr (I < N) f(std::integral_constant<int, I>{}); \
        break;
        DDB_SLOT(0) DDB_SLOT(1) DDB_SLOT(2) DDB_SLOT(3) DDB_SLOT(4) DDB_SLOT(5) DDB_SLOT(6) DDB_SLOT(7)
        DDB_SLOT(8) DDB_SLOT(9) DDB_SLOT(10) DDB_SLOT(11) DDB_SLOT(12) DDB_SLOT(13) DDB_SLOT(14) DDB_SLOT(15)
#undef DDB_SLOT
        default: break;
    }
}

struct TilePub {              // written by the pivot row's warp, read by everybody after the barrier
    double p;                 // pivot entry
    int k;                    // crash / phase 1: entering column (-1: none); phase 2: NEXT entering column (-1: optimal)
    int pad;
};

template <int W, int R, int C>
struct TileCfg {
    static constexpr int NT = W * 32;
    static constexpr int RW = 4 * R;                 // tile rows per warp
    static constexpr int RT = W * RW;                // tile rows
    static constexpr int CT = 8 * C;                 // tile columns (the last one is the right-hand side)
    static constexpr int CP = (C + 1) & ~1;          // column-group pitch in shared memory (doubles)
    static constexpr int PR = 8 * CP;                // pitch of a row in shared memory (doubles)
    static constexpr int RP = (R + 1) & ~1;          // row-group pitch of the per-row vectors (doubles)
    static constexpr int VT = W * 4 * RP;            // length of a per-row vector in shared memory
    static constexpr int QS = (CT + 31) / 32;        // lane-distributed slots over the columns
};

struct TileLayout {
    size_t D, order, colvar0, pivcol, basic_tile, cvsm, rowvar;            // persistent per LP
    size_t prow, pub, keys, crow, gsm, ghsm, fcol, svec;                   // pivot loops
    size_t gbuf, gnn, sig, xbuf, dsig, red;                                // stage 0 / 4 scratch (aliases the above)
    size_t total;
};
__host__ __device__ inline size_t tl_align(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline TileLayout make_tile_layout(int m, int n, int W, int R, int C) {
    TileLayout L;
    const int CP = (C + 1) & ~1, PR = 8 * CP, RP = (R + 1) & ~1, VT = W * 4 * RP, CT = 8 * C;
    size_t off = 0;
    L.D = off;          off += tl_align((size_t)n * PR * 8);
    L.order = off;      off += tl_align((size_t)m * 4);
    L.colvar0 = off;    off += tl_align((size_t)n * 4);
    L.pivcol = off;     off += tl_align((size_t)n * 4);
    L.basic_tile = off; off += tl_align((size_t)m * 4);
    L.cvsm = off;       off += tl_align((size_t)CT * 4);
    L.rowvar = off;     off += tl_align((size_t)W * 4 * R * 4);
    L.svec = off;       off += tl_align((size_t)VT * 8);
    const size_t u0 = off;
    L.prow = off;       off += tl_align((size_t)2 * PR * 8);
    L.pub = off;        off += tl_align((size_t)2 * sizeof(TilePub));
    L.keys = off;       off += tl_align((size_t)W * 8);
    L.crow = off;       off += tl_align((size_t)W * 4);
    L.gsm = off;        off += tl_align((size_t)PR * 8);
    L.ghsm = off;       off += tl_align((size_t)PR * 8);
    L.fcol = off;       off += tl_align((size_t)VT * 8);
    const size_t u1 = off;
    off = u0;
    L.gbuf = off;       off += tl_align((size_t)m * 8);
    L.gnn = off;        off += tl_align((size_t)m * 8);
    L.sig = off;        off += tl_align((size_t)PR * 8);
    L.xbuf = off;       off += tl_align((size_t)(n > 128 ? n : 128) * 8);
    L.dsig = off;       off += tl_align((size_t)PR * 8);
    L.red = off;        off += tl_align((size_t)(3 * W + 4) * 4);
    L.total = off > u1 ? off : u1;
    return L;
}

template <int W, int R, int C, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) simplex_tile2d_kernel(SolveArgs a) {
    using Cfg = TileCfg<W, R, C>;
    constexpr int NT = Cfg::NT, RW = Cfg::RW, RT = Cfg::RT, CT = Cfg::CT, CP = Cfg::CP, PR = Cfg::PR;
    constexpr int RP = Cfg::RP, QS = Cfg::QS;
    constexpr int RHS = CT - 1;               // tile column of the right-hand side: (lc = 7, cs = C - 1)
    static_assert(RW <= 32, "one tile row per lane in the ratio tests");
    static_assert(W <= 32, "one header per lane");
    static_assert((R * C) / QS >= 1, "tile too small to stream A through it");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const TileLayout L = make_tile_layout(m, n, W, R, C);
    double* Dsm = reinterpret_cast<double*>(smem_raw + L.D);
    int* order = reinterpret_cast<int*>(smem_raw + L.order);
    int* colvar0 = reinterpret_cast<int*>(smem_raw + L.colvar0);
    int* pivcol = reinterpret_cast<int*>(smem_raw + L.pivcol);
    int* basic_tile = reinterpret_cast<int*>(smem_raw + L.basic_tile);
    int* cvsm = reinterpret_cast<int*>(smem_raw + L.cvsm);          // column -> constraint whose slack is nonbasic
    int* rowvar = reinterpret_cast<int*>(smem_raw + L.rowvar);      // tile row -> constraint whose slack is basic
    double* svec = reinterpret_cast<double*>(smem_raw + L.svec);    // right-hand side per tile row (row-group format)
    double* prow = reinterpret_cast<double*>(smem_raw + L.prow);    // [2][PR] published pivot row (raw)
    TilePub* pub = reinterpret_cast<TilePub*>(smem_raw + L.pub);    // [2]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + L.keys);   // [W]
    int* crow = reinterpret_cast<int*>(smem_raw + L.crow);          // [W] candidate tile row of each warp
    double* gsm = reinterpret_cast<double*>(smem_raw + L.gsm);      // g: true reduced costs (column-group format)
    double* ghsm = reinterpret_cast<double*>(smem_raw + L.ghsm);    // ghat: artificial costs of phase 1
    double* fcol = reinterpret_cast<double*>(smem_raw + L.fcol);    // entering column per tile row (row-group format)
    double* gbuf = reinterpret_cast<double*>(smem_raw + L.gbuf);
    double* gnn = reinterpret_cast<double*>(smem_raw + L.gnn);
    double* sig = reinterpret_cast<double*>(smem_raw + L.sig);
    double* xbuf = reinterpret_cast<double*>(smem_raw + L.xbuf);
    double* dsig = reinterpret_cast<double*>(smem_raw + L.dsig);
    int* red = reinterpret_cast<int*>(smem_raw + L.red);
    __shared__ long long cur_lp;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 3, lc = lane & 7;
    const int row0 = (warp * 4 + lr) * R;                 // my first tile row
    const int vrow0 = (warp * 4 + lr) * RP;               // its position in the per-row vectors
    const int col0 = lc * C;                              // my first tile column
    const int fpos0 = lc * CP;                            // its position in a shared-memory row

    // position of tile column j in a shared-memory row / of tile row i in a per-row vector
    auto cpos = [&](int j) { return (j / C) * CP + (j % C); };
    auto rpos = [&](int i) { return (i / R) * RP + (i % R); };

    double T[R * C];          // my block: T[rs * C + cs]

    // ---- building blocks of a pivot -------------------------------------------------------------------------------
    // lanes (lr == lr_r) of the pivot row's warp: write row slot rs of my block to a shared-memory row
    auto publish_row = [&](int rs, double* dst) {
        slot_switch<R>(rs, [&](auto Rc) {
            constexpr int RS = decltype(Rc)::value;
            double2* d2 = reinterpret_cast<double2*>(dst + fpos0);
#pragma unroll
            for (int c2 = 0; c2 < C / 2; ++c2) d2[c2] = make_double2(T[RS * C + 2 * c2], T[RS * C + 2 * c2 + 1]);
            if constexpr (C & 1) d2[C / 2] = make_double2(T[RS * C + C - 1], 0.0);
        });
    };
    // lanes (lc == lc_k): my R entries of column slot cs -> e[], and into fcol (raw)
    auto extract_col = [&](int cs, double (&e)[R]) {
        slot_switch<C>(cs, [&](auto Cc) {
            constexpr int CS = decltype(Cc)::value;
#pragma unroll
            for (int rs = 0; rs < R; ++rs) {
                e[rs] = T[rs * C + CS];
                fcol[vrow0 + rs] = e[rs];
            }
        });
    };
    // lanes (lc == lc_k): overwrite column slot cs with v[]
    auto write_col = [&](int cs, const double (&v)[R]) {
        slot_switch<C>(cs, [&](auto Cc) {
            constexpr int CS = decltype(Cc)::value;
#pragma unroll
            for (int rs = 0; rs < R; ++rs) asm volatile("mov.f64 %0, %1;" : "=d"(T[rs * C + CS]) : "d"(v[rs]));
        });
    };
    // lanes (lr == lr_r) of the pivot row's warp: scale row slot rs by rp
    auto scale_row = [&](int rs, double rp) {
        slot_switch<R>(rs, [&](auto Rc) {
            constexpr int RS = decltype(Rc)::value;
#pragma unroll
            for (int cs = 0; cs < C; ++cs) asm volatile("mul.f64 %0, %0, %1;" : "+d"(T[RS * C + cs]) : "d"(rp));
        });
    };
    // T[rs][cs] -= f[rs] * q[cs] with q = my C entries of the shared-memory row `src`
    auto rank1 = [&](const double* src, const double (&f)[R]) {
        const double2* s2 = reinterpret_cast<const double2*>(src + fpos0);
#pragma unroll
        for (int c2 = 0; c2 < (C + 1) / 2; ++c2) {
            const double2 q = s2[c2];
#pragma unroll
            for (int rs = 0; rs < R; ++rs) {
                T[rs * C + 2 * c2] = fma(-f[rs], q.x, T[rs * C + 2 * c2]);
                if (2 * c2 + 1 < C) T[rs * C + 2 * c2 + 1] = fma(-f[rs], q.y, T[rs * C + 2 * c2 + 1]);
            }
        }
    };
    // The common tail of every pivot: the raw pivot row r is in `src`, the entering column is k, the pivot p.
    //   have_e: the entering column was already extracted into e[] / fcol by the caller (phase 2)
    auto apply_pivot = [&](const double* src, int r, int k, double p, bool have_e, double (&e)[R], int nrows) {
        const double rp = fast_rcp(p);
        const int lck = k / C, csk = k - lck * C;
        const bool colk = (lc == lck);
        if (!have_e) {
            if (colk) extract_col(csk, e);
            __syncwarp();
        }
        double f[R];
        {
            const double2* f2 = reinterpret_cast<const double2*>(fcol + vrow0);
#pragma unroll
            for (int r2 = 0; r2 < (R + 1) / 2; ++r2) {
                const double2 v = f2[r2];
                f[2 * r2] = v.x * rp;
                if (2 * r2 + 1 < R) f[2 * r2 + 1] = v.y * rp;
            }
#pragma unroll
            for (int rs = 0; rs < R; ++rs)
                if (row0 + rs == r || row0 + rs >= nrows) f[rs] = 0.0;      // the pivot row is not updated, it is scaled
        }
        rank1(src, f);
        const int wr = r / RW, lrr = (r - wr * RW) / R, rsr = r - wr * RW - lrr * R;
        if (warp == wr && lr == lrr) scale_row(rsr, rp);
        if (colk) {
            double v[R];
#pragma unroll
            for (int rs = 0; rs < R; ++rs) v[rs] = (row0 + rs == r) ? rp : ((row0 + rs < nrows) ? -e[rs] * rp : 0.0);
            write_col(csk, v);
        }
    };
    // lanes lc == 7: publish the right-hand sides of my rows
    auto publish_rhs = [&]() {
        if (lc == 7) {
#pragma unroll
            for (int rs = 0; rs < R; ++rs) svec[vrow0 + rs] = T[rs * C + C - 1];
        }
    };
    // T used as a streaming buffer: dot products of the rows of A with a lane-distributed vector (4 slots of 32
    // columns per row); all loads of a batch are in flight together.  out1[i] = a_i . v ; out2[i] = a_i . a_i
    auto row_dots = [&](const double* Ag, const double (&vl)[QS], double* out1, double* out2) {
        constexpr int NBQ = (R * C) / QS;
        for (int base = 0; base < m; base += NBQ * W) {
#pragma unroll
            for (int rr = 0; rr < NBQ; ++rr) {
                const int i = base + rr * W + warp;
#pragma unroll
                for (int qs = 0; qs < QS; ++qs) {
                    const int j = lane + 32 * qs;
                    T[rr * QS + qs] = (i < m && j < n) ? __ldg(Ag + (size_t)i * n + j) : 0.0;
                }
            }
#pragma unroll
            for (int rr = 0; rr < NBQ; ++rr) {
                const int i = base + rr * W + warp;
                double dot = 0.0, nn = 0.0;
#pragma unroll
                for (int qs = 0; qs < QS; ++qs) {
                    const double v = T[rr * QS + qs];
                    dot = fma(v, vl[qs], dot);
                    nn = fma(v, v, nn);
                }
                dot = warp_sum(dot);
                if (out2) nn = warp_sum(nn);
                if (lane == 0 && i < m) {
                    out1[i] = dot;
                    if (out2) out2[i] = nn;
                }
            }
        }
    };

    for (;;) {
        if (tid == 0) cur_lp = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long lp = cur_lp;
        if (lp >= a.B) break;
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg = a.c + (size_t)lp * n;
        const uint8_t* mask = a.row_mask ? a.row_mask + (size_t)lp * m : nullptr;
#ifdef DDB_TIMING
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = clock64();
#endif

        // ---- stage 0: crash order ---------------------------------------------------------------------------
        {
            double cl[QS];
#pragma unroll
            for (int qs = 0; qs < QS; ++qs) {
                const int j = lane + 32 * qs;
                cl[qs] = (j < n) ? __ldg(cg + j) : 0.0;
            }
            row_dots(Ag, cl, gbuf, gnn);
        }
        __syncthreads();
        for (int i = tid; i < m; i += NT) {
            const bool excl = mask && mask[i] == 0;
            const double dot = gbuf[i], nn = gnn[i];
            gnn[i] = excl ? kInf : (nn > 0.0 ? dot / sqrt(nn) : kInf * 0.5);
        }
        __syncthreads();
        for (int i = tid; i < m; i += NT) {
            const double v = gnn[i];
            int rank = 0;
            for (int i2 = 0; i2 < m; ++i2) {
                const double v2 = gnn[i2];
                rank += (v2 < v) || (v2 == v && i2 < i);
            }
            order[rank] = i;
            basic_tile[i] = -1;
        }
        int m_eff = m;
        if (mask) {
            m_eff = 0;
            for (int i = 0; i < m; ++i) m_eff += (gnn[i] < kInf);   // uniform, only for reduced LPs
        }
        __syncthreads();
        const int nN = m_eff - n;
        bool need_generic = (nN < 0) || (nN > RT) || (n > RT) || (n > CT - 1);

        int npiv_crash = 0, npiv_p1 = 0, npiv_p2 = 0;
        int status = ST_OPTIMAL;
        double e[R];
        TSTAGE(0);                                             // stage 0: scores + ranking

        if (!need_generic) {
            // ---- stage 1: tile row t < n  <-  row order[t] of [A | b]; Gauss-Jordan to the inverse -------------------
#pragma unroll
            for (int rs = 0; rs < R; ++rs) {
                const int i = row0 + rs;
                const bool have = i < n;
                const int row = have ? order[i] : 0;
#pragma unroll
                for (int cs = 0; cs < C; ++cs) {
                    const int j = col0 + cs;
                    double v = 0.0;
                    if (have && j < n) v = __ldg(Ag + (size_t)row * n + j);
                    if (have && j == RHS) v = __ldg(bg + row);
                    T[rs * C + cs] = v;
                }
                e[rs] = 0.0;
            }
            for (int j = tid; j < CT; j += NT) {
                gsm[cpos(j)] = (j < n) ? __ldg(cg + j) : 0.0;
                cvsm[j] = (j < n) ? -1 : -2;                     // -1: still a free x_j (crash), -2: not a column
            }
            __syncthreads();

            for (int t = 0; t < n; ++t) {
                const int nb = t & 1;
                double* pr = prow + nb * PR;
                const int wt = t / RW, lrt = (t - wt * RW) / R, rst = t - wt * RW - lrt * R;
                if (warp == wt) {
                    // the 8 lanes that hold tile row t publish it (raw); the warp picks the pivot column (largest
                    // |entry| among the free columns) and updates the cost row
                    if (lr == lrt) publish_row(rst, pr);
                    __syncwarp();
                    double rv[QS], gq[QS];
                    unsigned long long best = 0ull;
                    int bq = 0;
#pragma unroll
                    for (int qs = 0; qs < QS; ++qs) {
                        const int j = lane + 32 * qs;
                        const int jp = cpos(j);
                        rv[qs] = (j < n) ? pr[jp] : 0.0;
                        gq[qs] = (j < n) ? gsm[jp] : 0.0;
                        const unsigned long long kk =
                            (j < n && cvsm[j] == -1) ? (unsigned long long)__double_as_longlong(fabs(rv[qs])) : 0ull;
                        if (kk > best) { best = kk; bq = qs; }
                    }
                    unsigned long long kmin;
                    const int kl = warp_argmin_key(~best, kmin);   // argmax through the complemented key
                    const int k = kl + 32 * __shfl_sync(FULL, bq, kl);
                    const double pabs = __longlong_as_double((long long)~kmin);
                    const double p = pr[cpos(k)];
                    const double fg = gsm[cpos(k)] * fast_rcp(p);
                    __syncwarp();
#pragma unroll
                    for (int qs = 0; qs < QS; ++qs) {
                        const int j = lane + 32 * qs;
                        if (j < n) gsm[cpos(j)] = (j == k) ? -fg : fma(-fg, rv[qs], gq[qs]);
                    }
                    if (lane == 0) {
                        pub[nb].p = p;
                        pub[nb].k = (pabs >= kTolCrash) ? k : -1;
                        pivcol[t] = k;
                        cvsm[k] = 0;                  // no longer free
                    }
                }
                __syncthreads();
                const int k = pub[nb].k;
                if (k < 0) { need_generic = true; break; }
                apply_pivot(pr, t, k, pub[nb].p, false, e, n);
                ++npiv_crash;
            }
        }

        TSTAGE(1);                                             // crash
        if (!need_generic) {
            __syncthreads();
            // dump D' (row of x_k stored at index k; the rhs slot holds the x-vertex) and the column -> constraint map
#pragma unroll
            for (int rs = 0; rs < R; ++rs) {
                const int i = row0 + rs;
                if (i < n) {
                    const int k = pivcol[i];
                    double2* d2 = reinterpret_cast<double2*>(Dsm + (size_t)k * PR + fpos0);
#pragma unroll
                    for (int c2 = 0; c2 < C / 2; ++c2) d2[c2] = make_double2(T[rs * C + 2 * c2], T[rs * C + 2 * c2 + 1]);
                    if constexpr (C & 1) d2[C / 2] = make_double2(T[rs * C + C - 1], 0.0);
                    if (lc == 0) {
                        colvar0[k] = order[i];
                        cvsm[k] = order[i];
                    }
                }
            }
            for (int j = tid; j < CT; j += NT) ghsm[cpos(j)] = (j < n) ? 1.0 : 0.0;
            __syncthreads();

            // ---- stage 2: my block of P_N = -A_N D, s_N = b_N - A_N xv -----------------------------------------
            {
                int arow[R];
#pragma unroll
                for (int rs = 0; rs < R; ++rs) {
                    const int u = row0 + rs;
                    arow[rs] = (u < nN) ? order[n + u] : -1;
                    if (lc == 0 && u < RT) rowvar[u] = arow[rs];
                }
#pragma unroll
                for (int q = 0; q < R * C; ++q) T[q] = 0.0;
                if (warp * RW < nN) {
                    // A_N entries of my rows, two columns per step, prefetched one step ahead (they come from L2)
                    double an[R][2];
#pragma unroll
                    for (int rs = 0; rs < R; ++rs) {
                        an[rs][0] = (arow[rs] >= 0) ? __ldg(Ag + (size_t)arow[rs] * n) : 0.0;
                        an[rs][1] = (arow[rs] >= 0 && 1 < n) ? __ldg(Ag + (size_t)arow[rs] * n + 1) : 0.0;
                    }
                    for (int k0 = 0; k0 < n; k0 += 2) {
                        double av[R][2];
#pragma unroll
                        for (int rs = 0; rs < R; ++rs) {
                            av[rs][0] = an[rs][0];
                            av[rs][1] = an[rs][1];
                            an[rs][0] = (arow[rs] >= 0 && k0 + 2 < n) ? __ldg(Ag + (size_t)arow[rs] * n + k0 + 2) : 0.0;
                            an[rs][1] = (arow[rs] >= 0 && k0 + 3 < n) ? __ldg(Ag + (size_t)arow[rs] * n + k0 + 3) : 0.0;
                        }
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            if (k0 + kk < n) {
                                double f[R];
#pragma unroll
                                for (int rs = 0; rs < R; ++rs) f[rs] = av[rs][kk];
                                rank1(Dsm + (size_t)(k0 + kk) * PR, f);
                            }
                        }
                    }
                }
                if (lc == 7) {
#pragma unroll
                    for (int rs = 0; rs < R; ++rs)
                        if (arow[rs] >= 0) T[rs * C + C - 1] += __ldg(bg + arow[rs]);
                }
                publish_rhs();
            }
            __syncthreads();
            TSTAGE(2);                                         // dump + GEMM

            // ---- stage 3a: phase 1 (most negative slack leaves; ratio test along its row) -------------------
            int nb = 0;
            for (;;) {
                {
                    const int u = warp * RW + lane;
                    const double s = (lane < RW && u < nN) ? svec[rpos(u)] : 0.0;
                    unsigned long long kmin;
                    const int ll = warp_argmin_key((s < -kTolFeas) ? dkey(s) : KEY_INF, kmin);
                    if (lane == 0) {
                        keys[warp] = kmin;
                        crow[warp] = warp * RW + ll;
                    }
                }
                __syncthreads();
                unsigned long long kmin;
                const int ww = warp_argmin_key((lane < W) ? keys[lane] : KEY_INF, kmin);
                if (kmin == KEY_INF) break;                       // s >= 0 everywhere: phase 1 finished
                if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int r = crow[ww];
                double* pr = prow + nb * PR;
                const int wr = r / RW, lrr = (r - wr * RW) / R, rsr = r - wr * RW - lrr * R;
                if (warp == wr) {
                    if (lr == lrr) publish_row(rsr, pr);
                    __syncwarp();
                    // ratio test along the row: min ghat_j / (-e_j) over e_j < -tol; then ghat, g, column map
                    double rv[QS], gh[QS], gq[QS];
                    double bn = 0.0, bd = 0.0;   // best numerator / denominator (bd == 0: none)
                    int bq = 0;
#pragma unroll
                    for (int qs = 0; qs < QS; ++qs) {
                        const int j = lane + 32 * qs;
                        const int jp = cpos(j);
                        rv[qs] = (j < n) ? pr[jp] : 0.0;
                        gh[qs] = (j < n) ? ghsm[jp] : 0.0;
                        gq[qs] = (j < n) ? gsm[jp] : 0.0;
                        const double ej = -rv[qs];
                        if (ej > kTolPivot) {
                            const double num = fmax(gh[qs], 0.0);
                            if (bd == 0.0 || num * bd < bn * ej) { bn = num; bd = ej; bq = qs; }
                        }
                    }
                    const double ratio = bn * fast_rcp(bd > 0.0 ? bd : 1.0);
                    unsigned long long kmin2;
                    const int kl = warp_argmin_key((bd > 0.0) ? dkey(ratio) : KEY_INF, kmin2);
                    const bool none = (kmin2 == KEY_INF);
                    const int k = none ? 0 : kl + 32 * __shfl_sync(FULL, bq, kl);
                    const double p = none ? 1.0 : pr[cpos(k)];
                    const double rp = fast_rcp(p);
                    const double fv = ghsm[cpos(k)] * rp, fg = gsm[cpos(k)] * rp;
                    __syncwarp();
                    if (!none) {
#pragma unroll
                        for (int qs = 0; qs < QS; ++qs) {
                            const int j = lane + 32 * qs;
                            if (j < n) {
                                ghsm[cpos(j)] = (j == k) ? -fv : fma(-fv, rv[qs], gh[qs]);
                                gsm[cpos(j)] = (j == k) ? -fg : fma(-fg, rv[qs], gq[qs]);
                            }
                        }
                    }
                    if (lane == 0) {
                        pub[nb].p = p;
                        pub[nb].k = none ? -1 : k;
                        if (!none) {
                            const int cv = cvsm[k];       // becomes basic in row r
                            cvsm[k] = rowvar[r];          // the old basic slack becomes nonbasic in column k
                            rowvar[r] = cv;
                        }
                    }
                }
                __syncthreads();
                const int k = pub[nb].k;
                if (k < 0) { status = ST_INFEASIBLE; break; }
                apply_pivot(pr, r, k, pub[nb].p, false, e, nN);
                publish_rhs();
                nb ^= 1;
                ++npiv_p1;
                __syncwarp();
            }
            __syncthreads();
            TSTAGE(3);                                         // phase 1

            // ---- stage 3b: phase 2 (Dantzig) ---------------------------------------------------------------
            // The pivot row's warp prices the NEXT entering column while it holds the row.  First column: every warp
            // computes it (read-only).
            int k = -1;
            if (status == ST_OPTIMAL) {
                double gmin = kInf;
                int bq = 0;
#pragma unroll
                for (int qs = 0; qs < QS; ++qs) {
                    const int j = lane + 32 * qs;
                    const double g = (j < n) ? gsm[cpos(j)] : kInf;
                    if (g < gmin) { gmin = g; bq = qs; }
                }
                unsigned long long kmin;
                const int kl = warp_argmin_key(dkey(gmin), kmin);
                if (kmin < dkey(-kTolFeas)) k = kl + 32 * __shfl_sync(FULL, bq, kl);
            }
            while (status == ST_OPTIMAL && k >= 0) {
                // entering column -> fcol; ratio test: one tile row per lane
                {
                    const int lck = k / C;
                    if (lc == lck) extract_col(k - lck * C, e);
                    __syncwarp();
                    const int u = warp * RW + lane;
                    const bool in = (lane < RW && u < nN);
                    const double et = in ? fcol[rpos(u)] : 0.0;
                    const double sc = in ? fmax(svec[rpos(u)], 0.0) : 0.0;
                    const bool ok = in && et > kTolPivot;
                    const double ratio = sc * fast_rcp(ok ? et : 1.0);
                    unsigned long long kmin;
                    const int ll = warp_argmin_key(ok ? dkey(ratio) : KEY_INF, kmin);
                    if (lane == 0) {
                        keys[warp] = kmin;
                        crow[warp] = warp * RW + ll;
                    }
                }
                __syncthreads();
                unsigned long long kmin;
                const int ww = warp_argmin_key((lane < W) ? keys[lane] : KEY_INF, kmin);
                if (kmin == KEY_INF) { status = ST_UNBOUNDED; break; }
                if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int r = crow[ww];
                double* pr = prow + nb * PR;
                const int wr = r / RW, lrr = (r - wr * RW) / R, rsr = r - wr * RW - lrr * R;
                if (warp == wr) {
                    if (lr == lrr) publish_row(rsr, pr);
                    __syncwarp();
                    const double p = pr[cpos(k)];
                    const double fg = gsm[cpos(k)] * fast_rcp(p);
                    double gn[QS];
                    double gmin = kInf;
                    int bq = 0;
#pragma unroll
                    for (int qs = 0; qs < QS; ++qs) {
                        const int j = lane + 32 * qs;
                        const int jp = cpos(j);
                        gn[qs] = kInf;
                        if (j < n) gn[qs] = (j == k) ? -fg : fma(-fg, pr[jp], gsm[jp]);
                        if (gn[qs] < gmin) { gmin = gn[qs]; bq = qs; }
                    }
                    __syncwarp();
#pragma unroll
                    for (int qs = 0; qs < QS; ++qs) {
                        const int j = lane + 32 * qs;
                        if (j < n) gsm[cpos(j)] = gn[qs];
                    }
                    unsigned long long kmin2;
                    const int kl = warp_argmin_key(dkey(gmin), kmin2);
                    const int knext = (kmin2 < dkey(-kTolFeas)) ? kl + 32 * __shfl_sync(FULL, bq, kl) : -1;
                    if (lane == 0) {
                        pub[nb].p = p;
                        pub[nb].k = knext;
                        const int cv = cvsm[k];
                        cvsm[k] = rowvar[r];
                        rowvar[r] = cv;
                    }
                }
                __syncthreads();
                apply_pivot(pr, r, k, pub[nb].p, true, e, nN);
                publish_rhs();
                k = pub[nb].k;
                nb ^= 1;
                ++npiv_p2;
                __syncwarp();
            }
        }

        // ---- stage 4: x, objective, slacks, labels -----------------------------------------------------------------
        __syncthreads();
        TSTAGE(4);                                             // phase 2
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0, nref = 0;
        if (need_generic) {
            status = -1;   // re-solved by the generic kernel (capi.cu)
        } else if (status == ST_OPTIMAL) {
            // where does every constraint sit now?
            for (int u = tid; u < nN; u += NT) {
                const int q = rowvar[u];
                if (q >= 0) basic_tile[q] = u;
            }
            __syncthreads();
            for (int j = tid; j < n; j += NT) {
                const int bt = basic_tile[colvar0[j]];
                sig[j] = (bt >= 0) ? svec[rpos(bt)] : 0.0;
            }
            __syncthreads();
            {
                double sl[QS];
#pragma unroll
                for (int qs = 0; qs < QS; ++qs) {
                    const int j = lane + 32 * qs;
                    sl[qs] = (j < n) ? sig[j] : 0.0;
                }
                for (int k2 = warp; k2 < n; k2 += W) {
                    double acc = 0.0;
#pragma unroll
                    for (int qs = 0; qs < QS; ++qs) {
                        const int j = lane + 32 * qs;
                        if (j < n) acc = fma(Dsm[(size_t)k2 * PR + cpos(j)], sl[qs], acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) xbuf[k2] = Dsm[(size_t)k2 * PR + cpos(RHS)] - acc;
                }
            }
            __syncthreads();
            double xl[QS];
#pragma unroll
            for (int qs = 0; qs < QS; ++qs) {
                const int j = lane + 32 * qs;
                xl[qs] = (j < n) ? xbuf[j] : 0.0;
            }
            if (warp == 0) {
                double acc = 0.0;
#pragma unroll
                for (int qs = 0; qs < QS; ++qs) {
                    const int j = lane + 32 * qs;
                    if (j < n) acc = fma(__ldg(cg + j), xl[qs], acc);
                }
                acc = warp_sum(acc);
                if (lane == 0 && a.obj) a.obj[lp] = acc;
            }
            if (a.x)
                for (int j = tid; j < n; j += NT) a.x[(size_t)lp * n + j] = xbuf[j];
            row_dots(Ag, xl, gbuf, nullptr);          // gbuf[i] = a_i . x
            __syncthreads();
            for (int i = tid; i < m; i += NT) {
                const double slack = __ldg(bg + i) - gbuf[i];
                const double as = fabs(slack);
                const int active = as <= a.thr;
                lab[i] = (uint8_t)active;
                nact += active;
                int tie = (as >= a.thr * 0.1 && as <= a.thr * 10.0);
                const bool excl = mask && mask[i] == 0;
                if (!excl) tie |= (active != (basic_tile[i] < 0));
                nties += tie;
                nviol += (slack < -a.thr);
                nref += (!excl && basic_tile[i] < 0 && as > a.thr * 0.01);   // active row with a visible residual
            }
        }
        // An optimal instance whose active rows do not have (numerically) zero slack at the computed x -- an
        // ill-conditioned vertex -- is handed to the generic kernel, which holds the tableau in memory and can run a
        // step of iterative refinement on the final active set (simplex_generic.cu).
        if (__syncthreads_or(nref > 0) && status == ST_OPTIMAL) status = -1;
        if (!need_generic && status != ST_OPTIMAL) {
            for (int i = tid; i < m; i += NT) lab[i] = 0;
            if (a.x)
                for (int j = tid; j < n; j += NT) a.x[(size_t)lp * n + j] = 0.0;
            if (tid == 0 && a.obj) a.obj[lp] = __longlong_as_double(0x7ff8000000000000ll);
        }
        nact = __reduce_add_sync(FULL, nact);
        nties = __reduce_add_sync(FULL, nties);
        nviol = __reduce_add_sync(FULL, nviol);
        __syncthreads();
        if (lane == 0) {
            red[warp * 3 + 0] = nact;
            red[warp * 3 + 1] = nties;
            red[warp * 3 + 2] = nviol;
        }
        __syncthreads();
        if (tid == 0) {
            int t0 = 0, t1 = 0, t2 = 0;
            for (int w = 0; w < W; ++w) {
                t0 += red[w * 3 + 0];
                t1 += red[w * 3 + 1];
                t2 += red[w * 3 + 2];
            }
            a.status[lp] = status;
            if (status == -1) atomicAdd(a.flag_count, 1);
            if (status != -1) {
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
        __syncthreads();
#ifdef DDB_TIMING
        TSTAGE(5);                                             // stage 4
        if (a.gtab && tid == 0) {
            double* o = a.gtab + (size_t)lp * 16;
            for (int q = 0; q < 6; ++q) o[q] = (double)tacc[q];
            o[6] = npiv_crash; o[7] = npiv_p1; o[8] = npiv_p2; o[9] = status;
        }
#endif
    }
    (void)dsig;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
namespace {
struct TileVariant {
    int W, R, C;
    cudaError_t (*launch)(const SolveArgs&, int, cudaStream_t);
    int rows() const { return W * 4 * R; }
    int cols() const { return 8 * C - 1; }     // structural columns (one tile column is the rhs)
};

template <int W, int R, int C, int MINB>
cudaError_t launch_tile_variant(const SolveArgs& a, int sm_count, cudaStream_t st) {
    auto kern = simplex_tile2d_kernel<W, R, C, MINB>;
    const size_t smem = make_tile_layout(a.m, a.n, W, R, C).total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)sm_count * per_sm;
    if (grid > a.B) grid = a.B;
    kern<<<(int)grid, W * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// Picked: the first variant whose tile holds max(n, m - n) rows and n + 1 columns.
const TileVariant kTileVariants[] = {
    {1, 2, 1, launch_tile_variant<1, 2, 1, 16>},      //   8 rows x   7 cols : (10,5)
    {1, 8, 3, launch_tile_variant<1, 8, 3, 12>},      //  32 rows x  23 cols : (50,20)
    {2, 8, 5, launch_tile_variant<2, 8, 5, 4>},       //  64 rows x  39 cols
    {4, 8, 9, launch_tile_variant<4, 8, 9, 2>},       // 128 rows x  71 cols
    {4, 7, 13, launch_tile_variant<4, 7, 13, 2>},     // 112 rows x 103 cols : (200,100)
    {8, 7, 13, launch_tile_variant<8, 7, 13, 1>},     // 224 rows x 103 cols : m up to 324 at n = 100
};

const TileVariant* pick_tile_variant(int m, int n) {
    const int rows = (m - n > n) ? (m - n) : n;
    for (const TileVariant& v : kTileVariants)
        if (n <= v.cols() && rows <= v.rows()) return &v;
    return nullptr;
}
}  // namespace

bool tile2d_supported(int m, int n) { return m >= n && pick_tile_variant(m, n) != nullptr; }

cudaError_t launch_simplex_tile2d(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const TileVariant* v = pick_tile_variant(a.m, a.n);
    if (!v) return cudaErrorInvalidValue;
    return v->launch(a, sm_count, st);
}

}  // namespace ddb
