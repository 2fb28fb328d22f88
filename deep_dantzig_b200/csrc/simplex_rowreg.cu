// Row-per-thread batched fp64 simplex (plan 0): one LP per CTA, ONE TABLEAU ROW PER THREAD, held in registers.
//
// Why this layout: a pivot is a rank-1 update  T[i][c] -= f_i * prow[c].  With a whole row in one thread's registers
//   * the update is NC DFMAs per thread whose second operand (the pivot row) is the same for every thread: it is
//     read from shared memory with broadcast LDS.128 (one wavefront per two entries), so the instruction stream of
//     a pivot is ~NC DFMA + NC/2 LDS per warp and nothing else scales with the row count;
//   * everything that is "per row" (the right-hand side, the entry in the entering column, the ratio test, the
//     most-negative-slack search of phase 1) is one value PER LANE, so a ratio test over 32 rows is a handful of
//     instructions followed by a redux-based warp argmin -- not a loop;
//   * everything that is "per column" (pricing vectors ghat / g, the column -> constraint map) is kept
//     lane-distributed (column j in lane j%32, slot j/32) and replicated in every warp, so pricing is also a
//     handful of instructions + one warp argmin, identical in every warp, with no communication.
// The only dynamically indexed register accesses are "my entry in the entering column k" (read) and "the new
// column k" (write): warp-uniform branch trees (reg_get / reg_set below).
//
// One barrier per pivot: before the barrier EVERY warp speculatively publishes its own best candidate row and
// prices the cost vector that would result if that candidate won (per-warp copies of g in shared memory); after
// the barrier all warps pick the same winner from W small headers and go straight into the rank-1 update.
//
// At (m,n) = (200,100): NC = 101 registers-pairs per thread, 4 warps per LP, 2 LPs resident per SM (the register
// file is the limit: 2 x 128 x 255 registers), 404 warp-DFMAs per pivot = 202 clk of one SM's fp64 pipe.
//
// Stages per LP (same algorithm and tolerances as simplex_generic.cu):
//   0. crash order by cosine score (A streamed once from HBM through the idle tableau registers)
//   1. crash as an explicit inverse: Gauss-Jordan on the n x n block A_B0 (thread t owns row t)
//   2. remaining rows enter through  P_N = -A_N D  (D rows broadcast from shared memory)
//   3. phase 1 (most negative slack leaves, ratio test along the published row), phase 2 (Dantzig)
//   4. x = xv - D sigma, slack = b - A x from the caller's A, labels = |slack| <= threshold
// Instances the tile cannot hold or whose static crash basis is singular are flagged status = -1 and re-solved by
// the generic kernel on the device (capi.cu); nothing ever falls back to the CPU.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace ddb {

// Optional per-section cycle accounting of the phase-2 loop (debug builds only: -DDDB_TIMING; results in a.gtab).
#ifdef DDB_TIMING
#define TSTAMP(i)                                   \
    do {                                            \
        const long long _t = clock64();             \
        tacc[i] += _t - tlast;                      \
        tlast = _t;                                 \
    } while (0)
#else
#define TSTAMP(i) do { } while (0)
#endif

// ---- dynamic (warp-uniform) register index -> uniform branch tree ----------------------------------------------
// The asm volatile leaves keep the compiler from if-converting the switch into a select per register (3
// instructions per tableau column) and from copying registers around at the merge points.
#define DDB_R8(M, b) M(b + 0) M(b + 1) M(b + 2) M(b + 3) M(b + 4) M(b + 5) M(b + 6) M(b + 7)
#define DDB_R128(M)                                                                                       \
    DDB_R8(M, 0) DDB_R8(M, 8) DDB_R8(M, 16) DDB_R8(M, 24) DDB_R8(M, 32) DDB_R8(M, 40) DDB_R8(M, 48)       \
    DDB_R8(M, 56) DDB_R8(M, 64) DDB_R8(M, 72) DDB_R8(M, 80) DDB_R8(M, 88) DDB_R8(M, 96) DDB_R8(M, 104)    \
    DDB_R8(M, 112) DDB_R8(M, 120)

template <int NC>
__device__ __forceinline__ double reg_get(const double (&T)[NC], int k) {
    static_assert(NC <= 128, "the switch covers 128 registers");
    double v = 0.0;
    switch (k) {
#define DDB_CASE(I)                                                                         \
    case (I):                                                                               \
        if constexpr ((I) < NC) asm volatile("mov.f64 %0, %1;" : "=d"(v) : "d"(T[(I) < NC ? (I) : 0])); \
        break;
        DDB_R128(DDB_CASE)
#undef DDB_CASE
        default: break;
    }
    return v;
}
template <int NC>
__device__ __forceinline__ void reg_set(double (&T)[NC], int k, double v) {
    switch (k) {
#define DDB_CASE(I)                                                                         \
    case (I):                                                                               \
        if constexpr ((I) < NC) asm volatile("mov.f64 %0, %1;" : "=d"(T[(I) < NC ? (I) : 0]) : "d"(v)); \
        break;
        DDB_R128(DDB_CASE)
#undef DDB_CASE
        default: break;
    }
}

struct RowCand {              // one per (buffer, warp): that warp's candidate pivot row
    double p;                 // pivot entry (stored scale)
    double il;                // 1 / lam of the candidate row
    int row;                  // tile row (= thread index)
    int k;                    // crash / phase 1: entering column (-1: none); phase 2: NEXT entering column (-1: optimal)
    int var;                  // constraint whose slack is basic in the candidate row
    int pad;
};

// Pitch (in doubles) of the rows kept in shared memory: even (16-byte rows for LDS.128) with pitch/2 odd, so that
// 8 lanes reading the same 16-byte column of 8 consecutive rows hit 8 different bank groups.
__host__ __device__ constexpr int row_pitch(int NC) {
    int pd = (NC + 1) & ~1;
    if (((pd / 2) & 1) == 0) pd += 2;
    return pd;
}

struct RowLayout {
    // persistent per LP
    size_t D, order, colvar0, pivcol, basic_tile, cvsm;
    // pivot loops (crash, phase 1, phase 2) -- aliased with the stage 0 / stage 4 scratch below
    size_t prow, cand, keys, gbufs, ghbufs;
    // stage 0 / stage 4 scratch
    size_t gbuf, gnn, sval, sig, xbuf, red;
    size_t total;
};
__host__ __device__ inline size_t rr_align(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline RowLayout make_row_layout(int m, int n, int NC, int W) {
    RowLayout L;
    const int PD = row_pitch(NC);
    const int CT = 32 * ((NC + 31) / 32);
    size_t off = 0;
    L.D = off;          off += rr_align((size_t)n * PD * 8);
    L.order = off;      off += rr_align((size_t)m * 4);
    L.colvar0 = off;    off += rr_align((size_t)n * 4);
    L.pivcol = off;     off += rr_align((size_t)n * 4);
    L.basic_tile = off; off += rr_align((size_t)m * 4);
    L.cvsm = off;       off += rr_align((size_t)CT * 4);
    const size_t u0 = off;
    L.prow = off;       off += rr_align((size_t)2 * W * PD * 8);
    L.cand = off;       off += rr_align((size_t)2 * W * sizeof(RowCand));
    L.keys = off;       off += rr_align((size_t)2 * W * 8);
    L.gbufs = off;      off += rr_align((size_t)2 * W * CT * 8);
    L.ghbufs = off;     off += rr_align((size_t)2 * W * CT * 8);
    const size_t u1 = off;
    off = u0;
    L.gbuf = off;       off += rr_align((size_t)m * 8);
    L.gnn = off;        off += rr_align((size_t)m * 8);
    L.sval = off;       off += rr_align((size_t)W * 32 * 8);
    L.sig = off;        off += rr_align((size_t)n * 8);
    L.xbuf = off;       off += rr_align((size_t)(n > CT ? n : CT) * 8);
    L.red = off;        off += rr_align((size_t)(3 * W + 4) * 4);
    L.total = off > u1 ? off : u1;
    return L;
}

template <int NC, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) simplex_rowreg_kernel(SolveArgs a) {
    constexpr int CS = (NC + 31) / 32;      // slots of the lane-distributed column vectors
    constexpr int CT = 32 * CS;
    constexpr int PD = row_pitch(NC);
    constexpr int NT = W * 32;              // threads = tile rows
    constexpr int RHS = NC - 1;             // register / column that holds the right-hand side
    constexpr int RB = NC / CS;             // rows per register batch when T is used as a streaming buffer
    constexpr int NCH = (NC + 7) / 8;       // 8-column chunks of a row
    static_assert(W <= 32, "one header per lane");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const RowLayout L = make_row_layout(m, n, NC, W);
    double* Dsm = reinterpret_cast<double*>(smem_raw + L.D);
    int* order = reinterpret_cast<int*>(smem_raw + L.order);
    int* colvar0 = reinterpret_cast<int*>(smem_raw + L.colvar0);
    int* pivcol = reinterpret_cast<int*>(smem_raw + L.pivcol);
    int* basic_tile = reinterpret_cast<int*>(smem_raw + L.basic_tile);
    int* cvsm = reinterpret_cast<int*>(smem_raw + L.cvsm);          // column -> constraint whose slack is nonbasic
    double* prow = reinterpret_cast<double*>(smem_raw + L.prow);    // [2][W][PD] candidate pivot rows
    RowCand* cand = reinterpret_cast<RowCand*>(smem_raw + L.cand);  // [2][W]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + L.keys);   // [2][W]
    double* gbufs = reinterpret_cast<double*>(smem_raw + L.gbufs);    // [2][W][CT] g (true reduced costs) candidates
    double* ghbufs = reinterpret_cast<double*>(smem_raw + L.ghbufs);  // [2][W][CT] ghat (phase-1 costs) candidates
    double* gbuf = reinterpret_cast<double*>(smem_raw + L.gbuf);
    double* gnn = reinterpret_cast<double*>(smem_raw + L.gnn);
    double* sval = reinterpret_cast<double*>(smem_raw + L.sval);
    double* sig = reinterpret_cast<double*>(smem_raw + L.sig);
    double* xbuf = reinterpret_cast<double*>(smem_raw + L.xbuf);
    int* red = reinterpret_cast<int*>(smem_raw + L.red);
    __shared__ long long cur_lp;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // ---- register state -------------------------------------------------------------------------------------
    double T[NC];            // my tableau row; T[RHS] is its right-hand side
    int rowvar = -1;         // constraint whose slack is basic in my row
    double lam = 1.0, ilam = 1.0;   // my row's lazy scale: true row = lam * T

    // Rows are stored LAZILY NORMALISED: the true tableau row is lam * T (ilam = 1 / lam).  A pivot (r, k) never
    // rescales the pivot row: with p = T_r[k] (stored), rp = 1/p, il = ilam_r (before the pivot)
    //     rows i != r :  f = T_i[k] * rp;  T_i[c] -= f * T_r[c] (c != k);  T_i[k] = -f * il       (lam_i unchanged)
    //     row r       :  T_r unchanged except T_r[k] = il;  lam_r = rp, ilam_r = p
    //     costs       :  g[c] -= g_k rp T_r[c] (c != k);  g[k] = -g_k rp il
    // so the owner just publishes its raw registers and every other thread runs one FMA per entry.

    // plain rank-1 update of my row from a row in shared memory (broadcast reads):  T[c] -= f * pr[c]
    auto rank1 = [&](const double* pr, double f) {
        const double nf = -f;
        const double2* p2 = reinterpret_cast<const double2*>(pr);
#pragma unroll
        for (int c2 = 0; c2 < NC / 2; ++c2) {
            const double2 v = p2[c2];
            T[2 * c2] = fma(nf, v.x, T[2 * c2]);
            T[2 * c2 + 1] = fma(nf, v.y, T[2 * c2 + 1]);
        }
        if constexpr (NC & 1) T[NC - 1] = fma(nf, pr[NC - 1], T[NC - 1]);
    };
    // rank-1 update, then write column kset (T[kset] = newval) and read the updated column kget (warp-uniform
    // indices, -1 = none): two uniform branch trees around the pure-FMA loop.
    auto rank1_fused = [&](const double* pr, double f, int kset, double newval, int kget, double& eget) {
        rank1(pr, f);
        if (kset >= 0) reg_set<NC>(T, kset, newval);
        if (kget >= 0) eget = reg_get<NC>(T, kget);
    };
    auto get_col = [&](int k) -> double { return reg_get<NC>(T, k); };
    auto publish = [&](double* pr) {
        double2* p2 = reinterpret_cast<double2*>(pr);
#pragma unroll
        for (int c2 = 0; c2 < NC / 2; ++c2) p2[c2] = make_double2(T[2 * c2], T[2 * c2 + 1]);
        if constexpr (NC & 1) pr[NC - 1] = T[NC - 1];
    };
    // winner among the W candidate keys of buffer nb (lowest warp on ties); kmin = its key
    auto pick_winner = [&](int nb, unsigned long long& kmin) -> int {
        if constexpr (W <= 4) {
            const unsigned long long* kk = keys + nb * W;
            int ww = 0;
            kmin = kk[0];
#pragma unroll
            for (int w = 1; w < W; ++w) {
                const unsigned long long kw = kk[w];
                if (kw < kmin) { kmin = kw; ww = w; }
            }
            return ww;
        } else {
            return warp_argmin_key((lane < W) ? keys[nb * W + lane] : KEY_INF, kmin);
        }
    };
    // T used as a streaming buffer: dot products of up to RB*W rows of A (from `base`) with a lane-distributed
    // vector; all loads of a batch are in flight together.  out1[i] = a_i . v ; out2[i] = a_i . a_i (optional)
    auto row_dots = [&](const double* Ag, const double (&vl)[CS], double* out1, double* out2) {
        for (int base = 0; base < m; base += RB * W) {
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int i = base + r * W + warp;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    T[r * CS + cs] = (i < m && j < n) ? __ldg(Ag + (size_t)i * n + j) : 0.0;
                }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int i = base + r * W + warp;
                double dot = 0.0, nn = 0.0;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const double v = T[r * CS + cs];
                    dot = fma(v, vl[cs], dot);
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

        // ---- stage 0: crash order ---------------------------------------------------------------------------
        {
            double cl[CS];
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
                const int j = lane + 32 * cs;
                cl[cs] = (j < n) ? __ldg(cg + j) : 0.0;
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
        bool need_generic = (nN < 0) || (nN > NT) || (n > NT) || (n > NC - 1);

        int npiv_crash = 0, npiv_p1 = 0, npiv_p2 = 0;
        int status = ST_OPTIMAL;
        int par = 0, gw = 0;     // the current cost vectors are gbufs / ghbufs [par][gw]

        if (!need_generic) {
            // ---- stage 1: thread t < n loads row order[t] of [A | b]; Gauss-Jordan to the inverse --------------
            {
                const bool have = tid < n;
                const int row = have ? order[tid] : 0;
                const double* Ar = Ag + (size_t)row * n;
                if ((n & 1) == 0 && (reinterpret_cast<size_t>(a.A) & 15) == 0) {
                    const double2* Ar2 = reinterpret_cast<const double2*>(Ar);
#pragma unroll
                    for (int c2 = 0; c2 < (NC - 1) / 2; ++c2) {
                        double2 v = make_double2(0.0, 0.0);
                        if (have && 2 * c2 < n) v = __ldg(Ar2 + c2);
                        T[2 * c2] = v.x;
                        T[2 * c2 + 1] = v.y;
                    }
                    if constexpr (((NC - 1) & 1) != 0) T[NC - 2] = 0.0;
                } else {
#pragma unroll
                    for (int c = 0; c < NC - 1; ++c) T[c] = (have && c < n) ? __ldg(Ar + c) : 0.0;
                }
                T[RHS] = have ? __ldg(bg + row) : 0.0;
                rowvar = have ? row : -1;
                lam = 1.0;
                ilam = 1.0;
            }
            for (int j = tid; j < CT; j += NT) {
                gbufs[j] = (j < n) ? __ldg(cg + j) : 0.0;       // [0][0]
                cvsm[j] = (j < n) ? -1 : -2;                     // -1: still a free x_j (crash), -2: not a column
            }
            __syncthreads();

            for (int t = 0; t < n; ++t) {
                const int nb = t & 1;
                double* pr = prow + (size_t)(nb * W) * PD;
                RowCand* cd = cand + nb * W;
                const bool own = (tid == t);
                if (warp == (t >> 5)) {
                    // the owner publishes its raw row; its warp finds the pivot column (largest |entry| among the
                    // free columns; a row is a pivot row once in the crash, so its lam is still 1) and updates g
                    if (own) publish(pr);
                    __syncwarp();
                    double rv[CS];
                    unsigned long long best = 0ull;
                    int bq = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        rv[cs] = (j < n) ? pr[j] : 0.0;
                        const unsigned long long kk =
                            (cvsm[j] == -1) ? (unsigned long long)__double_as_longlong(fabs(rv[cs])) : 0ull;
                        if (kk > best) { best = kk; bq = cs; }
                    }
                    unsigned long long kmin;
                    const int kl = warp_argmin_key(~best, kmin);   // argmax through the complemented key
                    const int k = kl + 32 * __shfl_sync(FULL, bq, kl);
                    const double pabs = __longlong_as_double((long long)~kmin);
                    const double p = pr[k];
                    const double fg = gbufs[k] * fast_rcp(p);
                    __syncwarp();
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        const double g = fma(-fg, rv[cs], gbufs[j]);
                        gbufs[j] = (j == k) ? -fg : g;
                    }
                    if (lane == 0) {
                        cd->p = p;
                        cd->k = (pabs >= kTolCrash) ? k : -1;
                        pivcol[t] = k;
                        cvsm[k] = 0;                  // no longer free
                    }
                }
                __syncthreads();
                const int k = cd->k;
                if (k < 0) { need_generic = true; break; }
                const double p = cd->p;
                const double rp = fast_rcp(p);
                const double e = get_col(k);
                const double f = own ? 0.0 : e * rp;
                double dummy;
                rank1_fused(pr, f, k, own ? 1.0 : -f, -1, dummy);
                if (own) { lam = rp; ilam = p; }
                ++npiv_crash;
            }
        }

        if (!need_generic) {
            __syncthreads();
            // dump D' (row of x_k stored at index k; column RHS holds the x-vertex) and the column -> constraint map
            if (tid < n) {
                const int k = pivcol[tid];
#pragma unroll
                for (int c = 0; c < NC; ++c) T[c] *= lam;      // true rows of the inverse
                publish(Dsm + (size_t)k * PD);
                colvar0[k] = rowvar;
                cvsm[k] = rowvar;
            }
            for (int j = tid; j < CT; j += NT) ghbufs[j] = (j < n) ? 1.0 : 0.0;     // [0][0]: artificial costs
            __syncthreads();

            // ---- stage 2: my row of P_N = -A_N D, s_N = b_N - A_N xv -------------------------------------------
            const bool live = tid < nN;
            {
                const int myrow = live ? order[n + tid] : 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) T[c] = 0.0;
                if (warp * 32 < nN) {
                    const double* Ar = Ag + (size_t)myrow * n;
                    double an[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) an[q] = (live && q < n) ? __ldg(Ar + q) : 0.0;
                    for (int k0 = 0; k0 < n; k0 += 4) {
                        double av[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            av[q] = an[q];
                            an[q] = (live && k0 + 4 + q < n) ? __ldg(Ar + k0 + 4 + q) : 0.0;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (k0 + q < n) rank1(Dsm + (size_t)(k0 + q) * PD, av[q]);
                    }
                }
                if (live) T[RHS] += __ldg(bg + myrow);
                rowvar = live ? myrow : -1;
                lam = 1.0;
                ilam = 1.0;
            }

            // ---- stage 3a: phase 1 (most negative slack leaves; ratio test along its row) -------------------
            // Every warp publishes its own most-negative row, runs the ratio test along it and writes the cost
            // vectors that pivot would produce into its private candidate buffers; after the barrier the winner
            // (most negative slack over all warps) is adopted by everybody.
            for (;;) {
                const int nb = par ^ 1;
                const double* gcur = gbufs + (size_t)(par * W + gw) * CT;
                const double* ghcur = ghbufs + (size_t)(par * W + gw) * CT;
                const double s = lam * T[RHS];
                unsigned long long kmin;
                const int ll = warp_argmin_key((live && s < -kTolFeas) ? dkey(s) : KEY_INF, kmin);
                if (kmin != KEY_INF) {
                    double* prw = prow + (size_t)(nb * W + warp) * PD;
                    if (lane == ll) publish(prw);
                    __syncwarp();
                    const double lam_r = __shfl_sync(FULL, lam, ll);
                    const double il = __shfl_sync(FULL, ilam, ll);
                    const int var_r = __shfl_sync(FULL, rowvar, ll);
                    double rv[CS], gh[CS];
                    double bn = 0.0, bd = 0.0;   // best numerator / denominator (bd == 0: none)
                    int bq = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        rv[cs] = (j < n) ? prw[j] : 0.0;
                        gh[cs] = ghcur[j];
                        const double e = -lam_r * rv[cs];       // true entry of the candidate row
                        if (e > kTolPivot) {
                            const double num = fmax(gh[cs], 0.0);
                            if (bd == 0.0 || num * bd < bn * e) { bn = num; bd = e; bq = cs; }
                        }
                    }
                    const double ratio = bn * fast_rcp(bd > 0.0 ? bd : 1.0);
                    unsigned long long kmin2;
                    const int kl = warp_argmin_key((bd > 0.0) ? dkey(ratio) : KEY_INF, kmin2);
                    const bool none = (kmin2 == KEY_INF);
                    const int k = none ? 0 : kl + 32 * __shfl_sync(FULL, bq, kl);
                    const double p = none ? 1.0 : prw[k];
                    const double rp = fast_rcp(p);
                    const double fv = ghcur[k] * rp, fg = gcur[k] * rp;
                    double* gnext = gbufs + (size_t)(nb * W + warp) * CT;
                    double* ghnext = ghbufs + (size_t)(nb * W + warp) * CT;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        const double h = fma(-fv, rv[cs], gh[cs]);
                        const double g = fma(-fg, rv[cs], gcur[j]);
                        ghnext[j] = (j == k) ? -fv * il : h;
                        gnext[j] = (j == k) ? -fg * il : g;
                    }
                    if (lane == 0) {
                        RowCand* cd = cand + nb * W + warp;
                        cd->p = p;
                        cd->il = il;
                        cd->row = warp * 32 + ll;
                        cd->k = none ? -1 : k;
                        cd->var = var_r;
                    }
                }
                if (lane == 0) keys[nb * W + warp] = kmin;
                __syncthreads();
                const int ww = pick_winner(nb, kmin);
                if (kmin == KEY_INF) break;                       // s >= 0 everywhere: phase 1 finished
                if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const RowCand* cd = cand + nb * W + ww;
                const int k = cd->k;
                if (k < 0) { status = ST_INFEASIBLE; break; }
                const double p = cd->p, il = cd->il;
                const bool own = (tid == cd->row);
                const double rp = fast_rcp(p);
                const double e = get_col(k);
                const double f = own ? 0.0 : e * rp;
                double dummy;
                rank1_fused(prow + (size_t)(nb * W + ww) * PD, f, k, own ? il : -f * il, -1, dummy);
                if (own) {
                    lam = rp;
                    ilam = p;
                    const int cv = cvsm[k];       // becomes basic in my row
                    cvsm[k] = rowvar;             // my old slack becomes nonbasic in column k
                    rowvar = cv;
                }
                par = nb;
                gw = ww;
                ++npiv_p1;
            }
            __syncthreads();

            // ---- stage 3b: phase 2 (Dantzig) ---------------------------------------------------------------
            // Same scheme; a candidate also carries the NEXT entering column (priced on the cost vector its pivot
            // would produce), and the rank-1 update extracts my entry in that column on the way.
            int k = -1;
            double e = 0.0;
            if (status == ST_OPTIMAL) {
                const double* gcur = gbufs + (size_t)(par * W + gw) * CT;
                double gmin = kInf;
                int bq = 0;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    const double g = gcur[j];
                    if (j < n && g < gmin) { gmin = g; bq = cs; }
                }
                unsigned long long kmin;
                const int kl = warp_argmin_key(dkey(gmin), kmin);
                if (kmin < dkey(-kTolFeas)) k = kl + 32 * __shfl_sync(FULL, bq, kl);
                if (k >= 0) e = get_col(k);
            }
#ifdef DDB_TIMING
            long long tacc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long tlast = clock64();
#endif
            while (status == ST_OPTIMAL && k >= 0) {
                TSTAMP(7);
                const int nb = par ^ 1;
                const double* gcur = gbufs + (size_t)(par * W + gw) * CT;
                // ratio test: one row per lane
                const double et = lam * e;                         // true entry / right-hand side of my row
                const double sc = fmax(lam * T[RHS], 0.0);
                const bool ok = live && et > kTolPivot;
                const double ratio = sc * fast_rcp(ok ? et : 1.0);
                unsigned long long kmin;
                const int ll = warp_argmin_key(ok ? dkey(ratio) : KEY_INF, kmin);
                TSTAMP(0);                                         // ratio test + warp argmin
                if (kmin != KEY_INF) {
                    double* prw = prow + (size_t)(nb * W + warp) * PD;
                    if (lane == ll) publish(prw);
                    __syncwarp();
                    TSTAMP(1);                                     // publish
                    const double il = __shfl_sync(FULL, ilam, ll);
                    const int var_r = __shfl_sync(FULL, rowvar, ll);
                    const double p = prw[k];
                    const double fg = gcur[k] * fast_rcp(p);
                    double* gnext = gbufs + (size_t)(nb * W + warp) * CT;
                    double gmin = kInf;
                    int bq = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        const double rvj = (j < n) ? prw[j] : 0.0;
                        double g = fma(-fg, rvj, gcur[j]);
                        if (j == k) g = -fg * il;
                        gnext[j] = g;
                        if (j < n && g < gmin) { gmin = g; bq = cs; }
                    }
                    unsigned long long kmin2;
                    const int kl = warp_argmin_key(dkey(gmin), kmin2);
                    const int knext = (kmin2 < dkey(-kTolFeas)) ? kl + 32 * __shfl_sync(FULL, bq, kl) : -1;
                    if (lane == 0) {
                        RowCand* cd = cand + nb * W + warp;
                        cd->p = p;
                        cd->il = il;
                        cd->row = warp * 32 + ll;
                        cd->k = knext;
                        cd->var = var_r;
                    }
                    TSTAMP(2);                                     // speculative pricing
                }
                if (lane == 0) keys[nb * W + warp] = kmin;
                __syncthreads();
                const int ww = pick_winner(nb, kmin);
                TSTAMP(3);                                         // barrier + winner
                if (kmin == KEY_INF) { status = ST_UNBOUNDED; break; }
                if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const RowCand* cd = cand + nb * W + ww;
                const double p = cd->p, il = cd->il;
                const int knext = cd->k;
                const bool own = (tid == cd->row);
                const double rp = fast_rcp(p);
                const double f = own ? 0.0 : e * rp;
                double enext = 0.0;
                TSTAMP(4);                                         // header, rcp
                rank1_fused(prow + (size_t)(nb * W + ww) * PD, f, k, own ? il : -f * il, knext, enext);
                TSTAMP(5);                                         // rank-1 update
                if (own) {
                    lam = rp;
                    ilam = p;
                    const int cv = cvsm[k];
                    cvsm[k] = rowvar;
                    rowvar = cv;
                }
                par = nb;
                gw = ww;
                k = knext;
                e = enext;
                ++npiv_p2;
                TSTAMP(6);                                         // bookkeeping
            }
#ifdef DDB_TIMING
            if (a.gtab && lane == 0) {
                double* o = a.gtab + ((size_t)lp * W + warp) * 16;
                for (int q = 0; q < 14; ++q) o[q] = (double)tacc[q];
                o[14] = (double)npiv_p2;
            }
#endif
        }

        // ---- stage 4: x, objective, slacks, labels -----------------------------------------------------------------
        __syncthreads();
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0;
        if (need_generic) {
            status = -1;   // re-solved by the generic kernel (capi.cu)
        } else if (status == ST_OPTIMAL) {
            // where does every constraint sit now?
            if (tid < nN) {
                sval[tid] = lam * T[RHS];
                if (rowvar >= 0) basic_tile[rowvar] = tid;
            }
            __syncthreads();
            for (int j = tid; j < n; j += NT) {
                const int bt = basic_tile[colvar0[j]];
                sig[j] = (bt >= 0) ? sval[bt] : 0.0;
            }
            __syncthreads();
            {
                double sl[CS];
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    sl[cs] = (j < n) ? sig[j] : 0.0;
                }
                for (int k = warp; k < n; k += W) {
                    double acc = 0.0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        if (j < n) acc = fma(Dsm[(size_t)k * PD + j], sl[cs], acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) xbuf[k] = Dsm[(size_t)k * PD + RHS] - acc;
                }
            }
            __syncthreads();
            double xl[CS];
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
                const int j = lane + 32 * cs;
                xl[cs] = (j < n) ? xbuf[j] : 0.0;
            }
            if (warp == 0) {
                double acc = 0.0;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    if (j < n) acc = fma(__ldg(cg + j), xl[cs], acc);
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
                nviol += (slack < -a.thr * 10.0);
            }
        }
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
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
namespace {
struct RowVariant {
    int NC, W, MINB;
    cudaError_t (*launch)(const SolveArgs&, int, cudaStream_t);
};

template <int NC, int W, int MINB>
cudaError_t launch_row_variant(const SolveArgs& a, int sm_count, cudaStream_t st) {
    auto kern = simplex_rowreg_kernel<NC, W, MINB>;
    const size_t smem = make_row_layout(a.m, a.n, NC, W).total;
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

// (columns incl. rhs, warps, min CTAs/SM).  Picked: smallest NC >= n + 1, then smallest W with 32 W >= max(n, m - n).
const RowVariant kRowVariants[] = {
    {8, 1, 32, launch_row_variant<8, 1, 32>},
    {24, 1, 16, launch_row_variant<24, 1, 16>},
    {24, 2, 8, launch_row_variant<24, 2, 8>},
    {48, 2, 5, launch_row_variant<48, 2, 5>},
    {48, 4, 3, launch_row_variant<48, 4, 3>},
    {72, 4, 2, launch_row_variant<72, 4, 2>},
    {101, 4, 2, launch_row_variant<101, 4, 2>},
    {101, 8, 1, launch_row_variant<101, 8, 1>},
};

const RowVariant* pick_row_variant(int m, int n) {
    const int rows = (m - n > n) ? (m - n) : n;
    for (const RowVariant& v : kRowVariants)
        if (n + 1 <= v.NC && rows <= 32 * v.W) return &v;
    return nullptr;
}
}  // namespace

bool rowreg_supported(int m, int n) { return m >= n && pick_row_variant(m, n) != nullptr; }

cudaError_t launch_simplex_rowreg(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const RowVariant* v = pick_row_variant(a.m, a.n);
    if (!v) return cudaErrorInvalidValue;
    return v->launch(a, sm_count, st);
}

}  // namespace ddb
