// Row-per-thread batched fp64 simplex, SOFTWARE-PIPELINED pivots (plan 5, experimental): one LP per CTA, one tableau
// row per thread in registers (layout, lazy row normalisation and dynamic register indexing as in simplex_rowreg.cu,
// the default plan 0).  Parity-green and bit-identical in its pivot path, but MEASURED SLOWER than plan 0 (370 k vs
// 444 k LP/s at (200,100)): the schedule below removes the serial publish and one barrier, yet executes 690 instead of
// 450 instructions per warp per pivot, and at two warps per scheduler the kernel runs at ~10 clk per instruction per
// warp whatever they are (profiles/ncu_r01_rowpipe_vs_rowreg.txt).  Kept for A/B measurements.
//
// What changes against simplex_rowreg.cu is the schedule of a pivot.  There, the owner lane of the pivot row stores its
// 101 registers to shared memory (51 STS.128 from ONE lane, 600-900 clk measured) while every other warp of the LP
// waits, and pricing sits between two more barriers.  Here the update of pivot (r, k) stays PENDING while the next
// pivot is chosen from quantities that cost O(1) per thread:
//   * phase 2: the costs after the pending pivot (every warp prices all columns redundantly: no barrier) give the next
//     entering column k'; my entry in column k' and my right-hand side after the pending pivot are two FMAs, so the
//     ratio test -- hence the next pivot row r' -- is known BEFORE the rank-1 update runs;
//   * phase 1: the right-hand side after the pending pivot is one FMA, so the next leaving row r' is known before the
//     update; the ratio test along row r' runs in every warp once the row is in shared memory;
//   * crash: the pivot rows come in a static order.
// The owner of r' then stores its updated row into the second row buffer WHILE it updates it (predicated STS.128
// interleaved with the LDS.128 / DFMA stream), so publishing costs no serial time.  Two barriers per pivot in phases
// 1 / 2 -- (A) candidates -> r', (B) row r' is in shared memory -- and one in the crash.  The pivot path is bit-identical
// to the other kernels: every quantity is computed with the same FMAs, only earlier.
//
// Replaces, for a whole batch: LinProg model build + optimize + get_statuscode + get_active_constraints
// (reference src/data/gurobi_lp.py:11-29, 370-465) and the label assembly of create_lp_problem
// (reference src/data/randomlp_dataset.py:88-106); algorithm and tolerances: DESIGN.md section 3.
#include <cstdlib>
#include <type_traits>

#include "../common.cuh"

// dev-only stage accounting (tools/rowreg_timing.cu builds with -DDDB_TIMING; never defined in the library build)
#ifdef DDB_TIMING
#define DDB_TSTAMP(i)                          \
    do {                                       \
        if (tid == 0) {                        \
            const long long t_ = clock64();    \
            tacc[i] += (double)(t_ - tlast);   \
            tlast = t_;                        \
        }                                      \
    } while (0)
#else
#define DDB_TSTAMP(i)
#endif

namespace ddb {

// ---- dynamic (warp-uniform) register index -> jump table -------------------------------------------------------
#define DDB_P8(M, b) M(b + 0) M(b + 1) M(b + 2) M(b + 3) M(b + 4) M(b + 5) M(b + 6) M(b + 7)
#define DDB_P128(M)                                                                                       \
    DDB_P8(M, 0) DDB_P8(M, 8) DDB_P8(M, 16) DDB_P8(M, 24) DDB_P8(M, 32) DDB_P8(M, 40) DDB_P8(M, 48)       \
    DDB_P8(M, 56) DDB_P8(M, 64) DDB_P8(M, 72) DDB_P8(M, 80) DDB_P8(M, 88) DDB_P8(M, 96) DDB_P8(M, 104)    \
    DDB_P8(M, 112) DDB_P8(M, 120)

template <int NC>
__device__ __forceinline__ double preg_get(const double (&T)[NC], int k) {
    static_assert(NC <= 128, "jump table covers 128 registers");
    double v = 0.0;
    switch (k) {
#define DDB_PCASE(I)                                    \
    case (I):                                          \
        if constexpr ((I) < NC) v = T[(I) < NC ? (I) : 0]; \
        break;
        DDB_P128(DDB_PCASE)
#undef DDB_PCASE
        default: break;
    }
    return v;
}
// Write at a warp-uniform dynamic index.  The asm volatile leaves keep the compiler from if-converting the switch
// into a select per register (which costs 3 instructions per tableau column); what remains is a uniform branch tree.
template <int NC>
__device__ __forceinline__ void preg_set(double (&T)[NC], int k, double v) {
    switch (k) {
#define DDB_PCASE(I)                                                                         \
    case (I):                                                                               \
        if constexpr ((I) < NC) asm volatile("mov.f64 %0, %1;" : "=d"(T[(I) < NC ? (I) : 0]) : "d"(v)); \
        break;
        DDB_P128(DDB_PCASE)
#undef DDB_PCASE
        default: break;
    }
}

// shared-memory accesses whose order the source fixes (volatile, no memory clobber, immediate offsets: see apply())
template <int OFF>
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0+%3], {%1, %2};" ::"r"(addr), "d"(x), "d"(y), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ void sts64(uint32_t addr, double x) {
    asm volatile("st.shared.f64 [%0+%2], %1;" ::"r"(addr), "d"(x), "n"(OFF));
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

struct PipePub {               // what the pivot row's owner (warp) publishes beside the row itself
    double p;                 // pivot entry (stored scale)
    double il;                // 1 / lam of the pivot row before the pivot
    int k;                    // entering column (-1: none -> infeasible / singular crash basis)
    int var;                  // constraint whose slack was basic in the pivot row
};

struct PipeHdr {               // one per warp: its candidate row
    unsigned long long key;   // dkey(slack) in phase 1, dkey(ratio) in phase 2, KEY_INF = no candidate
    int row;                  // candidate tile row (= thread index)
    int pad;
};

// Pitch (in doubles) of the rows kept in shared memory: even (16-byte rows for LDS.128) with pitch/2 odd, so that
// 8 lanes reading the same 16-byte column of 8 consecutive rows hit 8 different bank groups.
__host__ __device__ constexpr int pipe_pitch(int NC) {
    int pd = (NC + 1) & ~1;
    if (((pd / 2) & 1) == 0) pd += 2;
    return pd;
}

struct PipeLayout {
    size_t D, prow, pub, hdr, part, gsm, ghsm, cvsm, order, colvar0, pivcol, basic_tile, sval, sig, xbuf, gbuf, gnn, red, total;
};
__host__ __device__ inline size_t pp_align(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline PipeLayout make_pipe_layout(int m, int n, int NC, int W) {
    PipeLayout L;
    const int PD = pipe_pitch(NC);
    const int CT = 32 * ((NC + 31) / 32);
    size_t off = 0;
    L.D = off;          off += pp_align((size_t)n * PD * 8);
    L.prow = off;       off += pp_align((size_t)2 * PD * 8);
    L.pub = off;        off += pp_align((size_t)2 * sizeof(PipePub));
    L.hdr = off;        off += pp_align((size_t)W * sizeof(PipeHdr));
    L.part = off;       off += pp_align((size_t)W * sizeof(PipeHdr));
    L.gsm = off;        off += pp_align((size_t)2 * CT * 8);
    L.ghsm = off;       off += pp_align((size_t)2 * CT * 8);
    L.cvsm = off;       off += pp_align((size_t)CT * 4);
    L.order = off;      off += pp_align((size_t)m * 4);
    L.colvar0 = off;    off += pp_align((size_t)n * 4);
    L.pivcol = off;     off += pp_align((size_t)n * 4);
    L.basic_tile = off; off += pp_align((size_t)m * 4);
    L.sval = off;       off += pp_align((size_t)W * 32 * 8);
    L.sig = off;        off += pp_align((size_t)n * 8);
    L.xbuf = off;       off += pp_align((size_t)(n > CT ? n : CT) * 8);
    L.gbuf = off;       off += pp_align((size_t)m * 8);
    L.gnn = off;        off += pp_align((size_t)m * 8);
    L.red = off;        off += pp_align((size_t)(3 * W + 4) * 4);
    L.total = off;
    return L;
}

template <int NC, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) simplex_rowpipe_kernel(SolveArgs a) {
    constexpr int CS = (NC + 31) / 32;      // slots of the lane-distributed column vectors
    constexpr int CT = 32 * CS;
    constexpr int PD = pipe_pitch(NC);
    constexpr int NT = W * 32;              // threads = tile rows
    constexpr int RHS = NC - 1;             // register / column that holds the right-hand side
    constexpr int RB = NC / CS;             // rows per register batch when T is used as a streaming buffer
    static_assert(W <= 32, "one header per lane");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const PipeLayout L = make_pipe_layout(m, n, NC, W);
    double* Dsm = reinterpret_cast<double*>(smem_raw + L.D);
    double* prow = reinterpret_cast<double*>(smem_raw + L.prow);
    PipePub* pub = reinterpret_cast<PipePub*>(smem_raw + L.pub);
    PipeHdr* hdr = reinterpret_cast<PipeHdr*>(smem_raw + L.hdr);
    PipeHdr* part = reinterpret_cast<PipeHdr*>(smem_raw + L.part);    // phase 2: per-warp partial pricing result
    double* gsm = reinterpret_cast<double*>(smem_raw + L.gsm);     // g: true reduced costs ([2][CT], phase 2 double-buffers)
    double* ghsm = reinterpret_cast<double*>(smem_raw + L.ghsm);   // ghat: artificial costs of phase 1
    int* cvsm = reinterpret_cast<int*>(smem_raw + L.cvsm);         // column -> constraint whose slack is nonbasic
    int* order = reinterpret_cast<int*>(smem_raw + L.order);
    int* colvar0 = reinterpret_cast<int*>(smem_raw + L.colvar0);
    int* pivcol = reinterpret_cast<int*>(smem_raw + L.pivcol);
    int* basic_tile = reinterpret_cast<int*>(smem_raw + L.basic_tile);
    double* sval = reinterpret_cast<double*>(smem_raw + L.sval);
    double* sig = reinterpret_cast<double*>(smem_raw + L.sig);
    double* xbuf = reinterpret_cast<double*>(smem_raw + L.xbuf);
    double* gbuf = reinterpret_cast<double*>(smem_raw + L.gbuf);
    double* gnn = reinterpret_cast<double*>(smem_raw + L.gnn);
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
    // rank-1 update of my row from the raw pivot row in shared memory (broadcast reads):  T[c] -= f * prow[c]
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
    auto publish = [&](double* pr) {
        double2* p2 = reinterpret_cast<double2*>(pr);
#pragma unroll
        for (int c2 = 0; c2 < NC / 2; ++c2) p2[c2] = make_double2(T[2 * c2], T[2 * c2 + 1]);
        if constexpr (NC & 1) pr[NC - 1] = T[NC - 1];
    };
    // Pending pivot (r, k) applied to my row:  T[c] -= f * prow[c],  T[k] = own ? il : -f * il.  The owner of the NEXT
    // pivot row (`me`) stores every updated pair to the other row buffer as it is produced (predicated STS.128).
    // Loads and stores are volatile asm without a memory clobber and with immediate offsets: their order is exactly the
    // source order (the compiler cannot prove that the two row buffers do not alias and would otherwise serialise
    // load -> FMA -> store), while the DFMAs float freely.  Schedule per chunk of CH pairs: loads of chunk i + 2, FMAs
    // of chunk i, stores of chunk i - 1 (so a store never waits for the FMA issued just before it).
    auto apply = [&](const double* pr, double* dst, double f, int k, double il, double rp, double p, bool own, bool me) {
        const double nf = -f;
        const uint32_t src = smem_u32(pr), dd = smem_u32(dst);
        constexpr int NP = NC / 2;            // double2 pairs in a row
        constexpr int CH = 2;                 // pairs per chunk
        constexpr int NCH = (NP + CH - 1) / CH;
        double2 v[3][CH];
        static_for<0, CH>([&](auto q) {
            if constexpr (q.value < NP) v[0][q.value] = lds128<16 * q.value>(src);
            if constexpr (CH + q.value < NP) v[1][q.value] = lds128<16 * (CH + q.value)>(src);
        });
        static_for<0, NCH + 1>([&](auto chc) {
            constexpr int ch = chc.value;
            if constexpr (ch + 2 < NCH) {
                static_for<0, CH>([&](auto q) {
                    constexpr int c2 = (ch + 2) * CH + q.value;
                    if constexpr (c2 < NP) v[(ch + 2) % 3][q.value] = lds128<16 * c2>(src);
                });
            }
            if constexpr (ch < NCH) {
                static_for<0, CH>([&](auto q) {
                    constexpr int c2 = ch * CH + q.value;
                    if constexpr (c2 < NP) {
                        T[2 * c2] = fma(nf, v[ch % 3][q.value].x, T[2 * c2]);
                        T[2 * c2 + 1] = fma(nf, v[ch % 3][q.value].y, T[2 * c2 + 1]);
                    }
                });
            }
            if constexpr (ch >= 1) {
                static_for<0, CH>([&](auto q) {
                    constexpr int c2 = (ch - 1) * CH + q.value;
                    if constexpr (c2 < NP) {
                        if (me) sts128<16 * c2>(dd, T[2 * c2], T[2 * c2 + 1]);
                    }
                });
            }
        });
        if constexpr (NC & 1) {
            T[NC - 1] = fma(nf, lds64<8 * (NC - 1)>(src), T[NC - 1]);
            if (me) sts64<8 * (NC - 1)>(dd, T[NC - 1]);
        }
        const double nk = own ? il : -f * il;
        preg_set<NC>(T, k, nk);
        if (me) sts64<0>(dd + 8 * k, nk);
        if (own) { lam = rp; ilam = p; }
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
#ifdef DDB_TIMING
        double tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = clock64();
#endif

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
        int buf = 0;
        int gpar = 0, hpar = 0;       // current buffers of the cost vectors g / ghat

        DDB_TSTAMP(0);
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
            // cost vectors: shared memory, double-buffered; EVERY warp computes every update (identical values), so no
            // warp ever waits for another one to price
            gpar = 0;
            hpar = 0;
            for (int j = tid; j < CT; j += NT) {
                gsm[j] = (j < n) ? __ldg(cg + j) : 0.0;
                ghsm[j] = (j < n) ? 1.0 : 0.0;
            }
            unsigned freebits = 0u;            // bit cs: column lane + 32 cs is still a free x_j (one copy per warp)
#pragma unroll
            for (int cs = 0; cs < CS; ++cs)
                if (lane + 32 * cs < n) freebits |= 1u << cs;
            buf = 0;
            if (tid == 0) publish(prow);
            __syncthreads();

            for (int t = 0; t < n; ++t) {
                const double* pr = prow + buf * PD;
                double* dst = prow + (buf ^ 1) * PD;
                const bool own = (tid == t);
                // every warp: pivot column = largest |entry| of the published row among the free columns (a row is a
                // pivot row once in the crash, so its lam is still 1), and the cost update g -= (g_k / p) row
                const double* gcur = gsm + gpar * CT;
                double* gnext = gsm + (gpar ^ 1) * CT;
                double pl[CS];
                unsigned long long best = 0ull;
                int bq = 0;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    pl[cs] = (j < n) ? pr[j] : 0.0;
                    const unsigned long long kk =
                        ((freebits >> cs) & 1u) ? (unsigned long long)__double_as_longlong(fabs(pl[cs])) : 0ull;
                    if (kk > best) { best = kk; bq = cs; }
                }
                unsigned long long kmin;
                const int kl = warp_argmin_key(~best, kmin);   // argmax through the complemented key
                const int k = kl + 32 * __shfl_sync(FULL, bq, kl);
                const double pabs = __longlong_as_double((long long)~kmin);
                if (!(pabs >= kTolCrash)) { need_generic = true; break; }   // same data in every warp: uniform
                const double p = pr[k];
                const double rp = fast_rcp(p);
                const double fg = gcur[k] * rp;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    if (j < n) gnext[j] = (j == k) ? -fg : fma(-fg, pl[cs], gcur[j]);
                }
                gpar ^= 1;
                if ((k & 31) == lane) freebits &= ~(1u << (k >> 5));
                if (tid == 0) pivcol[t] = k;
                const double e = preg_get<NC>(T, k);
                const double f = own ? 0.0 : e * rp;
                apply(pr, dst, f, k, 1.0, rp, p, own, (tid == t + 1) && (t + 1 < n));   // thread t + 1 publishes its updated row as it goes
                ++npiv_crash;
                buf ^= 1;
                __syncthreads();
            }
        }

        DDB_TSTAMP(1);
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

            DDB_TSTAMP(2);
            // ---- stage 3a: phase 1 (most negative slack leaves; ratio test along its row) -------------------
            // Software-pipelined pivots.  The update of pivot (r, k) is PENDING while the next pivot row is chosen: the
            // right-hand side after the pending update costs one FMA per thread, so the next leaving row r' is known
            // before the rank-1 update runs, and its owner stores its updated row into the other row buffer WHILE it
            // updates it.  Nothing is published serially; two barriers per pivot: (A) candidates -> r', (B) row r' is
            // in shared memory.  The ratio test along row r' and the cost updates are computed by every warp.
            {
                bool pending = false, own = false;
                int k = 0;
                double f = 0.0, p = 1.0, il = 1.0, rp = 1.0;
                for (;;) {
                    const double* pr = prow + buf * PD;
                    double* dst = prow + (buf ^ 1) * PD;
                    double rhs2 = T[RHS], lam2 = lam;
                    if (pending) {
                        if (own) lam2 = rp; else rhs2 = fma(-f, pr[RHS], rhs2);
                    }
                    const double s = lam2 * rhs2;
                    unsigned long long kmin;
                    const int ll = warp_argmin_key((live && s < -kTolFeas) ? dkey(s) : KEY_INF, kmin);
                    if (lane == 0) {
                        hdr[warp].key = kmin;
                        hdr[warp].row = warp * 32 + ll;
                    }
                    __syncthreads();                                   // (A)
                    const int ww = warp_argmin_key((lane < W) ? hdr[lane].key : KEY_INF, kmin);
                    const bool done = (kmin == KEY_INF);               // s >= 0 everywhere after the pending pivot
                    if (!done && npiv_p1 + (pending ? 1 : 0) >= a.max_iter) { status = ST_ITERATION_LIMIT; npiv_p1 += pending ? 1 : 0; break; }
                    const int r2 = done ? -1 : hdr[ww].row;
                    const bool me = (tid == r2);
                    if (pending) {
                        // cost vectors after the pending pivot (every warp, identical values, other buffer)
                        const double* gcur = gsm + gpar * CT;
                        double* gnext = gsm + (gpar ^ 1) * CT;
                        const double* hcur = ghsm + hpar * CT;
                        double* hnext = ghsm + (hpar ^ 1) * CT;
                        const double fv = hcur[k] * rp, fg = gcur[k] * rp;
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) {
                            const int j = lane + 32 * cs;
                            if (j < n) {
                                const double plj = pr[j];
                                hnext[j] = (j == k) ? -fv * il : fma(-fv, plj, hcur[j]);
                                gnext[j] = (j == k) ? -fg * il : fma(-fg, plj, gcur[j]);
                            }
                        }
                        gpar ^= 1;
                        hpar ^= 1;
                        if (own) {
                            const int cv = cvsm[k];       // becomes basic in my row
                            cvsm[k] = rowvar;             // my old slack becomes nonbasic in column k
                            rowvar = cv;
                        }
                        apply(pr, dst, f, k, il, rp, p, own, me);
                        ++npiv_p1;
                    } else if (me) {
                        publish(dst);
                    }
                    if (done) break;
                    if (me) {
                        pub[buf ^ 1].p = lam;             // phase 1 publishes lam of the pivot row here
                        pub[buf ^ 1].il = ilam;
                    }
                    __syncthreads();                                   // (B)
                    buf ^= 1;
                    own = me;
                    // ratio test along the true row lam_r * row (min ghat_j / (-e_j) over e_j < -tol): every warp
                    {
                        const double* prn = prow + buf * PD;
                        const double* hcur = ghsm + hpar * CT;
                        const double lam_r = pub[buf].p;
                        il = pub[buf].il;
                        double bn = 0.0, bd = 0.0;   // best numerator / denominator (bd == 0: none)
                        int bq = 0;
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) {
                            const int j = lane + 32 * cs;
                            const double plj = (j < n) ? prn[j] : 0.0;
                            const double e = -lam_r * plj;
                            if (j < n && e > kTolPivot) {
                                const double num = fmax(hcur[j], 0.0);
                                if (bd == 0.0 || num * bd < bn * e) { bn = num; bd = e; bq = cs; }
                            }
                        }
                        const double ratio = bn * fast_rcp(bd > 0.0 ? bd : 1.0);
                        const int kl = warp_argmin_key((bd > 0.0) ? dkey(ratio) : KEY_INF, kmin);
                        if (kmin == KEY_INF) { status = ST_INFEASIBLE; break; }
                        k = kl + 32 * __shfl_sync(FULL, bq, kl);
                        p = prn[k];
                        rp = fast_rcp(p);
                        const double e = preg_get<NC>(T, k);
                        f = own ? 0.0 : e * rp;
                        pending = true;
                    }
                }
            }
            __syncthreads();

            DDB_TSTAMP(3);
            // ---- stage 3b: phase 2 (Dantzig) ---------------------------------------------------------------
            // Same pipeline, column first: the costs after the pending pivot give the next entering column k' (every
            // warp prices all columns), the entries of column k' and the right-hand sides after the pending pivot cost
            // two FMAs per thread, so the ratio test -- hence the next pivot row r' -- precedes the rank-1 update and the
            // owner of r' publishes its row while updating it.  Barriers: (A) candidates -> r', (B) row r' published.
            if (status == ST_OPTIMAL) {
                bool pending = false, own = false;
                int k = 0, k2 = -1;
                double f = 0.0, p = 1.0, il = 1.0, rp = 1.0;
                {
                    const double* gcur = gsm + gpar * CT;
                    double gmin = kInf;
                    int bj = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        if (j < n) {
                            const double g = gcur[j];
                            if (g < gmin) { gmin = g; bj = j; }
                        }
                    }
                    unsigned long long kmin;
                    const int kl = warp_argmin_key(dkey(gmin), kmin);
                    const int bjw = __shfl_sync(FULL, bj, kl);
                    if (kmin < dkey(-kTolFeas)) k2 = bjw;
                }
                for (;;) {
                    const double* pr = prow + buf * PD;
                    double* dst = prow + (buf ^ 1) * PD;
                    const bool done = (k2 < 0);                        // g >= 0 after the pending pivot: optimal
                    int r2 = -1;
                    double e2 = 0.0;
                    if (!done) {
                        // ratio test on column k2 as it will be after the pending pivot: one row per lane
                        const double eraw = preg_get<NC>(T, k2);
                        double rhs2 = T[RHS], lam2 = lam;
                        e2 = eraw;
                        if (pending) {
                            if (own) {
                                lam2 = rp;
                                if (k2 == k) e2 = il;
                            } else {
                                rhs2 = fma(-f, pr[RHS], rhs2);
                                e2 = (k2 == k) ? -f * il : fma(-f, pr[k2], eraw);
                            }
                        }
                        const double et = lam2 * e2;                   // true entry / right-hand side of my row
                        const double sc = fmax(lam2 * rhs2, 0.0);
                        const bool cand = live && et > kTolPivot;
                        const double ratio = sc * fast_rcp(cand ? et : 1.0);
                        unsigned long long kmin;
                        const int ll = warp_argmin_key(cand ? dkey(ratio) : KEY_INF, kmin);
                        if (lane == 0) {
                            hdr[warp].key = kmin;
                            hdr[warp].row = warp * 32 + ll;
                        }
                        __syncthreads();                               // (A)
                        const int ww = warp_argmin_key((lane < W) ? hdr[lane].key : KEY_INF, kmin);
                        if (kmin == KEY_INF) { status = ST_UNBOUNDED; npiv_p2 += pending ? 1 : 0; break; }
                        if (npiv_p2 + (pending ? 1 : 0) >= a.max_iter) { status = ST_ITERATION_LIMIT; npiv_p2 += pending ? 1 : 0; break; }
                        r2 = hdr[ww].row;
                    }
                    const bool me = (tid == r2);
                    if (pending) {
                        if (own) {
                            const int cv = cvsm[k];       // becomes basic in my row
                            cvsm[k] = rowvar;             // my old slack becomes nonbasic in column k
                            rowvar = cv;
                        }
                        apply(pr, dst, f, k, il, rp, p, own, me);
                        ++npiv_p2;
                    } else if (me) {
                        publish(dst);
                    }
                    if (done) break;
                    if (me) {
                        pub[buf ^ 1].p = e2;              // my (stored) entry in column k2 is the next pivot
                        pub[buf ^ 1].il = ilam;
                    }
                    __syncthreads();                                   // (B)
                    buf ^= 1;
                    own = me;
                    k = k2;
                    p = pub[buf].p;
                    il = pub[buf].il;
                    rp = fast_rcp(p);
                    f = own ? 0.0 : e2 * rp;
                    pending = true;
                    // price the columns as they will be after this pivot: g' = g - (g_k / p) row, g'_k = -g_k / (p lam_r)
                    {
                        const double* prn = prow + buf * PD;
                        const double* gcur = gsm + gpar * CT;
                        double* gnext = gsm + (gpar ^ 1) * CT;
                        const double fg = gcur[k] * rp;
                        double gmin = kInf;
                        int bj = 0;
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) {
                            const int j = lane + 32 * cs;
                            if (j < n) {
                                const double g = (j == k) ? -fg * il : fma(-fg, prn[j], gcur[j]);
                                gnext[j] = g;
                                if (g < gmin) { gmin = g; bj = j; }
                            }
                        }
                        gpar ^= 1;
                        unsigned long long kmin;
                        const int kl = warp_argmin_key(dkey(gmin), kmin);
                        const int bjw = __shfl_sync(FULL, bj, kl);
                        k2 = (kmin < dkey(-kTolFeas)) ? bjw : -1;
                    }
                }
            }
        }

        DDB_TSTAMP(4);
        // ---- stage 4: x, objective, slacks, labels -----------------------------------------------------------------
        __syncthreads();
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0, nref = 0;
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
        DDB_TSTAMP(5);
#ifdef DDB_TIMING
        if (tid == 0)
            for (int q = 0; q < 8; ++q) a.gtab[(size_t)lp * 8 + q] = tacc[q];
#endif
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
struct PipeVariant {
    int NC, W, MINB;
    cudaError_t (*launch)(const SolveArgs&, int, cudaStream_t);
};

template <int NC, int W, int MINB>
cudaError_t launch_pipe_variant(const SolveArgs& a, int sm_count, cudaStream_t st) {
    auto kern = simplex_rowpipe_kernel<NC, W, MINB>;
    const size_t smem = make_pipe_layout(a.m, a.n, NC, W).total;
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
const PipeVariant kPipeVariants[] = {
    {8, 1, 32, launch_pipe_variant<8, 1, 32>},
    {24, 1, 16, launch_pipe_variant<24, 1, 16>},
    {24, 2, 8, launch_pipe_variant<24, 2, 8>},
    {48, 2, 5, launch_pipe_variant<48, 2, 5>},
    {48, 4, 3, launch_pipe_variant<48, 4, 3>},
    {72, 4, 2, launch_pipe_variant<72, 4, 2>},
    {101, 4, 2, launch_pipe_variant<101, 4, 2>},
    {101, 8, 1, launch_pipe_variant<101, 8, 1>},
};

const PipeVariant* pick_pipe_variant(int m, int n) {
    const int rows = (m - n > n) ? (m - n) : n;
    for (const PipeVariant& v : kPipeVariants)
        if (n + 1 <= v.NC && rows <= 32 * v.W) return &v;
    return nullptr;
}
}  // namespace

bool rowpipe_supported(int m, int n) { return m >= n && pick_pipe_variant(m, n) != nullptr; }

cudaError_t launch_simplex_rowpipe(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const PipeVariant* v = pick_pipe_variant(a.m, a.n);
    if (!v) return cudaErrorInvalidValue;
    return v->launch(a, sm_count, st);
}

}  // namespace ddb
