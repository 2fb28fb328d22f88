// Row-per-thread batched fp64 simplex (plan 0): one LP per CTA, ONE TABLEAU ROW PER THREAD.
//
// Why this layout: a pivot is a rank-1 update  T[i][c] -= f_i * prow[c].  With a whole row owned by one thread
//   * the update is NC DFMAs per thread whose second operand (the pivot row) is the same for every thread: it is
//     read from shared memory with broadcast LDS.128 (one wavefront per two entries), so the instruction stream of
//     a pivot is ~NC DFMA + NC/2 LDS per warp and nothing else scales with the row count;
//   * everything that is "per row" (the right-hand side, the entry in the entering column, the ratio test, the
//     most-negative-slack search of phase 1) is one value PER LANE, so a ratio test over 32 rows is a handful of
//     instructions followed by a redux-based warp argmin -- not a loop;
//   * everything that is "per column" (pricing vectors ghat / g, the column -> constraint map) lives in shared
//     memory, one copy per CTA, and is maintained by the warp that owns the pivot row.
//
// HYBRID ROWS (TS > 0, the (200,100) variant): a thread keeps the logical columns [TS, NC) of its row in registers
// (TR = NC - TS of them, the right-hand side last) and the columns [0, TS) in shared memory (its own 16-byte aligned
// row of pitch PS, conflict-free for LDS.128 / STS.128).  That brings a thread from 255 to <= 168 registers, so THREE LPs
// are resident per SM instead of two -- the kernel is latency-bound (one pivot is a serial chain of a branch tree, two
// warp argmins, the owner's publish and two or three barriers), so a third LP per SM is nearly free throughput.  It
// also shortens the chain: the owner lane publishes only its TR register columns (the shared-memory part of a pivot row
// is read in place by everybody), and a dynamic column index below TS is an address instead of a branch tree.
// The crash inverse D (n x NC doubles, 81 KB at n = 100) no longer fits beside three LPs: it is parked in a per-CTA global
// scratch that stays in L2 and streamed back for the product P_N = -A_N D by 1-D bulk TMA (cp.async.bulk -> mbarrier)
// through a ring of NSLOT x G rows -- two passes over D (first the shared-memory columns, accumulated in the idle
// registers and stored once, then the register columns), so the product never read-modify-writes shared memory.
//
// Rows are stored lazily normalised (true row = lam * stored row), so a pivot never rescales the pivot row: its
// owner publishes its raw registers and every other thread runs one FMA per entry.
//
// GEN (fused generate -> solve -> label): the CTA draws its instance itself (philox.cuh: generate_instance_cta) into
// the caller's A / b / c or, when those are not asked for, into a per-CTA slab that is rewritten by every LP and
// therefore lives in L2 -- A never makes an HBM round trip and the crash scores come out of the generator's tile.
//
// Stages per LP (same algorithm and tolerances as simplex_generic.cu):
//   0. crash order by cosine score (A streamed once from HBM through the idle tableau registers; GEN: from the tile)
//   1. crash as an explicit inverse: Gauss-Jordan on the n x n block A_B0 (thread t owns row t), 1 barrier/pivot
//   2. remaining rows enter through  P_N = -A_N D
//   3. phase 1 (most negative slack leaves, ratio test along the published row), phase 2 (Dantzig)
//   4. x = xv - D sigma, slack = b - A x from the caller's A, labels = |slack| <= threshold
// Instances the tile cannot hold or whose static crash basis is singular are flagged status = -1 and re-solved by
// the generic kernel on the device (capi.cu); nothing ever falls back to the CPU.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "philox.cuh"

// dev-only stage accounting (tools/row_timing.cu builds with -DDDB_TIMING; never defined in the library build)
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
#define DDB_R8(M, b) M(b + 0) M(b + 1) M(b + 2) M(b + 3) M(b + 4) M(b + 5) M(b + 6) M(b + 7)
#define DDB_R128(M)                                                                                       \
    DDB_R8(M, 0) DDB_R8(M, 8) DDB_R8(M, 16) DDB_R8(M, 24) DDB_R8(M, 32) DDB_R8(M, 40) DDB_R8(M, 48)       \
    DDB_R8(M, 56) DDB_R8(M, 64) DDB_R8(M, 72) DDB_R8(M, 80) DDB_R8(M, 88) DDB_R8(M, 96) DDB_R8(M, 104)    \
    DDB_R8(M, 112) DDB_R8(M, 120)

template <int NR>
__device__ __forceinline__ double reg_get(const double (&T)[NR], int k) {
    static_assert(NR <= 128, "jump table covers 128 registers");
    double v = 0.0;
    switch (k) {
#define DDB_CASE(I)                                    \
    case (I):                                          \
        if constexpr ((I) < NR) v = T[(I) < NR ? (I) : 0]; \
        break;
        DDB_R128(DDB_CASE)
#undef DDB_CASE
        default: break;
    }
    return v;
}
// Write at a warp-uniform dynamic index.  The asm volatile leaves keep the compiler from if-converting the switch
// into a select per register (which costs 3 instructions per tableau column); what remains is a uniform branch tree.
template <int NR>
__device__ __forceinline__ void reg_set(double (&T)[NR], int k, double v) {
    switch (k) {
#define DDB_CASE(I)                                                                         \
    case (I):                                                                               \
        if constexpr ((I) < NR) asm volatile("mov.f64 %0, %1;" : "=d"(T[(I) < NR ? (I) : 0]) : "d"(v)); \
        break;
        DDB_R128(DDB_CASE)
#undef DDB_CASE
        default: break;
    }
}

struct RowPub {               // what the pivot row's owner (warp) publishes beside the row itself
    double p;                 // pivot entry (stored scale)
    double il;                // 1 / lam of the pivot row before the pivot
    int k;                    // entering column (-1: none -> infeasible / singular crash basis)
    int var;                  // constraint whose slack was basic in the pivot row
};

struct RowHdr {               // one per warp: its candidate row
    unsigned long long key;   // dkey(slack) in phase 1, dkey(ratio) in phase 2, KEY_INF = no candidate
    int row;                  // candidate tile row (= thread index)
    int pad;
};

// Pitch (in doubles) of rows kept in shared memory: even (16-byte rows for LDS.128) with pitch/2 odd, so that
// 8 lanes reading the same 16-byte column of 8 consecutive rows hit 8 different bank groups.
__host__ __device__ constexpr int row_pitch(int NC) {
    int pd = (NC + 1) & ~1;
    if (((pd / 2) & 1) == 0) pd += 2;
    return pd;
}

constexpr int kRingSlots = 4;   // hybrid rows: slots of the bulk-TMA ring that streams D back for the product
constexpr int kRingRows = 4;    // rows of D per slot

struct RowLayout {
    size_t D, ts, ring, bars, prow, pub, hdr, part, gsm, ghsm, cvsm, order, colvar0, pivcol, basic_tile, sval, sig, xbuf, gbuf,
        gnn, red, total;
};
__host__ __device__ inline size_t rr_align(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline RowLayout make_row_layout(int m, int n, int NC, int TS, int W, bool gen) {
    RowLayout L;
    const int PD = row_pitch(NC);
    const int CT = 32 * ((NC + 31) / 32);
    const int rows = (m - n > n) ? (m - n) : n;
    size_t off = 0;
    // region 0 (dead during stage 0, so the in-solver generator's tile [32][n] + x0[n] + c[n] overlays it):
    //   TS == 0: the crash inverse D;   TS > 0: the shared-memory part of the rows + the TMA ring for D
    L.D = off;
    if (TS == 0) off += rr_align((size_t)n * PD * 8);
    L.ts = off;
    if (TS > 0) off += rr_align((size_t)rows * row_pitch(TS) * 8);
    L.ring = off;
    if (TS > 0) off += rr_align((size_t)kRingSlots * kRingRows * PD * 8);
    if (gen) {
        const size_t need = rr_align(gen_smem_doubles(m, n) * 8);
        if (off < need) off = need;
    }
    L.bars = off;       off += rr_align((size_t)2 * kRingSlots * 8);
    L.prow = off;       off += rr_align((size_t)2 * PD * 8);
    L.pub = off;        off += rr_align((size_t)2 * sizeof(RowPub));
    L.hdr = off;        off += rr_align((size_t)W * sizeof(RowHdr));
    L.part = off;       off += rr_align((size_t)W * sizeof(RowHdr));
    L.gsm = off;        off += rr_align((size_t)2 * CT * 8);
    L.ghsm = off;       off += rr_align((size_t)CT * 8);
    L.cvsm = off;       off += rr_align((size_t)CT * 4);
    L.order = off;      off += rr_align((size_t)m * 4);
    L.colvar0 = off;    off += rr_align((size_t)n * 4);
    L.pivcol = off;     off += rr_align((size_t)n * 4);
    L.basic_tile = off; off += rr_align((size_t)m * 4);
    L.sval = off;       off += rr_align((size_t)W * 32 * 8);
    L.sig = off;        off += rr_align((size_t)n * 8);
    L.xbuf = off;       off += rr_align((size_t)(n > CT ? n : CT) * 8);
    L.gbuf = off;       off += rr_align((size_t)m * 8);
    L.gnn = off;        off += rr_align((size_t)m * 8);
    L.red = off;        off += rr_align((size_t)(3 * W + 4) * 4);
    L.total = off;
    return L;
}

// Rare path (ill-conditioned vertex, ~0.06 % of the instances): dot product of a parked tableau row (columns [0, ns) at
// `rs`, columns [ns, n) at `rg`) with a shared-memory vector.  Not inlined, so that it does not take part in the register
// allocation of the pivot loops.
static __device__ __noinline__ double saved_row_dot(const double* rs, int ns, const double* rg, const double* vec, int n) {
    double d = 0.0;
    for (int c = 0; c < n; ++c) d = fma((c < ns) ? rs[c] : rg[c - ns], vec[c], d);
    return d;
}

template <bool GEN>
__device__ __forceinline__ double ldin(const double* p) {
    // instance data: read-only path when the caller supplied it; plain loads when this kernel wrote it (GEN)
    if constexpr (GEN) return *p;
    else return __ldg(p);
}
template <bool GEN>
__device__ __forceinline__ double2 ldin2(const double2* p) {
    if constexpr (GEN) return *p;
    else return __ldg(p);
}

template <int NC, int TS, int W, int MINB, bool GEN>
__global__ void __launch_bounds__(W * 32, MINB) simplex_rowreg_kernel(SolveArgs a) {
    constexpr bool HYB = TS > 0;
    constexpr int TR = NC - TS;             // register-resident columns: logical [TS, NC), right-hand side last
    constexpr int CS = (NC + 31) / 32;      // slots of the lane-distributed column vectors
    constexpr int CT = 32 * CS;
    constexpr int PD = row_pitch(NC);       // pitch of a full logical row (D rows, parked rows when TS == 0)
    constexpr int PRD = (TR + 1) & ~1;      // pitch of a published register part
    constexpr int PS = HYB ? row_pitch(TS) : 2;   // pitch of a thread's shared-memory part
    constexpr int NT = W * 32;              // threads = tile rows
    constexpr int RHSR = TR - 1;            // register that holds the right-hand side
    constexpr int RB = TR / CS;             // rows per register batch when T is used as a streaming buffer
    static_assert(W <= 32, "one header per lane");
    static_assert((TS & 1) == 0 && TS <= TR, "shared-memory part: even, not wider than the register part");
    static_assert(RB >= 1, "streaming buffer");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const RowLayout L = make_row_layout(m, n, NC, TS, W, GEN);
    double* Dsm = reinterpret_cast<double*>(smem_raw + L.D);       // TS == 0 only
    double* Ts = reinterpret_cast<double*>(smem_raw + L.ts);       // TS > 0: [rows][PS]
    double* ring = reinterpret_cast<double*>(smem_raw + L.ring);   // TS > 0: [kRingSlots][kRingRows][PD]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);   // full[kRingSlots], empty[kRingSlots]
    double* prow = reinterpret_cast<double*>(smem_raw + L.prow);
    RowPub* pub = reinterpret_cast<RowPub*>(smem_raw + L.pub);
    RowHdr* hdr = reinterpret_cast<RowHdr*>(smem_raw + L.hdr);
    RowHdr* part = reinterpret_cast<RowHdr*>(smem_raw + L.part);    // phase 2: per-warp partial pricing result
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
    double* Ts_own = Ts + (size_t)tid * PS;     // only dereferenced by threads that own a row

    // ---- register state -------------------------------------------------------------------------------------
    double T[TR];            // the register part of my tableau row; T[RHSR] is its right-hand side
    int rowvar = -1;         // constraint whose slack is basic in my row
    double lam = 1.0, ilam = 1.0;   // my row's lazy scale: true row = lam * T
    int ring_it = 0;         // HYB: iterations of the D ring so far (the mbarrier phases run on across LPs)
    int dk = -1;             // HYB: deferred write of the owner into the shared-memory part of its own row (the others read
    double dv = 0.0;         //      that row in place during the update, so the owner patches column k after the next barrier)

    // Rows are stored LAZILY NORMALISED: the true tableau row is lam * T (ilam = 1 / lam).  A pivot (r, k) never
    // rescales the pivot row: with p = T_r[k] (stored), rp = 1/p, il = ilam_r (before the pivot)
    //     rows i != r :  f = T_i[k] * rp;  T_i[c] -= f * T_r[c] (c != k);  T_i[k] = -f * il       (lam_i unchanged)
    //     row r       :  T_r unchanged except T_r[k] = il;  lam_r = rp, ilam_r = p
    //     costs       :  g[c] -= g_k rp T_r[c] (c != k);  g[k] = -g_k rp il
    // so the owner just publishes its raw registers and every other thread runs one FMA per entry.
    // rank-1 update of my row from the raw pivot row (broadcast reads):  T[c] -= f * prow[c]
    //   prr: register part of the pivot row (published), prs: its shared-memory part (read in place); do_s: I own a
    //   shared-memory part and am not the pivot row's owner
    auto rank1 = [&](const double* prr, const double* prs, double f, bool do_s) {
        const double nf = -f;
        const double2* p2 = reinterpret_cast<const double2*>(prr);
#pragma unroll
        for (int c2 = 0; c2 < TR / 2; ++c2) {
            const double2 v = p2[c2];
            T[2 * c2] = fma(nf, v.x, T[2 * c2]);
            T[2 * c2 + 1] = fma(nf, v.y, T[2 * c2 + 1]);
        }
        if constexpr (TR & 1) T[TR - 1] = fma(nf, prr[TR - 1], T[TR - 1]);
        if constexpr (HYB) {
            if (do_s) {
                const double2* s2 = reinterpret_cast<const double2*>(prs);
                double2* o2 = reinterpret_cast<double2*>(Ts_own);
#pragma unroll
                for (int c2 = 0; c2 < TS / 2; ++c2) {
                    const double2 v = s2[c2];
                    double2 o = o2[c2];
                    o.x = fma(nf, v.x, o.x);
                    o.y = fma(nf, v.y, o.y);
                    o2[c2] = o;
                }
            }
        }
    };
    auto publish = [&](double* pr) {
        double2* p2 = reinterpret_cast<double2*>(pr);
#pragma unroll
        for (int c2 = 0; c2 < TR / 2; ++c2) p2[c2] = make_double2(T[2 * c2], T[2 * c2 + 1]);
        if constexpr (TR & 1) pr[TR - 1] = T[TR - 1];
    };
    // logical column j of pivot row r: its shared-memory part in place, its register part as published at pr
    auto rowval = [&](int r, const double* pr, int j) -> double {
        if constexpr (HYB) {
            const double* src = (j < TS) ? (Ts + (size_t)r * PS + j) : (pr + (j - TS));
            return *src;
        } else {
            return pr[j];
        }
    };
    // my entry in logical column k (warp-uniform k)
    auto col_get = [&](int k) -> double {
        if constexpr (HYB) {
            if (k < TS) return Ts_own[k];
            return reg_get<TR>(T, k - TS);
        } else {
            return reg_get<TR>(T, k);
        }
    };
    // write my entry in logical column k; the owner of the pivot row defers a write into its shared-memory part
    auto col_set = [&](int k, double v, bool own, bool has_s) {
        if constexpr (HYB) {
            if (k < TS) {
                if (own) { dk = k; dv = v; }
                else if (has_s) Ts_own[k] = v;
            } else {
                reg_set<TR>(T, k - TS, v);
            }
        } else {
            reg_set<TR>(T, k, v);
        }
    };
    auto flush_deferred = [&]() {
        if constexpr (HYB) {
            if (dk >= 0) { Ts_own[dk] = dv; dk = -1; }
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
                    T[r * CS + cs] = (i < m && j < n) ? ldin<GEN>(Ag + (size_t)i * n + j) : 0.0;
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

    uint64_t* full = bars;                  // HYB: ring slot filled (1 arrival + transaction bytes)
    uint64_t* empty = bars + kRingSlots;    // HYB: ring slot released (one arrival per warp)
    if constexpr (HYB) {
        if (tid == 0) {
            for (int s = 0; s < kRingSlots; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], W);
            }
            fence_mbar_init();
        }
    }

    for (;;) {
        if (tid == 0) cur_lp = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long lp = cur_lp;
        if (lp >= a.B) break;
        const double* Ag;
        const double* bg;
        const double* cg;
        const uint8_t* mask = a.row_mask ? a.row_mask + (size_t)lp * m : nullptr;
#ifdef DDB_TIMING
        double tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = clock64();
#endif

        // ---- stage 0: crash order ---------------------------------------------------------------------------
        if constexpr (GEN) {
            // draw the instance here: into the caller's arrays when asked for, else into this CTA's slab (L2-resident)
            double* slab = a.slab + (size_t)blockIdx.x * slab_doubles(m, n);
            double* Aw = a.A ? const_cast<double*>(a.A) + (size_t)lp * m * n : slab;
            double* bw = a.A ? const_cast<double*>(a.b) + (size_t)lp * m : slab + slab_b_offset(m, n);
            double* cw = a.A ? const_cast<double*>(a.c) + (size_t)lp * n : slab + slab_c_offset(m, n);
            generate_instance_cta<NC * 1000 + TS * 10 + W>((uint64_t)a.gen_key, (uint64_t)(a.gen_first + lp), m, n, a.gen_density, Aw, bw, cw, nullptr,
                                  reinterpret_cast<double*>(smem_raw), gbuf, gnn);
            Ag = Aw; bg = bw; cg = cw;
        } else {
            Ag = a.A + (size_t)lp * m * n;
            bg = a.b + (size_t)lp * m;
            cg = a.c + (size_t)lp * n;
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
        const bool crow = tid < n;         // I own a row of the crash block
        double* Dg = nullptr;              // HYB: global home of the crash inverse
        if constexpr (HYB) Dg = a.dscr + (size_t)blockIdx.x * n * PD;

        DDB_TSTAMP(0);
        if (!need_generic) {
            // ---- stage 1: thread t < n loads row order[t] of [A | b]; Gauss-Jordan to the inverse --------------
            {
                const bool have = crow;
                const int row = have ? order[tid] : 0;
                const double* Ar = Ag + (size_t)row * n;
                if ((n & 1) == 0 && (reinterpret_cast<size_t>(Ag) & 15) == 0) {
                    const double2* Ar2 = reinterpret_cast<const double2*>(Ar);
                    if constexpr (HYB) {
                        if (have) {
                            double2* o2 = reinterpret_cast<double2*>(Ts_own);
#pragma unroll
                            for (int c2 = 0; c2 < TS / 2; ++c2)
                                o2[c2] = (2 * c2 < n) ? ldin2<GEN>(Ar2 + c2) : make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int c2 = 0; c2 < (TR - 1) / 2; ++c2) {
                        double2 v = make_double2(0.0, 0.0);
                        if (have && TS + 2 * c2 < n) v = ldin2<GEN>(Ar2 + TS / 2 + c2);
                        T[2 * c2] = v.x;
                        T[2 * c2 + 1] = v.y;
                    }
                    if constexpr (((TR - 1) & 1) != 0) T[TR - 2] = 0.0;
                } else {
                    if constexpr (HYB) {
                        if (have)
                            for (int c = 0; c < TS; ++c) Ts_own[c] = (c < n) ? ldin<GEN>(Ar + c) : 0.0;
                    }
#pragma unroll
                    for (int c = 0; c < TR - 1; ++c) T[c] = (have && TS + c < n) ? ldin<GEN>(Ar + TS + c) : 0.0;
                }
                T[RHSR] = have ? ldin<GEN>(bg + row) : 0.0;
                rowvar = have ? row : -1;
                lam = 1.0;
                ilam = 1.0;
            }
            // column vectors live in shared memory (one copy per CTA, maintained by the pivot row's warp)
            for (int j = tid; j < 32 * CS; j += NT) {
                gsm[j] = (j < n) ? ldin<GEN>(cg + j) : 0.0;
                ghsm[j] = (j < n) ? 1.0 : 0.0;
                cvsm[j] = (j < n) ? -1 : -2;          // -1: still a free x_j (crash), -2: not a column
            }
            __syncthreads();

            for (int t = 0; t < n; ++t) {
                double* pr = prow + buf * PD;
                const bool own = (tid == t);
                if (warp == (t >> 5)) {
                    // the owner publishes its raw row; its warp finds the pivot column (largest |entry| among the
                    // free columns; a row is a pivot row once in the crash, so its lam is still 1) and updates g
                    if (own) publish(pr);
                    __syncwarp();
                    double pl[CS];
                    unsigned long long best = 0ull;
                    int bq = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        pl[cs] = (j < n) ? rowval(t, pr, j) : 0.0;
                        const unsigned long long kk =
                            (cvsm[j] == -1) ? (unsigned long long)__double_as_longlong(fabs(pl[cs])) : 0ull;
                        if (kk > best) { best = kk; bq = cs; }
                    }
                    unsigned long long kmin;
                    const int kl = warp_argmin_key(~best, kmin);   // argmax through the complemented key
                    const int k = kl + 32 * __shfl_sync(FULL, bq, kl);
                    const double pabs = __longlong_as_double((long long)~kmin);
                    const double p = rowval(t, pr, k);
                    const double fg = gsm[k] * fast_rcp(p);
                    __syncwarp();
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        if (j < n) gsm[j] = (j == k) ? -fg : fma(-fg, pl[cs], gsm[j]);
                    }
                    if (lane == 0) {
                        pub[buf].p = p;
                        pub[buf].k = (pabs >= kTolCrash) ? k : -1;
                        pivcol[t] = k;
                        cvsm[k] = 0;                  // no longer free
                    }
                }
                __syncthreads();
                flush_deferred();
                const int k = pub[buf].k;
                if (k < 0) { need_generic = true; break; }
                const double p = pub[buf].p;
                const double rp = fast_rcp(p);
                const double e = crow ? col_get(k) : 0.0;
                const double f = own ? 0.0 : e * rp;
                rank1(pr, Ts + (size_t)t * PS, f, crow && !own);
                col_set(k, own ? 1.0 : -f, own, crow);
                if (own) { lam = rp; ilam = p; }
                buf ^= 1;
                ++npiv_crash;
            }
        }

        DDB_TSTAMP(1);
        const bool live = !need_generic && tid < nN;
        if (!need_generic) {
            __syncthreads();
            flush_deferred();
            // dump D' (row of x_k stored at index k; column NC-1 holds the x-vertex) and the column -> constraint map
            if (crow) {
                const int k = pivcol[tid];
                if constexpr (HYB) {
                    // (an L2 evict_last policy on these stores and on the bulk copies that read them back was measured: DRAM traffic
                    // per LP 392 -> 384 KB, throughput -1 % -- the working set of 444 resident LPs, A 71 MB + D 36 MB + parked rows
                    // 25 MB, exceeds the 126 MB L2 either way, and DRAM runs at 2 % of its bandwidth -- so plain stores it is)
                    double* drow = Dg + (size_t)k * PD;
                    double2* d2 = reinterpret_cast<double2*>(drow);
                    const double2* o2 = reinterpret_cast<const double2*>(Ts_own);
#pragma unroll
                    for (int c2 = 0; c2 < TS / 2; ++c2) {
                        const double2 v = o2[c2];
                        d2[c2] = make_double2(v.x * lam, v.y * lam);
                    }
#pragma unroll
                    for (int c2 = 0; c2 < TR / 2; ++c2) d2[TS / 2 + c2] = make_double2(T[2 * c2] * lam, T[2 * c2 + 1] * lam);
                    if constexpr (TR & 1) drow[NC - 1] = T[TR - 1] * lam;
                    fence_proxy_async_all();          // my generic-proxy stores before the async-proxy (TMA) reads below
                } else {
#pragma unroll
                    for (int c = 0; c < TR; ++c) T[c] *= lam;      // true rows of the inverse
                    publish(Dsm + (size_t)k * PD);
                }
                colvar0[k] = rowvar;
                cvsm[k] = rowvar;
            }
            __syncthreads();

            // ---- stage 2: my row of P_N = -A_N D, s_N = b_N - A_N xv -------------------------------------------
            {
                const int myrow = live ? order[n + tid] : 0;
                const double* Ar = Ag + (size_t)myrow * n;
                if constexpr (HYB) {
                    // D comes back from its global home through the bulk-TMA ring.  The mbarriers are initialised once per
                    // kernel and their phases run on across LPs (ring_it), so EVERY warp takes part in every iteration
                    // (waits for the slot, releases it) whether or not it owns live rows.
                    const int ngroups = (n + kRingRows - 1) / kRingRows;
                    const int total = 2 * ngroups;
                    const int nwl = (nN + 31) >> 5;       // warps that own live rows
                    auto issue = [&](int it) {          // tid 0: bulk copy of group ((it - ring_it) % ngroups) of D into slot it % kRingSlots
                        const int g = (it - ring_it) % ngroups, slot = it % kRingSlots;
                        if (it >= kRingSlots)             // every warp has released the slot's previous contents
                            mbar_wait(&empty[slot], (uint32_t)(((it / kRingSlots) - 1) & 1));
                        const int rws = (n - g * kRingRows < kRingRows) ? (n - g * kRingRows) : kRingRows;
                        const uint32_t bytes = (uint32_t)((size_t)rws * PD * 8);
                        mbar_expect_tx(&full[slot], bytes);
                        tma_load_1d(ring + (size_t)slot * kRingRows * PD, Dg + (size_t)g * kRingRows * PD, bytes, &full[slot]);
                    };
                    if (nN > 0) {
                        if (tid == 0)
                            for (int j = 0; j < kRingSlots - 1 && j < total; ++j) issue(ring_it + j);
                        const bool work = warp < nwl;
#pragma unroll 1
                        for (int pass = 0; pass < 2; ++pass) {
                            // pass 0: the shared-memory columns [0, TS), accumulated in the idle registers and stored once;
                            // pass 1: the register columns [TS, NC)
#pragma unroll
                            for (int c = 0; c < TR; ++c) T[c] = 0.0;
                            double an[kRingRows];
#pragma unroll
                            for (int q = 0; q < kRingRows; ++q) an[q] = (live && q < n) ? ldin<GEN>(Ar + q) : 0.0;
#pragma unroll 1
                            for (int g = 0; g < ngroups; ++g) {
                                const int lit = pass * ngroups + g, it = ring_it + lit, slot = it % kRingSlots;
                                // keep kRingSlots - 1 groups in flight: request the group that reuses the slot released last
                                if (tid == 0 && lit + kRingSlots - 1 < total) issue(it + kRingSlots - 1);
                                __syncwarp();
                                double av[kRingRows];
#pragma unroll
                                for (int q = 0; q < kRingRows; ++q) {
                                    av[q] = an[q];
                                    const int kn = (g + 1) * kRingRows + q;
                                    an[q] = (live && kn < n) ? ldin<GEN>(Ar + kn) : 0.0;
                                }
                                mbar_wait(&full[slot], (uint32_t)((it / kRingSlots) & 1));
                                const double* Dr = ring + (size_t)slot * kRingRows * PD;
                                if (work) {
                                    if (pass == 0) {
#pragma unroll
                                        for (int q = 0; q < kRingRows; ++q) {
                                            if (g * kRingRows + q < n) {
                                                const double nf = -av[q];
                                                const double2* p2 = reinterpret_cast<const double2*>(Dr + (size_t)q * PD);
#pragma unroll
                                                for (int c2 = 0; c2 < TS / 2; ++c2) {
                                                    const double2 v = p2[c2];
                                                    T[2 * c2] = fma(nf, v.x, T[2 * c2]);
                                                    T[2 * c2 + 1] = fma(nf, v.y, T[2 * c2 + 1]);
                                                }
                                            }
                                        }
                                    } else {
#pragma unroll
                                        for (int q = 0; q < kRingRows; ++q) {
                                            if (g * kRingRows + q < n) {
                                                const double nf = -av[q];
                                                const double* dr = Dr + (size_t)q * PD + TS;
                                                const double2* p2 = reinterpret_cast<const double2*>(dr);
#pragma unroll
                                                for (int c2 = 0; c2 < TR / 2; ++c2) {
                                                    const double2 v = p2[c2];
                                                    T[2 * c2] = fma(nf, v.x, T[2 * c2]);
                                                    T[2 * c2 + 1] = fma(nf, v.y, T[2 * c2 + 1]);
                                                }
                                                if constexpr (TR & 1) T[TR - 1] = fma(nf, dr[TR - 1], T[TR - 1]);
                                            }
                                        }
                                    }
                                }
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&empty[slot]);
                            }
                            if (pass == 0 && live) {
                                double2* o2 = reinterpret_cast<double2*>(Ts_own);
#pragma unroll
                                for (int c2 = 0; c2 < TS / 2; ++c2) o2[c2] = make_double2(T[2 * c2], T[2 * c2 + 1]);
                            }
                        }
                        ring_it += total;
                    } else {
#pragma unroll
                        for (int c = 0; c < TR; ++c) T[c] = 0.0;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < TR; ++c) T[c] = 0.0;
                    if (warp * 32 < nN) {
                        double an[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) an[q] = (live && q < n) ? ldin<GEN>(Ar + q) : 0.0;
                        for (int k0 = 0; k0 < n; k0 += 4) {
                            double av[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                av[q] = an[q];
                                an[q] = (live && k0 + 4 + q < n) ? ldin<GEN>(Ar + k0 + 4 + q) : 0.0;
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (k0 + q < n) rank1(Dsm + (size_t)(k0 + q) * PD, nullptr, av[q], false);
                        }
                    }
                }
                if (live) T[RHSR] += ldin<GEN>(bg + myrow);
                rowvar = live ? myrow : -1;
                lam = 1.0;
                ilam = 1.0;
            }

            DDB_TSTAMP(2);
            // ---- stage 3a: phase 1 (most negative slack leaves; ratio test along its row) -------------------
            for (;;) {
                const double s = lam * T[RHSR];
                unsigned long long kmin;
                const int ll = warp_argmin_key((live && s < -kTolFeas) ? dkey(s) : KEY_INF, kmin);
                if (lane == 0) {
                    hdr[warp].key = kmin;
                    hdr[warp].row = warp * 32 + ll;
                }
                __syncthreads();
                flush_deferred();
                const int ww = warp_argmin_key((lane < W) ? hdr[lane].key : KEY_INF, kmin);
                if (kmin == KEY_INF) break;                       // s >= 0 everywhere: phase 1 finished
                if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int r = hdr[ww].row;
                double* pr = prow + buf * PD;
                const bool own = (tid == r);
                if (warp == (r >> 5)) {
                    // owner publishes its raw row; its warp runs the ratio test along the true row lam_r * T_r
                    // (min ghat_j / (-e_j) over e_j < -tol) and updates ghat, g and the column map
                    if (own) publish(pr);
                    __syncwarp();
                    const double lam_r = __shfl_sync(FULL, lam, r & 31);
                    const double il = __shfl_sync(FULL, ilam, r & 31);
                    const int var_r = __shfl_sync(FULL, rowvar, r & 31);
                    double pl[CS], gh[CS];
                    double bn = 0.0, bd = 0.0;   // best numerator / denominator (bd == 0: none)
                    int bq = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        pl[cs] = (j < n) ? rowval(r, pr, j) : 0.0;
                        gh[cs] = ghsm[j];
                        const double e = -lam_r * pl[cs];
                        if (j < n && e > kTolPivot) {
                            const double num = fmax(gh[cs], 0.0);
                            if (bd == 0.0 || num * bd < bn * e) { bn = num; bd = e; bq = cs; }
                        }
                    }
                    const double ratio = bn * fast_rcp(bd > 0.0 ? bd : 1.0);
                    const int kl = warp_argmin_key((bd > 0.0) ? dkey(ratio) : KEY_INF, kmin);
                    const bool none = (kmin == KEY_INF);
                    const int k = none ? 0 : kl + 32 * __shfl_sync(FULL, bq, kl);
                    const double p = none ? 1.0 : rowval(r, pr, k);
                    const double rp = fast_rcp(p);
                    const double fv = ghsm[k] * rp, fg = gsm[k] * rp;
                    const int cv = cvsm[k];
                    __syncwarp();
                    if (!none) {
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) {
                            const int j = lane + 32 * cs;
                            if (j < n) {
                                ghsm[j] = (j == k) ? -fv * il : fma(-fv, pl[cs], gh[cs]);
                                gsm[j] = (j == k) ? -fg * il : fma(-fg, pl[cs], gsm[j]);
                            }
                        }
                    }
                    if (lane == 0) {
                        pub[0].p = p;
                        pub[0].il = il;
                        pub[0].k = none ? -1 : k;
                        pub[0].var = cv;              // becomes basic in the pivot row
                        if (!none) cvsm[k] = var_r;   // becomes nonbasic in column k
                    }
                }
                __syncthreads();
                const int k = pub[0].k;
                if (k < 0) { status = ST_INFEASIBLE; break; }
                const double p = pub[0].p, il = pub[0].il;
                const double rp = fast_rcp(p);
                const double e = live ? col_get(k) : 0.0;
                const double f = own ? 0.0 : e * rp;
                rank1(pr, Ts + (size_t)r * PS, f, live && !own);
                col_set(k, own ? il : -f * il, own, live);
                if (own) { lam = rp; ilam = p; rowvar = pub[0].var; }
                buf ^= 1;
                ++npiv_p1;
            }
            __syncthreads();
            flush_deferred();

            DDB_TSTAMP(3);
            // ---- stage 3b: phase 2 (Dantzig) ---------------------------------------------------------------
            // Three short barriers per pivot: (A) candidates -> winner row, (B) the owner lane has published its raw
            // row, (C) rank-1 update done and the next entering column priced.  Pricing is spread over the warps
            // (warp w updates / scans the columns of slot w of a double-buffered cost vector) so that the only
            // serial section of a pivot is the owner lane's stores.
            int k = -1;
            int gpar = 0;                 // current cost vector = gsm + gpar * CT
            if (status == ST_OPTIMAL) {
                double gmin = kInf;
                int bq = 0;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    const double g = gsm[j];
                    if (j < n && g < gmin) { gmin = g; bq = cs; }
                }
                unsigned long long kmin;
                const int kl = warp_argmin_key(dkey(gmin), kmin);
                if (kmin < dkey(-kTolFeas)) k = kl + 32 * __shfl_sync(FULL, bq, kl);
            }
            while (status == ST_OPTIMAL && k >= 0) {
                // ratio test: one row per lane
                const double e = live ? col_get(k) : 0.0;
                const double et = lam * e;                         // true entry / right-hand side of my row
                const double sc = fmax(lam * T[RHSR], 0.0);
                const bool cand = live && et > kTolPivot;
                const double ratio = sc * fast_rcp(cand ? et : 1.0);
                unsigned long long kmin;
                const int ll = warp_argmin_key(cand ? dkey(ratio) : KEY_INF, kmin);
                if (lane == 0) {
                    hdr[warp].key = kmin;
                    hdr[warp].row = warp * 32 + ll;
                }
                __syncthreads();                                   // (A)
                const int ww = warp_argmin_key((lane < W) ? hdr[lane].key : KEY_INF, kmin);
                if (kmin == KEY_INF) { status = ST_UNBOUNDED; break; }
                if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int r = hdr[ww].row;
                double* pr = prow + buf * PD;
                const bool own = (tid == r);
                double rp_own = 0.0;
                if (own) {
                    publish(pr);
                    pub[0].p = e;                                  // my entry in column k is the pivot
                    pub[0].il = ilam;
                    const int cv = cvsm[k];                        // becomes basic in my row
                    cvsm[k] = rowvar;                              // my old slack becomes nonbasic in column k
                    rowvar = cv;
                    rp_own = 1.0;
                }
                __syncthreads();                                   // (B)
                const double p = pub[0].p, il = pub[0].il;
                const double rp = fast_rcp(p);
                {
                    // my warp's share of the pricing: g' = g - (g_k / p) row, g'_k = -g_k / (p lam_r)
                    const double* gcur = gsm + gpar * CT;
                    double* gnext = gsm + (gpar ^ 1) * CT;
                    const double fg = gcur[k] * rp;
                    double gmin = kInf;
                    int bj = 0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        if ((cs % W) == warp) {
                            const int j = lane + 32 * cs;
                            if (j < n) {
                                const double g = (j == k) ? -fg * il : fma(-fg, rowval(r, pr, j), gcur[j]);
                                gnext[j] = g;
                                if (g < gmin) { gmin = g; bj = j; }
                            }
                        }
                    }
                    const int kl = warp_argmin_key(dkey(gmin), kmin);
                    const int bjw = __shfl_sync(FULL, bj, kl);
                    if (lane == 0) {
                        part[warp].key = kmin;
                        part[warp].row = bjw;
                    }
                }
                const double f = own ? 0.0 : e * rp;
                rank1(pr, Ts + (size_t)r * PS, f, live && !own);
                col_set(k, own ? il : -f * il, own, live);
                if (rp_own != 0.0) { lam = rp; ilam = p; }
                __syncthreads();                                   // (C)
                flush_deferred();
                const int w2 = warp_argmin_key((lane < W) ? part[lane].key : KEY_INF, kmin);
                k = (kmin < dkey(-kTolFeas)) ? part[w2].row : -1;
                gpar ^= 1;
                buf ^= 1;
                ++npiv_p2;
            }
        }

        DDB_TSTAMP(4);
        // ---- stage 4: x, objective, slacks, labels -----------------------------------------------------------------
        __syncthreads();
        flush_deferred();
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0;
        if (need_generic) {
            status = -1;   // re-solved by the generic kernel (capi.cu)
        } else if (status == ST_OPTIMAL) {
            const double* Dr = HYB ? Dg : Dsm;      // crash inverse: global (L2) for hybrid rows, shared memory otherwise
            // where does every constraint sit now?
            if (tid < nN) {
                sval[tid] = lam * T[RHSR];
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
                        if (j < n) acc = fma(Dr[(size_t)k * PD + j], sl[cs], acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) xbuf[k] = Dr[(size_t)k * PD + NC - 1] - acc;
                }
            }
            __syncthreads();
            double xl[CS];
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
                const int j = lane + 32 * cs;
                xl[cs] = (j < n) ? xbuf[j] : 0.0;
            }
            // the register part of my row goes to the per-CTA scratch (fire-and-forget stores) before T becomes a streaming
            // buffer: the rare ill-conditioned instance reloads it for one step of iterative refinement
            // (stores carry an L2 evict_last policy: the per-CTA scratch is rewritten by every LP and should stay in the
            // 126 MB L2 instead of being written back to HBM behind the streaming reads of A)
            double* tsave = a.gtab + ((size_t)blockIdx.x * NT + tid) * PRD;
            if (tid < nN) {
                uint64_t pol;
                asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
                for (int c2 = 0; c2 < TR / 2; ++c2)
                    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(tsave + 2 * c2), "d"(T[2 * c2]),
                                 "d"(T[2 * c2 + 1]), "l"(pol)
                                 : "memory");
                if constexpr (TR & 1)
                    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(tsave + TR - 1), "d"(T[TR - 1]), "l"(pol) : "memory");
            }
            auto write_x_obj = [&]() {
                if (warp == 0) {
                    double acc = 0.0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        if (j < n) acc = fma(ldin<GEN>(cg + j), xl[cs], acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0 && a.obj) a.obj[lp] = acc;
                }
                if (a.x)
                    for (int j = tid; j < n; j += NT) a.x[(size_t)lp * n + j] = xbuf[j];
            };
            // labels exactly as gurobi_lp.py:435-443 from the caller's A; returns whether an active (nonbasic) row has a
            // visible residual at this x
            auto label_pass = [&]() -> int {
                row_dots(Ag, xl, gbuf, nullptr);          // gbuf[i] = a_i . x   (T is a streaming buffer from here on)
                __syncthreads();
                nact = 0; nties = 0; nviol = 0;
                int nref = 0;
                for (int i = tid; i < m; i += NT) {
                    const double slack = ldin<GEN>(bg + i) - gbuf[i];
                    const double as = fabs(slack);
                    const int active = as <= a.thr;
                    lab[i] = (uint8_t)active;
                    nact += active;
                    int tie = (as >= a.thr * 0.1 && as <= a.thr * 10.0);
                    const bool excl = mask && mask[i] == 0;
                    if (!excl) tie |= (active != (basic_tile[i] < 0));
                    nties += tie;
                    nviol += (slack < -a.thr);
                    nref += (!excl && basic_tile[i] < 0 && as > a.thr * 0.01);
                }
                return nref;
            };
            write_x_obj();
            const int nref = label_pass();
            if (__syncthreads_or(nref > 0)) {
                // ---- one step of iterative refinement on the final active set (same arithmetic as simplex_generic.cu) ------
                // rho_j = slack of the active (nonbasic) constraint of column j at the computed x; it should be 0.  At an
                // ill-conditioned vertex it is not: move the nonbasic slacks from rho to 0 through the tableau and correct x
                // through the crash inverse (second-order accurate).
                double* rho = gnn;            // the crash scores are dead by now
                int* colpos = order;          // so is the crash order: constraint -> column where its slack is nonbasic
                for (int j = tid; j < n; j += NT) {
                    const int q = cvsm[j];
                    rho[j] = ldin<GEN>(bg + q) - gbuf[q];
                    colpos[q] = j;
                }
                __syncthreads();
                if (tid < nN) sval[tid] = lam * saved_row_dot(Ts_own, TS, tsave, rho, n);      // true row . rho
                __syncthreads();
                for (int j0 = tid; j0 < n; j0 += NT) {
                    const int q0 = colvar0[j0];
                    const int bt = basic_tile[q0];
                    sig[j0] = (bt >= 0) ? sval[bt] : -rho[colpos[q0]];
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
                            if (j < n) acc = fma(Dr[(size_t)k * PD + j], sl[cs], acc);
                        }
                        acc = warp_sum(acc);
                        if (lane == 0) xbuf[k] -= acc;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    xl[cs] = (j < n) ? xbuf[j] : 0.0;
                }
                write_x_obj();
                label_pass();
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
        DDB_TSTAMP(5);
#ifdef DDB_TIMING
        if (tid == 0)
            for (int q = 0; q < 8; ++q) a.gtab[(size_t)gridDim.x * NT * PRD + (size_t)lp * 8 + q] = tacc[q];   // behind the saved rows
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
// host side (shared by simplex_rowreg.cu = caller-supplied instances and simplex_rowreg_gen.cu = in-solver generator)
// ---------------------------------------------------------------------------------------------------------
struct RowVariant {
    int NC, TS, W, MINB;
    cudaError_t (*launch)(const SolveArgs&, int, cudaStream_t);
    int (*ctas_per_sm)(int m, int n);
};

template <int NC, int TS, int W, int MINB, bool GEN>
int row_variant_ctas_per_sm(int m, int n) {
    auto kern = simplex_rowreg_kernel<NC, TS, W, MINB, GEN>;
    const size_t smem = make_row_layout(m, n, NC, TS, W, GEN).total;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem) != cudaSuccess) return 0;
    return per_sm;
}

template <int NC, int TS, int W, int MINB, bool GEN>
cudaError_t launch_row_variant(const SolveArgs& a, int sm_count, cudaStream_t st) {
    auto kern = simplex_rowreg_kernel<NC, TS, W, MINB, GEN>;
    const size_t smem = make_row_layout(a.m, a.n, NC, TS, W, GEN).total;
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

#define DDB_ROW_VARIANT(NC, TS, W, MINB, GEN) \
    { NC, TS, W, MINB, launch_row_variant<NC, TS, W, MINB, GEN>, row_variant_ctas_per_sm<NC, TS, W, MINB, GEN> }

// (columns incl. rhs, shared-memory columns, warps, min CTAs/SM).  Picked: smallest NC >= n + 1, then smallest W with
// 32 W >= max(n, m - n).  Hybrid variants (TS > 0) come first within their NC; DDB_ROWREG_HYBRID=0 skips them.
template <bool GEN>
const RowVariant* row_variants(int* count) {
    static const RowVariant v[] = {
#ifndef DDB_ROWREG_ONLY_BIG
        DDB_ROW_VARIANT(8, 0, 1, 32, GEN),  DDB_ROW_VARIANT(24, 0, 1, 16, GEN), DDB_ROW_VARIANT(24, 0, 2, 8, GEN),
        DDB_ROW_VARIANT(48, 0, 2, 5, GEN),  DDB_ROW_VARIANT(48, 0, 4, 3, GEN),  DDB_ROW_VARIANT(72, 0, 4, 2, GEN),
#endif
        DDB_ROW_VARIANT(101, 46, 4, 3, GEN), DDB_ROW_VARIANT(101, 0, 4, 2, GEN),
#ifndef DDB_ROWREG_ONLY_BIG
        DDB_ROW_VARIANT(101, 46, 6, 2, GEN),      // up to 192 live rows: hybrid rows, two LPs per SM (DDB_ROWREG_HYBRID=0: the next one)
        DDB_ROW_VARIANT(101, 0, 8, 1, GEN),
        // up to 384 live rows ((400,100): the m/n = 4 cells of the configs[2] sweep): hybrid rows at the register budget of
        // the three-LPs-per-SM variant (65 536 / 384 = 170), one LP per SM, the shared-memory columns take rows x 46 x 8 bytes
        DDB_ROW_VARIANT(101, 46, 12, 1, GEN),
        // 100 < n <= 150 ((300,150), (260,130)): hybrid rows, 77 columns in registers and 74 in shared memory, one LP per SM
        DDB_ROW_VARIANT(151, 74, 5, 1, GEN), DDB_ROW_VARIANT(151, 74, 8, 1, GEN),
#endif
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}

template <bool GEN>
const RowVariant* pick_row_variant(int m, int n) {
    static const bool hybrid = [] { const char* e = getenv("DDB_ROWREG_HYBRID"); return !(e && e[0] == '0'); }();
    const int rows = (m - n > n) ? (m - n) : n;
    int cnt = 0;
    const RowVariant* v = row_variants<GEN>(&cnt);
    for (int i = 0; i < cnt; ++i) {
        if (v[i].TS > 0 && !hybrid) continue;
        if (n + 1 <= v[i].NC && rows <= 32 * v[i].W) return &v[i];
    }
    return nullptr;
}

}  // namespace ddb
