// Column-block-per-warp batched fp64 simplex (plan 7): one LP per CTA of four warps, the condensed tableau in the
// REGISTER FILE as a 2-D distribution -- tableau ROWS over the LANES (four rows per lane), tableau COLUMNS over the WARPS
// (25 columns per warp plus a private copy of the right-hand side).  A thread owns a 4 x 26 register tile.
//
// Why (measured on plan 0, profiles/ncu_r02_rowreg_summary.txt): with one whole row per thread every FMA of the rank-1
// update needs its own broadcast operand from shared memory (97 shared-memory instructions per 101 DFMAs per warp and
// pivot, load/store pipe 65 % busy), every warp repeats the per-pivot bookkeeping, and one lane publishes a whole row.
// Here
//   * a broadcast LDS.128 of the pivot row feeds EIGHT DFMAs (two columns x four rows): 13 loads per 104 DFMAs;
//   * a warp needs the pivot row only in ITS columns, and it owns them: the owner lane publishes 26 doubles per warp
//     (13 STS.128) into a warp-private buffer behind a __syncwarp -- no block barrier between choosing the row and
//     using it;
//   * the only quantity that crosses warps is the column of multipliers f_i = T_i[k] / p (one double per row).  Every
//     warp computes it SPECULATIVELY for its own best candidate column before the pivot's single __syncthreads;
//     afterwards all warps pick the same winner and read its multipliers.  One barrier per pivot in all three stages,
//     and no warp idles while another one runs a serial section;
//   * per-row state (lazy row scale lam, 1 / lam) is replicated per warp, per-column state (reduced costs g, ghat) sits
//     in the lane that owns the column; the right-hand side copies stay bitwise identical because every warp applies
//     the same operations to them, so phase 1's leaving row is found by every warp on its own.
//
// Same algorithm, tolerances and arithmetic per tableau entry as rowreg_kernel.cuh: static-order
// crash as an explicit inverse (kept in shared memory, 4 blocks of 26 per row), P_N = -A_N D, phase 1 (most negative
// slack leaves), phase 2 (Dantzig), lazily normalised rows; exact ties go to the lowest column / lowest tile row.
// Covers n <= 100, m - n <= 128, with or without row masks (reduced LPs); instances it cannot finish (singular static crash basis,
// ill-conditioned vertex) are flagged status = -1 and re-solved by the generic kernel on the device (capi.cu).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "philox.cuh"

namespace ddb {
namespace {

constexpr int QW = 4;                 // warps = column blocks
constexpr int QC = 25;                // tableau columns per warp
constexpr int QP = QC + 1;            // + right-hand side: a thread's row slice, 13 x 16 bytes
constexpr int QR = 4;                 // rows per lane
constexpr int QNT = QW * 32;          // threads
constexpr int QROWS = 32 * QR;        // tile rows
constexpr int QNMAX = QW * QC;        // columns
constexpr int QMMAX = 256;            // constraints (order / score arrays)
constexpr int QDP = QW * QP;          // pitch of a row of D in shared memory
constexpr unsigned long long QINF = 0x7ff0000000000000ull;   // bit pattern of +inf: "no candidate" among non-negative ratios
constexpr int kQuadGenChains = 1;     // Box-Muller chains per thread of the in-solver generator (measured: 1 chain 522 k LP/s, 2 chains 495 k, 4 chains 459 k)
constexpr int QCS = 4;                // column slots of the lane-distributed vectors of stages 0 and 4 (j = lane + 32 cs)

struct QCand {                        // one per (buffer, warp): the warp's speculative candidate
    unsigned long long unused;        // (the keys live in their own packed array keyS)
    int k;                            // entering column
    int r;                            // phase 2: leaving tile row (-1: column unblocked)
    double p;                         // pivot entry (stored scale)
    double gk;                        // g[k]
    double ghk;                       // ghat[k] (phase 1)
    double rp;                        // 1 / p as the producer computed it
    double pad[2];
};
static_assert(sizeof(QCand) == 64, "candidate record");

// fixed shared-memory layout (every offset is an immediate)
constexpr size_t qalign(size_t v) { return (v + 15) / 16 * 16; }
constexpr size_t Q_D = 0;
constexpr size_t Q_PROW = Q_D + qalign((size_t)QNMAX * QDP * 8);
constexpr size_t Q_F = Q_PROW + qalign((size_t)QW * QP * 8);
constexpr size_t Q_CAND = Q_F + qalign((size_t)2 * QW * QROWS * 8);
constexpr size_t Q_KEYS = Q_CAND + qalign((size_t)2 * QW * sizeof(QCand));
constexpr size_t Q_LAM = Q_KEYS + qalign((size_t)2 * QW * 8);
constexpr size_t Q_ILAM = Q_LAM + qalign((size_t)QW * QROWS * 8);
constexpr size_t Q_SVAL = Q_ILAM + qalign((size_t)QW * QROWS * 8);
constexpr size_t Q_SIG = Q_SVAL + qalign((size_t)QROWS * 8);
constexpr size_t Q_XBUF = Q_SIG + qalign((size_t)QROWS * 8);
constexpr size_t Q_GBUF = Q_XBUF + qalign((size_t)QROWS * 8);
constexpr size_t Q_GNN = Q_GBUF + qalign((size_t)QMMAX * 8);
constexpr size_t Q_ROWVAR = Q_GNN + qalign((size_t)QMMAX * 8);
constexpr size_t Q_CV = Q_ROWVAR + qalign((size_t)QROWS * 4);
constexpr size_t Q_COLVAR0 = Q_CV + qalign((size_t)QROWS * 4);
constexpr size_t Q_PIVCOL = Q_COLVAR0 + qalign((size_t)QROWS * 4);
constexpr size_t Q_ORDER = Q_PIVCOL + qalign((size_t)QROWS * 4);
constexpr size_t Q_BASIC = Q_ORDER + qalign((size_t)QMMAX * 4);
constexpr size_t Q_RED = Q_BASIC + qalign((size_t)QMMAX * 4);
constexpr size_t Q_CUR = Q_RED + qalign((size_t)(3 * QW + 4) * 4);
constexpr size_t Q_TOTAL = Q_CUR + 16;

#define QCOLS(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16) M(17) M(18) \
    M(19) M(20) M(21) M(22) M(23) M(24)

// my four entries in column kk of my block (warp-uniform kk): a jump table whose leaves are four moves
__device__ __forceinline__ void col_get4(const double (&T)[QR][QP], int kk, double (&e)[QR]) {
    e[0] = e[1] = e[2] = e[3] = 0.0;
    switch (kk) {
#define QCASE(I)                                                   \
    case I:                                                        \
        asm volatile("mov.f64 %0, %1;" : "=d"(e[0]) : "d"(T[0][I])); \
        asm volatile("mov.f64 %0, %1;" : "=d"(e[1]) : "d"(T[1][I])); \
        asm volatile("mov.f64 %0, %1;" : "=d"(e[2]) : "d"(T[2][I])); \
        asm volatile("mov.f64 %0, %1;" : "=d"(e[3]) : "d"(T[3][I])); \
        break;
        QCOLS(QCASE)
#undef QCASE
        default: break;
    }
}
__device__ __forceinline__ void col_set4(double (&T)[QR][QP], int kk, const double (&v)[QR]) {
    switch (kk) {
#define QCASE(I)                                                   \
    case I:                                                        \
        asm volatile("mov.f64 %0, %1;" : "=d"(T[0][I]) : "d"(v[0])); \
        asm volatile("mov.f64 %0, %1;" : "=d"(T[1][I]) : "d"(v[1])); \
        asm volatile("mov.f64 %0, %1;" : "=d"(T[2][I]) : "d"(v[2])); \
        asm volatile("mov.f64 %0, %1;" : "=d"(T[3][I]) : "d"(v[3])); \
        break;
        QCOLS(QCASE)
#undef QCASE
        default: break;
    }
}
__device__ __forceinline__ void store_row(double* pr, const double (&row)[QP]) {
    double2* p2 = reinterpret_cast<double2*>(pr);
#pragma unroll
    for (int c2 = 0; c2 < QP / 2; ++c2) p2[c2] = make_double2(row[2 * c2], row[2 * c2 + 1]);
}
// the lane that owns tile row (lane, slot q) stores its slice of that row (warp-uniform q)
__device__ __forceinline__ void publish_slot(double* pr, const double (&T)[QR][QP], int q) {
    switch (q) {
        case 0: store_row(pr, T[0]); break;
        case 1: store_row(pr, T[1]); break;
        case 2: store_row(pr, T[2]); break;
        default: store_row(pr, T[3]); break;
    }
}
__device__ __forceinline__ double sel4(const double (&v)[QR], int q) {
    const double a = (q & 1) ? v[1] : v[0], b = (q & 1) ? v[3] : v[2];
    return (q & 2) ? b : a;
}

// Sums of RB per-lane values over the warp with a halving butterfly: RB - 1 + (5 - log2 RB) shuffles instead of 5 RB.  The
// addition tree of every sum is the xor-16, 8, 4, 2, 1 tree of warp_sum (fp addition is commutative), so the results are
// bitwise the same.  The lane ends up with the sum of value  rsel = its lane bits 4.. (RB = 8: bits 4, 3, 2; RB = 2: bit 4).
template <int RB>
__device__ __forceinline__ double reduce_rows(const double (&v)[RB], int lane, int& rsel) {
    static_assert(RB == 8 || RB == 2, "batch sizes of row_dots");
    double w1;
    if constexpr (RB == 8) {
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        double w4[4], w2[2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double send = b4 ? v[j] : v[j + 4], keep = b4 ? v[j + 4] : v[j];
            w4[j] = keep + __shfl_xor_sync(FULL, send, 16);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double send = b3 ? w4[j] : w4[j + 2], keep = b3 ? w4[j + 2] : w4[j];
            w2[j] = keep + __shfl_xor_sync(FULL, send, 8);
        }
        {
            const double send = b2 ? w2[0] : w2[1], keep = b2 ? w2[1] : w2[0];
            w1 = keep + __shfl_xor_sync(FULL, send, 4);
        }
        rsel = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
    } else {
        const bool b4 = lane & 16;
        const double send = b4 ? v[0] : v[1], keep = b4 ? v[1] : v[0];
        w1 = keep + __shfl_xor_sync(FULL, send, 16);
        w1 += __shfl_xor_sync(FULL, w1, 8);
        w1 += __shfl_xor_sync(FULL, w1, 4);
        rsel = b4 ? 1 : 0;
    }
    w1 += __shfl_xor_sync(FULL, w1, 2);
    w1 += __shfl_xor_sync(FULL, w1, 1);
    return w1;
}

// instance data: read-only path when the caller supplied it; plain loads when this kernel wrote it (GEN)
template <bool GEN>
__device__ __forceinline__ double qld(const double* p) {
    if constexpr (GEN) return *p;
    else return __ldg(p);
}

// GEN (fused generate -> solve -> label): the CTA draws its instance itself (philox.cuh: generate_instance_cta, the same
// counters and summation order as the generator kernels, so the same bits) into the caller's A / b / c or, when those are
// not asked for, into a per-CTA slab that is rewritten by every LP and therefore lives in L2 -- A never makes an HBM round
// trip, no generator kernel runs, and the crash scores come out of the generator's shared-memory tile.  More independent Box-Muller
// chains per thread were measured slower here too (rounding-up waste, instruction-cache pressure), although the tile registers are idle at stage 0.
// MASK: reduced LPs (row masks) -- its own instantiation, so that the unmasked kernel carries no trace of it (the kernel sits
// at the 255-register limit: the two live registers of the mask pointer cost 3.7 % of the throughput)
template <bool GEN, bool MASK>
__global__ void __launch_bounds__(QNT, 2) simplex_quadcol_kernel(SolveArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Dsm = reinterpret_cast<double*>(smem_raw + Q_D);           // [n][QW][QP]: crash inverse, x-vertex in slot QC of every block
    double* prowS = reinterpret_cast<double*>(smem_raw + Q_PROW);      // [QW][QP]: a warp's slice of the current pivot row
    double* fS = reinterpret_cast<double*>(smem_raw + Q_F);            // [2][QW][QROWS]: speculative multipliers
    QCand* candS = reinterpret_cast<QCand*>(smem_raw + Q_CAND);        // [2][QW]
    unsigned long long* keyS = reinterpret_cast<unsigned long long*>(smem_raw + Q_KEYS);   // [2][QW]: the candidates' keys, packed
    double* lamS = reinterpret_cast<double*>(smem_raw + Q_LAM);        // [QW][QROWS]: per-warp copy of the lazy row scales
    double* ilamS = reinterpret_cast<double*>(smem_raw + Q_ILAM);      // [QW][QROWS]: 1 / lam
    double* sval = reinterpret_cast<double*>(smem_raw + Q_SVAL);
    double* sig = reinterpret_cast<double*>(smem_raw + Q_SIG);
    double* xbuf = reinterpret_cast<double*>(smem_raw + Q_XBUF);
    double* gbuf = reinterpret_cast<double*>(smem_raw + Q_GBUF);
    double* gnn = reinterpret_cast<double*>(smem_raw + Q_GNN);
    int* rowvarS = reinterpret_cast<int*>(smem_raw + Q_ROWVAR);        // tile row -> constraint whose slack is basic there
    int* cvsm = reinterpret_cast<int*>(smem_raw + Q_CV);               // column -> constraint whose slack is nonbasic there
    int* colvar0 = reinterpret_cast<int*>(smem_raw + Q_COLVAR0);
    int* pivcol = reinterpret_cast<int*>(smem_raw + Q_PIVCOL);
    int* order = reinterpret_cast<int*>(smem_raw + Q_ORDER);
    int* basic_tile = reinterpret_cast<int*>(smem_raw + Q_BASIC);
    int* red = reinterpret_cast<int*>(smem_raw + Q_RED);
    long long* cur_lp = reinterpret_cast<long long*>(smem_raw + Q_CUR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = a.m, n = a.n;
    const int c0 = warp * QC;                     // first column of my block
    const int col = c0 + lane;                    // the column whose costs I hold (lane < QC)
    const bool hascol = lane < QC && col < n;
    const int t0 = QR * lane;                     // my first tile row
    double* prow = prowS + warp * QP;
    double* lamW = lamS + warp * QROWS;
    double* ilamW = ilamS + warp * QROWS;

    double T[QR][QP];                             // my rows t0 .. t0+3, columns c0 .. c0+24, slot QC = right-hand side

    // rank-1 update of my tile from my warp's slice of the (raw) pivot row:  T[q][c] -= f[q] * prow[c]
    auto rank1 = [&](const double* pr, const double (&f)[QR]) {
        const double2* p2 = reinterpret_cast<const double2*>(pr);
        const double n0 = -f[0], n1 = -f[1], n2 = -f[2], n3 = -f[3];
#pragma unroll
        for (int c2 = 0; c2 < QP / 2; ++c2) {
            const double2 v = p2[c2];
            T[0][2 * c2] = fma(n0, v.x, T[0][2 * c2]);
            T[0][2 * c2 + 1] = fma(n0, v.y, T[0][2 * c2 + 1]);
            T[1][2 * c2] = fma(n1, v.x, T[1][2 * c2]);
            T[1][2 * c2 + 1] = fma(n1, v.y, T[1][2 * c2 + 1]);
            T[2][2 * c2] = fma(n2, v.x, T[2][2 * c2]);
            T[2][2 * c2 + 1] = fma(n2, v.y, T[2][2 * c2 + 1]);
            T[3][2 * c2] = fma(n3, v.x, T[3][2 * c2]);
            T[3][2 * c2 + 1] = fma(n3, v.y, T[3][2 * c2 + 1]);
        }
    };
    auto load4 = [&](const double* p, double (&v)[QR]) {          // p 32-byte aligned
        const double2 u0 = reinterpret_cast<const double2*>(p)[0], u1 = reinterpret_cast<const double2*>(p)[1];
        v[0] = u0.x; v[1] = u0.y; v[2] = u1.x; v[3] = u1.y;
    };
    auto store4 = [&](double* p, const double (&v)[QR]) {
        reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
        reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
    };
    // dot products of the rows of A with a lane-distributed vector, eight rows per warp in flight
    auto row_dots = [&](auto rbt, const double* Ag, const double (&vl)[QCS], double* out1, double* out2) {
        constexpr int RB = decltype(rbt)::value;
        for (int base = 0; base < m; base += RB * QW) {
            double v[RB][QCS];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int i = base + r * QW + warp;
#pragma unroll
                for (int cs = 0; cs < QCS; ++cs) {
                    const int j = lane + 32 * cs;
                    v[r][cs] = (i < m && j < n) ? qld<GEN>(Ag + (size_t)i * n + j) : 0.0;
                }
            }
            double dot[RB], nn[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                dot[r] = 0.0;
                nn[r] = 0.0;
#pragma unroll
                for (int cs = 0; cs < QCS; ++cs) {
                    dot[r] = fma(v[r][cs], vl[cs], dot[r]);
                    nn[r] = fma(v[r][cs], v[r][cs], nn[r]);
                }
            }
            // all RB row sums with one halving butterfly (same addition tree as warp_sum, so the same bits): afterwards the
            // lane holds the sum of row  rsel  of the batch
            int rsel;
            const double d1 = reduce_rows<RB>(dot, lane, rsel);
            const int i = base + rsel * QW + warp;
            const bool writer = (lane & (RB == 8 ? 3 : 15)) == 0;      // one lane per row of the batch
            if (out2) {
                const double n1 = reduce_rows<RB>(nn, lane, rsel);
                if (writer && i < m) out2[i] = n1;
            }
            if (writer && i < m) out1[i] = d1;
        }
    };
    // winner of the four candidates of buffer b: smallest key, lowest warp on ties
    auto pick_min = [&](int b, unsigned long long& kbest) -> int {
        // (opaque loads: otherwise the compiler proves the keys warp-uniform and runs the comparisons on the uniform
        //  datapath -- 32 instructions with R2UR round trips instead of 12)
        unsigned long long k0, k1, k2, k3;
        const uint32_t ka32 = smem_u32(keyS + b * QW);
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "r"(ka32));
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(k2), "=l"(k3) : "r"(ka32));
        const bool b1 = k1 < k0, b3 = k3 < k2;
        const unsigned long long ka = b1 ? k1 : k0, kb = b3 ? k3 : k2;
        const bool bb = kb < ka;
        kbest = bb ? kb : ka;
        return bb ? (b3 ? 3 : 2) : (b1 ? 1 : 0);
    };

    for (;;) {
        if (tid == 0) *cur_lp = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long lp = *cur_lp;
        if (lp >= a.B) break;
        const double* Ag;
        const double* bg;
        const double* cg;
        const uint8_t* mask = MASK ? a.row_mask + (size_t)lp * m : nullptr;            // reduced LP: rows with mask 0 are left out

        // ---- stage 0: crash order by cosine score ------------------------------------------------------------------
        if constexpr (GEN) {
            double* slab = a.slab + (size_t)blockIdx.x * slab_doubles(m, n);
            double* Aw = a.A ? const_cast<double*>(a.A) + (size_t)lp * m * n : slab;
            double* bw = a.A ? const_cast<double*>(a.b) + (size_t)lp * m : slab + slab_b_offset(m, n);
            double* cw = a.A ? const_cast<double*>(a.c) + (size_t)lp * n : slab + slab_c_offset(m, n);
            generate_instance_cta<7, kQuadGenChains>((uint64_t)a.gen_key, (uint64_t)(a.gen_first + lp), m, n, a.gen_density, Aw, bw, cw,
                                                     nullptr, Dsm, gbuf, gnn);          // the tile overlays the (dead) crash inverse
            Ag = Aw; bg = bw; cg = cw;
        } else {
            Ag = a.A + (size_t)lp * m * n;
            bg = a.b + (size_t)lp * m;
            cg = a.c + (size_t)lp * n;
            double cl[QCS];
#pragma unroll
            for (int cs = 0; cs < QCS; ++cs) {
                const int j = lane + 32 * cs;
                cl[cs] = (j < n) ? qld<GEN>(cg + j) : 0.0;
            }
            row_dots(std::integral_constant<int, 8>{}, Ag, cl, gbuf, gnn);
        }
        __syncthreads();
        for (int i = tid; i < m; i += QNT) {
            const double dot = gbuf[i], nn = gnn[i];
            const bool excl = mask && mask[i] == 0;
            gnn[i] = excl ? kInf : (nn > 0.0 ? dot / sqrt(nn) : kInf * 0.5);        // excluded rows rank last
        }
        __syncthreads();
        for (int i = tid; i < m; i += QNT) {
            const double v = gnn[i];
            int rank = 0;
            for (int i2 = 0; i2 < m; ++i2) {
                const double v2 = gnn[i2];
                rank += (v2 < v) || (v2 == v && i2 < i);
            }
            order[rank] = i;
            basic_tile[i] = -1;
        }
        __syncthreads();
        int m_eff = m;
        if constexpr (MASK) {
            m_eff = 0;
            for (int i = 0; i < m; ++i) m_eff += (gnn[i] < kInf);     // uniform, only for reduced LPs
        }
        const int nN = m_eff - n;                                      // live rows; order[m_eff ..) are the excluded ones
        bool need_generic = MASK ? (nN < 0) : false;                   // fewer kept rows than columns: the generic kernel's business
        int npiv_crash = 0, npiv_p1 = 0, npiv_p2 = 0;
        int status = ST_OPTIMAL;
        int buf = 0;
        double gj = 0.0, ghj = 0.0;               // reduced cost / phase-1 cost of my column

        // ---- stage 1: rows order[0..n) of [A | b]; Gauss-Jordan to the inverse ----------------------------------------
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const int t = t0 + q;
            const bool have = t < n;
            const int row = have ? order[t] : 0;
            const double* Ar = Ag + (size_t)row * n + c0;
#pragma unroll
            for (int c = 0; c < QC; ++c) T[q][c] = (have && c0 + c < n) ? qld<GEN>(Ar + c) : 0.0;
            T[q][QC] = have ? qld<GEN>(bg + row) : 0.0;
            lamW[t] = 1.0;
            ilamW[t] = 1.0;
            if (warp == 0) rowvarS[t] = have ? row : -1;
        }
        gj = hascol ? qld<GEN>(cg + col) : 0.0;
        ghj = hascol ? 1.0 : 0.0;
        bool colfree = hascol;                    // my column still holds a free x_j
        if (tid < QROWS) cvsm[tid] = (tid < n) ? -1 : -2;
        __syncthreads();

        for (int t = 0; t < n; ++t) {
            if constexpr (MASK) {
                if (need_generic) break;
            }
            const int lt = t >> 2, qt = t & 3;
            if (lane == lt) publish_slot(prow, T, qt);
            __syncwarp();
            // my warp's candidate: largest |entry| of row t among its free columns (lowest column on ties)
            const double pl = (lane < QC) ? prow[lane] : 0.0;
            unsigned long long kmin;
            const unsigned long long key = colfree ? (unsigned long long)__double_as_longlong(fabs(pl)) : 0ull;
            const int kkw = warp_argmin_key(~key, kmin);
            const unsigned long long kabs = ~kmin;
            const double pw = prow[kkw];
            double e[QR], f[QR];
            col_get4(T, kkw, e);
            const double rpw = fast_rcp(kabs ? pw : 1.0);
#pragma unroll
            for (int q = 0; q < QR; ++q) f[q] = (t0 + q == t) ? 0.0 : e[q] * rpw;
            store4(fS + ((size_t)buf * QW + warp) * QROWS + t0, f);
            const double gkw = __shfl_sync(FULL, gj, kkw);
            if (lane == 0) {
                QCand* cd = candS + buf * QW + warp;
                keyS[buf * QW + warp] = ~kabs;    // smallest complemented key = largest |pivot|
                cd->k = c0 + kkw;
                cd->p = pw;
                cd->gk = gkw;
                cd->rp = rpw;
            }
            __syncthreads();
            unsigned long long kbest;
            const int wk = pick_min(buf, kbest);
            if (__longlong_as_double((long long)~kbest) < kTolCrash) { need_generic = true; break; }
            const QCand* cd = candS + buf * QW + wk;
            const int k = cd->k;
            const double p = cd->p, gk = cd->gk;
            const double rp = cd->rp;
            load4(fS + ((size_t)buf * QW + wk) * QROWS + t0, f);
            rank1(prow, f);
            const double fg = gk * rp;
            if (hascol) gj = (col == k) ? -fg : fma(-fg, pl, gj);
            if (col == k) colfree = false;
            if (warp == wk) {
                double v[QR];
#pragma unroll
                for (int q = 0; q < QR; ++q) v[q] = (t0 + q == t) ? 1.0 : -f[q];
                col_set4(T, k - c0, v);
            }
            if (lane == lt) {
                lamW[t] = rp;
                ilamW[t] = p;
            }
            if (tid == 0) {
                pivcol[t] = k;
                cvsm[k] = 0;
            }
            __syncwarp();                         // everybody is done with prow before the next row's owner overwrites it
            buf ^= 1;
            ++npiv_crash;
        }

        if (!need_generic) {
            __syncthreads();
            // dump D' (row of x_k stored at index k, true rows lam * T; slot QC of every block = the x-vertex) and the maps
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const int t = t0 + q;
                if (t < n) {
                    const int k = pivcol[t];
                    const double lam = lamW[t];
                    double* dr = Dsm + (size_t)k * QDP + warp * QP;
#pragma unroll
                    for (int c = 0; c < QP; ++c) dr[c] = T[q][c] * lam;
                    if (warp == 0) {
                        colvar0[k] = rowvarS[t];
                        cvsm[k] = rowvarS[t];
                    }
                }
            }
            __syncthreads();

            // ---- stage 2: my rows of P_N = -A_N D, s_N = b_N - A_N xv ----------------------------------------------------
            {
                int rowq[QR];
                bool liveq[QR];
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    liveq[q] = (t0 + q) < nN;
                    rowq[q] = liveq[q] ? order[n + t0 + q] : 0;
#pragma unroll
                    for (int c = 0; c < QP; ++c) T[q][c] = 0.0;
                }
                // (staging A_N through shared memory -- one 32-byte sector per row and chunk, two LDS.128 per column instead of
                //  four scattered 8-byte loads -- was measured: 578 k against 604 k LP/s; the per-chunk barriers cost more than
                //  the loads, which the other LP of the SM hides)
                if (t0 < nN) {                    // lanes beyond the live rows keep a zero tile
                    double an[QR], av[QR];
                    int offq[QR];
#pragma unroll
                    for (int q = 0; q < QR; ++q) {
                        offq[q] = rowq[q] * n;
                        an[q] = liveq[q] ? qld<GEN>(Ag + offq[q]) : 0.0;
                    }
                    for (int k = 0; k < n; ++k) {
#pragma unroll
                        for (int q = 0; q < QR; ++q) {
                            av[q] = an[q];
                            an[q] = (liveq[q] && k + 1 < n) ? qld<GEN>(Ag + (offq[q] + k + 1)) : 0.0;
                        }
                        rank1(Dsm + (size_t)k * QDP + warp * QP, av);
                    }
                }
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    if (liveq[q]) T[q][QC] += qld<GEN>(bg + rowq[q]);
                    lamW[t0 + q] = 1.0;
                    ilamW[t0 + q] = 1.0;
                    if (warp == 0) rowvarS[t0 + q] = liveq[q] ? rowq[q] : -1;
                }
            }
            __syncthreads();

            // ---- stage 3a: phase 1 (most negative slack leaves; ratio test along its row) ---------------------------------
            for (;;) {
                double lam[QR];
                load4(lamW + t0, lam);
                // (rows beyond the live ones hold zeros: they are never candidates, here or in the ratio tests)
                double sbest = 0.0;
                unsigned long long kmin;
                int qloc = 0;
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const double s = lam[q] * T[q][QC];
                    if (s < -kTolFeas && s < sbest) { sbest = s; qloc = q; }
                }
                // dkey of a negative number is the complement of its bit pattern
                const int ll = warp_argmin_key(sbest < 0.0 ? ~(unsigned long long)__double_as_longlong(sbest) : KEY_INF, kmin);
                if (kmin == KEY_INF) break;                       // s >= 0 everywhere: phase 1 finished
                if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int qsel = __shfl_sync(FULL, qloc, ll);
                const int r = QR * ll + qsel;
                if (lane == ll) publish_slot(prow, T, qsel);
                __syncwarp();
                const double lam_r = lamW[r], il = ilamW[r];
                // ratio test over my columns: min ghat_j / (-e_j) over e_j < -tol (true row lam_r * T_r)
                const double pl = (lane < QC) ? prow[lane] : 0.0;
                const double en = -lam_r * pl;
                const bool ok = hascol && en > kTolPivot;
                const double ratio = ((__double_as_longlong(ghj) < 0) ? 0.0 : ghj) * fast_rcp(ok ? en : 1.0);
                unsigned long long kminc;                          // ratios are >= +0: bit patterns order like the values
                const int kl = warp_argmin_key(ok ? (unsigned long long)__double_as_longlong(ratio) : QINF, kminc);
                const bool none = (kminc == QINF);
                const int kkw = none ? 0 : kl;
                const double pw = prow[kkw];
                const double rpw = fast_rcp(none ? 1.0 : pw);
                double e[QR], f[QR];
                col_get4(T, kkw, e);
#pragma unroll
                for (int q = 0; q < QR; ++q) f[q] = (t0 + q == r) ? 0.0 : e[q] * rpw;
                store4(fS + ((size_t)buf * QW + warp) * QROWS + t0, f);
                const double gkw = __shfl_sync(FULL, gj, kkw), ghkw = __shfl_sync(FULL, ghj, kkw);
                if (lane == 0) {
                    QCand* cd = candS + buf * QW + warp;
                    keyS[buf * QW + warp] = kminc;
                    cd->k = c0 + kkw;
                    cd->p = pw;
                    cd->gk = gkw;
                    cd->ghk = ghkw;
                    cd->rp = rpw;
                }
                __syncthreads();
                unsigned long long kbest;
                const int wk = pick_min(buf, kbest);
                if (kbest == QINF) { status = ST_INFEASIBLE; break; }
                const QCand* cd = candS + buf * QW + wk;
                const int k = cd->k;
                const double p = cd->p, gk = cd->gk, ghk = cd->ghk;
                const double rp = cd->rp;
                load4(fS + ((size_t)buf * QW + wk) * QROWS + t0, f);
                rank1(prow, f);
                const double fv = ghk * rp, fg = gk * rp;
                if (hascol) {
                    ghj = (col == k) ? -fv * il : fma(-fv, pl, ghj);
                    gj = (col == k) ? -fg * il : fma(-fg, pl, gj);
                }
                if (warp == wk) {
                    double v[QR];
#pragma unroll
                    for (int q = 0; q < QR; ++q) v[q] = (t0 + q == r) ? il : -f[q] * il;
                    col_set4(T, k - c0, v);
                }
                if (lane == ll) {
                    lamW[r] = rp;
                    ilamW[r] = p;
                }
                if (tid == ll) {                                  // warp 0's owner lane keeps the maps
                    const int cv = cvsm[k];
                    cvsm[k] = rowvarS[r];                         // the old slack becomes nonbasic in column k
                    rowvarS[r] = cv;                              // column k's constraint becomes basic in row r
                }
                __syncwarp();
                buf ^= 1;
                ++npiv_p1;
            }
            __syncthreads();

            // ---- stage 3b: phase 2 (Dantzig) ---------------------------------------------------------------------------
            // every warp prices its own columns, runs the ratio test on its best column and leaves the multipliers of
            // that pivot in fS; after the barrier all warps take the globally best column's record
            auto speculate = [&]() {
                unsigned long long kmin;
                const int kl = warp_argmin_key(hascol ? dkey(gj) : KEY_INF, kmin);
                const bool has = kmin < dkey(-kTolFeas);
                const int kkw = has ? kl : 0;
                double e[QR], lam[QR], f[QR];
                col_get4(T, kkw, e);
                load4(lamW + t0, lam);
                const double pinf = __longlong_as_double(0x7ff0000000000000ll);
                double rbest = pinf;
                unsigned long long krow;
                int qloc = 0;
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const double et = lam[q] * e[q];                               // true entry / right-hand side of the row
                    const double sr = lam[q] * T[q][QC];
                    const double sc = (__double_as_longlong(sr) < 0) ? 0.0 : sr;     // max(sr, 0) without the NaN handling of fmax
                    const bool cand = et > kTolPivot;
                    const double ratio = cand ? sc * fast_rcp(cand ? et : 1.0) : pinf;
                    if (ratio < rbest) { rbest = ratio; qloc = q; }
                }
                // ratios are >= +0: their bit patterns order like the values (same winner as with dkey)
                const int ll = warp_argmin_key((unsigned long long)__double_as_longlong(rbest), krow);
                const bool none = (krow == QINF);
                const int qsel = __shfl_sync(FULL, qloc, ll);
                const int rw = none ? -1 : QR * ll + qsel;
                const double pw = __shfl_sync(FULL, sel4(e, qsel), ll);
                const double rpw = fast_rcp(none ? 1.0 : pw);
#pragma unroll
                for (int q = 0; q < QR; ++q) f[q] = (t0 + q == rw) ? 0.0 : e[q] * rpw;
                store4(fS + ((size_t)buf * QW + warp) * QROWS + t0, f);
                const double gkw = __shfl_sync(FULL, gj, kkw);
                if (lane == 0) {
                    QCand* cd = candS + buf * QW + warp;
                    keyS[buf * QW + warp] = has ? kmin : KEY_INF;
                    cd->k = c0 + kkw;
                    cd->r = rw;
                    cd->p = pw;
                    cd->gk = gkw;
                    cd->rp = rpw;
                }
            };
            if (status == ST_OPTIMAL) {
                speculate();
                __syncthreads();
            }
            while (status == ST_OPTIMAL) {
                unsigned long long kbest;
                const int wk = pick_min(buf, kbest);
                if (kbest == KEY_INF) break;                       // g >= 0: optimal
                const QCand* cd = candS + buf * QW + wk;
                const int k = cd->k, r = cd->r;
                if (r < 0) { status = ST_UNBOUNDED; break; }
                if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const double p = cd->p, gk = cd->gk;
                const int ll = r >> 2, qsel = r & 3;
                if (lane == ll) publish_slot(prow, T, qsel);
                __syncwarp();
                const double il = ilamW[r];
                const double rp = cd->rp;
                double f[QR];
                load4(fS + ((size_t)buf * QW + wk) * QROWS + t0, f);
                rank1(prow, f);
                const double pl = (lane < QC) ? prow[lane] : 0.0;
                const double fg = gk * rp;
                if (hascol) gj = (col == k) ? -fg * il : fma(-fg, pl, gj);
                if (warp == wk) {
                    double v[QR];
#pragma unroll
                    for (int q = 0; q < QR; ++q) v[q] = (t0 + q == r) ? il : -f[q] * il;
                    col_set4(T, k - c0, v);
                }
                if (lane == ll) {
                    lamW[r] = rp;
                    ilamW[r] = p;
                }
                if (tid == ll) {
                    const int cv = cvsm[k];
                    cvsm[k] = rowvarS[r];
                    rowvarS[r] = cv;
                }
                __syncwarp();
                buf ^= 1;
                ++npiv_p2;
                speculate();
                __syncthreads();
            }
        }

        // ---- stage 4: x, objective, slacks, labels ------------------------------------------------------------------------
        __syncthreads();
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0;
        if (need_generic) {
            status = -1;   // re-solved by the generic kernel (capi.cu)
        } else if (status == ST_OPTIMAL) {
            if (warp == 0) {
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const int t = t0 + q;
                    if (t < nN) {
                        sval[t] = lamW[t] * T[q][QC];
                        const int rv = rowvarS[t];
                        if (rv >= 0) basic_tile[rv] = t;
                    }
                }
            }
            __syncthreads();
            for (int j = tid; j < n; j += QNT) {
                const int bt = basic_tile[colvar0[j]];
                sig[j] = (bt >= 0) ? sval[bt] : 0.0;
            }
            __syncthreads();
            {
                double sl[QCS];
                int doff[QCS];
#pragma unroll
                for (int cs = 0; cs < QCS; ++cs) {
                    const int j = lane + 32 * cs;
                    sl[cs] = (j < n) ? sig[j] : 0.0;
                    const int jj = (j < n) ? j : 0;
                    doff[cs] = (jj / QC) * QP + (jj % QC);
                }
                for (int k = warp; k < n; k += QW) {
                    double acc = 0.0;
#pragma unroll
                    for (int cs = 0; cs < QCS; ++cs)
                        if (lane + 32 * cs < n) acc = fma(Dsm[(size_t)k * QDP + doff[cs]], sl[cs], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) xbuf[k] = Dsm[(size_t)k * QDP + QC] - acc;
                }
            }
            __syncthreads();
            double xl[QCS];
            auto take_x = [&]() {
#pragma unroll
                for (int cs = 0; cs < QCS; ++cs) {
                    const int j = lane + 32 * cs;
                    xl[cs] = (j < n) ? xbuf[j] : 0.0;
                }
                if (warp == 0) {
                    double acc = 0.0;
#pragma unroll
                    for (int cs = 0; cs < QCS; ++cs) {
                        const int j = lane + 32 * cs;
                        if (j < n) acc = fma(qld<GEN>(cg + j), xl[cs], acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0 && a.obj) a.obj[lp] = acc;
                }
                if (a.x)
                    for (int j = tid; j < n; j += QNT) a.x[(size_t)lp * n + j] = xbuf[j];
            };
            // labels exactly as gurobi_lp.py:435-443 from the caller's A; returns whether an active (nonbasic) row has a
            // visible residual at this x
            auto label_pass = [&]() -> int {
                row_dots(std::integral_constant<int, 2>{}, Ag, xl, gbuf, nullptr);          // gbuf[i] = a_i . x (two rows in flight: the tile is alive)
                __syncthreads();
                nact = 0; nties = 0; nviol = 0;
                int nref = 0;
                for (int i = tid; i < m; i += QNT) {
                    const double slack = qld<GEN>(bg + i) - gbuf[i];
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
            take_x();
            const int nref = label_pass();
            if (__syncthreads_or(nref > 0)) {
                // ---- one step of iterative refinement on the final active set (ill-conditioned vertex, ~0.06 % of the instances;
                // same scheme as rowreg_kernel.cuh / simplex_generic.cu).  rho_j = slack of the nonbasic constraint of column j
                // at the computed x; it should be 0: move the nonbasic slacks from rho to 0 through the tableau -- the tile is
                // still in the registers (the label pass streams A two rows at a time), a row's dot product with rho is four per-warp partial sums -- and correct x through
                // the crash inverse.
                double* rho = gnn;            // the crash scores are dead by now
                int* colpos = order;          // so is the crash order: constraint -> column where its slack is nonbasic
                double* part = fS;            // [QW][QROWS]
                for (int j = tid; j < QNMAX; j += QNT) {
                    double rv = 0.0;
                    if (j < n) {
                        const int q = cvsm[j];
                        rv = qld<GEN>(bg + q) - gbuf[q];
                        colpos[q] = j;
                    }
                    rho[j] = rv;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    double d = 0.0;
#pragma unroll
                    for (int c = 0; c < QC; ++c) d = fma(T[q][c], rho[c0 + c], d);
                    part[warp * QROWS + t0 + q] = d;
                }
                __syncthreads();
                if (tid < nN)
                    sval[tid] = lamS[tid] * (((part[tid] + part[QROWS + tid]) + part[2 * QROWS + tid]) + part[3 * QROWS + tid]);
                __syncthreads();
                for (int j0 = tid; j0 < n; j0 += QNT) {
                    const int q0 = colvar0[j0];
                    const int bt = basic_tile[q0];
                    sig[j0] = (bt >= 0) ? sval[bt] : -rho[colpos[q0]];
                }
                __syncthreads();
                {
                    double sl[QCS];
                    int doff[QCS];
#pragma unroll
                    for (int cs = 0; cs < QCS; ++cs) {
                        const int j = lane + 32 * cs;
                        sl[cs] = (j < n) ? sig[j] : 0.0;
                        const int jj = (j < n) ? j : 0;
                        doff[cs] = (jj / QC) * QP + (jj % QC);
                    }
                    for (int k = warp; k < n; k += QW) {
                        double acc = 0.0;
#pragma unroll
                        for (int cs = 0; cs < QCS; ++cs)
                            if (lane + 32 * cs < n) acc = fma(Dsm[(size_t)k * QDP + doff[cs]], sl[cs], acc);
                        acc = warp_sum(acc);
                        if (lane == 0) xbuf[k] -= acc;
                    }
                }
                __syncthreads();
                take_x();
                label_pass();
            }
        }
        if (status != -1 && status != ST_OPTIMAL) {
            for (int i = tid; i < m; i += QNT) lab[i] = 0;
            if (a.x)
                for (int j = tid; j < n; j += QNT) a.x[(size_t)lp * n + j] = 0.0;
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
            int s0 = 0, s1 = 0, s2 = 0;
            for (int w = 0; w < QW; ++w) {
                s0 += red[w * 3 + 0];
                s1 += red[w * 3 + 1];
                s2 += red[w * 3 + 2];
            }
            a.status[lp] = status;
            if (status == -1) atomicAdd(a.flag_count, 1);
            if (status != -1) {
                if (a.n_active) a.n_active[lp] = s0;
                if (a.ties) a.ties[lp] = s1;
                if (a.violations) a.violations[lp] = s2;
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

}  // namespace


bool quadcol_supported(int m, int n) { return n >= 1 && n <= QNMAX && m >= n && m - n <= QROWS && m <= QMMAX; }

template <bool GEN, bool MASK>
static int quadcol_ctas_per_sm() {
    static int per_sm = -1;
    if (per_sm < 0) {
        int v = 0;
        if (cudaFuncSetAttribute(simplex_quadcol_kernel<GEN, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_TOTAL) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, simplex_quadcol_kernel<GEN, MASK>, QNT, Q_TOTAL) != cudaSuccess)
            v = 0;
        per_sm = v;
    }
    return per_sm;
}

int quadcol_grid(int sm_count) {
    const int per_sm = quadcol_ctas_per_sm<false, false>();
    return sm_count * (per_sm > 0 ? per_sm : 1);
}
int quadcol_gen_grid(int sm_count) {
    const int per_sm = quadcol_ctas_per_sm<true, false>();
    return sm_count * (per_sm > 0 ? per_sm : 1);
}
// the in-solver generator draws pairs: even n; its tile (32 rows + x0 + c + eps) overlays the crash inverse
bool quadcol_gen_supported(int m, int n) {
    return quadcol_supported(m, n) && (n & 1) == 0 && gen_smem_doubles(m, n) * sizeof(double) <= (size_t)QNMAX * QDP * 8;
}

template <bool GEN, bool MASK>
static cudaError_t launch_quadcol(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const int per_sm = quadcol_ctas_per_sm<GEN, MASK>();
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)sm_count * per_sm;
    // DDB_QUADCOL_ONE_PER_SM=1: measurement switch (one LP per SM: how much of the throughput is latency hiding between LPs)
    static const bool one = [] { const char* e = getenv("DDB_QUADCOL_ONE_PER_SM"); return e && e[0] == '1'; }();
    if (one) grid = sm_count;
    if (grid > a.B) grid = a.B;
    simplex_quadcol_kernel<GEN, MASK><<<(int)grid, QNT, Q_TOTAL, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_simplex_quadcol(const SolveArgs& a, int sm_count, cudaStream_t st) {
    return a.row_mask ? launch_quadcol<false, true>(a, sm_count, st) : launch_quadcol<false, false>(a, sm_count, st);
}
cudaError_t launch_simplex_quadcol_gen(const SolveArgs& a, int sm_count, cudaStream_t st) { return launch_quadcol<true, false>(a, sm_count, st); }

}  // namespace ddb
