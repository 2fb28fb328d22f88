// Register-tiled batched fp64 simplex (plan 0): one LP per CTA, the condensed tableau lives in the REGISTER FILE.
//
// Why registers: at (m,n) = (200,100) a pivot touches ~10^4 tableau entries.  In shared memory that is 16 bytes of
// LDS+STS traffic per entry (128 B/clk/SM -> >= 1250 clk per pivot); in registers it is one DFMA per entry
// (64/clk/SM -> ~160 clk).  B200's 256 KB register file per SM is as large as its shared memory, so the tableau of
// the simplex stage (80 KB at this shape) fits with room to spare.
//
// Layout: NW warps per CTA.  Tile row ti lives in warp (ti % NW), register slot (ti / NW); tile column j lives in lane
// (j % 32), register slot (j / 32).  A warp therefore owns whole rows, a lane owns whole columns of its warp's rows.
// The last tile column (lane 31, slot CS-1) is the right-hand side, the last tile row (warp NW-1, slot RS-1) is the
// cost row g.  Everything indexed dynamically (pivot column slot, candidate row slot) is turned into a warp-uniform
// switch so that register indices stay static.
//
// One __syncthreads per pivot: before the barrier every warp finds its own best candidate row and *speculatively*
// publishes that row, already normalised, to shared memory; after the barrier every warp picks the same winner
// (min key, lowest warp on ties) and applies the rank-1 update from the winner's published row.
//
// Stages per LP (same algorithm and tolerances as simplex_generic.cu):
//   0. crash order by cosine score (global loads of A, one warp per row)
//   1. crash as an explicit inverse: Gauss-Jordan on the n x n block A_B0 held in the tile (n pivots, row-wise
//      pivot search is warp-local), result D = inverse + x-vertex dumped to shared memory
//   2. remaining rows enter through a register-blocked GEMM  P_N = -A_N D  (D from shared memory, A_N from L1/L2)
//   3. phase 1 (row-first, artificial costs), phase 2 (column-first Dantzig)
//   4. x = xv - D sigma, slack = b - A x from the caller's A, labels = |slack| <= threshold
// Instances the tile cannot hold or whose static crash basis is singular are flagged status = -1 and re-solved by
// the generic kernel on the device (capi.cu); nothing ever falls back to the CPU.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace ddb {

// Warp-uniform dynamic index -> compile-time index (jump table instead of chains of predicated moves).
template <int N, class F>
__device__ __forceinline__ void static_switch(int i, F&& f) {
    switch (i) {
#define DDB_CASE(I) \
    case I:         \
        if constexpr (I < N) f(std::integral_constant<int, I>{}); \
        break;
        DDB_CASE(0) DDB_CASE(1) DDB_CASE(2) DDB_CASE(3) DDB_CASE(4) DDB_CASE(5) DDB_CASE(6) DDB_CASE(7)
        DDB_CASE(8) DDB_CASE(9) DDB_CASE(10) DDB_CASE(11) DDB_CASE(12) DDB_CASE(13) DDB_CASE(14) DDB_CASE(15)
#undef DDB_CASE
        default: break;
    }
}

struct PubHdr {          // one per (buffer, warp): the speculative candidate
    unsigned long long key;   // dkey(s_min) in phase 1, dkey(ratio) in phase 2, KEY_INF = no candidate
    double srow;              // right-hand side of the normalised candidate row
    int var;                  // constraint whose slack is basic in the candidate row
    int k;                    // entering column
    int tile_row;             // candidate tile row
    int flag;                 // 0 ok, 2 = row proves infeasibility (phase 1)
};

template <int NW, int RS, int CS>
struct RegCfg {
    static constexpr int RT = NW * RS;     // tile rows (last one is the cost row)
    static constexpr int CT = 32 * CS;     // tile columns (last one is the right-hand side)
    static constexpr int THREADS = NW * 32;
};

struct RegLayout {
    size_t D, pub_row, pub_hdr, order, colvar0, pivcol, basic_tile, sval, sig, xbuf, gbuf, red, total;
};
__host__ __device__ inline size_t ralign(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ inline RegLayout make_reg_layout(int m, int n, int NW, int CT) {
    RegLayout L;
    size_t off = 0;
    L.D = off;          off += ralign((size_t)n * (n + 1) * 8 + 16);   // + one zero word for padding lanes
    L.pub_row = off;    off += ralign((size_t)2 * NW * CT * 8);
    L.pub_hdr = off;    off += ralign((size_t)2 * NW * sizeof(PubHdr));
    L.order = off;      off += ralign((size_t)m * 4);
    L.colvar0 = off;    off += ralign((size_t)n * 4);
    L.pivcol = off;     off += ralign((size_t)n * 4);
    L.basic_tile = off; off += ralign((size_t)m * 4);
    L.sval = off;       off += ralign((size_t)NW * 32 * 8);            // >= RT doubles (RS <= 32)
    L.sig = off;        off += ralign((size_t)n * 8);
    L.xbuf = off;       off += ralign((size_t)n * 8);
    L.gbuf = off;       off += ralign((size_t)(m > CT ? m : CT) * 8);  // scores (m) / g broadcast (CT)
    L.red = off;        off += 3 * 32 * 4;
    L.total = off;
    return L;
}

template <int NW, int RS, int CS, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) simplex_regtile_kernel(SolveArgs a) {
    using Cfg = RegCfg<NW, RS, CS>;
    constexpr int RT = Cfg::RT, CT = Cfg::CT;
    static_assert(RS <= 16 && CS <= 16, "static_switch covers 16 cases");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const RegLayout L = make_reg_layout(m, n, NW, CT);
    double* Dsm = reinterpret_cast<double*>(smem_raw + L.D);
    double* pub_row = reinterpret_cast<double*>(smem_raw + L.pub_row);
    PubHdr* pub_hdr = reinterpret_cast<PubHdr*>(smem_raw + L.pub_hdr);
    int* order = reinterpret_cast<int*>(smem_raw + L.order);
    int* colvar0 = reinterpret_cast<int*>(smem_raw + L.colvar0);
    int* pivcol = reinterpret_cast<int*>(smem_raw + L.pivcol);
    int* basic_tile = reinterpret_cast<int*>(smem_raw + L.basic_tile);
    double* sval = reinterpret_cast<double*>(smem_raw + L.sval);
    double* sig = reinterpret_cast<double*>(smem_raw + L.sig);
    double* xbuf = reinterpret_cast<double*>(smem_raw + L.xbuf);
    double* gbuf = reinterpret_cast<double*>(smem_raw + L.gbuf);
    int* red = reinterpret_cast<int*>(smem_raw + L.red);
    __shared__ long long cur_lp;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int PD = n + 1;                       // pitch of D in shared memory; column n is the x-vertex
    const int ZERO_OFF = n * (n + 1);           // a shared-memory word that always holds 0.0

    // ---- register state -----------------------------------------------------------------------------------
    double T[RS][CS];        // the tile
    double sv[RS];           // right-hand sides of this warp's rows, replicated in every lane
    double vec[CS];          // ghat (phase 1) / g (phase 2), replicated in every warp, lane-distributed by column
    int colvar[CS];          // constraint whose slack is nonbasic in my columns (-1: not a structural column)
    int rowvar_l = -1;       // lane rs < RS: constraint whose slack is basic in row slot rs of this warp

    // column bookkeeping that never changes
    bool col_struct[CS];     // a structural column (j < n)
#pragma unroll
    for (int cs = 0; cs < CS; ++cs) col_struct[cs] = (lane + 32 * cs) < n;

    // dynamic-slot helpers: warp-uniform jump tables keep every register index static
    const bool has_g = (warp == NW - 1);         // this warp owns the cost row (slot RS-1)
    // publish row `slot` scaled by rp into dst[0..CT) with the pivot entry replaced by rp; returns its raw rhs
    auto publish_row = [&](int slot, double rp, int k, double* dst) -> double {
        double sraw = 0.0;
        static_switch<RS>(slot, [&](auto Rc) {
            constexpr int R = decltype(Rc)::value;
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) dst[lane + 32 * cs] = T[R][cs] * rp;
            sraw = sv[R];
        });
        if (lane == (k & 31)) dst[k] = rp;       // same thread wrote dst[k] above: program order makes this one win
        return sraw;
    };
    auto set_row = [&](int slot, const double (&in)[CS], double s_new) {
        static_switch<RS>(slot, [&](auto Rc) {
            constexpr int R = decltype(Rc)::value;
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) T[R][cs] = in[cs];
            sv[R] = s_new;
        });
    };
    // rank-1 update of this warp's live rows (slots < nslots, plus the cost row) from the published pivot row.
    // The pivot row itself is updated too (garbage) and restored by set_row afterwards.
    auto update_crash = [&](int kq, int kl, const double (&pr)[CS], int nslots) {
        const bool is_kl = (lane == kl);
        static_switch<CS>(kq, [&](auto Kc) {
            constexpr int KQ = decltype(Kc)::value;
#pragma unroll
            for (int rs = 0; rs < RS; ++rs) {
                if (rs < nslots || (rs == RS - 1 && has_g)) {
                    const double f = __shfl_sync(FULL, T[rs][KQ], kl);
                    if (is_kl) T[rs][KQ] = 0.0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) T[rs][cs] = fma(-f, pr[cs], T[rs][cs]);
                }
            }
        });
    };
    // same, plus the replicated right-hand sides, the replicated pricing vector and the column bookkeeping;
    // returns the constraint that was nonbasic in column k (it becomes basic in the pivot row)
    auto update_simplex = [&](int kq, int kl, const double (&pr)[CS], double srow, int nslots, int var_r) -> int {
        const bool is_kl = (lane == kl);
        int cv = -1;
        static_switch<CS>(kq, [&](auto Kc) {
            constexpr int KQ = decltype(Kc)::value;
#pragma unroll
            for (int rs = 0; rs < RS; ++rs) {
                if (rs < nslots || (rs == RS - 1 && has_g)) {
                    const double f = __shfl_sync(FULL, T[rs][KQ], kl);
                    if (is_kl) T[rs][KQ] = 0.0;
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) T[rs][cs] = fma(-f, pr[cs], T[rs][cs]);
                    sv[rs] = fma(-f, srow, sv[rs]);
                }
            }
            const double vk = __shfl_sync(FULL, vec[KQ], kl);
            cv = __shfl_sync(FULL, colvar[KQ], kl);
            if (is_kl) { vec[KQ] = 0.0; colvar[KQ] = var_r; }
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) vec[cs] = fma(-vk, pr[cs], vec[cs]);
        });
        return cv;
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
        for (int i = warp; i < m; i += NW) {
            double dot = 0.0, nn = 0.0;
            for (int j = lane; j < n; j += 32) {
                const double v = __ldg(Ag + (size_t)i * n + j);
                dot = fma(v, __ldg(cg + j), dot);
                nn = fma(v, v, nn);
            }
            dot = warp_sum(dot);
            nn = warp_sum(nn);
            if (lane == 0) {
                const bool excl = mask && mask[i] == 0;
                gbuf[i] = excl ? kInf : (nn > 0.0 ? dot / sqrt(nn) : kInf * 0.5);
            }
        }
        if (tid == 0) Dsm[ZERO_OFF] = 0.0;
        __syncthreads();
        int m_eff = 0;
        for (int i = tid; i < m; i += NW * 32) {
            const double v = gbuf[i];
            int rank = 0;
            for (int i2 = 0; i2 < m; ++i2) {
                const double v2 = gbuf[i2];
                rank += (v2 < v) || (v2 == v && i2 < i);
            }
            order[rank] = i;
            basic_tile[i] = -1;
        }
        __syncthreads();
        if (mask) {
            for (int i = 0; i < m; ++i) m_eff += (gbuf[i] < kInf);   // uniform, only for reduced LPs
        } else {
            m_eff = m;
        }
        const int nN = m_eff - n;
        bool need_generic = (nN < 0) || (nN > RT - 1) || (n > RT - 1) || (n > CT - 1);

        int npiv_crash = 0, npiv_p1 = 0, npiv_p2 = 0;
        int status = ST_OPTIMAL;
        int buf = 0;

        if (!need_generic) {
            // ---- stage 1: load A_B0 | b_B0 and the cost row, Gauss-Jordan to the inverse -----------------------
#pragma unroll
            for (int rs = 0; rs < RS; ++rs) {
                const int t = rs * NW + warp;
                if (t < n) {
                    const int row = order[t];
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        T[rs][cs] = (j < n) ? __ldg(Ag + (size_t)row * n + j) : 0.0;
                    }
                    if (lane == 31) T[rs][CS - 1] = __ldg(bg + row);
                } else if (t == RT - 1) {
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        T[rs][cs] = (j < n) ? __ldg(cg + j) : 0.0;
                    }
                } else {
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) T[rs][cs] = 0.0;
                }
                sv[rs] = 0.0;
            }
            unsigned freemask = 0;   // bit cs set: column (lane + 32 cs) is still a free x_j
#pragma unroll
            for (int cs = 0; cs < CS; ++cs)
                if (col_struct[cs]) freemask |= 1u << cs;
            const int nslots_c = (n > warp) ? (n - warp + NW - 1) / NW : 0;

            for (int t = 0; t < n; ++t) {
                const int wo = t % NW, so = t / NW;
                if (warp == wo) {
                    // largest |entry| of my pivot row among free columns
                    unsigned long long best = 0ull;
                    int bq = 0;
                    static_switch<RS>(so, [&](auto Rc) {
                        constexpr int R = decltype(Rc)::value;
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) {
                            const unsigned long long kk =
                                (freemask >> cs & 1u) ? (unsigned long long)__double_as_longlong(fabs(T[R][cs])) : 0ull;
                            if (kk > best) { best = kk; bq = cs; }
                        }
                    });
                    unsigned long long kmin;
                    const int kl = warp_argmin_key(~best, kmin);   // argmax through the complemented key
                    const int kq = __shfl_sync(FULL, bq, kl);
                    const double p = __longlong_as_double((long long)~kmin);   // |pivot|; the sign is folded in below
                    const bool bad = !(p >= kTolCrash);
                    double* pr = pub_row + (size_t)(buf * NW) * CT;
                    // sign of the pivot: read it back from the owning lane
                    double pv = 0.0;
                    static_switch<RS>(so, [&](auto Rc) {
                        constexpr int R = decltype(Rc)::value;
                        double mine = T[R][0];
#pragma unroll
                        for (int cs = 1; cs < CS; ++cs)
                            if (kq == cs) mine = T[R][cs];
                        pv = __shfl_sync(FULL, mine, kl);
                    });
                    const double rp = fast_rcp(pv);
                    publish_row(so, rp, kl + 32 * kq, pr);
                    if (lane == 0) {
                        PubHdr* h = pub_hdr + buf * NW;
                        h->k = kl + 32 * kq;
                        h->flag = bad ? 2 : 0;
                        pivcol[t] = kl + 32 * kq;
                    }
                }
                __syncthreads();
                const PubHdr* h = pub_hdr + buf * NW;
                if (h->flag) { need_generic = true; break; }
                const int k = h->k;
                const int kl = k & 31, kq = k >> 5;
                double pr[CS];
                const double* prs = pub_row + (size_t)(buf * NW) * CT;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) pr[cs] = prs[lane + 32 * cs];
                if (lane == kl) freemask &= ~(1u << kq);
                update_crash(kq, kl, pr, nslots_c);
                if (warp == wo) set_row(so, pr, 0.0);   // restore the normalised pivot row
                buf ^= 1;
                ++npiv_crash;
            }
        }

        if (!need_generic) {
            __syncthreads();
            // dump D' (row of x_k stored at index k) and the column -> constraint map
#pragma unroll
            for (int rs = 0; rs < RS; ++rs) {
                const int t = rs * NW + warp;
                if (t < n) {
                    const int k = pivcol[t];
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) {
                        const int j = lane + 32 * cs;
                        if (j < n) Dsm[(size_t)k * PD + j] = T[rs][cs];
                    }
                    if (lane == 31) Dsm[(size_t)k * PD + n] = T[rs][CS - 1];
                    if (lane == 0) colvar0[k] = order[t];
                }
            }
            __syncthreads();
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
                const int j = lane + 32 * cs;
                colvar[cs] = (j < n) ? colvar0[j] : -1;
                vec[cs] = (j < n) ? 1.0 : 0.0;
            }

            // ---- stage 2: P_N = -A_N D, s_N = b_N - A_N xv ---------------------------------------------------
            {
                int doff[CS], dstr[CS];
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) {
                    const int j = lane + 32 * cs;
                    const bool rhs = (cs == CS - 1) && (lane == 31);
                    const bool valid = (j < n) || rhs;
                    doff[cs] = valid ? (rhs ? n : j) : ZERO_OFF;
                    dstr[cs] = valid ? PD : 0;
                }
                int arow[RS];
#pragma unroll
                for (int rs = 0; rs < RS; ++rs) {
                    const int u = rs * NW + warp;
                    arow[rs] = (u < nN) ? order[n + u] * n : -1;
                    if (u < nN) {
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) T[rs][cs] = 0.0;
                        if (lane == 31) T[rs][CS - 1] = __ldg(bg + order[n + u]);
                        if (lane == rs) rowvar_l = order[n + u];
                    } else if (u != RT - 1) {
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) T[rs][cs] = 0.0;
                        if (lane == rs) rowvar_l = -1;
                    }
                }
#pragma unroll 2
                for (int k = 0; k < n; ++k) {
                    double d[CS];
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) d[cs] = Dsm[doff[cs] + k * dstr[cs]];
#pragma unroll
                    for (int rs = 0; rs < RS; ++rs) {
                        if (arow[rs] >= 0) {
                            const double av = __ldg(Ag + arow[rs] + k);
#pragma unroll
                            for (int cs = 0; cs < CS; ++cs) T[rs][cs] = fma(-av, d[cs], T[rs][cs]);
                        }
                    }
                }
            }
#pragma unroll
            for (int rs = 0; rs < RS; ++rs) sv[rs] = __shfl_sync(FULL, T[rs][CS - 1], 31);
            const int nslots = (nN > warp) ? (nN - warp + NW - 1) / NW : 0;

            // ---- stage 3a: phase 1 ----------------------------------------------------------------------------
            for (;;) {
                // my most negative right-hand side (in-register loop, identical in every lane)
                double smin = kInf;
                int slot = -1;
#pragma unroll
                for (int rs = 0; rs < RS; ++rs)
                    if (rs < nslots && sv[rs] < smin) { smin = sv[rs]; slot = rs; }
                PubHdr* myh = pub_hdr + buf * NW + warp;
                if (slot >= 0 && smin < -kTolFeas) {
                    // ratio test along the row: min ghat_j / (-e_j) over e_j < -tol
                    double bn = 0.0, bd = 0.0;   // best numerator / denominator (bd == 0: none)
                    int bq = 0;
                    static_switch<RS>(slot, [&](auto Rc) {
                        constexpr int R = decltype(Rc)::value;
#pragma unroll
                        for (int cs = 0; cs < CS; ++cs) {
                            const double e = -T[R][cs];
                            if (colvar[cs] >= 0 && e > kTolPivot) {
                                const double num = fmax(vec[cs], 0.0);
                                if (bd == 0.0 || num * bd < bn * e) { bn = num; bd = e; bq = cs; }
                            }
                        }
                    });
                    const double ratio = bn * fast_rcp(bd > 0.0 ? bd : 1.0);
                    unsigned long long kmin;
                    const int kl = warp_argmin_key((bd > 0.0) ? dkey(ratio) : KEY_INF, kmin);
                    const bool none = (kmin == KEY_INF);
                    const int kq = __shfl_sync(FULL, bq, kl);
                    const double p = -__shfl_sync(FULL, bd, kl);       // the pivot entry itself
                    const double rp = fast_rcp(p);
                    publish_row(slot, rp, kl + 32 * kq, pub_row + (size_t)(buf * NW + warp) * CT);
                    const int var = __shfl_sync(FULL, rowvar_l, slot);
                    if (lane == 0) {
                        myh->key = dkey(smin);
                        myh->srow = smin * rp;
                        myh->var = var;
                        myh->k = kl + 32 * kq;
                        myh->tile_row = slot * NW + warp;
                        myh->flag = none ? 2 : 0;
                    }
                } else if (lane == 0) {
                    myh->key = KEY_INF;
                }
                __syncthreads();
                unsigned long long kmin;
                const unsigned long long mykey = (lane < NW) ? pub_hdr[buf * NW + lane].key : KEY_INF;
                const int ww = warp_argmin_key(mykey, kmin);
                if (kmin == KEY_INF) break;                       // s >= 0 everywhere: phase 1 finished
                const PubHdr* h = pub_hdr + buf * NW + ww;
                if (h->flag == 2) { status = ST_INFEASIBLE; break; }
                if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const int k = h->k, kl = k & 31, kq = k >> 5;
                const int tr = h->tile_row, var_r = h->var;
                const double srow = h->srow;
                double pr[CS];
                const double* prs = pub_row + (size_t)(buf * NW + ww) * CT;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) pr[cs] = prs[lane + 32 * cs];
                const int cv = update_simplex(kq, kl, pr, srow, nslots, var_r);
                if (warp == ww) {
                    const int myslot = tr / NW;
                    set_row(myslot, pr, srow);
                    if (lane == myslot) rowvar_l = cv;
                }
                buf ^= 1;
                ++npiv_p1;
            }

            // ---- g becomes the replicated vector ---------------------------------------------------------------
            if (status == ST_OPTIMAL) {
                __syncthreads();
                if (warp == NW - 1) {
#pragma unroll
                    for (int cs = 0; cs < CS; ++cs) gbuf[lane + 32 * cs] = T[RS - 1][cs];
                }
                __syncthreads();
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) vec[cs] = gbuf[lane + 32 * cs];
            }

            // ---- stage 3b: phase 2 ----------------------------------------------------------------------------
            while (status == ST_OPTIMAL) {
                // entering column: most negative g (identical decision in every warp)
                double gmin = kInf;
                int bq = 0;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs)
                    if (colvar[cs] >= 0 && vec[cs] < gmin) { gmin = vec[cs]; bq = cs; }
                unsigned long long kmin;
                const int kl = warp_argmin_key(dkey(gmin), kmin);
                if (kmin >= dkey(-kTolFeas)) break;                // optimal
                const int kq = __shfl_sync(FULL, bq, kl);
                // ratio test over my rows, evaluated in the lane that owns column k (cross-multiplied, no division)
                double bs = 1.0, be = 0.0;                         // best ratio bs/be; be == 0: none yet (ratio +inf)
                int slot = -1;
                static_switch<CS>(kq, [&](auto Kc) {
                    constexpr int KQ = decltype(Kc)::value;
#pragma unroll
                    for (int rs = 0; rs < RS; ++rs) {
                        const double e = T[rs][KQ];
                        const double sc = fmax(sv[rs], 0.0);
                        if (rs < nslots && e > kTolPivot && sc * be < bs * e) { bs = sc; be = e; slot = rs; }
                    }
                });
                slot = __shfl_sync(FULL, slot, kl);
                PubHdr* myh = pub_hdr + buf * NW + warp;
                if (slot >= 0) {
                    bs = __shfl_sync(FULL, bs, kl);
                    be = __shfl_sync(FULL, be, kl);
                    const double rp = fast_rcp(be);
                    const double sraw = publish_row(slot, rp, kl + 32 * kq, pub_row + (size_t)(buf * NW + warp) * CT);
                    const int var = __shfl_sync(FULL, rowvar_l, slot);
                    if (lane == 0) {
                        myh->key = dkey(bs * rp);
                        myh->srow = sraw * rp;
                        myh->var = var;
                        myh->k = kl + 32 * kq;
                        myh->tile_row = slot * NW + warp;
                        myh->flag = 0;
                    }
                } else if (lane == 0) {
                    myh->key = KEY_INF;
                }
                __syncthreads();
                const unsigned long long mykey = (lane < NW) ? pub_hdr[buf * NW + lane].key : KEY_INF;
                const int ww = warp_argmin_key(mykey, kmin);
                if (kmin == KEY_INF) { status = ST_UNBOUNDED; break; }
                if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
                const PubHdr* h = pub_hdr + buf * NW + ww;
                const int tr = h->tile_row, var_r = h->var;
                const double srow = h->srow;
                double pr[CS];
                const double* prs = pub_row + (size_t)(buf * NW + ww) * CT;
#pragma unroll
                for (int cs = 0; cs < CS; ++cs) pr[cs] = prs[lane + 32 * cs];
                const int cv = update_simplex(kq, kl, pr, srow, nslots, var_r);
                if (warp == ww) {
                    const int myslot = tr / NW;
                    set_row(myslot, pr, srow);
                    if (lane == myslot) rowvar_l = cv;
                }
                buf ^= 1;
                ++npiv_p2;
            }
        }

        // ---- stage 4: x, objective, slacks, labels -----------------------------------------------------------------
        __syncthreads();
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0, nref = 0;
        if (need_generic) {
            status = -1;   // re-solved by the generic kernel (capi.cu)
        } else if (status == ST_OPTIMAL) {
            // where does every constraint sit now?
#pragma unroll
            for (int rs = 0; rs < RS; ++rs) {
                const int u = rs * NW + warp;
                if (u < nN && lane == 0) sval[u] = sv[rs];
            }
            if (lane < RS) {
                const int u = lane * NW + warp;
                if (u < nN && rowvar_l >= 0) basic_tile[rowvar_l] = u;
            }
            __syncthreads();
            for (int j = tid; j < n; j += NW * 32) {
                const int bt = basic_tile[colvar0[j]];
                sig[j] = (bt >= 0) ? sval[bt] : 0.0;
            }
            __syncthreads();
            for (int k = warp; k < n; k += NW) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(Dsm[(size_t)k * PD + j], sig[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) xbuf[k] = Dsm[(size_t)k * PD + n] - acc;
            }
            __syncthreads();
            if (warp == 0) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(__ldg(cg + j), xbuf[j], acc);
                acc = warp_sum(acc);
                if (lane == 0 && a.obj) a.obj[lp] = acc;
            }
            if (a.x)
                for (int j = tid; j < n; j += NW * 32) a.x[(size_t)lp * n + j] = xbuf[j];
            for (int i = warp; i < m; i += NW) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(__ldg(Ag + (size_t)i * n + j), xbuf[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) {
                    const double slack = __ldg(bg + i) - acc;
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
        }
        // An optimal instance whose active rows do not have (numerically) zero slack at the computed x -- an
        // ill-conditioned vertex -- is handed to the generic kernel, which holds the tableau in memory and can run a
        // step of iterative refinement on the final active set (simplex_generic.cu).
        if (__syncthreads_or(nref > 0) && status == ST_OPTIMAL) status = -1;
        if (!need_generic && status != ST_OPTIMAL) {
            for (int i = tid; i < m; i += NW * 32) lab[i] = 0;
            if (a.x)
                for (int j = tid; j < n; j += NW * 32) a.x[(size_t)lp * n + j] = 0.0;
            if (tid == 0 && a.obj) a.obj[lp] = __longlong_as_double(0x7ff8000000000000ll);
        }
        __syncthreads();
        if (lane == 0) {
            red[warp * 3 + 0] = nact;
            red[warp * 3 + 1] = nties;
            red[warp * 3 + 2] = nviol;
        }
        __syncthreads();
        if (tid == 0) {
            int t0 = 0, t1 = 0, t2 = 0;
            for (int w = 0; w < NW; ++w) {
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
struct Variant {
    int NW, RS, CS;
    int max_rows() const { return NW * RS - 1; }
    int max_cols() const { return 32 * CS - 1; }
};
constexpr Variant kSmall{4, 8, 1};    // n <= 31, m - n <= 31   : (10,5), (50,20)
constexpr Variant kMid{8, 8, 2};      // n <= 63, m - n <= 63
constexpr Variant kLarge{16, 7, 4};   // n <= 111, m - n <= 111 : (200,100)

const Variant* pick_variant(int m, int n) {
    const int rows = (m - n > n) ? (m - n) : n;
    for (const Variant* v : {&kSmall, &kMid, &kLarge})
        if (n <= v->max_cols() && rows <= v->max_rows() && n <= v->max_rows()) return v;
    return nullptr;
}

template <int NW, int RS, int CS, int MINB>
cudaError_t launch_variant(const SolveArgs& a, int sm_count, cudaStream_t st) {
    auto kern = simplex_regtile_kernel<NW, RS, CS, MINB>;
    const size_t smem = make_reg_layout(a.m, a.n, NW, 32 * CS).total;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)sm_count * per_sm;
    if (grid > a.B) grid = a.B;
    kern<<<(int)grid, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}
}  // namespace

bool regtile_supported(int m, int n) { return m >= n && pick_variant(m, n) != nullptr; }
size_t regtile_scratch_bytes(int, int, int) { return 0; }

cudaError_t launch_simplex_regtile(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const Variant* v = pick_variant(a.m, a.n);
    if (!v) return cudaErrorInvalidValue;
    if (v == &kSmall) return launch_variant<4, 8, 1, 6>(a, sm_count, st);
    if (v == &kMid) return launch_variant<8, 8, 2, 2>(a, sm_count, st);
    static const int alt = [] { const char* e = getenv("DDB_REGTILE_ALT"); return e ? atoi(e) : 0; }();
    if (alt == 1) return launch_variant<8, 14, 4, 1>(a, sm_count, st);
    return launch_variant<16, 7, 4, 1>(a, sm_count, st);
}

}  // namespace ddb
