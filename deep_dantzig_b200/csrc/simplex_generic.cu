// Generic batched fp64 simplex: one LP per thread block, condensed tableau in shared memory (plan 1, staged
// by 1-D bulk TMA) or in a per-CTA global-memory slab that lives in L2/HBM (plan 2, for shapes whose tableau
// exceeds 227 KB).  Persistent CTAs pull instance indices from an atomic work queue because the work per LP
// is bimodal (about half the instances at m = 2n are unbounded and exit early).
//
// Replaces, for a whole batch: LinProg model build + optimize + get_statuscode + get_active_constraints
// (reference src/data/gurobi_lp.py:11-29, 370-465) and the label assembly of create_lp_problem
// (reference src/data/randomlp_dataset.py:88-106).  The reference delegates the solve to Gurobi; the
// algorithm below is ours, parity is on results.
//
// Dictionary kept per LP (primal orientation, condensed):
//     sigma_i = s_i - sum_j P[i][j] nu_j        row i: constraint whose slack sigma_i is basic
//     z       = z0  + sum_j g[j]  nu_j          nu_j: free x_j before the crash, then an active slack
// Stages: (1) static-order crash: n Gauss-Jordan pivots that make every free x_j basic, rows taken in the
// order of the cosine score a_i.c/|a_i| (rows most opposed to c first), column = largest |entry|;
// (2) phase 1: dual-simplex-type pivots with artificial costs ghat = 1 until s >= 0;
// (3) phase 2: Dantzig primal simplex until g >= 0; ties everywhere broken by lowest variable index (Bland);
// (4) x from the rows frozen at the end of the crash, one step of iterative refinement on the final active set when
//     its residual is not negligible, slack = b - A x recomputed from the caller's A, labels = |slack| <= threshold
//     exactly as gurobi_lp.py:435-443.
#include <cstdlib>
#include "philox.cuh"

namespace ddb {

__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

struct Sel {
    double p;      // pivot element
    double gk;     // g[k] before the pivot
    double ghk;    // ghat[k] before the pivot
    int r, k;
    int flag;      // 0 = pivot, 1 = stage finished, 2 = infeasible / unbounded, 3 = skip row (crash)
};

struct Layout {
    size_t tab, s, g, gh, colbuf, xbuf, sig, rowvar, colvar, colvar0, where, order, rowfree, liveidx, rowstate, red, cands, sel, bar, total;
    size_t ring, rbar;   // global-memory plan, 512-thread instantiation: per-warp ring of tableau rows filled by bulk TMA
    int ring_k;          // slots per warp (0: no ring -- odd n or no room -- the update loads rows through registers)
};

constexpr int kRingWarps = 16;        // warps of the 512-thread instantiation
constexpr int kRingMaxSlots = 8;
constexpr size_t kSmemBudget = 232448 - 1024 - 64;   // 227 KB opt-in minus the launcher's margin and the static variables

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline Layout make_layout(int m, int n, bool smem_tab, int ring_slots = 0) {
    Layout L;
    size_t off = 0;
    L.tab = off;      off += smem_tab ? align_up((size_t)m * n * 8, 16) : 0;
    L.s = off;        off += align_up((size_t)m * 8, 16);
    L.g = off;        off += align_up((size_t)n * 8, 16);
    L.gh = off;       off += align_up((size_t)n * 8, 16);
    L.colbuf = off;   off += align_up((size_t)m * 8, 16);
    L.xbuf = off;     off += align_up((size_t)n * 8, 16);
    L.sig = off;      off += align_up((size_t)n * 8, 16);
    L.rowvar = off;   off += align_up((size_t)m * 4, 16);
    L.colvar = off;   off += align_up((size_t)n * 4, 16);
    L.colvar0 = off;  off += align_up((size_t)n * 4, 16);
    L.where = off;    off += align_up((size_t)m * 4, 16);
    L.order = off;    off += align_up((size_t)m * 4, 16);
    L.rowfree = off;  off += align_up((size_t)m * 4, 16);
    L.liveidx = off;  off += align_up((size_t)m * 4, 16);
    L.rowstate = off; off += align_up((size_t)m, 16);
    L.red = off;      off += 3 * 32 * 4;
    L.cands = off;    off += align_up(32 * sizeof(Cand), 16);
    L.sel = off;      off += align_up(sizeof(Sel), 16);
    L.bar = off;      off += 16;
    L.ring = L.rbar = 0;
    L.ring_k = 0;
    if (ring_slots > 0 && !smem_tab && (long long)m * n > 16384 && n % 2 == 0) {
        const size_t base = align_up(off, 128);
        const size_t bars = (size_t)kRingWarps * kRingMaxSlots * 8;
        const size_t row = (size_t)n * 8;
        if (base + bars + 2 * kRingWarps * row <= kSmemBudget) {
            size_t k = (kSmemBudget - base - bars) / (kRingWarps * row);
            if (k > (size_t)kRingMaxSlots) k = kRingMaxSlots;
            if (k > (size_t)ring_slots) k = ring_slots;
            L.ring_k = (int)k;
            L.rbar = base;
            L.ring = base + bars;
            off = L.ring + (size_t)kRingWarps * k * row;
        }
    }
    L.total = off;
    return L;
}

// DDB_PLAN2_RING (experiment, default 0): 0 = rows through registers, 1 = bulk-TMA row ring, 2 = ring + the leading rows of the
// next pivot requested ahead, 3 = as 2 with state-space-qualified proxy fences; DDB_PLAN2_RING_K caps the slots per warp.
// The ring is only laid out when it is used: its shared memory comes out of the L1 carve-out, which the register path wants.
static int generic_ring_mode() {
    static const int v = [] { const char* e = getenv("DDB_PLAN2_RING"); return e ? atoi(e) : 0; }();
    return v;
}
static int generic_ring_slots() {
    static const int v = [] { const char* e = getenv("DDB_PLAN2_RING_K"); const int k = e ? atoi(e) : kRingMaxSlots; return k < 2 ? 2 : k; }();
    return generic_ring_mode() > 0 ? v : 0;
}
size_t generic_smem_bytes(int m, int n, bool smem_tab) { return make_layout(m, n, smem_tab, generic_ring_slots()).total; }

// NTMAX: largest block the instantiation is launched with (1024, or 512 for the wide global-memory shapes, which trades
// warps for registers: more loads in flight per lane)
template <bool kSmemTab, int CPL, int NTMAX, bool kRing>
__global__ void __launch_bounds__(NTMAX, 1) simplex_generic_kernel(SolveArgs a, int ring_mode, int ring_slots) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = a.m, n = a.n;
    const Layout L = make_layout(m, n, kSmemTab, ring_slots);
    double* P = kSmemTab ? reinterpret_cast<double*>(smem_raw + L.tab)
                         : a.gtab + (size_t)blockIdx.x * ((size_t)m * n);
    double* s = reinterpret_cast<double*>(smem_raw + L.s);
    double* g = reinterpret_cast<double*>(smem_raw + L.g);
    double* gh = reinterpret_cast<double*>(smem_raw + L.gh);
    double* colbuf = reinterpret_cast<double*>(smem_raw + L.colbuf);
    double* xbuf = reinterpret_cast<double*>(smem_raw + L.xbuf);
    double* sig = reinterpret_cast<double*>(smem_raw + L.sig);
    int* rowvar = reinterpret_cast<int*>(smem_raw + L.rowvar);
    int* colvar = reinterpret_cast<int*>(smem_raw + L.colvar);
    int* colvar0 = reinterpret_cast<int*>(smem_raw + L.colvar0);
    int* where = reinterpret_cast<int*>(smem_raw + L.where);
    int* order = reinterpret_cast<int*>(smem_raw + L.order);
    int* rowfree = reinterpret_cast<int*>(smem_raw + L.rowfree);
    int* liveidx = reinterpret_cast<int*>(smem_raw + L.liveidx);      // live rows after the crash (phases 1 / 2 iterate over these)
    Cand* cands = reinterpret_cast<Cand*>(smem_raw + L.cands);         // per-warp candidates of the block-wide ratio test
    __shared__ int nlive_sm;
    uint8_t* rowstate = smem_raw + L.rowstate;
    Sel* sel = reinterpret_cast<Sel*>(smem_raw + L.sel);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    __shared__ long long cur_lp;

    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    uint32_t bar_parity = 0;

    if (kSmemTab) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    // Streaming update of the global-memory plan: every warp owns ring_k row slots and one mbarrier per slot; its lane 0
    // bulk-copies (cp.async.bulk, SASS UBLKCP) the next rows of the warp's share into free slots while the warp runs the
    // FMAs of the oldest one out of shared memory and stores the result straight back to the slab.  The copies need no
    // registers, so ring_k rows (up to 16 KB) per warp are in flight instead of the 32 loads per lane of the register path.
    static_assert(!kRing || (!kSmemTab && NTMAX <= 512), "the row ring belongs to the 512-thread global-memory instantiation");
    int ring_k = 0;
    if (kRing && nw == kRingWarps && ring_mode > 0) ring_k = L.ring_k;
    const bool ring_ahead = ring_mode > 1;   // request the leading rows of the next pivot at the end of this one
    double* wring = reinterpret_cast<double*>(smem_raw + L.ring) + (size_t)warp * ring_k * n;
    uint64_t* wbar = reinterpret_cast<uint64_t*>(smem_raw + L.rbar) + warp * kRingMaxSlots;
    unsigned r_issued = 0, r_done = 0;     // rows this warp has requested / consumed since the kernel started (warp-uniform)
    int npre = 0;                          // leading rows of this warp's share already requested for the NEXT pivot
    if (kRing && ring_k > 0) {
        if (lane == 0) {
            for (int q = 0; q < ring_k; ++q) mbar_init(&wbar[q], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }

    // One pivot on (r, k).  Every thread calls it with the same arguments (read from *sel after a barrier).
    // One pivot on (r, k).  Every thread calls it with the same arguments (read from *sel after a barrier).
    // have_col: colbuf already holds column k of the live rows (the block-wide ratio test of phase 2 gathered it), so the
    // strided re-read of the column is skipped.
    auto do_pivot = [&](int r, int k, double p, bool crash_mode, bool have_col) {
        const double rp = 1.0 / p;
        // B: normalise the pivot row, pull the pivot column out into colbuf and zero it in place
        for (int j = tid; j < n; j += nt) P[(size_t)r * n + j] = (j == k) ? rp : P[(size_t)r * n + j] * rp;
        if (have_col) {
            const int nl = nlive_sm;
            for (int q = tid; q < nl; q += nt) {
                const int i = liveidx[q];
                // (ring mode: the update takes column k of a row as 0 itself, see below, so it is not zeroed in place)
                if (i == r) colbuf[i] = 0.0; else if (!(kRing && ring_k > 0)) P[(size_t)i * n + k] = 0.0;
            }
        } else {
            for (int i = tid; i < m; i += nt) {
                double f = 0.0;
                const uint8_t st = rowstate[i];
                if (i != r && (st == ROW_LIVE || (crash_mode && st == ROW_CRASHED))) {
                    f = P[(size_t)i * n + k];
                    if (!(kRing && ring_k > 0)) P[(size_t)i * n + k] = 0.0;
                }
                colbuf[i] = f;
            }
        }
        if (tid == 0) {
            s[r] *= rp;
            sel->gk = g[k];
            sel->ghk = gh[k];
            g[k] = 0.0;
            gh[k] = 0.0;
        }
        // every generic-proxy store of this thread to the slab (this step's and the previous update's) is ordered before
        // the bulk copies (async proxy) that are issued after the barrier
        if (kRing && ring_k > 0) { if (ring_mode == 3) fence_proxy_async_global(); else fence_proxy_async_all(); }
        __syncthreads();
        // C: rank-1 update of every other row, of s, g and ghat
        double pr[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int j = lane + 32 * q;
            pr[q] = (j < n) ? P[(size_t)r * n + j] : 0.0;
        }
        const double sr = s[r];
        // RB rows per step, so that RB * CPL loads are in flight per lane (the global-memory tableau is bound by L2
        // round-trip latency): the register budget decides RB -- 64 registers at 1024 threads, 128 at 512
        constexpr int RB = (NTMAX <= 512) ? ((CPL <= 8) ? 4 : 2) : ((CPL <= 2) ? 2 : 1);
        const int nrows = crash_mode ? m : nlive_sm;           // the crash updates every row, phases 1 / 2 the live ones
        bool streamed = false;
        if constexpr (kRing) {
        if (ring_k > 0) {
            streamed = true;
            // Rows of this warp's share: list positions warp, warp + nw, ...  The first npre of them were requested at the
            // end of the previous pivot (their copies predate this pivot's step B, which only touched column k -- taken as 0
            // here for every row -- and the pivot row, which is skipped); the others are requested as slots free up.
            const uint32_t rowbytes = (uint32_t)n * 8u;
            int tp = warp + npre * nw;                         // producer cursor over this warp's share of the row list
            int t = 0;
            for (int q0 = warp; q0 < nrows; q0 += nw, ++t) {
                const int ix = crash_mode ? q0 : liveidx[q0];
                const double f = colbuf[ix];
                const bool pre = t < npre;
                if (f == 0.0 && !pre) continue;                // warp-uniform
                while ((int)(r_issued - r_done) < ring_k && tp < nrows) {
                    const int ixp = crash_mode ? tp : liveidx[tp];
                    tp += nw;
                    if (colbuf[ixp] != 0.0) {
                        const unsigned sl = r_issued % (unsigned)ring_k;
                        if (lane == 0) {
                            mbar_expect_tx(&wbar[sl], rowbytes);
                            tma_load_1d(wring + (size_t)sl * n, P + (size_t)ixp * n, rowbytes, &wbar[sl]);
                        }
                        ++r_issued;
                    }
                }
                const unsigned sl = r_done % (unsigned)ring_k;
                mbar_wait(&wbar[sl], (r_done / (unsigned)ring_k) & 1u);
                if (f != 0.0) {
                    const double* src = wring + (size_t)sl * n;
                    double* Pi = P + (size_t)ix * n;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int j = lane + 32 * q;
                        if (j < n) Pi[j] = fma(-f, pr[q], (j == k) ? 0.0 : src[j]);
                    }
                    if (lane == 0) s[ix] = fma(-f, sr, s[ix]);
                }
                __syncwarp();                                  // every lane has read the slot before lane 0 refills it
                ++r_done;
            }
            // request the leading rows of the share for the next pivot now: the copies travel while the block selects it
            if (ring_mode == 3) fence_proxy_async_global(); else fence_proxy_async_all();   // my stores above before the async-proxy reads below
            __syncwarp();
            npre = 0;
            for (int q0 = warp; ring_ahead && q0 < nrows && npre < ring_k; q0 += nw, ++npre) {
                const int ixp = crash_mode ? q0 : liveidx[q0];
                const unsigned sl = r_issued % (unsigned)ring_k;
                if (lane == 0) {
                    mbar_expect_tx(&wbar[sl], rowbytes);
                    tma_load_1d(wring + (size_t)sl * n, P + (size_t)ixp * n, rowbytes, &wbar[sl]);
                }
                ++r_issued;
            }
        }
        }
        if (!streamed)
        for (int q0 = warp; q0 < nrows; q0 += RB * nw) {
            int ix[RB];
            double fx[RB];
            double vx[RB][CPL];
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const int qq = q0 + b * nw;
                const bool ok = qq < nrows;
                ix[b] = ok ? (crash_mode ? qq : liveidx[qq]) : 0;
                fx[b] = ok ? colbuf[ix[b]] : 0.0;
            }
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const double* Pi = P + (size_t)ix[b] * n;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int j = lane + 32 * q;
                    vx[b][q] = (j < n && fx[b] != 0.0) ? Pi[j] : 0.0;
                }
            }
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                if (fx[b] != 0.0) {
                    double* Pi = P + (size_t)ix[b] * n;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int j = lane + 32 * q;
                        if (j < n) Pi[j] = fma(-fx[b], pr[q], vx[b][q]);
                    }
                    if (lane == 0) s[ix[b]] = fma(-fx[b], sr, s[ix[b]]);
                }
            }
        }
        {
            const double gk = sel->gk, ghk = sel->ghk;
            for (int j = tid; j < n; j += nt) {
                const double prj = P[(size_t)r * n + j];
                g[j] = fma(-gk, prj, g[j]);
                gh[j] = fma(-ghk, prj, gh[j]);
            }
        }
        __syncthreads();
    };

    // the rows requested ahead for a pivot that never came (end of a stage: the row list changes, or the LP is finished)
    auto drain_ring = [&]() {
        if constexpr (kRing) {
            while (r_done != r_issued) {
                mbar_wait(&wbar[r_done % (unsigned)ring_k], (r_done / (unsigned)ring_k) & 1u);
                ++r_done;
            }
            npre = 0;
        }
    };

    if (a.only_flagged && *a.flag_count == 0) return;   // nothing was handed over by the register-tiled kernel

    for (;;) {
        if (tid == 0) cur_lp = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long lp = cur_lp;
        if (lp >= a.B) break;
        if (a.only_flagged && a.status[lp] != -1) {
            __syncthreads();   // everyone has read cur_lp before thread 0 overwrites it
            continue;
        }
        const double* Ag = a.A + (size_t)lp * m * n;
        const double* bg = a.b + (size_t)lp * m;
        const double* cg = a.c + (size_t)lp * n;
        if (a.gen) {
            // fused generate -> solve -> label: an instance the row-per-thread kernel handed over exists nowhere unless the
            // caller asked for A -- draw it again (same counters, same bits) into this CTA's slab
            double* slab = a.slab + (size_t)blockIdx.x * slab_doubles(m, n);
            double* Aw = a.A ? const_cast<double*>(Ag) : slab;
            double* bw = a.A ? const_cast<double*>(bg) : slab + slab_b_offset(m, n);
            double* cw = a.A ? const_cast<double*>(cg) : slab + slab_c_offset(m, n);
            generate_instance_cta_notile((uint64_t)a.gen_key, (uint64_t)(a.gen_first + lp), m, n, a.gen_density, Aw, bw, cw, xbuf);
            Ag = Aw; bg = bw; cg = cw;
            fence_proxy_async_all();       // generic-proxy stores of the slab before the bulk-TMA staging below reads it
            __syncthreads();
        }
        const uint8_t* mask = a.row_mask ? a.row_mask + (size_t)lp * m : nullptr;

        // ---- stage the instance ------------------------------------------------------------------------
        bool used_tma = false;
        if (kSmemTab) {
            const size_t bytesA = (size_t)m * n * 8;
            const bool aligned = (bytesA % 16 == 0) && ((reinterpret_cast<uintptr_t>(Ag) & 15) == 0);
            if (aligned) {
                used_tma = true;
                if (tid == 0) {
                    fence_proxy_async();   // order earlier generic-proxy accesses of the tableau before the async writes
                    mbar_expect_tx(bar, (uint32_t)bytesA);
                    const uint32_t chunk = 32768;
                    for (size_t off = 0; off < bytesA; off += chunk) {
                        const uint32_t nb = (uint32_t)((bytesA - off < chunk) ? (bytesA - off) : chunk);
                        tma_load_1d(reinterpret_cast<unsigned char*>(P) + off,
                                    reinterpret_cast<const unsigned char*>(Ag) + off, nb, bar);
                    }
                }
            }
        }
        if (!used_tma) {
            for (size_t e = tid; e < (size_t)m * n; e += nt) P[e] = Ag[e];
        }
        for (int i = tid; i < m; i += nt) {
            s[i] = bg[i];
            rowvar[i] = i;
            where[i] = i;   // >= 0: basic in that row; < 0: nonbasic in column -(w+1)
            rowfree[i] = -1;
            rowstate[i] = (mask && mask[i] == 0) ? ROW_EXCLUDED : ROW_LIVE;
        }
        for (int j = tid; j < n; j += nt) {
            g[j] = cg[j];
            gh[j] = 1.0;
            colvar[j] = -1;
        }
        if (used_tma) {
            mbar_wait(bar, bar_parity);
            bar_parity ^= 1;
        }
        __syncthreads();

        // ---- crash order: rank rows by a_i.c / |a_i| (ascending), excluded rows last -----------------------
        for (int i = warp; i < m; i += nw) {
            double dot = 0.0, nn = 0.0;
            for (int j = lane; j < n; j += 32) {
                const double v = P[(size_t)i * n + j];
                dot = fma(v, g[j], dot);
                nn = fma(v, v, nn);
            }
            dot = warp_sum(dot);
            nn = warp_sum(nn);
            if (lane == 0) colbuf[i] = (rowstate[i] == ROW_EXCLUDED) ? kInf : (nn > 0.0 ? dot / sqrt(nn) : kInf * 0.5);
        }
        __syncthreads();
        for (int i = tid; i < m; i += nt) {
            const double v = colbuf[i];
            int rank = 0;
            for (int i2 = 0; i2 < m; ++i2) {
                const double v2 = colbuf[i2];
                rank += (v2 < v) || (v2 == v && i2 < i);
            }
            order[rank] = i;
        }
        __syncthreads();

        int npiv_crash = 0, npiv_p1 = 0, npiv_p2 = 0;
        int status = ST_OPTIMAL;

        // ---- stage 1: crash --------------------------------------------------------------------------------
        for (int oi = 0; oi < m && npiv_crash < n; ++oi) {
            const int r = order[oi];
            if (rowstate[r] == ROW_EXCLUDED) break;   // uniform: shared memory read after a barrier
            if (warp == 0) {
                Cand cd{kInf, kBigVar, -1};
                double er[CPL];                    // the row's entries first: independent loads, one round trip
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int j = lane + 32 * q;
                    er[q] = (j < n) ? P[(size_t)r * n + j] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int j = lane + 32 * q;
                    if (j < n && colvar[j] < 0) {
                        const double v = -fabs(er[q]);
                        if (cand_better(v, j, cd.val, cd.var)) cd = Cand{v, j, j};
                    }
                }
                cd = warp_argmin(cd);
                if (lane == 0) {
                    sel->r = r;
                    sel->k = cd.idx;
                    if (cd.idx < 0 || -cd.val < kTolCrash) {
                        sel->flag = 3;
                    } else {
                        sel->flag = 0;
                        sel->p = P[(size_t)r * n + cd.idx];
                    }
                }
            }
            __syncthreads();
            const int flag = sel->flag, k = sel->k;
            const double p = sel->p;
            __syncthreads();   // everyone has read *sel before anyone rewrites it
            if (flag == 3) continue;
            do_pivot(r, k, p, true, false);
            if (tid == 0) {
                rowstate[r] = ROW_CRASHED;
                rowfree[r] = k;
                colvar[k] = r;
                where[r] = -(k + 1);
            }
            ++npiv_crash;
            __syncthreads();
        }
        drain_ring();
        for (int j = tid; j < n; j += nt) {
            colvar0[j] = colvar[j];
            gh[j] = 1.0;
        }
        if (tid == 0) {
            int nl = 0;
            for (int i = 0; i < m; ++i)
                if (rowstate[i] == ROW_LIVE) liveidx[nl++] = i;
            nlive_sm = nl;
        }
        __syncthreads();

        // ---- stage 2: phase 1 (row-first pivots until s >= 0) ------------------------------------------------
        for (;;) {
            if (warp == 0) {
                Cand cr{kInf, kBigVar, -1};
                for (int i = lane; i < m; i += 32) {
                    if (rowstate[i] == ROW_LIVE) {
                        const double v = s[i];
                        if (cand_better(v, rowvar[i], cr.val, cr.var)) cr = Cand{v, rowvar[i], i};
                    }
                }
                cr = warp_argmin(cr);
                int flag = 0, kk = -1;
                if (cr.idx < 0 || cr.val >= -kTolFeas) {
                    flag = 1;
                } else {
                    const int r = cr.idx;
                    Cand ck{kInf, kBigVar, -1};
                    double er[CPL];                // the row's entries first: independent loads, one round trip
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int j = lane + 32 * q;
                        er[q] = (j < n) ? P[(size_t)r * n + j] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int j = lane + 32 * q;
                        const double e = er[q];
                        if (j < n && colvar[j] >= 0 && e < -kTolPivot) {
                            const double ratio = fmax(gh[j], 0.0) / (-e);
                            if (cand_better(ratio, colvar[j], ck.val, ck.var)) ck = Cand{ratio, colvar[j], j};
                        }
                    }
                    ck = warp_argmin(ck);
                    kk = ck.idx;
                    if (kk < 0) flag = 2;
                }
                if (lane == 0) {
                    sel->flag = flag;
                    sel->r = cr.idx;
                    sel->k = kk;
                    if (flag == 0) sel->p = P[(size_t)cr.idx * n + kk];
                }
            }
            __syncthreads();
            const int flag = sel->flag, r = sel->r, k = sel->k;
            const double p = sel->p;
            __syncthreads();
            if (flag == 1) break;
            if (flag == 2) { status = ST_INFEASIBLE; break; }
            if (npiv_p1 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
            do_pivot(r, k, p, false, false);
            if (tid == 0) {
                const int vr = rowvar[r], vk = colvar[k];
                rowvar[r] = vk;
                colvar[k] = vr;
                where[vk] = r;
                where[vr] = -(k + 1);
            }
            ++npiv_p1;
            __syncthreads();
        }

        // ---- stage 3: phase 2 (column-first Dantzig pivots until g >= 0) -----------------------------------
        // Pricing (argmin g, shared memory) by warp 0; the ratio test by the WHOLE block: every thread reads the entry of
        // its live row(s) in column k -- one round trip to the tableau instead of m / 32 sequential ones for one warp --
        // and leaves it in colbuf for the pivot, then a two-level argmin (warp, block) with the same Bland tie-break.
        while (status == ST_OPTIMAL) {
            if (warp == 0) {
                Cand ck{kInf, kBigVar, -1};
                int free_unbounded = 0;
                for (int j = lane; j < n; j += 32) {
                    if (colvar[j] >= 0) {
                        const double v = g[j];
                        if (cand_better(v, colvar[j], ck.val, ck.var)) ck = Cand{v, colvar[j], j};
                    } else if (fabs(g[j]) > kTolFeas) {
                        free_unbounded = 1;   // a free direction with non-zero cost that no row constrains
                    }
                }
                ck = warp_argmin(ck);
                free_unbounded = __any_sync(0xffffffffu, free_unbounded);
                int flag = 0;
                if (free_unbounded) flag = 2;
                else if (ck.idx < 0 || ck.val >= -kTolFeas) flag = 1;
                if (lane == 0) {
                    sel->flag = flag;
                    sel->k = ck.idx;
                }
            }
            __syncthreads();
            const int flag0 = sel->flag, k = sel->k;
            if (flag0 == 1) break;
            if (flag0 == 2) { status = ST_UNBOUNDED; break; }
            {
                Cand cr{kInf, kBigVar, -1};
                const int nl = nlive_sm;
                for (int q = tid; q < nl; q += nt) {
                    const int i = liveidx[q];
                    const double e = P[(size_t)i * n + k];
                    colbuf[i] = e;
                    if (e > kTolPivot) {
                        const double ratio = fmax(s[i], 0.0) / e;
                        if (cand_better(ratio, rowvar[i], cr.val, cr.var)) cr = Cand{ratio, rowvar[i], i};
                    }
                }
                cr = warp_argmin(cr);
                if (lane == 0) cands[warp] = cr;
            }
            __syncthreads();
            if (warp == 0) {
                Cand cr = (lane < nw) ? cands[lane] : Cand{kInf, kBigVar, -1};
                cr = warp_argmin(cr);
                if (lane == 0) {
                    sel->r = cr.idx;
                    sel->flag = (cr.idx < 0) ? 2 : 0;
                    if (cr.idx >= 0) sel->p = colbuf[cr.idx];
                }
            }
            __syncthreads();
            const int flag = sel->flag, r = sel->r;
            const double p = sel->p;
            __syncthreads();
            if (flag == 2) { status = ST_UNBOUNDED; break; }
            if (npiv_p2 >= a.max_iter) { status = ST_ITERATION_LIMIT; break; }
            do_pivot(r, k, p, false, true);
            if (tid == 0) {
                const int vr = rowvar[r], vk = colvar[k];
                rowvar[r] = vk;
                colvar[k] = vr;
                where[vk] = r;
                where[vr] = -(k + 1);
            }
            ++npiv_p2;
            __syncthreads();
        }

        drain_ring();
        // ---- stage 4: x, objective, slacks, labels ---------------------------------------------------------
        uint8_t* lab = a.labels + (size_t)lp * m;
        int nact = 0, nties = 0, nviol = 0;
        if (status == ST_OPTIMAL) {
            for (int j = tid; j < n; j += nt) {
                const int q = colvar0[j];
                double v = 0.0;
                if (q >= 0) {
                    const int w = where[q];
                    if (w >= 0) v = s[w];   // that constraint left the active set: its current slack
                }
                sig[j] = v;
                xbuf[j] = 0.0;
            }
            __syncthreads();
            for (int i = warp; i < m; i += nw) {
                if (rowstate[i] == ROW_CRASHED) {
                    double acc = 0.0;
                    for (int j = lane; j < n; j += 32) acc = fma(P[(size_t)i * n + j], sig[j], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) xbuf[rowfree[i]] = s[i] - acc;
                }
            }
            __syncthreads();
            // ---- one step of iterative refinement on the final active set -------------------------------------------------
            // rho_j = slack of the active (nonbasic) constraint of column j at the computed x, from the caller's A, b; it
            // should be 0.  When it is not (an ill-conditioned vertex: the errors of a few hundred pivots amplified),
            // move the nonbasic slacks from rho to 0 through the tableau:  d sigma = +sum_j P_wj rho_j for a crash
            // constraint that is basic in row w, d sigma = -rho_k for one that is nonbasic in column k, and
            // x -= D d sigma.  The corrected slacks are accurate to second order.
            {
                double* rho = gh;     // the phase-1 cost vector is dead by now
                double* dsig = g;     // so is g
                double rmax = 0.0;
                for (int j = warp; j < n; j += nw) {
                    const int q = colvar[j];
                    double acc = 0.0;
                    if (q >= 0)
                        for (int c2 = lane; c2 < n; c2 += 32) acc = fma(Ag[(size_t)q * n + c2], xbuf[c2], acc);
                    acc = warp_sum(acc);
                    const double r = (q >= 0) ? bg[q] - acc : 0.0;
                    if (lane == 0) rho[j] = r;
                    rmax = fmax(rmax, fabs(r));
                }
                if (__syncthreads_or(rmax > a.thr * 0.01)) {
                    for (int j0 = warp; j0 < n; j0 += nw) {
                        const int q0 = colvar0[j0];
                        double d = 0.0;
                        if (q0 >= 0) {
                            const int w = where[q0];
                            if (w >= 0) {
                                for (int j = lane; j < n; j += 32) d = fma(P[(size_t)w * n + j], rho[j], d);
                                d = warp_sum(d);
                            } else {
                                d = -rho[-w - 1];
                            }
                        }
                        if (lane == 0) sig[j0] = d;
                    }
                    __syncthreads();
                    for (int i = warp; i < m; i += nw) {
                        if (rowstate[i] == ROW_CRASHED) {
                            double acc = 0.0;
                            for (int j = lane; j < n; j += 32) acc = fma(P[(size_t)i * n + j], sig[j], acc);
                            acc = warp_sum(acc);
                            if (lane == 0) xbuf[rowfree[i]] -= acc;
                        }
                    }
                    __syncthreads();
                }
                (void)dsig;
            }
            if (warp == 0) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(cg[j], xbuf[j], acc);
                acc = warp_sum(acc);
                if (lane == 0 && a.obj) a.obj[lp] = acc;
            }
            if (a.x)
                for (int j = tid; j < n; j += nt) a.x[(size_t)lp * n + j] = xbuf[j];
            // slack from the caller's A (not the tableau), as the reference does from the solver's x
            for (int i = warp; i < m; i += nw) {
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(Ag[(size_t)i * n + j], xbuf[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) {
                    const double slack = bg[i] - acc;
                    const double as = fabs(slack);
                    const int active = as <= a.thr;
                    lab[i] = (uint8_t)active;
                    nact += active;
                    int tie = (as >= a.thr * 0.1 && as <= a.thr * 10.0);
                    if (rowstate[i] != ROW_EXCLUDED) tie |= (active != (where[i] < 0));
                    nties += tie;
                    nviol += (slack < -a.thr);
                }
            }
        } else {
            for (int i = tid; i < m; i += nt) lab[i] = 0;
            if (a.x)
                for (int j = tid; j < n; j += nt) a.x[(size_t)lp * n + j] = 0.0;
            if (tid == 0 && a.obj) a.obj[lp] = __longlong_as_double(0x7ff8000000000000ll);
        }
        // block-wide sums of the three counters (lane 0 of each warp holds partials)
        __syncthreads();
        int* red = reinterpret_cast<int*>(smem_raw + L.red);
        if (lane == 0) {
            red[warp * 3 + 0] = nact;
            red[warp * 3 + 1] = nties;
            red[warp * 3 + 2] = nviol;
        }
        __syncthreads();
        if (tid == 0) {
            int t0 = 0, t1 = 0, t2 = 0;
            for (int w = 0; w < nw; ++w) {
                t0 += red[w * 3 + 0];
                t1 += red[w * 3 + 1];
                t2 += red[w * 3 + 2];
            }
            a.status[lp] = status;
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
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host-side launcher
// ---------------------------------------------------------------------------------------------------------
template <bool kSmemTab, int CPL, int NTMAX, bool kRing = false>
static cudaError_t launch_one(const SolveArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = simplex_generic_kernel<kSmemTab, CPL, NTMAX, kRing>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, block, smem, st>>>(a, generic_ring_mode(), generic_ring_slots());
    return cudaGetLastError();
}

// block size of the generic kernel for a shape: the large global-memory shapes run 512 threads (see NTMAX)
int generic_block_threads(int m, int n, bool smem_tab) {
    const long long e = (long long)m * n;
    if (e <= 2048) return 128;
    if (e <= 8192) return 256;
    if (e <= 16384) return 512;
    if (!smem_tab) return 512;          // global-memory tableau: trade warps for registers (loads in flight)
    return 1024;
}

cudaError_t launch_simplex_generic(const SolveArgs& a, bool smem_tab, int grid, int block, cudaStream_t st) {
    const size_t smem = generic_smem_bytes(a.m, a.n, smem_tab);
    const int cpl = (a.n + 31) / 32;
    if (smem_tab) {
        if (cpl <= 1) return launch_one<true, 1, 1024>(a, grid, block, smem, st);
        if (cpl <= 2) return launch_one<true, 2, 1024>(a, grid, block, smem, st);
        if (cpl <= 4) return launch_one<true, 4, 1024>(a, grid, block, smem, st);
        if (cpl <= 8) return launch_one<true, 8, 1024>(a, grid, block, smem, st);
        if (cpl <= 16) return launch_one<true, 16, 1024>(a, grid, block, smem, st);
        return cudaErrorInvalidValue;
    }
    if (block <= 512 && (long long)a.m * a.n > 16384) {      // generic_block_threads(): large global-memory shapes
        if (generic_ring_mode() > 0) {                        // experiment: rows streamed through a bulk-TMA ring
            if (cpl <= 4) return launch_one<false, 4, 512, true>(a, grid, block, smem, st);
            if (cpl <= 8) return launch_one<false, 8, 512, true>(a, grid, block, smem, st);
            if (cpl <= 16) return launch_one<false, 16, 512, true>(a, grid, block, smem, st);
            return cudaErrorInvalidValue;
        }
        if (cpl <= 4) return launch_one<false, 4, 512>(a, grid, block, smem, st);
        if (cpl <= 8) return launch_one<false, 8, 512>(a, grid, block, smem, st);
        if (cpl <= 16) return launch_one<false, 16, 512>(a, grid, block, smem, st);
        return cudaErrorInvalidValue;
    }
    if (cpl <= 1) return launch_one<false, 1, 1024>(a, grid, block, smem, st);
    if (cpl <= 2) return launch_one<false, 2, 1024>(a, grid, block, smem, st);
    if (cpl <= 4) return launch_one<false, 4, 1024>(a, grid, block, smem, st);
    if (cpl <= 8) return launch_one<false, 8, 1024>(a, grid, block, smem, st);
    if (cpl <= 16) return launch_one<false, 16, 1024>(a, grid, block, smem, st);
    return cudaErrorInvalidValue;
}

}  // namespace ddb
