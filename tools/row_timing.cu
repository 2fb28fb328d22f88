// Dev harness (not product code): runs the row-per-thread simplex kernel (plan 0) directly on Philox instances and prints
// throughput plus the per-stage cycle accounting (thread 0 of each CTA, -DDDB_TIMING).  DDB_ROWREG_HYBRID=0 selects the
// all-register variant (two LPs per SM) instead of the hybrid register + shared-memory rows (three LPs per SM); a
// third argument "gen" times the in-solver generator variant (fused generate -> solve -> label).
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -DDDB_ROWREG_ONLY_BIG [-DDDB_TIMING] tools/row_timing.cu -o tools/row_timing
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../deep_dantzig_b200/csrc/generate.cu"
#include "../deep_dantzig_b200/csrc/simplex_rowreg.cu"
#include "../deep_dantzig_b200/csrc/simplex_rowreg_gen.cu"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int m = argc > 1 ? atoi(argv[1]) : 200, n = argc > 2 ? atoi(argv[2]) : 100;
    const long long B = argc > 3 ? atoll(argv[3]) : 4736;
    const int reps = argc > 4 ? atoi(argv[4]) : 3;
    const bool gen = argc > 5 && !strcmp(argv[5], "gen");
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sm = prop.multiProcessorCount;
    double *A, *b, *c, *x, *obj, *dbg, *dscr, *slab; int *status, *nact, *piv, *ties, *viol; uint8_t* labels; unsigned long long* counter;
    CK(cudaMalloc(&A, B * m * n * 8)); CK(cudaMalloc(&b, B * m * 8)); CK(cudaMalloc(&c, B * n * 8));
    CK(cudaMalloc(&x, B * n * 8)); CK(cudaMalloc(&obj, B * 8)); CK(cudaMalloc(&status, B * 4)); CK(cudaMalloc(&nact, B * 4));
    CK(cudaMalloc(&piv, B * 16)); CK(cudaMalloc(&ties, B * 4)); CK(cudaMalloc(&viol, B * 4)); CK(cudaMalloc(&labels, B * m));
    CK(cudaMalloc(&counter, 64));
    const int grid = gen ? ddb::rowreg_gen_grid(m, n, sm) : ddb::rowreg_grid(m, n, sm);
    const size_t save_bytes = ddb::rowreg_rows_scratch_bytes(m, n, grid);     // saved rows first, stage timers behind them
    const size_t dbg_bytes = save_bytes + (size_t)B * 8 * 8;
    CK(cudaMalloc(&dbg, dbg_bytes)); CK(cudaMemset(dbg, 0, dbg_bytes));
    const size_t dbytes = ddb::rowreg_d_scratch_bytes(m, n, grid);
    CK(cudaMalloc(&dscr, dbytes ? dbytes : 16));
    CK(cudaMalloc(&slab, (size_t)grid * ddb::slab_doubles(m, n) * 8));
    int launches = 0;
    CK(ddb::launch_generate(42, 0, B, m, n, 1.0, A, b, c, nullptr, sm, 0, &launches));
    ddb::SolveArgs a{};
    a.m = m; a.n = n; a.B = B; a.A = gen ? nullptr : A; a.b = gen ? nullptr : b; a.c = gen ? nullptr : c; a.row_mask = nullptr; a.thr = 1e-7;
    a.status = status; a.x = x; a.obj = obj; a.labels = labels; a.n_active = nact; a.pivots = piv; a.ties = ties;
    a.violations = viol; a.counter = counter; a.gtab = dbg; a.dscr = dscr; a.slab = slab; a.max_iter = 50 * (m + n); a.only_flagged = 0;
    a.flag_count = reinterpret_cast<int*>(counter + 2);
    a.gen = gen ? 1 : 0; a.gen_key = 42; a.gen_first = 0; a.gen_density = 1.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("---- rowreg %s, grid %d (%d CTAs/SM) ----\n", gen ? "with in-solver generator" : "(instances from HBM)", grid, grid / sm);
    for (int r = 0; r < reps; ++r) {
        CK(cudaMemset(counter, 0, 64));
        cudaEventRecord(e0);
        if (gen) CK(ddb::launch_simplex_rowreg_gen(a, sm, 0));
        else CK(ddb::launch_simplex_rowreg(a, sm, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("(%d,%d) B=%lld: %.3f ms, %.0f LP/s\n", m, n, B, ms, B / ms * 1e3);
    }
    std::vector<int> hs(B), hp(B * 4);
    cudaMemcpy(hs.data(), status, B * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hp.data(), piv, B * 16, cudaMemcpyDeviceToHost);
    long long nopt = 0, flagged = 0; double pc = 0, p1 = 0, p2 = 0;
    for (long long i = 0; i < B; ++i) { nopt += hs[i] == 2; flagged += hs[i] == -1; pc += hp[i*4]; p1 += hp[i*4+1]; p2 += hp[i*4+2]; }
    printf("optimal %.3f flagged %lld mean pivots crash %.1f p1 %.1f p2 %.1f\n", (double)nopt / B, flagged, pc / B, p1 / B, p2 / B);
#ifdef DDB_TIMING
    std::vector<double> hd(B * 8);
    cudaMemcpy(hd.data(), reinterpret_cast<char*>(dbg) + save_bytes, B * 8 * 8, cudaMemcpyDeviceToHost);
    const char* names[6] = {"stage 0 (scores, ranking)", "crash", "dump + product", "phase 1", "phase 2", "stage 4 (x, labels)"};
    double tot[6] = {0}, all = 0;
    for (long long i = 0; i < B; ++i) for (int q = 0; q < 6; ++q) { tot[q] += hd[i * 8 + q]; all += hd[i * 8 + q]; }
    printf("cycles per LP (thread 0 view): total %.0f\n", all / B);
    for (int q = 0; q < 6; ++q) printf("  %-28s %10.0f  (%.1f%%)\n", names[q], tot[q] / B, 100.0 * tot[q] / all);
    printf("  per pivot: crash %.0f, phase 1 %.0f, phase 2 %.0f clk; product per column %.0f; stage 4 per optimal LP %.0f\n",
           tot[1] / pc, tot[3] / (p1 > 0 ? p1 : 1), tot[4] / (p2 > 0 ? p2 : 1), tot[2] / B / n, tot[5] / (nopt > 0 ? nopt : 1));
#endif
    return 0;
}
