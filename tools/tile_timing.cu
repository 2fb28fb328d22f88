// Dev harness (not product code): runs the row-per-thread simplex kernel directly on Philox instances and prints
// throughput plus, when built with -DDDB_TIMING, the per-section cycle accounting of the phase-2 loop.
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a [-DDDB_TIMING] tools/rowreg_timing.cu -o tools/rowreg_timing
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../deep_dantzig_b200/csrc/generate.cu"
#include "../deep_dantzig_b200/csrc/experiments/simplex_tile2d.cu"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int m = argc > 1 ? atoi(argv[1]) : 200, n = argc > 2 ? atoi(argv[2]) : 100;
    const long long B = argc > 3 ? atoll(argv[3]) : 4736;
    const int reps = argc > 4 ? atoi(argv[4]) : 3;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    double *A, *b, *c, *x, *obj, *dbg; int *status, *nact, *piv, *ties, *viol; uint8_t* labels; unsigned long long* counter;
    CK(cudaMalloc(&A, B * m * n * 8)); CK(cudaMalloc(&b, B * m * 8)); CK(cudaMalloc(&c, B * n * 8));
    CK(cudaMalloc(&x, B * n * 8)); CK(cudaMalloc(&obj, B * 8)); CK(cudaMalloc(&status, B * 4)); CK(cudaMalloc(&nact, B * 4));
    CK(cudaMalloc(&piv, B * 16)); CK(cudaMalloc(&ties, B * 4)); CK(cudaMalloc(&viol, B * 4)); CK(cudaMalloc(&labels, B * m));
    CK(cudaMalloc(&counter, 64)); CK(cudaMalloc(&dbg, B * 8 * 16 * 8)); CK(cudaMemset(dbg, 0, B * 8 * 16 * 8));
    int launches = 0;
    CK(ddb::launch_generate(42, 0, B, m, n, 1.0, A, b, c, nullptr, prop.multiProcessorCount, 0, &launches));
    ddb::SolveArgs a{};
    a.m = m; a.n = n; a.B = B; a.A = A; a.b = b; a.c = c; a.row_mask = nullptr; a.thr = 1e-7;
    a.status = status; a.x = x; a.obj = obj; a.labels = labels; a.n_active = nact; a.pivots = piv; a.ties = ties;
    a.violations = viol; a.counter = counter; a.gtab = dbg; a.max_iter = 50 * (m + n); a.only_flagged = 0;
    a.flag_count = reinterpret_cast<int*>(counter + 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < reps; ++r) {
        CK(cudaMemset(counter, 0, 64));
        cudaEventRecord(e0);
        CK(ddb::launch_simplex_tile2d(a, prop.multiProcessorCount, 0));
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
    std::vector<double> hd(B * 16);
    cudaMemcpy(hd.data(), dbg, B * 16 * 8, cudaMemcpyDeviceToHost);
    const char* names[6] = {"stage 0 (scores, ranking)", "crash", "dump + GEMM", "phase 1", "phase 2", "stage 4 (x, labels)"};
    double tot[6] = {0}, pv[3] = {0}; double nopt2 = 0, t4opt = 0;
    for (long long i = 0; i < B; ++i) { const double* o = &hd[i * 16]; for (int q = 0; q < 6; ++q) tot[q] += o[q]; for (int q = 0; q < 3; ++q) pv[q] += o[6 + q]; if (o[9] == 2) { nopt2 += 1; t4opt += o[5]; } }
    double all = 0; for (int q = 0; q < 6; ++q) all += tot[q];
    printf("cycles per LP (thread 0 view): total %.0f\n", all / B);
    for (int q = 0; q < 6; ++q) printf("  %-28s %10.0f  (%.1f%%)\n", names[q], tot[q] / B, 100 * tot[q] / all);
    printf("  per pivot: crash %.0f, phase 1 %.0f, phase 2 %.0f clk; GEMM per column %.0f; stage 4 per optimal LP %.0f\n", tot[1] / pv[0], tot[3] / pv[1], tot[4] / pv[2], tot[2] / B / n, t4opt / nopt2);
#endif
    return 0;
}
