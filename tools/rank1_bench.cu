// Dev microbenchmark (not product code): the rank-1 row update of simplex_rowreg.cu in isolation.
// Each thread holds NC doubles; per iteration: T[c] -= f * row[c] with row broadcast from shared memory.
#include <cstdio>
#include <cuda_runtime.h>
template <int NC, int MODE>
__global__ void __launch_bounds__(128, 2) k_rank1(double* out, long long* cyc, int iters, double f0) {
    __shared__ __align__(16) double row[2][128];
    double T[NC];
    for (int c = 0; c < NC; ++c) T[c] = threadIdx.x * 1e-3 + c;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&row[0][0])[i] = 1e-6 * i;
    __syncthreads();
    double f = f0 + threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double* pr = row[it & 1];
        const double nf = -f;
        if (MODE == 0) {            // LDS.128 broadcast
            const double2* p2 = reinterpret_cast<const double2*>(pr);
#pragma unroll
            for (int c2 = 0; c2 < NC / 2; ++c2) {
                const double2 v = p2[c2];
                T[2 * c2] = fma(nf, v.x, T[2 * c2]);
                T[2 * c2 + 1] = fma(nf, v.y, T[2 * c2 + 1]);
            }
            if (NC & 1) T[NC - 1] = fma(nf, pr[NC - 1], T[NC - 1]);
        } else if (MODE == 1) {     // LDS.64 broadcast
#pragma unroll
            for (int c = 0; c < NC; ++c) T[c] = fma(nf, pr[c], T[c]);
        } else {                    // no shared memory: multiplicand from a register
#pragma unroll
            for (int c = 0; c < NC; ++c) T[c] = fma(nf, f0, T[c]);
        }
        f = f * 0.999;
    }
    __syncthreads();        // time the LAST warp: with a greedy-then-oldest scheduler warp 0 alone finishes early (the first
                            // version of this benchmark read 1.58 clk per warp-DFMA per SMSP that way, above the hardware rate)
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < NC; ++c) s += T[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 296 * 128 * 8); cudaMalloc(&cyc, 296 * 8);
    const int iters = 2000;
    long long h[296];
#define RUN(NC, MODE, name)                                                                        \
    for (int grid : {148, 296}) {                                                                  \
        k_rank1<NC, MODE><<<grid, 128>>>(out, cyc, iters, 1e-7); cudaDeviceSynchronize();           \
        k_rank1<NC, MODE><<<grid, 128>>>(out, cyc, iters, 1e-7); cudaDeviceSynchronize();           \
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);                                      \
        double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;                   \
        printf("%-28s NC=%d CTAs/SM=%d: %.1f clk per rank-1 (per warp), %.2f clk per DFMA-warp-instr per SMSP\n", name, NC, grid / 148, avg / iters, avg / iters / (NC * (grid / 148))); \
    }
    RUN(101, 0, "LDS.128 broadcast")
    RUN(101, 1, "LDS.64 broadcast")
    RUN(101, 2, "register multiplicand")
    RUN(48, 0, "LDS.128 broadcast")
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
