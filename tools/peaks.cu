// On-chip peak microbenchmarks for the rooflines MEASURED_PEAKS.json does not hold:
// fp64 FMA throughput and shared-memory bandwidth of the read-modify-write pattern the pivot update uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

// cyc[2 * block] = SM cycles (clock64), cyc[2 * block + 1] = nanoseconds (globaltimer) of the same loop: their ratio is the SM
// clock the loop really ran at (a full-GPU fp64 loop runs power-capped well below the 1 965 MHz boost clock), so the
// wall-clock TFLOP/s can be split into "FMA lanes per clock per SM" x "clock" and compared with clock64-based microbenchmarks
// such as tools/rank1_bench.cu.
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a, double b, long long* cyc) {
    const long long c0 = clock64();
    const unsigned long long g0 = gtimer();
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();        // the LAST warp's finish: a greedy-then-oldest scheduler lets warp 0 finish well before its siblings
    if (cyc && threadIdx.x == 0) {
        cyc[2 * blockIdx.x] = clock64() - c0;
        cyc[2 * blockIdx.x + 1] = (long long)(gtimer() - g0);
    }
}

// each thread owns 8-byte words at conflict-free addresses: P[j] = fma(-f, pr, P[j]) repeated
__global__ void __launch_bounds__(1024) smem_rmw_kernel(double* out, int iters, double f) {
    extern __shared__ double sm[];
    const int words = 24576;   // 192 KB
    for (int j = threadIdx.x; j < words; j += blockDim.x) sm[j] = j * 1e-6;
    __syncthreads();
    const double pr = 1e-9;
    for (int it = 0; it < iters; ++it) {
        for (int j = threadIdx.x; j < words; j += blockDim.x) sm[j] = fma(-f, pr, sm[j]);
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 2 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // fp64
    const int iters = 1 << 14;
    double best_fp64 = 0;
    long long* cyc;
    cudaMalloc(&cyc, sizeof(long long) * sms * 4);
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8 * iters * (double)sms * 2 * 1024 / (ms * 1e-3) / 1e12;
        if (tf > best_fp64) best_fp64 = tf;
    }
    // smem
    cudaFuncSetAttribute(smem_rmw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    const int it2 = 2000;
    double best_smem = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        smem_rmw_kernel<<<sms, 1024, 196608>>>(out, it2, 0.5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tb = 16.0 * 24576 * it2 * (double)sms / (ms * 1e-3) / 1e12;
        if (tb > best_smem) best_smem = tb;
    }
    // per-clock view of the fp64 loop: cycles and nanoseconds of the last repetition
    long long hc[4 * 256];
    cudaMemcpy(hc, cyc, sizeof(long long) * sms * 4, cudaMemcpyDeviceToHost);
    double cyc_avg = 0, ns_avg = 0;
    for (int i = 0; i < sms * 2; ++i) { cyc_avg += hc[2 * i]; ns_avg += hc[2 * i + 1]; }
    cyc_avg /= sms * 2; ns_avg /= sms * 2;
    const double ghz = cyc_avg / ns_avg;                                   // SM clock during the loop
    const double lanes_per_clk_sm = 8.0 * iters * 2 * 1024 / cyc_avg;      // FMA lanes per clock per SM (two CTAs per SM)
    // the same loop on ONE SM (no power cap): what a clock64-based microbenchmark sees
    dfma_kernel<<<1, 1024>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, sizeof(long long) * 2, cudaMemcpyDeviceToHost);
    const double ghz1 = (double)hc[0] / (double)hc[1];
    const double lanes1 = 8.0 * iters * 1024 / (double)hc[0];          // one CTA of 1024 threads alone on its SM
    printf("{\"fp64_loop_sm_clock_ghz\": %.3f, \"fp64_fma_lanes_per_clk_per_sm\": %.2f, \"fp64_clk_per_warp_dfma_per_smsp\": %.3f, "
           "\"one_sm_clock_ghz\": %.3f, \"one_sm_fma_lanes_per_clk\": %.2f, \"fp64_tflops_at_boost_1965mhz\": %.2f}\n",
           ghz, lanes_per_clk_sm, 128.0 / lanes_per_clk_sm, ghz1, lanes1, lanes_per_clk_sm * 2 * sms * 1.965e9 / 1e12);
    cudaError_t e = cudaGetLastError();
    printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"fp64_tflops\": %.3f, \"smem_tbs\": %.3f, \"how\": \"tools/peaks.cu: "
           "8 independent DFMA chains/thread, 2x1024 threads/SM (2 flop per FMA); shared-memory read-modify-write of "
           "8-byte words, conflict-free, 1024 threads/SM (16 bytes per word per pass); best of 5\", \"cuda_error\": \"%s\"}\n",
           prop.name, sms, best_fp64, best_smem, cudaGetErrorString(e));
    return e != cudaSuccess;
}
