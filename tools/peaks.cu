// On-chip peak microbenchmarks for the rooflines MEASURED_PEAKS.json does not hold:
// fp64 FMA throughput and shared-memory bandwidth of the read-modify-write pattern the pivot update uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a, double b) {
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
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9);
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
    cudaError_t e = cudaGetLastError();
    printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"fp64_tflops\": %.3f, \"smem_tbs\": %.3f, \"how\": \"tools/peaks.cu: "
           "8 independent DFMA chains/thread, 2x1024 threads/SM (2 flop per FMA); shared-memory read-modify-write of "
           "8-byte words, conflict-free, 1024 threads/SM (16 bytes per word per pass); best of 5\", \"cuda_error\": \"%s\"}\n",
           prop.name, sms, best_fp64, best_smem, cudaGetErrorString(e));
    return e != cudaSuccess;
}
