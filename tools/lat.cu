// Dev microbenchmark (not product code): dependent-chain latencies of the instructions on the simplex critical path.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k_lat(double* out, long long* cyc, int dummy) {
    __shared__ double sm[1024];
    __shared__ unsigned long long skey[32];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double x = 1.0 + lane * 1e-9, y = 0.999999 + dummy;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // LDS chain (pointer chasing through shared memory)
    int idx = lane + dummy;
    int* smi = reinterpret_cast<int*>(sm);
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) smi[i] = (i * 7 + 3) & 2047;
    __syncthreads();
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) idx = smi[idx];
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // SHFL chain
    int v = idx;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // REDUX chain
    unsigned u = v;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) u = __reduce_min_sync(0xffffffffu, u + lane);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // ballot chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) u = __ballot_sync(0xffffffffu, (u >> (lane & 7)) & 1) + i;
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // MUFU.RCP64H + newton (fast_rcp) chain
    double r = x + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        double q;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r));
        double e = fma(-r, q, 1.0); q = fma(q, e, q); e = fma(-r, q, 1.0); q = fma(q, e, q);
        r = q + 1.5;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // STS -> syncwarp -> LDS round trip
    double w = r;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        sm[lane] = w;
        __syncwarp();
        w = sm[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // __syncthreads chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // 51 STS.128 from one lane (publish)
    __syncthreads();
    t0 = clock64();
    if (lane == 3) {
        double2* p2 = reinterpret_cast<double2*>(sm);
#pragma unroll
        for (int i = 0; i < 51; ++i) p2[i] = make_double2(w + i, x + i);
    }
    __syncwarp();
    double z = sm[lane];
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0 + (z == 12345.0);
    // DSETP+select chain (double compare)
    double a = w, b2 = x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { a = (a < b2) ? b2 + 1e-9 : a; b2 = (b2 < a) ? a : b2; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
    out[threadIdx.x] = x + v + u + r + w + a + b2 + z;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    const char* names[] = {"DFMA dep", "DMUL dep", "LDS dep (int)", "SHFL dep", "REDUX.MIN dep", "BALLOT dep", "fast_rcp+DADD (x64)", "STS->syncwarp->LDS+DADD (x64)", "__syncthreads (x64)", "51 STS.128 one lane + LDS", "DSETP+sel pair"};
    const int div[] = {N, N, N, N, N, N, 64, 64, 64, 1, N};
    for (int threads : {32, 128}) {
        k_lat<<<1, threads>>>(out, cyc, 0); cudaDeviceSynchronize();
        k_lat<<<1, threads>>>(out, cyc, 0); cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
        printf("threads/CTA = %d (one CTA, otherwise idle GPU)\n", threads);
        for (int i = 0; i < 11; ++i) printf("  %-32s %8.1f clk\n", names[i], (double)h[i] / div[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
