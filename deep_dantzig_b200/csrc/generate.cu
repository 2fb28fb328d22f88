// Counter-based instance generator (throughput mode).
//
// Replaces the serial per-seed loop of the reference (src/data/randomlp_dataset.py:58-63) and the generator part
// of create_lp_problem (:76-86):  A = randn(m,n); b = A.randn(n) + |randn(m)|; c = |randn(n)|.
// The reference's stream is numpy's legacy MT19937 + polar Gaussian, which a counter-based generator cannot
// reproduce bit for bit; parity mode therefore uploads numpy-generated instances, and this generator is pinned
// by Philox4x32-10 known-answer tests on the integer level (oracle/philox.py, tests/test_philox.py).
//
// Instance i is a pure function of (key, i): Philox4x32-10, key = (key_lo, key_hi),
// counter = (pair index, stream id, i_lo, i_hi).  One Philox block -> two 53-bit uniforms -> one Box-Muller
// pair -> normals for elements 2*pair and 2*pair+1 of that stream.
#include "philox.cuh"

namespace ddb {

// A: one thread per element pair, grid-stride, coalesced 16-byte stores when the pair is aligned.
__global__ void __launch_bounds__(256) generate_A_kernel(uint64_t key, long long first, long long B, int m, int n,
                                                         double density, double* __restrict__ A) {
    const long long per = (long long)m * n;
    const long long pairs_per = (per + 1) / 2;
    const long long total = B * pairs_per;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long k = t / pairs_per;
        const uint32_t pair = (uint32_t)(t - k * pairs_per);
        const uint64_t inst = (uint64_t)(first + k);
        double z0, z1;
        normal_pair(key, inst, STREAM_A, pair, z0, z1);
        if (density < 1.0) {
            uint32_t o[4];
            philox4x32_10(pair, STREAM_MASK, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)key,
                          (uint32_t)(key >> 32), o);
            if (u53(o[0], o[1]) >= density) z0 = 0.0;
            if (u53(o[2], o[3]) >= density) z1 = 0.0;
        }
        double* dst = A + k * per;
        const long long e = 2ll * pair;
        dst[e] = z0;
        if (e + 1 < per) dst[e + 1] = z1;
    }
}

// b, c, x0: one CTA per instance (A is read back, normally from L2).
__global__ void __launch_bounds__(256) generate_bc_kernel(uint64_t key, long long first, long long B, int m, int n,
                                                          const double* __restrict__ A, double* __restrict__ b,
                                                          double* __restrict__ c, double* __restrict__ x0out) {
    extern __shared__ double x0s[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (long long k = blockIdx.x; k < B; k += gridDim.x) {
        const uint64_t inst = (uint64_t)(first + k);
        for (int pr = tid; pr < (n + 1) / 2; pr += blockDim.x) {
            double z0, z1;
            normal_pair(key, inst, STREAM_X0, (uint32_t)pr, z0, z1);
            x0s[2 * pr] = z0;
            if (2 * pr + 1 < n) x0s[2 * pr + 1] = z1;
            normal_pair(key, inst, STREAM_C, (uint32_t)pr, z0, z1);
            c[k * n + 2 * pr] = fabs(z0);
            if (2 * pr + 1 < n) c[k * n + 2 * pr + 1] = fabs(z1);
        }
        __syncthreads();
        if (x0out)
            for (int j = tid; j < n; j += blockDim.x) x0out[k * n + j] = x0s[j];
        const double* Ak = A + k * (long long)m * n;
        for (int i = warp; i < m; i += nw) {
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc = fma(Ak[(long long)i * n + j], x0s[j], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                double z0, z1;
                normal_pair(key, inst, STREAM_EPS, (uint32_t)(i >> 1), z0, z1);
                b[k * m + i] = acc + fabs((i & 1) ? z1 : z0);
            }
        }
        __syncthreads();
    }
}

// Fused generator for even n: one CTA per instance, 32 rows at a time.  The CTA draws the tile's normals (pair index =
// (i n + j) / 2, the same counters as generate_A_kernel), stores them to A with 16-byte stores and keeps the tile in
// shared memory, from which b_i = a_i . x0 + |eps_i| is accumulated in exactly the order generate_bc_kernel uses
// (lane L: columns L, L + 32, ...; then the xor butterfly), so both paths produce the same bits and A is never read back.
constexpr int kGenRows = 32;
__global__ void __launch_bounds__(256) generate_fused_kernel(uint64_t key, long long first, long long B, int m, int n,
                                                             double density, double* __restrict__ A, double* __restrict__ b,
                                                             double* __restrict__ c, double* __restrict__ x0out) {
    extern __shared__ double gsm[];
    double* x0s = gsm;                       // [n]
    double* tile = gsm + ((n + 1) & ~1);     // [kGenRows][n]
    double* eps = tile + (size_t)kGenRows * n;   // [m]: |eps_i|, drawn up front by all threads, one Box-Muller pair per two rows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int half = n / 2;                  // pairs per row
    for (long long k = blockIdx.x; k < B; k += gridDim.x) {
        const uint64_t inst = (uint64_t)(first + k);
        for (int pr = tid; pr < half; pr += blockDim.x) {
            double z0, z1;
            normal_pair(key, inst, STREAM_X0, (uint32_t)pr, z0, z1);
            x0s[2 * pr] = z0;
            x0s[2 * pr + 1] = z1;
            normal_pair(key, inst, STREAM_C, (uint32_t)pr, z0, z1);
            c[k * n + 2 * pr] = fabs(z0);
            c[k * n + 2 * pr + 1] = fabs(z1);
        }
        // (round 2: this used to be one pair per ROW drawn by lane 0 of the row's warp inside the accumulation loop --
        //  200 warp-wide calls per instance instead of 100 / 32, 30 % of the kernel's instructions)
        for (int pr = tid; pr < (m + 1) / 2; pr += blockDim.x) {
            double z0, z1;
            normal_pair(key, inst, STREAM_EPS, (uint32_t)pr, z0, z1);
            eps[2 * pr] = fabs(z0);
            if (2 * pr + 1 < m) eps[2 * pr + 1] = fabs(z1);
        }
        double* Ak = A + k * (long long)m * n;
        for (int r0 = 0; r0 < m; r0 += kGenRows) {
            const int rows = (m - r0 < kGenRows) ? (m - r0) : kGenRows;
            __syncthreads();                 // x0s ready / previous tile consumed
            for (int t = tid; t < rows * half; t += blockDim.x) {
                const uint32_t pair = (uint32_t)(r0 * half + t);
                double z0, z1;
                normal_pair(key, inst, STREAM_A, pair, z0, z1);
                if (density < 1.0) {
                    uint32_t o[4];
                    philox4x32_10(pair, STREAM_MASK, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)key,
                                  (uint32_t)(key >> 32), o);
                    if (u53(o[0], o[1]) >= density) z0 = 0.0;
                    if (u53(o[2], o[3]) >= density) z1 = 0.0;
                }
                *reinterpret_cast<double2*>(Ak + (size_t)r0 * n + 2 * (size_t)t) = make_double2(z0, z1);
                *reinterpret_cast<double2*>(tile + 2 * (size_t)t) = make_double2(z0, z1);
            }
            __syncthreads();
            for (int ri = warp; ri < rows; ri += nw) {
                const int i = r0 + ri;
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(tile[ri * n + j], x0s[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) b[k * m + i] = acc + eps[i];
            }
        }
        if (x0out)
            for (int j = tid; j < n; j += blockDim.x) x0out[k * n + j] = x0s[j];
        __syncthreads();
    }
}

cudaError_t launch_generate(uint64_t key, long long first, long long B, int m, int n, double density, double* A,
                            double* b, double* c, double* x0, int sm_count, cudaStream_t st, int* launches) {
    // even n, 16-byte aligned A: the fused one-pass kernel (A written once and never read back)
    const size_t fsm = ((size_t)((n + 1) & ~1) + (size_t)kGenRows * n + (size_t)((m + 1) & ~1)) * sizeof(double);
    if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && fsm <= 200 * 1024) {
        cudaError_t e0 = cudaFuncSetAttribute(generate_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
        if (e0 != cudaSuccess) return e0;
        int per_sm = 0;
        e0 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generate_fused_kernel, 256, fsm);
        if (e0 != cudaSuccess) return e0;
        long long gf = (long long)sm_count * (per_sm > 0 ? per_sm : 1);
        if (gf > B) gf = B;
        if (gf < 1) gf = 1;
        generate_fused_kernel<<<(int)gf, 256, fsm, st>>>(key, first, B, m, n, density, A, b, c, x0);
        *launches += 1;
        return cudaGetLastError();
    }
    const long long pairs = B * (((long long)m * n + 1) / 2);
    long long blocks = (pairs + 255) / 256;
    const long long cap = (long long)sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    generate_A_kernel<<<(int)blocks, 256, 0, st>>>(key, first, B, m, n, density, A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    long long g2 = B < (long long)sm_count * 8 ? B : (long long)sm_count * 8;
    if (g2 < 1) g2 = 1;
    generate_bc_kernel<<<(int)g2, 256, (size_t)(n + 1) * sizeof(double), st>>>(key, first, B, m, n, A, b, c, x0);
    *launches += 2;
    return cudaGetLastError();
}

}  // namespace ddb
