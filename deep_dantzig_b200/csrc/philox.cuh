// Philox4x32-10 counter-based stream + Box-Muller of the instance generator (generate.cu) and of the in-solver generator
// (simplex_rowreg.cu, simplex_generic.cu: fused generate -> solve -> label without an HBM round trip of A).
//
// Instance i is a pure function of (key, i): key = (key_lo, key_hi), counter = (pair index, stream id, i_lo, i_hi).
// One Philox block -> two 53-bit uniforms -> one Box-Muller pair -> normals for elements 2*pair and 2*pair+1 of that
// stream.  Pinned on the integer level by Random123's known answers (oracle/philox.py, tests/test_philox.py).
#pragma once
#include "common.cuh"


namespace ddb {

enum : uint32_t { STREAM_A = 0, STREAM_X0 = 1, STREAM_EPS = 2, STREAM_C = 3, STREAM_MASK = 4 };

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// (0,1] uniform from two 32-bit words: 27 + 26 = 53 bits, plus one so that log() is finite.
__device__ __forceinline__ double u53_open0(uint32_t hi, uint32_t lo) {
    const unsigned long long v = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return (double)(v + 1ull) * 0x1.0p-53;
}
// [0,1) uniform.
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const unsigned long long v = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return (double)v * 0x1.0p-53;
}

__device__ __forceinline__ void normal_pair(uint64_t key, uint64_t inst, uint32_t stream, uint32_t pair, double& z0,
                                            double& z1) {
    uint32_t o[4];
    philox4x32_10(pair, stream, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)key, (uint32_t)(key >> 32), o);
    const double u1 = u53_open0(o[0], o[1]);
    const double u2 = u53(o[2], o[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}


// ---------------------------------------------------------------------------------------------------------
// One instance drawn by one CTA (any block size, even n): A, b, c written to global memory (the caller's output arrays
// or a per-CTA slab that stays in L2), 32 rows at a time through a shared-memory tile.  Same counters and the same
// summation order as generate_fused_kernel / generate_bc_kernel (lane L: columns L, L + 32, ...; xor butterfly), so
// every path produces the same bits.  Also leaves what the crash ranking needs -- dotc[i] = a_i . c and nn[i] = |a_i|^2,
// accumulated exactly as stage 0 of the solver kernels does from a materialised A -- so the solver does not read A
// again for it.  Shared memory: tile[kGenTileRows * n], x0s[n], cs[n] doubles.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGenTileRows = 32;
static __device__ __noinline__ void generate_instance_cta(uint64_t key, uint64_t inst, int m, int n, double density, double* Aw,
                                                   double* bw, double* cw, double* x0w, double* tile, double* x0s, double* cs,
                                                   double* dotc, double* nn) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int half = n / 2;
    for (int pr = tid; pr < half; pr += nt) {
        double z0, z1;
        normal_pair(key, inst, STREAM_X0, (uint32_t)pr, z0, z1);
        x0s[2 * pr] = z0;
        x0s[2 * pr + 1] = z1;
        normal_pair(key, inst, STREAM_C, (uint32_t)pr, z0, z1);
        z0 = fabs(z0); z1 = fabs(z1);
        cs[2 * pr] = z0;
        cs[2 * pr + 1] = z1;
        *reinterpret_cast<double2*>(cw + 2 * pr) = make_double2(z0, z1);
    }
    for (int r0 = 0; r0 < m; r0 += kGenTileRows) {
        const int rows = (m - r0 < kGenTileRows) ? (m - r0) : kGenTileRows;
        __syncthreads();                 // x0s / cs ready, previous tile consumed
        for (int t = tid; t < rows * half; t += nt) {
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
            *reinterpret_cast<double2*>(Aw + (size_t)r0 * n + 2 * (size_t)t) = make_double2(z0, z1);
            *reinterpret_cast<double2*>(tile + 2 * (size_t)t) = make_double2(z0, z1);
        }
        __syncthreads();
        for (int ri = warp; ri < rows; ri += nw) {
            const int i = r0 + ri;
            double acc = 0.0, dc = 0.0, q = 0.0;
            for (int j = lane; j < n; j += 32) {
                const double v = tile[ri * n + j];
                acc = fma(v, x0s[j], acc);
                dc = fma(v, cs[j], dc);
                q = fma(v, v, q);
            }
            acc = warp_sum(acc);
            dc = warp_sum(dc);
            q = warp_sum(q);
            if (lane == 0) {
                double z0, z1;
                normal_pair(key, inst, STREAM_EPS, (uint32_t)(i >> 1), z0, z1);
                bw[i] = acc + fabs((i & 1) ? z1 : z0);
                dotc[i] = dc;
                nn[i] = q;
            }
        }
    }
    if (x0w)
        for (int j = tid; j < n; j += nt) x0w[j] = x0s[j];
    __syncthreads();
}

// Same instance, same bits, without the tile (rare path: the generic kernel re-solving an instance the row-per-thread
// kernel flagged in fused mode): A goes to global memory first and b is accumulated from it.  Shared memory: x0s[n].
static __device__ __noinline__ void generate_instance_cta_notile(uint64_t key, uint64_t inst, int m, int n, double density, double* Aw,
                                                          double* bw, double* cw, double* x0s) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int half = n / 2;
    for (int pr = tid; pr < half; pr += nt) {
        double z0, z1;
        normal_pair(key, inst, STREAM_X0, (uint32_t)pr, z0, z1);
        x0s[2 * pr] = z0;
        x0s[2 * pr + 1] = z1;
        normal_pair(key, inst, STREAM_C, (uint32_t)pr, z0, z1);
        cw[2 * pr] = fabs(z0);
        cw[2 * pr + 1] = fabs(z1);
    }
    for (int t = tid; t < m * half; t += nt) {
        double z0, z1;
        normal_pair(key, inst, STREAM_A, (uint32_t)t, z0, z1);
        if (density < 1.0) {
            uint32_t o[4];
            philox4x32_10((uint32_t)t, STREAM_MASK, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)key, (uint32_t)(key >> 32), o);
            if (u53(o[0], o[1]) >= density) z0 = 0.0;
            if (u53(o[2], o[3]) >= density) z1 = 0.0;
        }
        Aw[2 * (size_t)t] = z0;
        Aw[2 * (size_t)t + 1] = z1;
    }
    __syncthreads();
    for (int i = warp; i < m; i += nw) {
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) acc = fma(Aw[(size_t)i * n + j], x0s[j], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            double z0, z1;
            normal_pair(key, inst, STREAM_EPS, (uint32_t)(i >> 1), z0, z1);
            bw[i] = acc + fabs((i & 1) ? z1 : z0);
        }
    }
    __syncthreads();
}

}  // namespace ddb
