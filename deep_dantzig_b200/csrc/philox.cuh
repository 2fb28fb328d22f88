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

// Branch-free fp64 elementary functions on the generator's restricted domains.  The CUDA library versions are accurate but
// carry data-dependent slow-path branches, which keep the compiler from interleaving independent Box-Muller chains -- and
// a solver CTA that draws its own instance has only four warps to hide ~100 dependent fp64 operations per pair behind.
// Accuracy (checked against extended precision on 2M samples each, tests/test_philox.py restates them): <= 1.5e-16
// relative for the logarithm, <= 2e-16 absolute for sine / cosine.

// ln(x) for x in (0, 1], x a normal double (fdlibm's e_log.c kernel: x = 2^k m, m in [sqrt(1/2), sqrt(2)), s = f / (2 + f))
__device__ __forceinline__ double gen_log01(double x) {
    const long long bits = __double_as_longlong(x);
    int k = (int)(bits >> 52) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
    const bool big = m > 1.4142135623730951;
    m = big ? m * 0.5 : m;
    k += big ? 1 : 0;
    const double f = m - 1.0, d = 2.0 + f;
    double r;                                         // 1 / d: hardware seed + two Newton steps, then s = f / d with a residual fix
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    double s = f * r;
    s = fma(fma(-d, s, f), r, s);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)k;
    return dk * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f);
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1): quadrant reduction (exact), fdlibm's k_sin.c / k_cos.c kernels on [-pi/4, pi/4]
__device__ __forceinline__ void gen_sincos_2pi(double u, double& sn, double& cs) {
    const double t = 2.0 * u;
    const double kq = rint(2.0 * t);                  // 0 .. 4
    const double r = fma(-0.5, kq, t);                // exact, in [-1/4, 1/4]
    const double x = fma(r, 1.2246467991473532e-16, r * 3.141592653589793116);
    const double z = x * x;
    const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08),
                                                 2.75573137070700676789e-06), -1.98412698298579493134e-04),
                                 8.33333333332248946124e-03), -1.66666666666666324348e-01);
    const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09),
                                                 -2.75573143513906633035e-07), 2.48015872894767294178e-05),
                                 -1.38888888888741095749e-03), 4.16666666666666019037e-02);
    const double s0 = fma(x * z, ps, x);
    const double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int q = (int)kq & 3;
    sn = (q == 0) ? s0 : ((q == 1) ? c0 : ((q == 2) ? -s0 : -c0));
    cs = (q == 0) ? c0 : ((q == 1) ? -s0 : ((q == 2) ? -c0 : s0));
}

// sqrt(a) for finite a >= 0: hardware reciprocal-square-root seed, two coupled Newton steps, residual correction
__device__ __forceinline__ double gen_sqrt(double a) {
    a = fmax(a, 1e-300);                              // a == 0 (u1 == 1, probability 2^-53) would make the seed infinite
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, a), h, g);
}

__device__ __forceinline__ void normal_pair(uint64_t key, uint64_t inst, uint32_t stream, uint32_t pair, double& z0,
                                            double& z1) {
    uint32_t o[4];
    philox4x32_10(pair, stream, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)key, (uint32_t)(key >> 32), o);
    const double u1 = u53_open0(o[0], o[1]);
    const double u2 = u53(o[2], o[3]);
    const double rad = gen_sqrt(-2.0 * gen_log01(u1));
    double sn, cs;
    gen_sincos_2pi(u2, sn, cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

// ---------------------------------------------------------------------------------------------------------
// One instance drawn by one CTA (any block size, even n): A, b, c written to global memory (the caller's output arrays
// or a per-CTA slab that stays in L2), 32 rows at a time through a shared-memory tile.  Same counters and the same
// summation order as generate_fused_kernel / generate_bc_kernel (lane L: columns L, L + 32, ...; xor butterfly), so
// every path produces the same bits.  Also leaves what the crash ranking needs -- dotc[i] = a_i . c and nn[i] = |a_i|^2,
// accumulated exactly as stage 0 of the solver kernels does from a materialised A -- so the solver does not read A
// again for it.  Shared memory: gen_smem_doubles(m, n) doubles at `tile`.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGenTileRows = 32;
#ifndef DDB_GEN_CHAINS
#define DDB_GEN_CHAINS 1   // measured on B200 at (200,100), in-solver generator: 1 chain 403 k LP/s, 2 chains 371 k, 4 chains 342 k
#endif
#ifdef DDB_GEN_INLINE
#define DDB_GEN_FN __forceinline__
#else
#define DDB_GEN_FN __noinline__
#endif
constexpr int kGenChains = DDB_GEN_CHAINS;
// Per-CTA instance slab of the fused mode: A[m n], b[m], c[n], each piece padded to an even count so that every piece of
// every CTA's slab is 16-byte aligned (the generator stores pairs).
__host__ __device__ constexpr size_t slab_doubles(int m, int n) {
    return (((size_t)m * n + 1) & ~(size_t)1) + (size_t)((m + 1) & ~1) + (size_t)((n + 1) & ~1);
}
__host__ __device__ constexpr size_t slab_b_offset(int m, int n) { return ((size_t)m * n + 1) & ~(size_t)1; }
__host__ __device__ constexpr size_t slab_c_offset(int m, int n) { return slab_b_offset(m, n) + (size_t)((m + 1) & ~1); }
// doubles of shared memory generate_instance_cta needs
__host__ __device__ constexpr size_t gen_smem_doubles(int m, int n) {
    return (size_t)kGenTileRows * n + 2 * (size_t)((n + 1) & ~1) + (size_t)((m + 1) & ~1);
}
// TAG: one copy per calling kernel, so that each copy is compiled under its caller's register budget (a shared copy would
// have to live with the tightest launch bound of all the solver variants).
template <int TAG, int CHAINS = kGenChains>
static __device__ DDB_GEN_FN void generate_instance_cta(uint64_t key, uint64_t inst, int m, int n, double density, double* Aw,
                                                          double* bw, double* cw, double* x0w, double* tile, double* dotc,
                                                          double* nn) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int half = n / 2;
    double* x0s = tile + (size_t)kGenTileRows * n;
    double* cs = x0s + ((n + 1) & ~1);
    double* eps = cs + ((n + 1) & ~1);      // |eps_i|, drawn up front by all threads (one Box-Muller pair per two rows)
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
    for (int pr = tid; pr < (m + 1) / 2; pr += nt) {
        double z0, z1;
        normal_pair(key, inst, STREAM_EPS, (uint32_t)pr, z0, z1);
        eps[2 * pr] = fabs(z0);
        if (2 * pr + 1 < m) eps[2 * pr + 1] = fabs(z1);
    }
    for (int r0 = 0; r0 < m; r0 += kGenTileRows) {
        const int rows = (m - r0 < kGenTileRows) ? (m - r0) : kGenTileRows;
        const int npair = rows * half;
        __syncthreads();                 // x0s / cs / eps ready, previous tile consumed
        // kGenChains independent Philox + Box-Muller chains per thread and trip.  More chains were measured SLOWER inside the
        // solver (the solver kernel runs at ~11 warp-cycles per issued instruction whatever the instruction, so time follows
        // the instruction count, and the unrolled body only adds rounding-up waste and instruction-cache misses)
        for (int t0 = tid; t0 < npair; t0 += CHAINS * nt) {
            double z0[CHAINS], z1[CHAINS];
#pragma unroll
            for (int u = 0; u < CHAINS; ++u) {
                const int t = t0 + u * nt;
                const uint32_t pair = (uint32_t)(r0 * half + (t < npair ? t : 0));
                normal_pair(key, inst, STREAM_A, pair, z0[u], z1[u]);
                if (density < 1.0) {
                    uint32_t o[4];
                    philox4x32_10(pair, STREAM_MASK, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)key,
                                  (uint32_t)(key >> 32), o);
                    if (u53(o[0], o[1]) >= density) z0[u] = 0.0;
                    if (u53(o[2], o[3]) >= density) z1[u] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < CHAINS; ++u) {
                const int t = t0 + u * nt;
                if (t < npair) {
                    *reinterpret_cast<double2*>(Aw + (size_t)r0 * n + 2 * (size_t)t) = make_double2(z0[u], z1[u]);
                    *reinterpret_cast<double2*>(tile + 2 * (size_t)t) = make_double2(z0[u], z1[u]);
                }
            }
        }
        __syncthreads();
        // four rows per warp and trip: twelve independent butterfly reductions in flight instead of three
        for (int rb = 4 * warp; rb < rows; rb += 4 * nw) {
            double acc[4], dc[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[u] = 0.0; dc[u] = 0.0; q[u] = 0.0;
                const int ri = (rb + u < rows) ? rb + u : rb;
                for (int j = lane; j < n; j += 32) {
                    const double v = tile[ri * n + j];
                    acc[u] = fma(v, x0s[j], acc[u]);
                    dc[u] = fma(v, cs[j], dc[u]);
                    q[u] = fma(v, v, q[u]);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], off);
                    dc[u] += __shfl_xor_sync(0xffffffffu, dc[u], off);
                    q[u] += __shfl_xor_sync(0xffffffffu, q[u], off);
                }
            }
            if (lane < 4 && rb + lane < rows) {
                const int i = r0 + rb + lane;
                const double a_ = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
                const double d_ = lane == 0 ? dc[0] : (lane == 1 ? dc[1] : (lane == 2 ? dc[2] : dc[3]));
                const double q_ = lane == 0 ? q[0] : (lane == 1 ? q[1] : (lane == 2 ? q[2] : q[3]));
                bw[i] = a_ + eps[i];
                dotc[i] = d_;
                nn[i] = q_;
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
