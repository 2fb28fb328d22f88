// Evaluation metrics of the classifier on the device: replaces the host round-trip of the reference's per-epoch
// evaluation -- sklearn roc_curve for the recall-1 threshold (reference src/ml/train.py:118-150: the first ROC threshold
// whose TPR is 1.0, i.e. the smallest predicted probability of a positive) and the confusion-matrix loop of
// performance() (train.py:174-246) / get_accuracy (src/ml/test.py:10-54) -- with one streaming pass over
// probs[N,2] + labels[N] (9 bytes per constraint node, HBM-bound).
//
// out[8] (fp64, overwritten): tp, fp, tn, fn at `thresh` (predicted positive <=> probs[:,1] >= thresh, train.py:203),
//   min over positives of probs[:,1] (+inf when there is no positive), weighted NLL sum (weights w0 / w1 by label, the
//   criterion of benchmark.py:70-75 evaluated as -w_y log(probs[:,y])), number of positives, number of negatives.
#include "common.cuh"
#include <math_constants.h>

namespace ddb {
namespace {

__device__ __forceinline__ double wsum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__global__ void __launch_bounds__(256) s2v_metrics_kernel(long long N, const float* __restrict__ logp,
                                                          const float* __restrict__ probs,
                                                          const uint8_t* __restrict__ labels, float thresh, float w0, float w1,
                                                          double* out, unsigned int* minbits) {
    __shared__ double red[8][6];
    __shared__ unsigned int redmin[8];
    double tp = 0, fp = 0, tn = 0, fn = 0, loss = 0, npos = 0;
    unsigned int mn = 0x7f800000u;   // +inf; probabilities are >= 0 so their bit patterns order like unsigned ints
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float2 pr = __ldg(reinterpret_cast<const float2*>(probs) + i);
        const int y = labels[i] ? 1 : 0;
        const bool pred = pr.y >= thresh;
        tp += (y && pred); fp += (!y && pred); tn += (!y && !pred); fn += (y && !pred);
        npos += y;
        if (y) mn = min(mn, __float_as_uint(pr.y));
        if (logp) {
            const float2 lq = __ldg(reinterpret_cast<const float2*>(logp) + i);
            loss -= y ? (double)(w1 * lq.y) : (double)(w0 * lq.x);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tp = wsum_d(tp); fp = wsum_d(fp); tn = wsum_d(tn); fn = wsum_d(fn); loss = wsum_d(loss); npos = wsum_d(npos);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if (lane == 0) {
        red[warp][0] = tp; red[warp][1] = fp; red[warp][2] = tn; red[warp][3] = fn; red[warp][4] = loss; red[warp][5] = npos;
        redmin[warp] = mn;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        const int slot = threadIdx.x < 4 ? threadIdx.x : (threadIdx.x == 4 ? 5 : 6);
        if (v != 0.0) atomicAdd(out + slot, v);
    }
    if (threadIdx.x == 6) {
        unsigned int v = 0x7f800000u;
        for (int w = 0; w < 8; ++w) v = min(v, redmin[w]);
        atomicMin(minbits, v);
    }
}

__global__ void s2v_metrics_finish_kernel(long long N, double* out, const unsigned int* minbits) {
    out[4] = (*minbits == 0x7f7f7f7fu) ? CUDART_INF : (double)__uint_as_float(*minbits);
    out[7] = (double)N - out[6];
}

}  // namespace

cudaError_t launch_s2v_metrics(long long N, const float* logp, const float* probs, const uint8_t* labels, float thresh,
                               float w0, float w1, double* out, unsigned int* minbits, int sm_count, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, 8 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(minbits, 0x7f, sizeof(unsigned int), st);   // 0x7f7f7f7f (3.4e38) stands for "no positive seen"
    if (e != cudaSuccess) return e;
    long long blocks = (N + 255) / 256;
    const long long cap = (long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    s2v_metrics_kernel<<<(int)blocks, 256, 0, st>>>(N, logp, probs, labels, thresh, w0, w1, out, minbits);
    s2v_metrics_finish_kernel<<<1, 1, 0, st>>>(N, out, minbits);
    return cudaGetLastError();
}

}  // namespace ddb
