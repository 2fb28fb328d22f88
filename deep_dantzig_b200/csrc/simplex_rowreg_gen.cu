// Row-per-thread simplex (plan 0) with the in-solver instance generator: the CTA that solves instance i draws it first
// (Philox, counter = global instance index), so the fused generate -> solve -> label entry point never moves A through HBM.
#include "rowreg_kernel.cuh"

namespace ddb {

bool rowreg_gen_supported(int m, int n) { return m >= n && (n & 1) == 0 && pick_row_variant<true>(m, n) != nullptr; }

int rowreg_gen_grid(int m, int n, int sm_count) {
    const RowVariant* v = pick_row_variant<true>(m, n);
    if (!v) return 0;
    const int per_sm = v->ctas_per_sm(m, n);
    return sm_count * (per_sm > 0 ? per_sm : 1);
}

cudaError_t launch_simplex_rowreg_gen(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const RowVariant* v = pick_row_variant<true>(a.m, a.n);
    if (!v) return cudaErrorInvalidValue;
    return v->launch(a, sm_count, st);
}

}  // namespace ddb
