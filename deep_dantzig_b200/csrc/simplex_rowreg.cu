// Row-per-thread simplex (plan 0), caller-supplied instances: instantiations and host entry points.
// The kernel itself lives in rowreg_kernel.cuh; simplex_rowreg_gen.cu holds the instantiations with the in-solver
// instance generator (fused generate -> solve -> label).
#include "rowreg_kernel.cuh"

namespace ddb {

bool rowreg_supported(int m, int n) { return m >= n && pick_row_variant<false>(m, n) != nullptr; }

int rowreg_grid(int m, int n, int sm_count) {
    const RowVariant* v = pick_row_variant<false>(m, n);
    if (!v) return 0;
    const int per_sm = v->ctas_per_sm(m, n);      // the same occupancy query the launch sizes its grid with
    return sm_count * (per_sm > 0 ? per_sm : 1);
}

// per-CTA scratch: the register part of one saved tableau row per thread (stage 4 parks the rows there before T becomes a
// streaming buffer) and, for hybrid rows, the crash inverse D (n rows of pitch row_pitch(NC))
size_t rowreg_rows_scratch_bytes(int m, int n, int grid) {
    const RowVariant* v = pick_row_variant<false>(m, n);
    if (!v) return 0;
    const int prd = (v->NC - v->TS + 1) & ~1;
    return (size_t)grid * (size_t)(v->W * 32) * prd * sizeof(double);
}
size_t rowreg_d_scratch_bytes(int m, int n, int grid) {
    const RowVariant* v = pick_row_variant<false>(m, n);
    if (!v || v->TS == 0) return 0;
    return (size_t)grid * (size_t)n * row_pitch(v->NC) * sizeof(double);
}

cudaError_t launch_simplex_rowreg(const SolveArgs& a, int sm_count, cudaStream_t st) {
    const RowVariant* v = pick_row_variant<false>(a.m, a.n);
    if (!v) return cudaErrorInvalidValue;
    return v->launch(a, sm_count, st);
}

// tools/row_timing.cu and older callers: both scratches in one block (saved rows first)
size_t rowreg_scratch_bytes(int m, int n, int sm_count) {
    const int grid = rowreg_grid(m, n, sm_count);
    return rowreg_rows_scratch_bytes(m, n, grid) + rowreg_d_scratch_bytes(m, n, grid);
}

}  // namespace ddb
