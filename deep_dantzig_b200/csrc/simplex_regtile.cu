// Register-tiled simplex (plan 0) -- placeholder until the kernel lands; the plan selector never picks it.
#include "common.cuh"
namespace ddb {
bool regtile_supported(int, int) { return false; }
size_t regtile_scratch_bytes(int, int, int) { return 0; }
cudaError_t launch_simplex_regtile(const SolveArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace ddb
