// bcr.cu -- banded coarsest-level direct solver by block cyclic reduction (placeholder until built).
#include "common.cuh"
namespace mgb {
int bcr_solve(const void *, const double *, double *, cudaStream_t) {
    return set_error(MG_ERR_UNSUPPORTED, "bcr_solve", "block cyclic reduction solver not built yet");
}
}  // namespace mgb
