// sell_modes_res.cu -- residual and fused residual-norm modes of the SELL-32 streaming kernels (sell_core.cuh).
#include "sell_core.cuh"

namespace mgb {

int sell_residual(const mg_sell *A, const double *x, const double *b, double *r, int64_t row0, int64_t row1,
                  const SellFuse *fuse, cudaStream_t st) {
    return launch_sell<RESID>(A, x, b, nullptr, r, 0.0, nullptr, row0, row1, st, "sell_residual", nullptr, fuse);
}
int sell_residual_partials(const mg_sell *A, const double *x, const double *b, double *partials, int64_t row0,
                           int64_t row1, int *nblocks, const SellFuse *fuse, cudaStream_t st) {
    return launch_sell<RESNORM>(A, x, b, nullptr, nullptr, 0.0, partials, row0, row1, st, "sell_residual_partials", nblocks, fuse);
}
}  // namespace mgb
