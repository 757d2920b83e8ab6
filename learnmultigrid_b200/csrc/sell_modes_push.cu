// sell_modes_push.cu -- colour sweeps of a partitioned level that push their own boundary values (producer-driven
// exchange, sell_gs_push_kernel of sell_core.cuh), plain and with a fused residual / norm.
#include "sell_core.cuh"

namespace mgb {

int sell_gs_rows_push(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *carry,
                      const SellPush *push, int tail, double *r_out, double *partials, int *nblocks, cudaStream_t st) {
    if (tail == TAIL_RESIDUAL) return launch_sell_push<GS_RES>(A, x, b, row0, row1, carry, push, r_out, nullptr, nullptr, st);
    if (tail == TAIL_NORM) return launch_sell_push<GS_NORM>(A, x, b, row0, row1, carry, push, nullptr, partials, nblocks, st);
    return launch_sell_push<GS>(A, x, b, row0, row1, carry, push, nullptr, nullptr, nullptr, st);
}

}  // namespace mgb
