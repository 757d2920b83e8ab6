// sell_modes_gs_tail.cu -- colour sweeps that also produce the residual / the squared residual norm of their rows
// (GS_RES, GS_NORM of sell_core.cuh).
#include "sell_core.cuh"

namespace mgb {

int sell_gs_rows_tail_launch(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *fuse,
                             int tail, double *r_out, double *partials, int *nblocks, cudaStream_t st) {
    if (tail == TAIL_RESIDUAL)
        return launch_sell<GS_RES>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows+residual", nullptr, fuse, r_out);
    return launch_sell<GS_NORM>(A, x, b, nullptr, x, 0.0, partials, row0, row1, st, "sell_gs_rows+norm", nblocks, fuse);
}

}  // namespace mgb
