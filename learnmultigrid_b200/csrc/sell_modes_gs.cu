// sell_modes_gs.cu -- the Gauss-Seidel modes of the SELL-32 streaming kernels (sell_core.cuh): plain colour sweep,
// sweep + residual of the swept rows, sweep + squared residual norm of the swept rows; with or without an exchange
// site riding along, with or without pushing the colour's boundary values.
#include "sell_core.cuh"

namespace mgb {

int sell_gs_rows(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *fuse,
                 int tail, double *r_out, double *partials, int *nblocks, cudaStream_t st) {
    if (tail == TAIL_RESIDUAL)
        return launch_sell<GS_RES>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows+residual", nullptr, fuse, r_out);
    if (tail == TAIL_NORM)
        return launch_sell<GS_NORM>(A, x, b, nullptr, x, 0.0, partials, row0, row1, st, "sell_gs_rows+norm", nblocks, fuse);
    return launch_sell<GS>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows", nullptr, fuse);
}

int sell_gs_rows_push(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *carry,
                      const SellPush *push, int tail, double *r_out, double *partials, int *nblocks, cudaStream_t st) {
    if (tail == TAIL_RESIDUAL) return launch_sell_push<GS_RES>(A, x, b, row0, row1, carry, push, r_out, nullptr, nullptr, st);
    if (tail == TAIL_NORM) return launch_sell_push<GS_NORM>(A, x, b, row0, row1, carry, push, nullptr, partials, nblocks, st);
    return launch_sell_push<GS>(A, x, b, row0, row1, carry, push, nullptr, nullptr, nullptr, st);
}

}  // namespace mgb
