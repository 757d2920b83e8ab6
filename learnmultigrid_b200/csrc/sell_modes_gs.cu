// sell_modes_gs.cu -- the Gauss-Seidel modes of the SELL-32 streaming kernels (sell_core.cuh): plain colour sweep,
// sweep + residual of the swept rows, sweep + squared residual norm of the swept rows; with or without an exchange
// site riding along.  (The tail modes are instantiated in sell_modes_gs_tail.cu -- a separate translation unit so that
// they compile in parallel.)
#include "sell_core.cuh"

namespace mgb {

int sell_gs_rows_tail_launch(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *fuse,
                             int tail, double *r_out, double *partials, int *nblocks, cudaStream_t st);

int sell_gs_rows(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *fuse,
                 int tail, double *r_out, double *partials, int *nblocks, cudaStream_t st) {
    if (tail != TAIL_NONE) return sell_gs_rows_tail_launch(A, x, b, row0, row1, fuse, tail, r_out, partials, nblocks, st);
    return launch_sell<GS>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows", nullptr, fuse);
}

}  // namespace mgb
