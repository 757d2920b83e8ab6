// sell_modes_vec.cu -- Jacobi, prolongation-plus-correction and SpMV-with-dot modes of the SELL-32 streaming kernels (sell_core.cuh).
#include "sell_core.cuh"

namespace mgb {

int sell_jacobi(const mg_sell *A, const double *dinv, const double *x, const double *b, double *xo,
                double omega, cudaStream_t st) {
    return launch_sell<JACOBI>(A, x, b, dinv, xo, omega, nullptr, 0, A->nrows, st, "sell_jacobi");
}
int sell_spmv_dot(const mg_sell *A, const double *x, const double *w, double *y, double *partials, int *nblocks,
                  const SellFuse *fuse, cudaStream_t st) {
    return launch_sell<SPMV_DOT>(A, x, nullptr, w, y, 0.0, partials, 0, A->nrows, st, "sell_spmv_dot", nblocks, fuse);
}
int sell_prolong(const mg_sell *Q, const double *e, const double *u, double *uo, int64_t row0, int64_t row1,
                 const SellFuse *fuse, cudaStream_t st) {
    return launch_sell<PROLONG>(Q, e, nullptr, u, uo, 0.0, nullptr, row0, row1, st, "sell_prolong", nullptr, fuse);
}

}  // namespace mgb
