/*
 * oracle/relax_kernels.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, serial) of the native relaxation / mat-vec kernels the
 * reference's V-cycle reaches through third-party libraries.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (learnmultigrid_b200/) never does.
 *
 * Third-party algorithm restated (absent from /root/reference, version unpinned there):
 *   PyAMG  amg_core::gauss_seidel  (published algorithm, pyamg/amg_core/relaxation.h):
 *   call sites  learn_multigrid/solvers/Multigrid.py:88,121 (and :257,299), import :7.
 *   SciPy  sparsetools csr_matvec: call sites Multigrid.py:62,90,93,115.
 *
 * Arithmetic is fp64, indices int32 (SciPy's index type at these sizes); every product
 * and sum is a separate IEEE operation (compile with -ffp-contract=off) so that the
 * CUDA kernels, which use __dmul_rn/__dadd_rn in the same order, can be compared
 * bit-for-bit.
 */
#include <stdint.h>
#include <stddef.h>

/* PyAMG gauss_seidel, forward sweep (row_start=0,row_stop=n,row_step=1), `iterations` times.
 * x_i <- (b_i - sum_{j!=i} A_ij x_j) / A_ii, rows in index order, in place, rows with a zero
 * diagonal are skipped.  (Multigrid.py:88,121 call it with iterations=smooth_steps.) */
void oracle_gauss_seidel(const int32_t *Ap, const int32_t *Aj, const double *Ax,
                         double *x, const double *b, int64_t n, int iterations)
{
    for (int it = 0; it < iterations; ++it) {
        for (int64_t i = 0; i < n; ++i) {
            double rsum = 0.0, diag = 0.0;
            for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) {
                int32_t j = Aj[jj];
                if (j == i) diag = Ax[jj];
                else        rsum += Ax[jj] * x[j];
            }
            if (diag != 0.0) x[i] = (b[i] - rsum) / diag;
        }
    }
}

/* Same update restricted to an explicit row list, Jacobi-style inside the list is NOT used:
 * rows are updated in place in list order.  With rows = the rows of one colour of a valid
 * colouring the order inside the list does not matter (no two rows are coupled).
 * This is the multicolour Gauss-Seidel oracle named by BASELINE.json north_star. */
void oracle_gauss_seidel_rows(const int32_t *Ap, const int32_t *Aj, const double *Ax,
                              double *x, const double *b,
                              const int32_t *rows, int64_t nrows)
{
    for (int64_t r = 0; r < nrows; ++r) {
        int64_t i = rows[r];
        double rsum = 0.0, diag = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) {
            int32_t j = Aj[jj];
            if (j == i) diag = Ax[jj];
            else        rsum += Ax[jj] * x[j];
        }
        if (diag != 0.0) x[i] = (b[i] - rsum) / diag;
    }
}

/* SciPy csr_matvec: y_i = sum_j A_ij x_j accumulated in storage order starting from 0
 * (Multigrid.py:62,90 via A.dot(x)); then r = b - y. */
void oracle_residual(const int32_t *Ap, const int32_t *Aj, const double *Ax,
                     const double *x, const double *b, double *r, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) sum += Ax[jj] * x[Aj[jj]];
        r[i] = b[i] - sum;
    }
}

void oracle_spmv(const int32_t *Ap, const int32_t *Aj, const double *Ax,
                 const double *x, double *y, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) sum += Ax[jj] * x[Aj[jj]];
        y[i] = sum;
    }
}

/* Damped Jacobi sweep, out of place: xo = x + omega * (dinv * (b - A x)).
 * omega = 1 is learn_multigrid/solvers/Jacobi.py:35 (solution += inv_d * residual). */
void oracle_jacobi(const int32_t *Ap, const int32_t *Aj, const double *Ax,
                   const double *dinv, const double *x, const double *b, double *xo,
                   double omega, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) sum += Ax[jj] * x[Aj[jj]];
        double r = b[i] - sum;
        xo[i] = x[i] + omega * (dinv[i] * r);
    }
}

/* y += A x in storage order: prolongation + correction u + Q e (Multigrid.py:115).
 * SciPy computes t = Q e (from 0) and then u + t; we do exactly that. */
void oracle_prolong_correct(const int32_t *Qp, const int32_t *Qj, const double *Qx,
                            const double *e, double *u, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int32_t jj = Qp[i]; jj < Qp[i + 1]; ++jj) sum += Qx[jj] * e[Qj[jj]];
        u[i] = u[i] + sum;
    }
}
