// sell_kernels.cu -- the streaming hot-path kernels on the SELL-32 format (see include/mgb200.h).
//
// One thread per row, one warp per slice.  Entry k of the 32 rows of a slice is contiguous
// (256 B of values + 128 B of columns per warp request), so the matrix streams fully coalesced; the x
// gathers go through L1/L2, which hold the few grid lines a structured stencil touches.
// Every kernel is HBM-bound: algorithmic bytes per row = 12*nnz_row + 4 (CSR yardstick of SURVEY 8d) plus
// the vector traffic listed at each launcher.
#include "exchange.cuh"

namespace mgb {

struct SellArgs {
    const int64_t *__restrict__ slice_ptr;
    const int32_t *__restrict__ cols;
    const double *__restrict__ vals;
    int64_t row_begin;   // first row this launch touches
    int64_t row_end;     // one past the last row
    int64_t first_row;   // row of thread 0 of block 0 (row_begin rounded down to a slice)
    int64_t nrows;       // rows of the matrix (rows of the last slice beyond it hold no data)
};

// accumulate CNT consecutive entries of a row: all cols/vals loads are issued first, then all x gathers, then the
// sums in storage order.  Straight-line code: CNT*12 bytes per thread in flight at ~4 registers per entry.
// `halo_wait` (fused launches only): this slice reads halo columns, so between the matrix loads (which do not depend on
// the halo and are already in flight) and the x gathers, wait until the launch's exchange CTAs have unpacked it.
template <int MODE, int CNT>
__device__ __forceinline__ void row_chunk(const int32_t *__restrict__ c, const double *__restrict__ v,
                                          const double *x, int64_t row, double &sum, double &diag,
                                          unsigned char halo_wait = 0, const ExArgs *fx = nullptr) {
    int32_t cc[CNT];
    double vv[CNT], xx[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
        cc[j] = ld_stream(c + j * kSlice);
        vv[j] = ld_stream(v + j * kSlice);
    }
    if (halo_wait) fused_wait_ready(*fx);
#pragma unroll
    for (int j = 0; j < CNT; ++j) xx[j] = x[cc[j]];
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
        if (MODE == GS) {
            if (cc[j] == row) { if (vv[j] != 0.0) diag = vv[j]; } else sum = mul_add_unfused(sum, vv[j], xx[j]);
        } else {
            sum = mul_add_unfused(sum, vv[j], xx[j]);
        }
    }
}

// LEN   : the matrix' longest slice when it is <= 8 (fast path: slices of exactly LEN entries run one straight-line
//         row_chunk<LEN>; shorter slices take a rolled loop), 0 for longer rows (chunks of 4 + rolled remainder).
// UNIFORM: every slice of the matrix has exactly LEN entries, so the slice offset is computed instead of loaded.
// The microbenchmark behind these choices is tools/sellbench.cu (profiles/r01_sellbench.log): occupancy x bytes in
// flight per thread decides; at 32 registers and 60 B per thread the fine-level sweep reaches the DRAM limit
// (6.8 TB/s algorithmic, ~7.1 TB/s of actual traffic), a rolled loop stays at 5.4 TB/s.
template <int MODE, int LEN, bool UNIFORM, bool FUSED>
__device__ __forceinline__ void
sell_body(const SellArgs &A, const double *x, const double *__restrict__ b, const double *aux, double *y, double omega,
          double *__restrict__ partials, int64_t bid, const ExArgs *fx, const unsigned char *__restrict__ mask) {
    const int64_t row = A.first_row + bid * kBlock + threadIdx.x;
    const bool active = row >= A.row_begin && row < A.row_end;
    double contrib = 0.0;
    if (row < A.row_end) {   // warp-uniform except in the last slice
        const int64_t slice = row >> 5;
        const int lane = (int)(row & 31);
        unsigned char hw = 0;      // fused launch: does this slice read halo columns? (load issued now, used in row_chunk)
        if (FUSED) hw = mask ? mask[slice] : 1;
        int64_t base;
        int len;
        if (UNIFORM) {
            base = slice * (int64_t)(kSlice * LEN);
            len = LEN;
        } else {
            base = A.slice_ptr[slice];
            len = (int)((A.slice_ptr[slice + 1] - base) >> 5);
        }
        const double *__restrict__ v = A.vals + base + lane;
        const int32_t *__restrict__ c = A.cols + base + lane;
        double sum = 0.0, diag = 0.0;
        if (LEN > 0 && (UNIFORM || len == LEN)) {
            row_chunk<MODE, (LEN > 0 ? LEN : 1)>(c, v, x, row, sum, diag, hw, fx);
        } else {
            int k = 0;
            if (LEN == 0)
                for (; k + 4 <= len; k += 4) {
                    row_chunk<MODE, 4>(c + k * kSlice, v + k * kSlice, x, row, sum, diag, hw, fx);
                    hw = 0;
                }
            for (; k < len; ++k) {
                row_chunk<MODE, 1>(c + k * kSlice, v + k * kSlice, x, row, sum, diag, hw, fx);
                hw = 0;
            }
        }
        if (active) {
            if (MODE == SPMV) {
                y[row] = sum;
            } else if (MODE == RESID) {
                y[row] = __dsub_rn(b[row], sum);
            } else if (MODE == RESNORM) {
                const double r = __dsub_rn(b[row], sum);
                contrib = r * r;
            } else if (MODE == JACOBI) {
                const double r = __dsub_rn(b[row], sum);
                y[row] = __dadd_rn(x[row], __dmul_rn(omega, __dmul_rn(aux[row], r)));
            } else if (MODE == GS) {
                if (diag != 0.0) y[row] = __ddiv_rn(__dsub_rn(b[row], sum), diag);
            } else if (MODE == PROLONG) {
                y[row] = __dadd_rn(aux[row], sum);   // aux = u (may alias y)
            }
        }
    }
    if (MODE == RESNORM) {
        const double s = block_sum<kBlock>(contrib);
        if (threadIdx.x == 0) partials[bid] = s;
    }
}

template <int MODE, int LEN, bool UNIFORM>
__global__ void __launch_bounds__(kBlock)
sell_kernel(SellArgs A, const double *x, const double *__restrict__ b, const double *aux,
            double *y, double omega, double *__restrict__ partials) {
    pdl_prologue();
    sell_body<MODE, LEN, UNIFORM, false>(A, x, b, aux, y, omega, partials, (int64_t)blockIdx.x, nullptr, nullptr);
}

// The same kernel carrying an exchange site (multi-GPU, csrc/comm.cu): the first npeers*ctas_per_peer CTAs push the
// boundary values the previous kernel produced, poll for the peers' packets and unpack them into the halo of x; the
// compute CTAs whose slice reads halo columns (mask) wait for that, all others start at once.  The exchange latency
// (NVLink flight + polling) is hidden behind the interior rows, and the site costs no launch of its own.
template <int MODE, int LEN, bool UNIFORM>
__global__ void __launch_bounds__(kBlock)
sell_kernel_fused(SellArgs A, const double *x, const double *__restrict__ b, const double *aux, double *y, double omega,
                  double *__restrict__ partials, const ExArgs fx, const unsigned char *__restrict__ mask) {
    pdl_prologue();
    const int nex = fx.npeers * fx.ctas_per_peer;
    if ((int)blockIdx.x < nex) {
        fused_exchange_cta(fx, (int)blockIdx.x);
        return;
    }
    sell_body<MODE, LEN, UNIFORM, true>(A, x, b, aux, y, omega, partials, (int64_t)blockIdx.x - nex, &fx, mask);
}

// ---- implied columns ------------------------------------------------------------------------------------------------
// On a structured stencil level nearly every slice is REGULAR: entry j of every one of its 32 rows has column
// row + off[j] with ONE offset table for the slice (mg_sell_slice_offsets; formats.sell_slice_offsets is the host twin).
// For those slices the kernel below computes the columns instead of streaming them: 4*LEN bytes of offsets per slice
// in place of 128*LEN bytes of column indices, i.e. 64 instead of 88 bytes per 5-point row.  The values, the gathers
// and the order of the additions are untouched, so the results are the same bits; slices that are not regular (a
// boundary node among the rows, the ragged tail) take the ordinary path inside the same kernel.  Uniform matrices
// with at most 8 entries per row only; opt-in (mg_set_implied_columns) until it has been measured.
constexpr int32_t kSliceIrregular = INT32_MIN;
static int g_implied_columns = 0;

__global__ void __launch_bounds__(kBlock)
sell_slice_offsets_kernel(int64_t nslices, int64_t nrows, int len, const int32_t *__restrict__ cols,
                          int32_t *__restrict__ off) {
    const int64_t w = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nslices) return;
    const int64_t row = w * kSlice + lane;
    bool regular = (w + 1) * kSlice <= nrows;
    for (int j = 0; j < len; ++j) {
        const int64_t rel = (int64_t)cols[(w * len + j) * kSlice + lane] - row;
        const int64_t rel0 = __shfl_sync(0xffffffffu, rel, 0);
        regular = __all_sync(0xffffffffu, rel == rel0) && regular;
        if (lane == 0) off[w * len + j] = (int32_t)rel0;
    }
    if (lane == 0 && !regular) off[w * len] = kSliceIrregular;
}

template <int MODE, int LEN, bool FUSED>
__device__ __forceinline__ void
sell_reg_body(const SellArgs &A, const int32_t *__restrict__ soff, const double *x, const double *__restrict__ b,
              const double *aux, double *y, double omega, double *__restrict__ partials, int64_t bid, const ExArgs *fx,
              const unsigned char *__restrict__ mask) {
    const int64_t row = A.first_row + bid * kBlock + threadIdx.x;
    const bool active = row >= A.row_begin && row < A.row_end;
    double contrib = 0.0;
    if (row < A.row_end) {
        const int64_t slice = row >> 5;
        const int lane = (int)(row & 31);
        unsigned char hw = 0;              // fused launch: does this slice read halo columns?
        if (FUSED) hw = mask ? mask[slice] : 1;
        const int64_t base = slice * (int64_t)(kSlice * LEN);
        const double *__restrict__ v = A.vals + base + lane;
        const int32_t *__restrict__ o = soff + slice * LEN;
        double sum = 0.0, diag = 0.0;
        const int32_t o0 = __ldg(o);
        if (o0 != kSliceIrregular) {          // warp-uniform
            int32_t oo[LEN];                   // the same for all 32 rows: column of entry j = row + oo[j]
            double vv[LEN], xx[LEN];
#pragma unroll
            for (int j = 0; j < LEN; ++j) vv[j] = ld_stream(v + j * kSlice);
            oo[0] = o0;
#pragma unroll
            for (int j = 1; j < LEN; ++j) oo[j] = __ldg(o + j);
            if (FUSED && hw) fused_wait_ready(*fx);
            const double *xr = x + row;
#pragma unroll
            for (int j = 0; j < LEN; ++j) xx[j] = xr[oo[j]];
#pragma unroll
            for (int j = 0; j < LEN; ++j) {
                if (MODE == GS) {              // the diagonal entry is the one with offset 0
                    if (oo[j] == 0) { if (vv[j] != 0.0) diag = vv[j]; } else sum = mul_add_unfused(sum, vv[j], xx[j]);
                } else {
                    sum = mul_add_unfused(sum, vv[j], xx[j]);
                }
            }
        } else {
            row_chunk<MODE, LEN>(A.cols + base + lane, v, x, row, sum, diag, hw, fx);
        }
        if (active) {      // the epilogues of sell_body
            if (MODE == SPMV) {
                y[row] = sum;
            } else if (MODE == RESID) {
                y[row] = __dsub_rn(b[row], sum);
            } else if (MODE == RESNORM) {
                const double r = __dsub_rn(b[row], sum);
                contrib = r * r;
            } else if (MODE == JACOBI) {
                const double r = __dsub_rn(b[row], sum);
                y[row] = __dadd_rn(x[row], __dmul_rn(omega, __dmul_rn(aux[row], r)));
            } else if (MODE == GS) {
                if (diag != 0.0) y[row] = __ddiv_rn(__dsub_rn(b[row], sum), diag);
            } else if (MODE == PROLONG) {
                y[row] = __dadd_rn(aux[row], sum);
            }
        }
    }
    if (MODE == RESNORM) {
        const double s = block_sum<kBlock>(contrib);
        if (threadIdx.x == 0) partials[bid] = s;
    }
}

template <int MODE, int LEN>
__global__ void __launch_bounds__(kBlock)
sell_kernel_reg(SellArgs A, const int32_t *__restrict__ soff, const double *x, const double *__restrict__ b,
                const double *aux, double *y, double omega, double *__restrict__ partials) {
    pdl_prologue();
    sell_reg_body<MODE, LEN, false>(A, soff, x, b, aux, y, omega, partials, (int64_t)blockIdx.x, nullptr, nullptr);
}

// the same carrying an exchange site as extra CTAs (see sell_kernel_fused)
template <int MODE, int LEN>
__global__ void __launch_bounds__(kBlock)
sell_kernel_reg_fused(SellArgs A, const int32_t *__restrict__ soff, const double *x, const double *__restrict__ b,
                      const double *aux, double *y, double omega, double *__restrict__ partials, const ExArgs fx,
                      const unsigned char *__restrict__ mask) {
    pdl_prologue();
    const int nex = fx.npeers * fx.ctas_per_peer;
    if ((int)blockIdx.x < nex) {
        fused_exchange_cta(fx, (int)blockIdx.x);
        return;
    }
    sell_reg_body<MODE, LEN, true>(A, soff, x, b, aux, y, omega, partials, (int64_t)blockIdx.x - nex, &fx, mask);
}

// Colour sweep of a partitioned level that PUSHES its own boundary values (producer-driven exchange, exchange.cuh):
// the compute CTAs run the rows next to the upper neighbour first (tail_first), every thread whose row is in the
// colour's send table stores its new value straight into the peer's staging slot, and the NVLink flight overlaps the
// interior rows of this very kernel.  CARRY: the launch also carries the previous site as extra CTAs, exactly like
// sell_kernel_fused.  Opt-in (mg_set_push_exchange); same values, same packets, same receiving code.
template <int LEN, bool UNIFORM, bool CARRY>
__global__ void __launch_bounds__(kBlock)
sell_gs_push_kernel(SellArgs A, double *x, const double *__restrict__ b, const ExArgs fx,
                    const unsigned char *__restrict__ mask, const SellPush push) {
    pdl_prologue();
    const int nex = CARRY ? fx.npeers * fx.ctas_per_peer : 0;
    if (CARRY && (int)blockIdx.x < nex) {
        fused_exchange_cta(fx, (int)blockIdx.x);
        return;
    }
    const int64_t nb = (int64_t)gridDim.x - nex;
    int64_t bid = (int64_t)blockIdx.x - nex;
    bid = bid < push.tail_first ? nb - 1 - bid : bid - push.tail_first;
    sell_body<GS, LEN, UNIFORM, CARRY>(A, x, b, nullptr, x, 0.0, nullptr, bid, &fx, mask);
    const int64_t row = A.first_row + bid * kBlock + threadIdx.x;
    if (row >= A.row_begin && row < A.row_end && !push.ex.dry && push.mask[row >> 5]) push_row_if_listed(push, row, x);
}

// ---- long rows: four warps per slice ---------------------------------------------------------------------------------
// With 19- / 37-point Galerkin stencils (quasi-L2 transfers) one thread walking a whole row is a chain of dependent
// (column -> x gather) round trips: ~15 us per launch however small, and half the DRAM rate on large levels.  Here
// the four warps of a slice take every fourth entry each, issue all their loads up front, and park the separately
// rounded products v*x in shared memory; one warp then adds them in STORAGE ORDER, so the row sums keep the oracle's
// bits (the additions are the only order-sensitive part, and they are a few hundred cycles of shared-memory reads).
constexpr int kWideU = 5;          // entries per warp and pass
constexpr int kWideMaxLen = 64;    // products kept in shared memory: 64 entries x 32 rows per slice

// WPS warps share one slice (kBlock/32/WPS slices per CTA); WPS*kWideU entries per pass: 4 warps cover the 19-point
// stencil in one pass, 8 warps the 37-point one.
template <int MODE, int WPS>
__global__ void __launch_bounds__(kBlock)
sell_wide_kernel(SellArgs A, int uniform_len, const double *x, const double *__restrict__ b, const double *aux,
                 double *y, double omega, double *__restrict__ partials) {
    pdl_prologue();
    constexpr int SPC = kBlock / 32 / WPS;
    __shared__ double prod[SPC][kWideMaxLen][kSlice];
    __shared__ unsigned char skip[SPC][kWideMaxLen][kSlice];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = warp / WPS, w = warp % WPS;
    const int64_t slice = (A.first_row >> 5) + (int64_t)blockIdx.x * SPC + sl;
    const int64_t row = slice * kSlice + lane;
    int len = 0;
    int64_t base = 0;
    if (slice * kSlice < A.row_end && row < A.nrows) {
        if (uniform_len > 0) {
            base = slice * (int64_t)kSlice * uniform_len;
            len = uniform_len;
        } else {
            base = A.slice_ptr[slice];
            len = (int)((A.slice_ptr[slice + 1] - base) >> 5);
        }
    }
    // the warp that will finish the rows fetches their vector entries now, under the shadow of the matrix loads
    const bool finisher = w == 0 && row >= A.row_begin && row < A.row_end;
    double bv = 0.0, av = 0.0;
    if (finisher) {
        if (MODE == RESID || MODE == RESNORM || MODE == JACOBI || MODE == GS) bv = b[row];
        if (MODE == JACOBI || MODE == PROLONG) av = aux[row];
    }
    const double *__restrict__ v = A.vals + base + lane;
    const int32_t *__restrict__ c = A.cols + base + lane;
    for (int k0 = w; k0 < len; k0 += WPS * kWideU) {
        int32_t cc[kWideU];
        double vv[kWideU], xx[kWideU];
#pragma unroll
        for (int j = 0; j < kWideU; ++j) {
            const int k = k0 + j * WPS;
            if (k < len) {
                cc[j] = ld_stream(c + (int64_t)k * kSlice);
                vv[j] = ld_stream(v + (int64_t)k * kSlice);
            }
        }
#pragma unroll
        for (int j = 0; j < kWideU; ++j)
            if (k0 + j * WPS < len) xx[j] = x[cc[j]];
#pragma unroll
        for (int j = 0; j < kWideU; ++j) {
            const int k = k0 + j * WPS;
            if (k < len) {
                const bool is_diag = (MODE == GS) && cc[j] == row;
                prod[sl][k][lane] = is_diag ? vv[j] : __dmul_rn(vv[j], xx[j]);
                if (MODE == GS) skip[sl][k][lane] = is_diag ? 1 : 0;
            }
        }
    }
    double xr = 0.0;
    if (MODE == JACOBI && finisher) xr = x[row];
    __syncthreads();
    double contrib = 0.0;
    if (finisher) {
        double sum = 0.0, diag = 0.0;
        for (int k = 0; k < len; ++k) {
            const double p = prod[sl][k][lane];
            if (MODE == GS && skip[sl][k][lane]) { if (p != 0.0) diag = p; }
            else sum = __dadd_rn(sum, p);
        }
        if (MODE == SPMV) {
            y[row] = sum;
        } else if (MODE == RESID) {
            y[row] = __dsub_rn(bv, sum);
        } else if (MODE == RESNORM) {
            const double r = __dsub_rn(bv, sum);
            contrib = r * r;
        } else if (MODE == JACOBI) {
            const double r = __dsub_rn(bv, sum);
            y[row] = __dadd_rn(xr, __dmul_rn(omega, __dmul_rn(av, r)));
        } else if (MODE == GS) {
            if (diag != 0.0) y[row] = __ddiv_rn(__dsub_rn(bv, sum), diag);
        } else if (MODE == PROLONG) {
            y[row] = __dadd_rn(av, sum);
        }
    }
    if (MODE == RESNORM) {
        const double s = block_sum<kBlock>(contrib);
        if (threadIdx.x == 0) partials[blockIdx.x] = s;
    }
}

// slices at least this long use sell_wide_kernel (0 = never)
static int64_t g_wide_min_len = 9;
// ... for launches of at most this many rows; larger launches have enough rows in flight for the thread-per-row kernel,
// which then streams at the DRAM limit (measured: profiles/r01_launches_c3_quasi_*.txt)
static int64_t g_wide_max_rows = 1 << 18;

// mask[s] = 1 if slice s holds a column >= first_halo_col (one warp per slice)
__global__ void __launch_bounds__(kBlock)
sell_halo_mask_kernel(int64_t nslices, const int64_t *__restrict__ slice_ptr, int64_t uniform_len,
                      const int32_t *__restrict__ cols, int64_t first_halo_col, unsigned char *__restrict__ mask) {
    const int64_t s = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int64_t base = uniform_len > 0 ? s * kSlice * uniform_len : slice_ptr[s];
    const int64_t end = uniform_len > 0 ? base + kSlice * uniform_len : slice_ptr[s + 1];
    int hit = 0;
    for (int64_t p = base + lane; p < end; p += 32) hit |= cols[p] >= first_halo_col;
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0) mask[s] = hit ? 1 : 0;
}

// second stage of the deterministic norm / dot: one CTA sums the per-block partials in a fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ partials, int64_t n,
                                                               double *__restrict__ out) {
    pdl_prologue();
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += partials[i];
    s = block_sum<1024>(s);
    if (threadIdx.x == 0) *out = s;
}

template <int MODE>
int launch_sell_tma(const mg_sell *M, int64_t max_len, const double *x, const double *b, const double *aux, double *y,
                    double omega, double *partials, int64_t row0, int64_t row1, int *grid_out, cudaStream_t st,
                    const char *name);

// rows per launch from which the bulk-async staged kernel (sell_tma.cu) is used; 0 disables it
static int64_t g_tma_min_rows = 0;   // off by default: the register-staged kernel is as fast (see sell_tma.cu)

// an exchange site riding on a SELL launch (prepared by comm_prepare)
struct SellFuse {
    ExArgs ex;
    int nex;                       // exchange CTAs in front of the compute CTAs
    const unsigned char *mask;     // per slice of the matrix: reads halo columns (NULL: assume every slice does)
};

// which launches can carry an exchange site: the thread-per-row kernel only
bool sell_fusable(const mg_sell *A, int64_t row0, int64_t row1) {
    if (row1 <= row0 || g_tma_min_rows > 0) return false;
    const int64_t ml = A->max_slice_len;
    if (g_wide_min_len > 0 && ml >= g_wide_min_len && ml <= kWideMaxLen && row1 - row0 <= g_wide_max_rows) return false;
    return true;
}

template <int MODE>
static int launch_sell(const mg_sell *A, const double *x, const double *b, const double *aux, double *y,
                       double omega, double *partials, int64_t row0, int64_t row1, cudaStream_t st,
                       const char *name, int *nblocks_out = nullptr, const SellFuse *fuse = nullptr) {
    if (fuse && !sell_fusable(A, row0, row1)) return set_error(MG_ERR_INVALID, name, "this launch cannot carry an exchange site");
    if (row1 <= row0) return MG_OK;
    if (g_tma_min_rows > 0 && row1 - row0 >= g_tma_min_rows && A->max_slice_len > 0) {
        int grid = 0;
        const int rc = launch_sell_tma<MODE>(A, A->max_slice_len, x, b, aux, y, omega, partials, row0, row1, &grid, st, name);
        if (rc <= 0) {
            if (nblocks_out) *nblocks_out = grid;
            return rc;
        }
    }
    SellArgs a;
    a.slice_ptr = A->d_slice_ptr;
    a.cols = A->d_cols;
    a.vals = A->d_vals;
    a.row_begin = row0;
    a.row_end = row1;
    a.first_row = row0 & ~(int64_t)(kSlice - 1);
    a.nrows = A->nrows;
    const int64_t nthreads = row1 - a.first_row;
    const int64_t ml = A->max_slice_len;
    const bool uni = A->uniform_len > 0 && A->uniform_len == ml;
    if (g_wide_min_len > 0 && ml >= g_wide_min_len && ml <= kWideMaxLen && row1 - row0 <= g_wide_max_rows) {
        const int wps = ml <= 4 * kWideU ? 4 : 8;
        const int spc = kBlock / 32 / wps;                            // slices per CTA
        const int64_t nsl = (nthreads + kSlice - 1) / kSlice;
        const int64_t wgrid = (nsl + spc - 1) / spc;
        if (wgrid > 0x7fffffffLL) return set_error(MG_ERR_OVERFLOW, name, "grid too large");
        if (wps == 4)
            launch_k(sell_wide_kernel<MODE, 4>, (unsigned)wgrid, kBlock, st, a, (int)(uni ? ml : 0), x, b, aux, y, omega, partials);
        else
            launch_k(sell_wide_kernel<MODE, 8>, (unsigned)wgrid, kBlock, st, a, (int)(uni ? ml : 0), x, b, aux, y, omega, partials);
        MG_CHECK_LAUNCH(name);
        if (nblocks_out) *nblocks_out = (int)wgrid;
        return MG_OK;
    }
    const int64_t grid = (nthreads + kBlock - 1) / kBlock;
    if (grid > 0x7fffffffLL) return set_error(MG_ERR_OVERFLOW, name, "grid too large");
    if (g_implied_columns && uni && A->d_slice_off && ml >= 1 && ml <= 8) {      // see sell_kernel_reg
#define MG_REG_CASE(L)                                                                                                  \
    case L:                                                                                                             \
        if (fuse) launch_k(sell_kernel_reg_fused<MODE, L>, (unsigned)(grid + fuse->nex), kBlock, st, a, A->d_slice_off, x, b, aux, y, omega, partials, fuse->ex, fuse->mask); \
        else launch_k(sell_kernel_reg<MODE, L>, (unsigned)grid, kBlock, st, a, A->d_slice_off, x, b, aux, y, omega, partials); \
        break
        switch (ml) {
            MG_REG_CASE(1); MG_REG_CASE(2); MG_REG_CASE(3); MG_REG_CASE(4);
            MG_REG_CASE(5); MG_REG_CASE(6); MG_REG_CASE(7); MG_REG_CASE(8);
        }
#undef MG_REG_CASE
        MG_CHECK_LAUNCH(name);
        if (nblocks_out) *nblocks_out = (int)grid;
        return MG_OK;
    }
#define MG_SELL_CASE(L)                                                                                      \
    do {                                                                                                     \
        if (fuse) {                                                                                          \
            if (uni) launch_k(sell_kernel_fused<MODE, L, true>, (unsigned)(grid + fuse->nex), kBlock, st, a, x, b, aux, y, omega, partials, fuse->ex, fuse->mask); \
            else launch_k(sell_kernel_fused<MODE, L, false>, (unsigned)(grid + fuse->nex), kBlock, st, a, x, b, aux, y, omega, partials, fuse->ex, fuse->mask);     \
        } else if (uni) launch_k(sell_kernel<MODE, L, true>, (unsigned)grid, kBlock, st, a, x, b, aux, y, omega, partials); \
        else launch_k(sell_kernel<MODE, L, false>, (unsigned)grid, kBlock, st, a, x, b, aux, y, omega, partials);     \
    } while (0)
    switch (ml) {
        case 1: MG_SELL_CASE(1); break;
        case 2: MG_SELL_CASE(2); break;
        case 3: MG_SELL_CASE(3); break;
        case 4: MG_SELL_CASE(4); break;
        case 5: MG_SELL_CASE(5); break;
        case 6: MG_SELL_CASE(6); break;
        case 7: MG_SELL_CASE(7); break;
        case 8: MG_SELL_CASE(8); break;
        default:   // long rows, or length unknown (0)
            if (fuse) launch_k(sell_kernel_fused<MODE, 0, false>, (unsigned)(grid + fuse->nex), kBlock, st, a, x, b, aux, y, omega, partials, fuse->ex, fuse->mask);
            else launch_k(sell_kernel<MODE, 0, false>, (unsigned)grid, kBlock, st, a, x, b, aux, y, omega, partials);
    }
#undef MG_SELL_CASE
    MG_CHECK_LAUNCH(name);
    if (nblocks_out) *nblocks_out = (int)grid;
    return MG_OK;
}

int sell_spmv(const mg_sell *A, const double *x, double *y, cudaStream_t st) {
    return launch_sell<SPMV>(A, x, nullptr, nullptr, y, 0.0, nullptr, 0, A->nrows, st, "sell_spmv");
}
int sell_residual(const mg_sell *A, const double *x, const double *b, double *r, cudaStream_t st) {
    return launch_sell<RESID>(A, x, b, nullptr, r, 0.0, nullptr, 0, A->nrows, st, "sell_residual");
}
// variants carrying an exchange site
int sell_spmv_fused(const mg_sell *A, const double *x, double *y, const SellFuse *f, cudaStream_t st) {
    return launch_sell<SPMV>(A, x, nullptr, nullptr, y, 0.0, nullptr, 0, A->nrows, st, "sell_spmv", nullptr, f);
}
int sell_residual_fused(const mg_sell *A, const double *x, const double *b, double *r, const SellFuse *f, cudaStream_t st) {
    return launch_sell<RESID>(A, x, b, nullptr, r, 0.0, nullptr, 0, A->nrows, st, "sell_residual", nullptr, f);
}
int sell_gs_rows_fused(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *f,
                       cudaStream_t st) {
    return launch_sell<GS>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows", nullptr, f);
}
// colour sweep that pushes its own boundary values (carry: the previous site riding along, or NULL)
int sell_gs_rows_push(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, const SellFuse *carry,
                      const SellPush *push, cudaStream_t st) {
    const char *name = "sell_gs_rows_push";
    if (!push || !sell_fusable(A, row0, row1)) return set_error(MG_ERR_INVALID, name, "this launch cannot push an exchange site");
    SellArgs a;
    a.slice_ptr = A->d_slice_ptr;
    a.cols = A->d_cols;
    a.vals = A->d_vals;
    a.row_begin = row0;
    a.row_end = row1;
    a.first_row = row0 & ~(int64_t)(kSlice - 1);
    a.nrows = A->nrows;
    const int64_t ml = A->max_slice_len;
    const bool uni = A->uniform_len > 0 && A->uniform_len == ml;
    const int64_t grid = (row1 - a.first_row + kBlock - 1) / kBlock;
    const int nex = carry ? carry->nex : 0;
    if (grid + nex > 0x7fffffffLL) return set_error(MG_ERR_OVERFLOW, name, "grid too large");
    SellPush p = *push;
    if (p.tail_first < 0 || p.tail_first > grid / 2) p.tail_first = 0;
    ExArgs none;
    memset(&none, 0, sizeof(none));
    const ExArgs &fx = carry ? carry->ex : none;
    const unsigned char *mask = carry ? carry->mask : nullptr;
#define MG_PUSH_CASE(L, U)                                                                                              \
    do {                                                                                                                \
        if (carry) launch_k(sell_gs_push_kernel<L, U, true>, (unsigned)(grid + nex), kBlock, st, a, x, b, fx, mask, p); \
        else launch_k(sell_gs_push_kernel<L, U, false>, (unsigned)grid, kBlock, st, a, x, b, fx, mask, p);              \
    } while (0)
#define MG_PUSH_LEN(L)                  \
    do {                                \
        if (uni) MG_PUSH_CASE(L, true); \
        else MG_PUSH_CASE(L, false);    \
    } while (0)
    switch (ml) {
        case 1: MG_PUSH_LEN(1); break;
        case 2: MG_PUSH_LEN(2); break;
        case 3: MG_PUSH_LEN(3); break;
        case 4: MG_PUSH_LEN(4); break;
        case 5: MG_PUSH_LEN(5); break;
        case 6: MG_PUSH_LEN(6); break;
        case 7: MG_PUSH_LEN(7); break;
        case 8: MG_PUSH_LEN(8); break;
        default: MG_PUSH_CASE(0, false);
    }
#undef MG_PUSH_LEN
#undef MG_PUSH_CASE
    MG_CHECK_LAUNCH(name);
    return MG_OK;
}
int sell_prolong_fused(const mg_sell *Q, const double *e, const double *u, double *uo, const SellFuse *f, cudaStream_t st) {
    return launch_sell<PROLONG>(Q, e, nullptr, u, uo, 0.0, nullptr, 0, Q->nrows, st, "sell_prolong", nullptr, f);
}
int sell_residual_norm2(const mg_sell *A, const double *x, const double *b, double *partials, double *out,
                        cudaStream_t st) {
    int nblocks = 0;
    int rc = launch_sell<RESNORM>(A, x, b, nullptr, nullptr, 0.0, partials, 0, A->nrows, st,
                                  "sell_residual_norm2", &nblocks);
    if (rc) return rc;
    launch_k(reduce_partials_kernel, 1u, 1024u, st, (const double *)partials, (int64_t)nblocks, out);
    MG_CHECK_LAUNCH("reduce_partials");
    return MG_OK;
}
// first stage only: per-block partial sums of ||b - A x||^2 (the partitioned norm adds its own second stage, comm.cu)
int sell_residual_partials(const mg_sell *A, const double *x, const double *b, double *partials, int *nblocks,
                           cudaStream_t st) {
    return launch_sell<RESNORM>(A, x, b, nullptr, nullptr, 0.0, partials, 0, A->nrows, st, "sell_residual_partials", nblocks);
}
int sell_jacobi(const mg_sell *A, const double *dinv, const double *x, const double *b, double *xo,
                double omega, cudaStream_t st) {
    return launch_sell<JACOBI>(A, x, b, dinv, xo, omega, nullptr, 0, A->nrows, st, "sell_jacobi");
}
int sell_gs_rows(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, cudaStream_t st) {
    return launch_sell<GS>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows");
}
int sell_prolong(const mg_sell *Q, const double *e, const double *u, double *uo, cudaStream_t st) {
    return launch_sell<PROLONG>(Q, e, nullptr, u, uo, 0.0, nullptr, 0, Q->nrows, st, "sell_prolong");
}

}  // namespace mgb

using namespace mgb;

extern "C" {

static int check_sell(const mg_sell *A) {
    if (!A || A->nrows < 0 || (A->nrows > 0 && (!A->d_slice_ptr || !A->d_cols || !A->d_vals)))
        return set_error(MG_ERR_INVALID, "mg_sell", "null or negative-sized SELL matrix");
    if (A->nslices != (A->nrows + kSlice - 1) / kSlice)
        return set_error(MG_ERR_INVALID, "mg_sell", "nslices != ceil(nrows/32)");
    return MG_OK;
}

int mg_sell_spmv(const mg_sell *A, const double *d_x, double *d_y, void *stream) {
    if (int rc = check_sell(A)) return rc;
    return sell_spmv(A, d_x, d_y, (cudaStream_t)stream);
}
int mg_sell_residual(const mg_sell *A, const double *d_x, const double *d_b, double *d_r, void *stream) {
    if (int rc = check_sell(A)) return rc;
    return sell_residual(A, d_x, d_b, d_r, (cudaStream_t)stream);
}
/* worst case over the kernels that write partials: the warps-per-slice kernel with eight warps per slice has one CTA
 * (one partial) per 32 rows */
int64_t mg_norm_workspace_size(int64_t n) { return (n + kSlice - 1) / kSlice + 1; }
/* rows per launch from which the bulk-async staged SELL kernel is used (0 = never); returns the old value */
int mg_sell_halo_mask(const mg_sell *A, int64_t first_halo_col, unsigned char *d_mask, void *stream) {
    MG_REQUIRE(A && d_mask && A->nrows > 0, "null argument");
    const int64_t ns = A->nslices;
    sell_halo_mask_kernel<<<(unsigned)((ns * 32 + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        ns, A->d_slice_ptr, A->uniform_len, A->d_cols, first_halo_col, d_mask);
    MG_CHECK_LAUNCH("sell_halo_mask");
    return MG_OK;
}
/* per-slice column offsets of a UNIFORM matrix (uniform_len entries per row): d_off[s * len + j]; slices that are not
 * regular get d_off[s * len] = INT32_MIN (see sell_kernel_reg) */
int mg_sell_slice_offsets(const mg_sell *A, int32_t *d_off, void *stream) {
    if (int rc = check_sell(A)) return rc;
    MG_REQUIRE(d_off && A->uniform_len > 0 && A->uniform_len == A->max_slice_len, "uniform SELL matrix expected");
    if (A->nslices == 0) return MG_OK;
    sell_slice_offsets_kernel<<<(unsigned)((A->nslices * 32 + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        A->nslices, A->nrows, (int)A->uniform_len, A->d_cols, d_off);
    MG_CHECK_LAUNCH("sell_slice_offsets");
    return MG_OK;
}
int mg_set_implied_columns(int enabled) {
    const int prev = g_implied_columns;
    g_implied_columns = enabled ? 1 : 0;
    return prev;
}
int64_t mg_set_wide_min_len(int64_t len) {
    const int64_t prev = g_wide_min_len;
    g_wide_min_len = len < 0 ? 0 : len;
    return prev;
}
int64_t mg_set_wide_max_rows(int64_t rows) {
    const int64_t prev = g_wide_max_rows;
    g_wide_max_rows = rows < 0 ? 0 : rows;
    return prev;
}
int64_t mg_set_tma_min_rows(int64_t rows) {
    const int64_t old = g_tma_min_rows;
    g_tma_min_rows = rows;
    return old;
}
int mg_sell_residual_norm2(const mg_sell *A, const double *d_x, const double *d_b, double *d_partials,
                           double *d_norm2, void *stream) {
    if (int rc = check_sell(A)) return rc;
    if (A->nrows == 0) return mg_fill(1, 0.0, d_norm2, stream);
    return sell_residual_norm2(A, d_x, d_b, d_partials, d_norm2, (cudaStream_t)stream);
}
int mg_sell_jacobi(const mg_sell *A, const double *d_dinv, const double *d_x, const double *d_b,
                   double *d_x_out, double omega, void *stream) {
    if (int rc = check_sell(A)) return rc;
    MG_REQUIRE(d_x != d_x_out, "Jacobi is out of place: x_out must not alias x");
    return sell_jacobi(A, d_dinv, d_x, d_b, d_x_out, omega, (cudaStream_t)stream);
}
int mg_sell_gs_rows(const mg_sell *A, double *d_x, const double *d_b, int64_t row0, int64_t row1,
                    void *stream) {
    if (int rc = check_sell(A)) return rc;
    MG_REQUIRE(row0 >= 0 && row1 <= A->nrows && row0 <= row1, "row range outside the matrix");
    return sell_gs_rows(A, d_x, d_b, row0, row1, (cudaStream_t)stream);
}
int mg_sell_prolong_correct(const mg_sell *Q, const double *d_e, const double *d_u, double *d_u_out,
                            void *stream) {
    if (int rc = check_sell(Q)) return rc;
    return sell_prolong(Q, d_e, d_u, d_u_out, (cudaStream_t)stream);
}

}  // extern "C"
