// tail.cuh -- interface between the V-cycle orchestration (cycle.cu) and the tail program (tail.cu)
#pragma once
#include "common.cuh"

namespace mgb {

enum TailKind { T_SELL = 0, T_FILL = 1, T_DIAG_SCALE = 2, T_COPY = 3 };

bool tail_recording();                       // operations are being recorded instead of launched (this thread)
bool tail_host_mode();                       // ... and will be executed on the host (CPU test-suite)
int64_t tail_max_rows();                     // mg_set_tail_max_rows threshold (0 = off)
void tail_begin(bool host, uint64_t shuffle);
void tail_end();
int tail_flush(cudaStream_t st);             // run what has been recorded so far, keep recording
void tail_stats_reset();
void tail_record_sell(int mode, const mg_sell *A, const double *x, const double *b, const double *aux, double *y,
                      double omega, int64_t row0, int64_t row1);
void tail_record_vector(int kind, int64_t n, double value, const double *src, const double *b, const double *aux,
                        double *y);

}  // namespace mgb
