// ct_internal.h -- pieces of ct_api.cu the other translation units of the library build on (the time-sharded drivers).
#pragma once
#include <thread>

#include "ct_ctx.h"

namespace cti {
int api_check_args(long T, int H, int W, const double* w_host, const double* thr_host, long thr_n, int in_dtype, int op);
// weights / special rows / thresholds on the device, row-indexed scratch for T planes
int api_prepare(ct_ctx* c, long T, int H, int W, const double* w_host, const double* thr_host, long thr_n, cudaStream_t st);
// planes [t0, t0 + nt) of the context's scratch from anom_dev (which points at the first of them)
int api_launch_threshold(ct_ctx* c, const void* anom_dev, int in_dtype, long t0, long nt, long thr_n, int thr_is_f32, int op,
                         cudaStream_t st, int all_bits);
int api_launch_paint(ct_ctx* c, long t0, long nt, int32_t* flag_dev, int sparse, cudaStream_t st);
int api_ensure_streams(ct_ctx* c);
// ordered phase with the host replays on tables that are already on the device (counts in the context)
int api_table_phase(ct_ctx* c, double overlap, int persistence, int twosided, int stage, long* n_features, cudaStream_t st);
int api_classic_tables(ct_ctx* c, cudaStream_t st);
bool api_plane_runs(ct_ctx* c, long plane, std::vector<cth::PlaneRun>& out, cudaStream_t st);
// host side of the host-buffer entry points (zeroing threads, expansion of the row-run table into the host cube)
void api_host_zero_start(ct_ctx* c, int32_t* flag_host, size_t cells, int share, std::vector<std::thread>& threads);
int api_host_expand_runs(ct_ctx* c, const uint32_t* h_x, const uint32_t* h_row, const int32_t* h_val, long R, long row_shift,
                         int W, int32_t* flag_host, int share);
}  // namespace cti
