// ct_extras.h -- launchers of ct_extras.cu that more than one translation unit uses.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace cte {

// sums / minima over the ranks of a time-sharded cube between the passes of quantile_time (null: single GPU)
struct QuantileReduce {
    int (*sum_u32)(void* user, uint32_t* buf, size_t n, cudaStream_t st);
    int (*min_i64)(void* user, long long* buf, size_t n, cudaStream_t st);
    void* user;
};
size_t quantile_scratch_bytes(long npts, int nq, int f64);
// np.nanquantile(x[:, y0:y1, :], q, axis=0) ('linear') of the local time steps -- of the whole time-sharded cube when `red`
// sums the tallies over the ranks; out [nq, y1 - y0, W] float64 (the same on every rank)
cudaError_t quantile_time(const void* x, int f64, long T, int H, int W, int y0, int y1, const double* q_dev, int nq, double* out,
                          void* scratch, const QuantileReduce* red, cudaStream_t st);

}  // namespace cte
