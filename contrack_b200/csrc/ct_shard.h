// ct_shard.h -- packed rank-local tables of a time-sharded run (layout) and the global tables they are merged into (ct_dist.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace cts {

// arrays of one rank's export buffer, in this order, each padded to 16 bytes (layout())
enum Arr {
    A_T = 0, A_Y0, A_Y1, A_X0, A_X1, A_CLS,           // int32 / uint32 [nc]   component plane, box, date-line class
    A_CONE, A_CONS, A_FE, A_FS,                       // float64 [nc]          class sums at the representative
    A_NSP, A_FNSP,                                    // uint32 [nc]
    A_PPTR,                                           // uint32 [nc + 1]       pair CSR over the plane-t component
    A_PB, A_PNPIX, A_PNSP,                            // uint32 [np]
    A_PE, A_PS,                                       // float64 [np]
    A_GT, A_GY0, A_GY1, A_GA, A_GB,                   // int32 / uint32 [ns]   date-line segments
    A_COUNT
};
size_t layout(long nc, long np, long ns, size_t off[A_COUNT]);      // returns the total bytes

struct GlobalTables {
    int32_t *t, *y0, *y1, *x0, *x1; uint32_t* cls;
    double *conE, *conS, *fE, *fS; uint32_t *nsp, *fnsp;
    uint32_t* pptr;
    uint32_t *p_b, *p_npix, *p_nsp; double *p_E, *p_S;
    int32_t *g_t, *g_y0, *g_y1; uint32_t *g_a, *g_b;
};

}  // namespace cts
