// ct_fast.h -- host side of the fast table pipeline (ct_fast.cu): plane kernel per time chunk, cooperative global kernel,
// O(events) host replay.  The tables land in the same device buffers (and layout) as the global-memory kernels' of
// ct_kernels.cu, so every slow path of ct_api.cu (debug stages, exact near-tie resolver, per-component replay) works on them.
#pragma once
#include "ct_ctx.h"
#include "ct_plane.h"

namespace ctf {

enum Outcome {
    FAST_OK = 0,        // the value of every component is in c_val on the device
    FAST_SLOW = 1,      // tables are valid (counts in the context); the caller continues with the ordered host phase
    FAST_RETRY = 2,     // a table was too small: buffers have been grown to the exact totals, build the tables again
    FAST_FALLBACK = 3,  // a plane does not fit the shared-memory budget: next_budget(), or the global-memory table kernels
    FAST_STATUS = 4     // global(..., raw_status): the status word was not zero, nothing was done (sharded run decides)
};

int begin(ct_ctx* c, long planes, cudaStream_t st);              // buffers, chain reset
int chunk(ct_ctx* c, long p0, long p1, cudaStream_t st);         // planes [p0, p1): one launch; chunks in time order
int finish(ct_ctx* c, cudaStream_t st);                          // totals of the chain -> control block
bool next_budget(ct_ctx* c);
int ensure_tables(ct_ctx* c, size_t runs, size_t comps, size_t pairs, size_t segs);
int ensure_control(ct_ctx* c);
int ensure_global_scratch(ct_ctx* c, size_t comps, size_t segs);
int totals_to_host(ct_ctx* c, cudaStream_t st, int* outcome);     // counts -> context (one synchronisation)
// global phase on the tables of `c` (T = planes of the cube they describe); ends with one stream synchronisation
int global(ct_ctx* c, long T, double overlap, int persistence, int twosided, long* n_features, cudaStream_t st, int* outcome,
           uint32_t* raw_status);

}  // namespace ctf
