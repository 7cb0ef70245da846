// ct_shard.cu -- layout of one rank's packed tables in the all-gather buffer of a time-sharded run (ct_dist.cu).
#include "ct_shard.h"

namespace cts {

size_t layout(long nc, long np, long ns, size_t off[A_COUNT]) {
    static const int elt[A_COUNT] = {4, 4, 4, 4, 4, 4, 8, 8, 8, 8, 4, 4, 4, 4, 4, 4, 8, 8, 4, 4, 4, 4, 4};
    size_t o = 0;
    for (int k = 0; k < A_COUNT; ++k) {
        const long n = k <= A_FNSP ? nc : k == A_PPTR ? nc + 1 : k <= A_PS ? np : ns;
        off[k] = o;
        o += (((size_t)n * elt[k] + 15) / 16) * 16;
    }
    return o;
}

}  // namespace cts
