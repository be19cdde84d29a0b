// placeholder while the aggregation pass is being written
#include "common.cuh"
extern "C" size_t al_aggregate_workspace_bytes(int dtype, int64_t n, int64_t E) { (void)dtype; (void)n; (void)E; return 256; }
extern "C" int al_aggregate(const void*, int, const uint32_t*, int64_t, int64_t, uint32_t*, uint32_t*, void*, int64_t*, void*, size_t, void*) {
    al_set_error("al_aggregate: not implemented yet");
    return AL_EINTERNAL;
}
