// placeholder until the tcgen05 kernel lands: reports "not supported" so callers take the fp32 path
#include "common.cuh"
#include "layer_tc.h"
int mvn_tc_layer_supported(int, int, int) { return 0; }
int mvn_tc_layer_fwd(const void*, const void*, void*, float*, const float*, const PackedLayout&, const Geo&, int, cudaStream_t) {
    mvn_set_error("tensor-core layer kernel not built"); return -1;
}
