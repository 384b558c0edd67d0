// Fourth translation unit of the tensor-core solve kernel: the four-warps-per-matrix instantiations for launches with fewer
// matrices than SMs (dispatch_ns_few).
#define B200DA_NS_LARGE 3
#include "ns_launch.cu"
