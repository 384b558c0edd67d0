// Third translation unit of the tensor-core solve kernel: the instantiations for 104 < k <= 120 (kts 14, 15).
#define B200DA_NS_LARGE 2
#include "ns_launch.cu"
