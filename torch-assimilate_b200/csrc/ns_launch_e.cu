// Fifth translation unit of the tensor-core solve kernel: the instantiation for 120 < k <= 128 (kts 16), split from
// ns_launch_c.cu so that a from-scratch parallel build stays below six minutes.
#define B200DA_NS_LARGE 4
#include "ns_launch.cu"
