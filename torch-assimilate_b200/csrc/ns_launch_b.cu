// Second translation unit of the tensor-core solve kernel: the instantiations for 80 < k <= 104 (kts 11..13), split from
// ns_launch.cu so that the parts compile in parallel.
#define B200DA_NS_LARGE 1
#include "ns_launch.cu"
