// DMMA Gram kernel with TWO extra rows on the DFMA pipe (ER = 2): k + 1 = 8 KT + 2 (k = 9, 17, ..., 49, ...).
#include "launch.cuh"
#include "gram_launch.cuh"

namespace b200da {

int dispatch_fused_er2(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) { B200DA_DISPATCH_ER(2) }

}  // namespace b200da
