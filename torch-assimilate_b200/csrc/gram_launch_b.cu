// DMMA Gram kernel with ONE extra row on the DFMA pipe (ER = 1): ensemble sizes that are a multiple of 8, KT = k / 8 tile rows
// for C, b = Y~ d~ by FMAs.  Its own translation unit so that the library builds in parallel.
#include "launch.cuh"
#include "gram_launch.cuh"

namespace b200da {

int dispatch_fused_er1(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) { B200DA_DISPATCH_ER(1) }

}  // namespace b200da
