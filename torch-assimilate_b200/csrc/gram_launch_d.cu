// DMMA Gram kernel with THREE extra rows on the DFMA pipe (ER = 3): k + 1 = 8 KT + 3 (k = 10, 18, ..., 50, ...) — the
// headline ensemble size k = 50 runs 21 DMMA tiles + 3 DFMA rows instead of 28 tiles.
#include "launch.cuh"
#include "gram_launch.cuh"

namespace b200da {

int dispatch_fused_er3(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    // experiment switch for the headline shape (k = 50): one warp per grid point (234 registers, no spills, 8 warps per CTA)
    // instead of two (128 registers); the default is whatever measured faster
    if (pl->kt == 7 && getenv("B200DA_GRAM_W1")) return launch_fused<6, 8, 1, 3>(pl, P, nblocks, st);
    B200DA_DISPATCH_ER(3)
}

}  // namespace b200da
