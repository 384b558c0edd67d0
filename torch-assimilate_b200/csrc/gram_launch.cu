// Instantiations and dispatch of the DMMA Gram kernel (letkf_kernel.cuh) for ensemble sizes that are NOT a multiple of 8
// (the innovation row rides in the padding of the last tile row): its own translation unit so that the library builds
// in parallel.  gram_launch_b.cu holds the multiples of 8.
#include "launch.cuh"
#include "gram_launch.cuh"

namespace b200da {

#define B200DA_KT_CASE(KT, G, WPG) case KT: return launch_fused<KT, G, WPG, false>(pl, P, nblocks, st);
int dispatch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    // kernelised plans need d.d, which only the in-tile innovation row produces (kernelise.cuh)
    if (pl->k % 8 == 0 && pl->kprog.n == 0) return dispatch_fused_brow(pl, P, nblocks, st);
    switch (pl->kt) {
        B200DA_KT_CASE(1, 8, 1) B200DA_KT_CASE(2, 8, 1) B200DA_KT_CASE(3, 8, 1) B200DA_KT_CASE(4, 8, 1)
        B200DA_KT_CASE(5, 8, 1) B200DA_KT_CASE(6, 8, 2) B200DA_KT_CASE(7, 8, 2) B200DA_KT_CASE(8, 4, 4)
        B200DA_KT_CASE(9, 4, 4) B200DA_KT_CASE(10, 4, 4) B200DA_KT_CASE(11, 2, 8) B200DA_KT_CASE(12, 2, 8)
        B200DA_KT_CASE(13, 2, 8) B200DA_KT_CASE(14, 2, 8) B200DA_KT_CASE(15, 2, 8) B200DA_KT_CASE(16, 2, 8)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

}  // namespace b200da
