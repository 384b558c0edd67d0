// Instantiations and dispatch of the DMMA Gram kernel (letkf_kernel.cuh) with every row of [Yn; d] inside the tiles (ER = 0:
// the innovation row rides in the padding of the last tile row): its own translation unit so that the library builds in
// parallel.  gram_launch_b.cu / _c.cu / _d.cu hold the variants with 1 / 2 / 3 trailing rows on the DFMA pipe.
#include "launch.cuh"
#include "gram_launch.cuh"

namespace b200da {

#define B200DA_KT_CASE(KT, G, WPG) case KT: return launch_fused<KT, G, WPG, 0>(pl, P, nblocks, st);
int dispatch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    switch (gram_extra_rows(pl)) {
        case 1: return dispatch_fused_er1(pl, P, nblocks, st);
        case 2: return dispatch_fused_er2(pl, P, nblocks, st);
        case 3: return dispatch_fused_er3(pl, P, nblocks, st);
        default: break;
    }
    switch (pl->kt) {
        B200DA_KT_CASE(1, 8, 1) B200DA_KT_CASE(2, 8, 1) B200DA_KT_CASE(3, 8, 1) B200DA_KT_CASE(4, 8, 1)
        B200DA_KT_CASE(5, 8, 1) B200DA_KT_CASE(6, 8, 2) B200DA_KT_CASE(7, 8, 2) B200DA_KT_CASE(8, 4, 4)
        B200DA_KT_CASE(9, 4, 4) B200DA_KT_CASE(10, 4, 4) B200DA_KT_CASE(11, 2, 8) B200DA_KT_CASE(12, 2, 8)
        B200DA_KT_CASE(13, 2, 8) B200DA_KT_CASE(14, 2, 8) B200DA_KT_CASE(15, 2, 8) B200DA_KT_CASE(16, 2, 8)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

}  // namespace b200da
