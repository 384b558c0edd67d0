// DMMA Gram kernel for ensemble sizes that are a multiple of 8: KT = k / 8 tile rows for C, b = Y~ d~ by FMAs (BROW).
// The grid points per block / warps per grid point follow the plan's kt = k / 8 + 1 so that blocks do not depend on
// which Gram variant runs.
#include "launch.cuh"
#include "gram_launch.cuh"

namespace b200da {

#define B200DA_KTB_CASE(KT, G, WPG) case KT + 1: return launch_fused<KT, G, WPG, true>(pl, P, nblocks, st);
int dispatch_fused_brow(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    switch (pl->kt) {
        B200DA_KTB_CASE(1, 8, 1) B200DA_KTB_CASE(2, 8, 1) B200DA_KTB_CASE(3, 8, 1) B200DA_KTB_CASE(4, 8, 1)
        B200DA_KTB_CASE(5, 8, 2) B200DA_KTB_CASE(6, 8, 2) B200DA_KTB_CASE(7, 4, 4) B200DA_KTB_CASE(8, 4, 4)
        B200DA_KTB_CASE(9, 4, 4) B200DA_KTB_CASE(10, 2, 8) B200DA_KTB_CASE(11, 2, 8) B200DA_KTB_CASE(12, 2, 8)
        B200DA_KTB_CASE(13, 2, 8) B200DA_KTB_CASE(14, 2, 8) B200DA_KTB_CASE(15, 2, 8) B200DA_KTB_CASE(16, 2, 8)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

}  // namespace b200da
