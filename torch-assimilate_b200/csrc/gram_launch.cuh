// Shared launch helper of the two Gram translation units.
#pragma once
#include "letkf_kernel.cuh"

namespace b200da {

template <typename T, int KT, int G, int WPG, int ER>
static int launch_fused_t(const LetkfParams& P, int nblocks, cudaStream_t st) {
    const size_t hdr = (sizeof(BlockHeader<G>) + 31) & ~size_t(31);
    const size_t smem = hdr + gram_smem_bytes<T, KT, G, WPG, ER>() + sizeof(double) * taper_tab_doubles(P.tt);
    if (smem > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    auto kern = k_letkf_gram<T, KT, G, WPG, ER>;
    B200DA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, G * WPG * 32, smem, st>>>(P);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

template <int KT, int G, int WPG, int ER>
static int launch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    return pl->dtype == B200DA_F32 ? launch_fused_t<float, KT, G, WPG, ER>(P, nblocks, st)
                                   : launch_fused_t<double, KT, G, WPG, ER>(P, nblocks, st);
}

// ER = 1, 2, 3 extra rows on the DFMA pipe (gram_launch_b.cu, _c.cu, _d.cu): k + 1 = 8 KT + ER.  The grid points per
// block / warps per grid point follow the plan's kt = KT + 1 so that blocks do not depend on which Gram variant runs.
int dispatch_fused_er1(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);
int dispatch_fused_er2(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);
int dispatch_fused_er3(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);

#define B200DA_KTE_CASE(KT, G, WPG, ER) case KT + 1: return launch_fused<KT, G, WPG, ER>(pl, P, nblocks, st);
#define B200DA_DISPATCH_ER(ER)                                                                                        \
    switch (pl->kt) {                                                                                                 \
        B200DA_KTE_CASE(1, 8, 1, ER) B200DA_KTE_CASE(2, 8, 1, ER) B200DA_KTE_CASE(3, 8, 1, ER) B200DA_KTE_CASE(4, 8, 1, ER)   \
        B200DA_KTE_CASE(5, 8, 2, ER) B200DA_KTE_CASE(6, 8, 2, ER) B200DA_KTE_CASE(7, 4, 4, ER) B200DA_KTE_CASE(8, 4, 4, ER)   \
        B200DA_KTE_CASE(9, 4, 4, ER) B200DA_KTE_CASE(10, 2, 8, ER) B200DA_KTE_CASE(11, 2, 8, ER) B200DA_KTE_CASE(12, 2, 8, ER) \
        B200DA_KTE_CASE(13, 2, 8, ER) B200DA_KTE_CASE(14, 2, 8, ER) B200DA_KTE_CASE(15, 2, 8, ER) B200DA_KTE_CASE(16, 2, 8, ER) \
        default: return B200DA_ERR_UNSUPPORTED;                                                                       \
    }

}  // namespace b200da
