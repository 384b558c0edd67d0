// Shared launch helper of the two Gram translation units.
#pragma once
#include "letkf_kernel.cuh"

namespace b200da {

template <typename T, int KT, int G, int WPG, bool BROW>
static int launch_fused_t(const LetkfParams& P, int nblocks, cudaStream_t st) {
    const size_t hdr = (sizeof(BlockHeader<G>) + 31) & ~size_t(31);
    const size_t smem = hdr + gram_smem_bytes<T, KT, G, WPG, BROW>();
    if (smem > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    auto kern = k_letkf_gram<T, KT, G, WPG, BROW>;
    B200DA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, G * WPG * 32, smem, st>>>(P);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

template <int KT, int G, int WPG, bool BROW>
static int launch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    return pl->dtype == B200DA_F32 ? launch_fused_t<float, KT, G, WPG, BROW>(P, nblocks, st)
                                   : launch_fused_t<double, KT, G, WPG, BROW>(P, nblocks, st);
}

int dispatch_fused_brow(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);

}  // namespace b200da
