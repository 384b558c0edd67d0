// Instantiations and dispatch of the Gram kernel (letkf_kernel.cuh): its own translation unit so that the
// library builds in parallel.
#include "launch.cuh"
#include "letkf_kernel.cuh"

namespace b200da {

template <typename T, int KT, int G, int WPG>
static int launch_fused_t(const LetkfParams& P, int nblocks, cudaStream_t st) {
    const size_t hdr = (sizeof(BlockHeader<G>) + 31) & ~size_t(31);
    const size_t smem = hdr + gram_smem_bytes<T, KT, G, WPG>();
    if (smem > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    auto kern = k_letkf_gram<T, KT, G, WPG>;
    B200DA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, G * WPG * 32, smem, st>>>(P);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

template <int KT, int G, int WPG>
static int launch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    return pl->dtype == B200DA_F32 ? launch_fused_t<float, KT, G, WPG>(P, nblocks, st)
                                   : launch_fused_t<double, KT, G, WPG>(P, nblocks, st);
}

#define B200DA_KT_CASE(KT, G, WPG) case KT: return launch_fused<KT, G, WPG>(pl, P, nblocks, st);
int dispatch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st) {
    switch (pl->kt) {
        B200DA_KT_CASE(1, 8, 1) B200DA_KT_CASE(2, 8, 1) B200DA_KT_CASE(3, 8, 1) B200DA_KT_CASE(4, 8, 1)
        B200DA_KT_CASE(5, 8, 1) B200DA_KT_CASE(6, 8, 2) B200DA_KT_CASE(7, 8, 2) B200DA_KT_CASE(8, 4, 4)
        B200DA_KT_CASE(9, 4, 4) B200DA_KT_CASE(10, 4, 4) B200DA_KT_CASE(11, 2, 8) B200DA_KT_CASE(12, 2, 8)
        B200DA_KT_CASE(13, 2, 8) B200DA_KT_CASE(14, 2, 8)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

}  // namespace b200da
