// Instantiations and dispatch of the tensor-core solve kernel (ns_solve_kernel.cuh): its own translation unit so
// that the library builds in parallel.
#include <algorithm>
#include "launch.cuh"
#include "ns_solve_kernel.cuh"

namespace b200da {

// ---- tensor-core solve (ns_solve_kernel.cuh): warps per matrix and matrices per CTA by ensemble size ----------------
template <int KT> struct NsPick {
    static constexpr int WPM = KT <= 7 ? 1 : (KT <= 10 ? 2 : 4);
    static constexpr size_t GB = NsCfg<KT, WPM>::GROUP_BYTES;
    static constexpr int FIT = (int)((kMaxSmem - 1024) / GB);
    static constexpr int GROUPS = FIT * WPM >= 8 ? 8 / WPM : (FIT < 1 ? 1 : FIT);
};

template <int KT>
static int launch_ns(const NsParams& P, cudaStream_t st) {
    constexpr int WPM = NsPick<KT>::WPM, GROUPS = NsPick<KT>::GROUPS;
    constexpr size_t smem = NsCfg<KT, WPM>::GROUP_BYTES * GROUPS;
    static_assert(smem <= kMaxSmem, "solve kernel does not fit in shared memory");
    auto kern = k_letkf_solve_ns<KT, WPM, GROUPS>;
    B200DA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((P.n_slots + GROUPS - 1) / GROUPS, 148);
    kern<<<grid, GROUPS * WPM * 32, smem, st>>>(P);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}
template <int KT> static size_t scratch_bytes() { return sizeof(double) * ns_scratch_doubles(KT) * (size_t)NsPick<KT>::GROUPS * 148; }
#define B200DA_NSS_CASE(KT) case KT: return scratch_bytes<KT>();
size_t ns_scratch_bytes(int kts) {
    switch (kts) {
        B200DA_NSS_CASE(1) B200DA_NSS_CASE(2) B200DA_NSS_CASE(3) B200DA_NSS_CASE(4) B200DA_NSS_CASE(5) B200DA_NSS_CASE(6)
        B200DA_NSS_CASE(7) B200DA_NSS_CASE(8) B200DA_NSS_CASE(9) B200DA_NSS_CASE(10) B200DA_NSS_CASE(11) B200DA_NSS_CASE(12)
        B200DA_NSS_CASE(13) B200DA_NSS_CASE(14) B200DA_NSS_CASE(15) B200DA_NSS_CASE(16)
        default: return 0;
    }
}
#define B200DA_NS_CASE(KT) case KT: return launch_ns<KT>(P, st);
int dispatch_ns(int kts, const NsParams& P, cudaStream_t st) {
    switch (kts) {
        B200DA_NS_CASE(1) B200DA_NS_CASE(2) B200DA_NS_CASE(3) B200DA_NS_CASE(4) B200DA_NS_CASE(5) B200DA_NS_CASE(6)
        B200DA_NS_CASE(7) B200DA_NS_CASE(8) B200DA_NS_CASE(9) B200DA_NS_CASE(10) B200DA_NS_CASE(11) B200DA_NS_CASE(12)
        B200DA_NS_CASE(13) B200DA_NS_CASE(14) B200DA_NS_CASE(15) B200DA_NS_CASE(16)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

}  // namespace b200da
