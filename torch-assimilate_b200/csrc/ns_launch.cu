// Instantiations and dispatch of the tensor-core solve kernel (ns_solve_kernel.cuh): its own translation unit so
// that the library builds in parallel.
#include <algorithm>
#include "launch.cuh"
#include "ns_solve_kernel.cuh"

namespace b200da {

// ---- tensor-core solve (ns_solve_kernel.cuh): warps per matrix and matrices per CTA by ensemble size ----------------
#ifndef B200DA_NS_WPM1_KT
#define B200DA_NS_WPM1_KT 5      // largest tile count solved by one warp per matrix
#endif
#ifndef B200DA_NS_MAXW1
#define B200DA_NS_MAXW1 8        // warps per CTA of the one-warp-per-matrix variants
#endif
template <int KT> struct NsPick {
#ifdef B200DA_NS_WPM1
    static constexpr int WPM = KT <= 7 ? 1 : (KT <= 10 ? 2 : 4);
    static constexpr int MAXW = 8;
#else
    // two warps per matrix from k > 40 on: with one warp per matrix only 5 (k = 50) warps fit next to their matrices in shared
    // memory and the DMMA pipe idles half of the time (ncu: 53 %); splitting the tiles of a matrix over two warps doubles the
    // warps per scheduler at the same shared-memory footprint (cfg3 solve 225 -> 192 ms; k = 40: 6.0 -> 6.3 ms, so not there)
    static constexpr int WPM = KT <= B200DA_NS_WPM1_KT ? 1 : (KT <= 10 ? 2 : 4);
    static constexpr int MAXW = WPM == 1 ? B200DA_NS_MAXW1 : 16;
#endif
    static constexpr size_t GB = NsCfg<KT, WPM>::GROUP_BYTES;
    static constexpr int FIT = (int)((kMaxSmem - 256) / GB);      // static shared memory: one slot word per group
    static constexpr int GROUPS = FIT * WPM >= MAXW ? MAXW / WPM : (FIT < 1 ? 1 : FIT);
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
#ifndef B200DA_NS_LARGE
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
#endif
#define B200DA_NS_CASE(KT) case KT: return launch_ns<KT>(P, st);
#ifndef B200DA_NS_LARGE
int dispatch_ns_large(int kts, const NsParams& P, cudaStream_t st);      // ns_launch_b.cu: 11 <= kts <= 16
int dispatch_ns_few(int kts, const NsParams& P, cudaStream_t st);        // ns_launch_d.cu: fewer matrices than SMs
int dispatch_ns(int kts, const NsParams& P, cudaStream_t st) {
    if (P.n_slots <= 148 && kts >= 5 && kts <= 7) return dispatch_ns_few(kts, P, st);
    switch (kts) {
        B200DA_NS_CASE(1) B200DA_NS_CASE(2) B200DA_NS_CASE(3) B200DA_NS_CASE(4) B200DA_NS_CASE(5) B200DA_NS_CASE(6)
        B200DA_NS_CASE(7) B200DA_NS_CASE(8) B200DA_NS_CASE(9) B200DA_NS_CASE(10)
        default: return dispatch_ns_large(kts, P, st);
    }
}
#elif B200DA_NS_LARGE == 1
int dispatch_ns_large2(int kts, const NsParams& P, cudaStream_t st);     // ns_launch_c.cu: kts 14, 15; ns_launch_e.cu: kts 16
int dispatch_ns_large(int kts, const NsParams& P, cudaStream_t st) {
    switch (kts) {
        B200DA_NS_CASE(11) B200DA_NS_CASE(12) B200DA_NS_CASE(13)
        default: return dispatch_ns_large2(kts, P, st);
    }
}
#elif B200DA_NS_LARGE == 2
int dispatch_ns_large3(int kts, const NsParams& P, cudaStream_t st);     // ns_launch_e.cu: kts 16
int dispatch_ns_large2(int kts, const NsParams& P, cudaStream_t st) {
    switch (kts) {
        B200DA_NS_CASE(14) B200DA_NS_CASE(15)
        default: return dispatch_ns_large3(kts, P, st);
    }
}
#elif B200DA_NS_LARGE == 4
int dispatch_ns_large3(int kts, const NsParams& P, cudaStream_t st) {
    switch (kts) {
        B200DA_NS_CASE(16)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}
#else
// Fewer matrices than SMs (the reference's own benchmark shape: 40 grid points): the launch is one matrix deep, so its time is
// the latency of ONE solve; four warps per matrix, one matrix per CTA, halve the tile rows every warp walks through per product.
template <int KT>
static int launch_ns_few(const NsParams& P, cudaStream_t st) {
    constexpr int WPM = 4, GROUPS = 1;
    constexpr size_t smem = NsCfg<KT, WPM>::GROUP_BYTES * GROUPS;
    auto kern = k_letkf_solve_ns<KT, WPM, GROUPS>;
    B200DA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)std::min<int64_t>(P.n_slots, 148), GROUPS * WPM * 32, smem, st>>>(P);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}
int dispatch_ns_few(int kts, const NsParams& P, cudaStream_t st) {
    switch (kts) {
        case 5: return launch_ns_few<5>(P, st);
        case 6: return launch_ns_few<6>(P, st);
        case 7: return launch_ns_few<7>(P, st);
        default: return B200DA_ERR_UNSUPPORTED;
    }
}
#endif

}  // namespace b200da
