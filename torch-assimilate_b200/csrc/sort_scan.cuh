// Small device-side building blocks of the binning stage: an exclusive scan and a segmented sort.
// Both run once per set_grid / bin_obs call (not per grid point), so they favour simplicity.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200da {

// ---- exclusive scan by ONE CTA that walks the array in chunks ------------------------------------------------
// out[i] = op-prefix of in[0..i-1]; out[n] = total.  MODE 0: sum, MODE 1: max (identity 0, inputs >= 0).
// `in` and `out` may alias when the element types match (each chunk is read before it is written).
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;

template <typename TIn, typename TOut, int MODE>
__global__ void __launch_bounds__(kScanThreads) k_exclusive_scan(const TIn* __restrict__ in, TOut* __restrict__ out,
                                                                 int64_t n) {
    __shared__ TOut warp_tot[32];
    __shared__ TOut carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    const int64_t chunk = (int64_t)kScanThreads * kScanItems;
    for (int64_t base = 0; base < n; base += chunk) {
        TOut v[kScanItems];
        TOut local = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int64_t idx = base + (int64_t)tid * kScanItems + i;
            v[i] = idx < n ? (TOut)in[idx] : (TOut)0;
            local = MODE == 0 ? local + v[i] : (local > v[i] ? local : v[i]);
        }
        // inclusive scan of `local` across the CTA
        TOut incl = local;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const TOut t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl = MODE == 0 ? incl + t : (incl > t ? incl : t);
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            TOut w = warp_tot[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const TOut t = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w = MODE == 0 ? w + t : (w > t ? w : t);
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const TOut carry = carry_s;
        TOut excl_thread;   // prefix of everything before this thread's items
        {
            const TOut before_warp = warp > 0 ? warp_tot[warp - 1] : (TOut)0;
            const TOut incl_prev = __shfl_up_sync(0xffffffffu, incl, 1);
            const TOut in_warp = lane > 0 ? incl_prev : (TOut)0;
            if (MODE == 0) excl_thread = carry + before_warp + in_warp;
            else {
                excl_thread = carry > before_warp ? carry : before_warp;
                excl_thread = excl_thread > in_warp ? excl_thread : in_warp;
            }
        }
        TOut run = excl_thread;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            const int64_t idx = base + (int64_t)tid * kScanItems + i;
            if (idx < n) out[idx] = run;
            run = MODE == 0 ? run + v[i] : (run > v[i] ? run : v[i]);
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = run;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

// ---- segmented ascending sort of uint64 keys: one CTA per segment --------------------------------------------
// Bitonic network in its "all comparators ascending" form (the first stage of every merge mirrors the
// partner index), which stays correct for a length that is not a power of two when the missing tail is
// treated as +infinity: a comparator whose upper index is >= len is simply skipped.
constexpr int kSegSortThreads = 128;
constexpr int kSegSortSmem = 4096;     // keys held in shared memory; longer segments sort in global memory

template <typename TOff>
__global__ void __launch_bounds__(kSegSortThreads) k_segmented_sort(unsigned long long* __restrict__ keys,
                                                                    const TOff* __restrict__ seg_off, int64_t n_seg) {
    __shared__ unsigned long long sk[kSegSortSmem];
    for (int64_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const int64_t beg = (int64_t)seg_off[seg];
        const int64_t len = (int64_t)seg_off[seg + 1] - beg;
        if (len <= 1) continue;                                   // uniform per CTA
        unsigned long long* a = keys + beg;
        const bool in_smem = len <= kSegSortSmem;
        if (in_smem) {
            for (int64_t i = threadIdx.x; i < len; i += kSegSortThreads) sk[i] = a[i];
            __syncthreads();
            a = sk;
        }
        int64_t npow = 1;
        while (npow < len) npow <<= 1;
        for (int64_t size = 2; size <= npow; size <<= 1) {
            for (int64_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (int64_t t = threadIdx.x; t < (npow >> 1); t += kSegSortThreads) {
                    int64_t lo, hi;
                    if (stride == (size >> 1)) {                  // mirrored first stage
                        const int64_t blk = t / stride, off = t % stride;
                        lo = blk * size + off;
                        hi = blk * size + size - 1 - off;
                    } else {
                        const int64_t blk = t / stride, off = t % stride;
                        lo = blk * (stride << 1) + off;
                        hi = lo + stride;
                    }
                    if (hi < len) {
                        const unsigned long long x = a[lo], y = a[hi];
                        if (x > y) { a[lo] = y; a[hi] = x; }
                    }
                }
                __syncthreads();                                   // orders global accesses of this CTA too
            }
        }
        if (in_smem) {
            unsigned long long* g = keys + beg;
            for (int64_t i = threadIdx.x; i < len; i += kSegSortThreads) g[i] = sk[i];
            __syncthreads();
        }
    }
}

}  // namespace b200da
