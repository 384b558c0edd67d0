// Kernel dispatchers that live in their own translation units (gram_launch.cu, ns_launch.cu).
#pragma once
#include "plan.cuh"

namespace b200da {

constexpr size_t kMaxSmem = 232448;     // 227 KB opt-in limit per CTA on sm_100

// rows of [Yn; d] the DMMA Gram accumulates by DFMA: (k + 1) mod 8 when that is 1..3; kernelised plans keep every row in the
// tiles (they need d.d, which only the in-tile innovation row produces, kernelise.cuh)
inline int gram_extra_rows(const b200da_plan* pl) {
    const int e = (pl->k + 1) % 8;
    if (pl->kprog.n > 0 || pl->k < 8 || e < 1 || e > 3) return 0;
    if (const char* v = getenv("B200DA_GRAM_ER_MAX")) { if (e > atoi(v)) return 0; }
    return e;
}

struct LetkfParams;
struct NsParams;
int dispatch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);
int dispatch_ns(int kts, const NsParams& P, cudaStream_t st);
size_t ns_scratch_bytes(int kts);     // global scratch of the solve kernel (two-level path of stiff matrices)
void tc_chunking(int k, int* n_cols, int* n_chunks, int* nc);
int launch_tc_gram(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);

}  // namespace b200da
