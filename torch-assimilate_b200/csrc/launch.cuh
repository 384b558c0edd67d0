// Kernel dispatchers that live in their own translation units (gram_launch.cu, ns_launch.cu).
#pragma once
#include "plan.cuh"

namespace b200da {

constexpr size_t kMaxSmem = 232448;     // 227 KB opt-in limit per CTA on sm_100

struct LetkfParams;
struct NsParams;
int dispatch_fused(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);
int dispatch_ns(int kts, const NsParams& P, cudaStream_t st);
size_t ns_scratch_bytes(int kts);     // global scratch of the solve kernel (two-level path of stiff matrices)
void tc_chunking(int k, int* n_cols, int* n_chunks, int* nc);
int launch_tc_gram(b200da_plan* pl, const LetkfParams& P, int nblocks, cudaStream_t st);

}  // namespace b200da
