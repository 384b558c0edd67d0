// The plan object behind the C ABI and host-side helpers (error handling, scratch buffers).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "common.cuh"
#include "kernelise.cuh"

namespace b200da {

extern thread_local std::string g_last_cuda_error;
extern int64_t g_launch_count;

#define B200DA_CUDA(expr)                                                                                    \
    do {                                                                                                      \
        cudaError_t err__ = (expr);                                                                           \
        if (err__ != cudaSuccess) {                                                                           \
            ::b200da::g_last_cuda_error = std::string(#expr) + ": " + cudaGetErrorString(err__);              \
            return B200DA_ERR_CUDA;                                                                           \
        }                                                                                                     \
    } while (0)

// B200DA_DEBUG_SYNC=1 in the environment: every launch is followed by a device synchronisation, so that a faulting kernel is
// named (file:line) in b200da_last_cuda_error instead of surfacing at a later API call.
extern int g_debug_sync;
#define B200DA_LAUNCH_CHECK()                                                                                \
    do {                                                                                                      \
        ++::b200da::g_launch_count;                                                                           \
        B200DA_CUDA(cudaGetLastError());                                                                      \
        if (::b200da::g_debug_sync) {                                                                         \
            cudaError_t err__ = cudaDeviceSynchronize();                                                      \
            if (err__ != cudaSuccess) {                                                                       \
                ::b200da::g_last_cuda_error = std::string("kernel launched at " __FILE__ ":") + std::to_string(__LINE__) + ": " + cudaGetErrorString(err__); \
                return B200DA_ERR_CUDA;                                                                       \
            }                                                                                                 \
        }                                                                                                     \
    } while (0)

// Grow-only device buffer owned by the plan.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return B200DA_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            g_last_cuda_error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
            p = nullptr;
            return B200DA_ERR_NOMEM;
        }
        cap = want;
        return B200DA_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace b200da

struct b200da_plan {
    // configuration
    int k = 0, n_slices = 1, n_coord = 1, dtype = B200DA_F64;
    int kt = 0;            // 8-row tiles of the augmented [Yn; d] matrix: ceil((k + 1) / 8)
    int kp = 0;            // padded row length of the staging copy: 8 * kt
    int gpb = 8;           // grid points per block (CTA) of the fused kernel
    bool use_tc = false;   // FP32 plan with k >= 32: tcgen05 Gram kernel, 128 grid points per block
    int gram_er = 0;       // rows of [Yn; d] the DMMA Gram accumulates by DFMA next to its tiles (letkf_kernel.cuh: ER)
    double rho = 1.0;
    b200da::Geometry geom{};
    // grid side
    int64_t n_grid = 0, n_blocks = 0;
    bool have_grid = false, have_obs = false;
    b200da::DevBuf gpos;        // Pos4[N]  block-sorted grid positions, id = original grid index
    b200da::DevBuf gorder;      // int32[N] slot -> original index
    b200da::DevBuf block_off;   // int32[n_blocks + 1]
    std::vector<int32_t> block_off_host;
    // obs side
    int64_t n_obs = 0;
    b200da::DevBuf opos;        // Pos4[M]  cell-sorted obs positions, id = original obs index
    b200da::DevBuf cell_start;  // int32[ncell + 2]
    b200da::DevBuf ys;          // T[M][kp] cell-sorted, obs-major [Yn; d; 0...]
    // scratch
    b200da::DevBuf tmp_keys, tmp_cell, tmp_count, tmp_a, tmp_b, tmp_pos;
    b200da::DevBuf host_stage_obs, host_stage_y, host_stage_d, host_stage_x, host_stage_xa;
    b200da::DevBuf etkf_partial, etkf_w, stats, cmat, counter, ns_scratch;
    b200da::DevBuf taper_tab;   // tabulated taper of the Gram kernels (b200da.cu: build_taper_table)
    b200da::TaperTab tt{};
    b200da::DevBuf tc_centre;   // FP32 tcgen05 plans: centring constant of every pair column (b200da.cu: tc_centre_constants)
    b200da::DevBuf gext, oext;  // extra coordinate columns in block- / cell-sorted order (b200da_plan_set_extra)
    // ambiguity protocol and device-side error flags (common.cuh: PlanStatus, PairRec)
    b200da::DevBuf devstat;     // PlanStatus
    b200da::DevBuf amb_list;    // PairRec[kAmbCapacity] recorded by the Gram kernel
    b200da::DevBuf over_list;   // PairRec[n_over] host decisions (b200da_plan_set_overrides)
    int n_over = 0;
    std::vector<int32_t> slot_of_host;   // original grid index -> block-sorted slot (lazily built for b200da_blocks_of_grid)
    double ns_stiff = 0.0;      // 0: default (2e3 for FP64 plans, 2e4 for FP32 plans); see NsParams::stiff
    int solver = B200DA_SOLVER_NEWTON_SCHULZ;
    b200da::KernelProgram kprog{};   // n > 0: kernelised ETKF (b200da_plan_set_kernel)
    bool collect_stats = false;
    // timing
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = -1.f, gram_ms = 0.f, solve_ms = 0.f;
    std::vector<cudaEvent_t> ev_pool;
    int n_ev = 0;
    int next_events(cudaEvent_t* a, cudaEvent_t* b, cudaEvent_t* c) {
        while ((int)ev_pool.size() < n_ev + 3) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return B200DA_ERR_CUDA;
            ev_pool.push_back(e);
        }
        *a = ev_pool[n_ev]; *b = ev_pool[n_ev + 1]; *c = ev_pool[n_ev + 2];
        n_ev += 3;
        return B200DA_OK;
    }
    std::string kernel_name;
};
