// Local-observation index lists (CSR) — the materialised form of `np.nonzero(use_obs)[0]` of
// GaspariCohn.localize_obs (pytassim/localization/gaspari_cohn.py:135) for every grid point.
// The fused LETKF kernel never materialises these lists; this path exists so that the neighbour search can
// be checked bit-exactly against the reference and for callers that want the lists.
// Same candidate sweep as the fused kernel (setup_block), one CTA per grid-point block.
#pragma once
#include "letkf_kernel.cuh"

namespace b200da {

struct NeighbourParams {
    Geometry g;
    const Pos4* gpos;
    const int* block_off;
    const Pos4* opos;
    const double* gext;                 // extra coordinates, see LetkfParams
    const double* oext;
    const int* cell_start;
    int64_t n_obs;
    double cut_pad;
    PlanStatus* status;                 // plan-owned error flags
    const PairRec* over;                // host decisions of the ambiguity protocol (MODE 0 / 1 lists reflect them)
    int n_over;
    // MODE 0: counts
    long long* counts;                  // [N] original grid order
    unsigned long long* n_ambiguous;    // or null
    // MODE 1: fill
    const long long* offsets;           // [N + 1]
    unsigned long long* keys;           // [nnz]  (orig obs id << 32) | sorted slot
    int* cursor;                        // [N] zero-initialised
    // MODE 2: ambiguous pairs
    long long capacity;
    long long* amb_grid; long long* amb_obs; double* amb_w;
    unsigned long long* amb_found;
};

template <int G, int MODE>
__global__ void __launch_bounds__(256) k_neighbours(const NeighbourParams P) {
    __shared__ BlockHeader<G> H;
    __shared__ int cnt[G];
    const Geometry& g = P.g;
    const int tid = threadIdx.x;
    setup_block<G>(H, g, P.gpos, P.block_off, P.cell_start, P.n_obs, P.cut_pad, blockIdx.x, P.status);
    if (tid < G) cnt[tid] = 0;
    __syncthreads();
    const int ng = H.ng, n_runs = H.n_runs, cand_total = H.cand_total;
    const double reach = (P.cut_pad + H.rb) * (1.0 + 1e-12);
    unsigned long long my_amb = 0;
    for (int c = tid; c < cand_total; c += blockDim.x) {
        int lo_ = 0, hi_ = n_runs;
        while (hi_ - lo_ > 1) {
            const int mid = (lo_ + hi_) >> 1;
            if (H.run_pref[mid] <= c) lo_ = mid; else hi_ = mid;
        }
        const int s = H.run_start[lo_] + (c - H.run_pref[lo_]);
        const Pos4 po = P.opos[s];
        if (!(bin_distance(g, H.cx, H.cy, H.cz, po.x, po.y, po.z) <= reach)) continue;
        for (int gi = 0; gi < ng; ++gi) {
            bool amb;
            const double* ge = P.gext + (size_t)(P.block_off[blockIdx.x] + gi) * g.n_ext;
            const double* oe = P.oext + (size_t)s * g.n_ext;
            double w = pair_weight(g, H.gp[gi].x, H.gp[gi].y, H.gp[gi].z, po.x, po.y, po.z, ge, oe, amb);
            if (MODE != 2 && P.n_over > 0) w = apply_override(P.over, P.n_over, H.gp[gi].id, po.id, w);
            if (MODE == 0) {
                if (w > 0.0) atomicAdd(&cnt[gi], 1);
                if (amb) ++my_amb;
            } else if (MODE == 1) {
                const long long og = H.gp[gi].id;
                // grid points the caller gave no room (offsets[og + 1] == offsets[og]: not selected) are skipped
                if (w > 0.0 && P.offsets[og + 1] > P.offsets[og]) {
                    const int at = atomicAdd(&P.cursor[og], 1);
                    P.keys[P.offsets[og] + at] = ((unsigned long long)(unsigned)po.id << 32) | (unsigned)s;
                }
            } else {
                if (amb) {
                    const unsigned long long at = atomicAdd(P.amb_found, 1ull);
                    if ((long long)at < P.capacity) {
                        P.amb_grid[at] = H.gp[gi].id; P.amb_obs[at] = po.id;
                        P.amb_w[at] = pair_weight_raw(g, H.gp[gi].x, H.gp[gi].y, H.gp[gi].z, po.x, po.y, po.z, ge, oe);
                    }
                }
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        if (tid < ng) P.counts[H.gp[tid].id] = cnt[tid];
        if (P.n_ambiguous && my_amb) atomicAdd(P.n_ambiguous, my_amb);
    }
}

// sorted keys -> obs indices, weights and ambiguity flags; one CTA per block-sorted grid slot
__global__ void k_neighbour_finalize(Geometry g, const Pos4* __restrict__ gpos, const Pos4* __restrict__ opos,
                                     const double* __restrict__ gext, const double* __restrict__ oext,
                                     const long long* __restrict__ offsets, const unsigned long long* __restrict__ keys,
                                     int64_t n_grid, int* __restrict__ idx, double* __restrict__ w_out,
                                     unsigned char* __restrict__ amb_out, const PairRec* __restrict__ over, int n_over) {
    for (int64_t slot = blockIdx.x; slot < n_grid; slot += gridDim.x) {
        const Pos4 gp = gpos[slot];
        const long long beg = offsets[gp.id], end = offsets[gp.id + 1];
        for (long long e = beg + threadIdx.x; e < end; e += blockDim.x) {
            const unsigned long long key = keys[e];
            idx[e] = (int)(key >> 32);
            if (w_out || amb_out) {
                const unsigned so = (unsigned)(key & 0xffffffffull);
                const Pos4 po = opos[so];
                bool amb;
                double w = pair_weight(g, gp.x, gp.y, gp.z, po.x, po.y, po.z, gext + (size_t)slot * g.n_ext,
                                       oext + (size_t)so * g.n_ext, amb);
                if (n_over > 0) w = apply_override(over, n_over, gp.id, po.id, w);
                if (w_out) w_out[e] = w;
                if (amb_out) amb_out[e] = amb ? 1 : 0;
            }
        }
    }
}

}  // namespace b200da
