// K1 — cell binning of grid points and observations (count -> scan -> fill -> per-cell sort) and the
// formation of grid-point blocks.  Replaces, for the B200 engine, the brute-force "every grid point looks at
// every observation" search of the reference (pytassim/localization/gaspari_cohn.py:124-136) and the
// per-grid-point boolean gather of interface/wrapper.py:91-97.
#pragma once
#include <algorithm>
#include <cmath>
#include "plan.cuh"
#include "sort_scan.cuh"

namespace b200da {

// ---- kernels --------------------------------------------------------------------------------------------------

__global__ void k_positions(Geometry g, const double* __restrict__ coord, int64_t n, Pos4* __restrict__ pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    bin_position(g, coord, n, i, x, y, z);
    Pos4 p; p.x = x; p.y = y; p.z = z; p.id = i;
    pos[i] = p;
}

// min / max of the positions per dimension by one CTA: out[0..2] = min, out[3..5] = max, out[6] = #non-finite
__global__ void __launch_bounds__(1024) k_bbox(const Pos4* __restrict__ pos, int64_t n, double* __restrict__ out) {
    __shared__ double red[7][32];
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    double bad = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const Pos4 p = pos[i];
        const double v[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (!isfinite(v[d])) { bad += 1.0; continue; }
            mn[d] = fmin(mn[d], v[d]); mx[d] = fmax(mx[d], v[d]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], off));
            mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], off));
        }
        bad += __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if (lane == 0) {
        for (int d = 0; d < 3; ++d) { red[d][warp] = mn[d]; red[3 + d][warp] = mx[d]; }
        red[6][warp] = bad;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        const int nw = blockDim.x >> 5;
        double r = red[threadIdx.x][0];
        for (int w = 1; w < nw; ++w) {
            const double v = red[threadIdx.x][w];
            r = threadIdx.x < 3 ? fmin(r, v) : (threadIdx.x < 6 ? fmax(r, v) : r + v);
        }
        out[threadIdx.x] = r;
    }
}

__device__ inline unsigned spread5(unsigned v) {   // 5 bits -> every third bit
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) r |= ((v >> b) & 1u) << (3 * b);
    return r;
}

// cell id and sort key of every point; counts[cell] += 1.  With `morton` the key orders the points of one cell
// along a Z-curve of 32^3 sub-cells (compact grid-point blocks); otherwise by original index only.
__global__ void k_cell_keys(Geometry g, const Pos4* __restrict__ pos, int64_t n, int morton,
                            int* __restrict__ cell_out, unsigned long long* __restrict__ key_out,
                            int* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Pos4 p = pos[i];
    const int c = cell_of(g, p.x, p.y, p.z);
    unsigned sub = 0;
    if (morton && c < g.ncell) {
        const double v[3] = {p.x, p.y, p.z};
        unsigned q[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double t = (v[d] - g.org[d]) / g.h[d];
            t = (t - floor(t)) * 32.0;
            int qi = (int)t;
            q[d] = (unsigned)min(max(qi, 0), 31);
        }
        sub = (spread5(q[0]) << 2) | (spread5(q[1]) << 1) | spread5(q[2]);
    }
    cell_out[i] = c;
    key_out[i] = ((unsigned long long)sub << 32) | (unsigned long long)(unsigned)i;
    atomicAdd(&counts[c], 1);
}

__global__ void k_fill_cells(const int* __restrict__ cell, const unsigned long long* __restrict__ key, int64_t n,
                             const int* __restrict__ start, int* __restrict__ cursor,
                             unsigned long long* __restrict__ sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell[i];
    const int slot = start[c] + atomicAdd(&cursor[c], 1);
    sorted[slot] = key[i];
}

// sorted keys -> block-sorted grid positions, order and the cell of every slot
__global__ void k_finalize_points(Geometry g, const unsigned long long* __restrict__ sorted, const Pos4* __restrict__ pos,
                                  int64_t n, Pos4* __restrict__ pos_sorted, int* __restrict__ order,
                                  int* __restrict__ cell_sorted) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned idx = (unsigned)(sorted[s] & 0xffffffffull);
    const Pos4 p = pos[idx];
    pos_sorted[s] = p;
    if (order) order[s] = (int)idx;
    if (cell_sorted) cell_sorted[s] = cell_of(g, p.x, p.y, p.z);
}

// extra coordinate columns (rows n_primary.. of the struct-of-arrays coordinates) in sorted order: out[s][e]
__global__ void k_gather_ext(const unsigned long long* __restrict__ sorted, const double* __restrict__ coord_ext, int64_t n,
                             int n_ext, double* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned idx = (unsigned)(sorted[s] & 0xffffffffull);
    for (int e = 0; e < n_ext; ++e) out[s * n_ext + e] = coord_ext[(int64_t)e * n + idx];
}

// flag[s] = s + 1 where a new segment of z-adjacent cells starts (0 elsewhere); the +1 keeps slot 0 non-zero
__global__ void k_segment_flags(Geometry g, const int* __restrict__ cell_sorted, int64_t n, int* __restrict__ flag) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    bool start = (s == 0);
    if (!start) {
        const int c = cell_sorted[s], cp = cell_sorted[s - 1];
        start = !(c == cp || (c == cp + 1 && (c % g.nc[2]) != 0));
    }
    flag[s] = start ? (int)(s + 1) : 0;
}

// seg_start_incl[s + 1] (exclusive max-scan shifted by one) -> block-start flags
__global__ void k_block_flags(const int* __restrict__ seg_scan, int64_t n, int gpb, int* __restrict__ bflag) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int seg_start = seg_scan[s + 1] - 1;      // inclusive max over flags[0..s], minus the +1 bias
    bflag[s] = ((s - seg_start) % gpb == 0) ? 1 : 0;
}

__global__ void k_block_scatter(const int* __restrict__ bflag, const int* __restrict__ bscan, int64_t n,
                                int* __restrict__ block_off) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (bflag[s]) block_off[bscan[s]] = (int)s;
    if (s == n - 1) block_off[bscan[n]] = (int)n;
}

// Cell-sorted, observation-major staging copy: ys[s][0..k-1] = Yn[:, j], ys[s][k] = d[j], rest 0 (j = obs of slot s)
template <typename T>
__global__ void k_gather_obs(const unsigned long long* __restrict__ sorted, const Pos4* __restrict__ pos, int64_t n,
                             const T* __restrict__ yn, const T* __restrict__ d, int k, int kp,
                             Pos4* __restrict__ pos_sorted, T* __restrict__ ys) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (s >= n) return;
    const unsigned j = (unsigned)(sorted[s] & 0xffffffffull);
    if (threadIdx.x == 0) pos_sorted[s] = pos[j];
    for (int m = threadIdx.x; m < kp; m += blockDim.x) {
        T v = (T)0;
        if (m < k) v = yn[(int64_t)m * n + j];
        else if (m == k) v = d[j];
        ys[s * kp + m] = v;
    }
}

// Observation-space variables of one stacked observation vector (interface/base.py:359-379 with a diagonal R,
// observation.py:241-245,277-279): mean_j = (sum_i hx[i][j]) / k summed in member order (bit-identical to numpy's
// mean over the leading axis), Yn[i][j] = (hx[i][j] - mean_j) * rc_j, d[j] = (y[j] - mean_j) * rc_j, rc_j = 1 / sqrt(var_j).
// One thread per observation: member rows are read coalesced, the second pass over the column hits L1 / L2.
template <typename T>
__global__ void __launch_bounds__(256) k_obs_prep(const T* __restrict__ hx, const T* __restrict__ y, const T* __restrict__ var,
                                                  int k, int64_t m, T* __restrict__ yn, T* __restrict__ d) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    T sum = (T)0;
    for (int i = 0; i < k; ++i) sum += hx[(int64_t)i * m + j];
    const T mean = sum / (T)k;
    const T rc = (T)1 / sqrt(var[j]);
    for (int i = 0; i < k; ++i) yn[(int64_t)i * m + j] = (hx[(int64_t)i * m + j] - mean) * rc;
    d[j] = (y[j] - mean) * rc;
}

// The same with the observation operator fused in front: for operators that SELECT grid columns of one
// state variable (obs_ops/lorenz_96/identity.py:88-92: `sel(var_name='x').sel(grid=points)`; examples/benchmark_letkf.py:
// 100-104: `sel(grid=obs_grid, method='nearest')`; obs_ops/base_ops.py:63-75 picks the observation times) the ensemble of
// observation equivalents is a gather from the pseudo state:  hx[i][j] = xp[src[j] + i * member_stride],  src[j] = offset of
// member 0 of (variable, time, grid column) of observation j in the (n_var, n_time, k, N) array.  The gathered values are
// exact copies and the arithmetic is that of k_obs_prep, so FP64 results are bit-identical to operator -> prep.
template <typename T>
__global__ void __launch_bounds__(256) k_obs_gather_prep(const T* __restrict__ xp, const long long* __restrict__ src,
                                                         int64_t member_stride, const T* __restrict__ y,
                                                         const T* __restrict__ var, int k, int64_t m, T* __restrict__ yn,
                                                         T* __restrict__ d) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const T* col = xp + src[j];
    T sum = (T)0;
    for (int i = 0; i < k; ++i) sum += col[(int64_t)i * member_stride];
    const T mean = sum / (T)k;
    const T rc = (T)1 / sqrt(var[j]);
    for (int i = 0; i < k; ++i) yn[(int64_t)i * m + j] = (col[(int64_t)i * member_stride] - mean) * rc;
    d[j] = (y[j] - mean) * rc;
}

// ---- host side --------------------------------------------------------------------------------------------------

inline int grid1d(int64_t n, int block) { return (int)std::max<int64_t>(1, (n + block - 1) / block); }

// Sort `n` points by (cell, key); returns cell_start (ncell + 2 ints: [ncell] = start of the discard bin,
// [ncell + 1] = n) in `start_out` and the sorted keys in plan->tmp_keys (second half).
inline int sort_points_by_cell(b200da_plan* pl, const Pos4* pos, int64_t n, int morton, int* start_out,
                               unsigned long long** sorted_out, cudaStream_t st) {
    const Geometry& g = pl->geom;
    const int64_t nseg = (int64_t)g.ncell + 1;
    int rc;
    if ((rc = pl->tmp_keys.ensure(sizeof(unsigned long long) * 2 * (size_t)std::max<int64_t>(n, 1)))) return rc;
    if ((rc = pl->tmp_cell.ensure(sizeof(int) * (size_t)std::max<int64_t>(n, 1)))) return rc;
    if ((rc = pl->tmp_count.ensure(sizeof(int) * 2 * (size_t)(nseg + 1)))) return rc;
    unsigned long long* keys = pl->tmp_keys.as<unsigned long long>();
    unsigned long long* sorted = keys + std::max<int64_t>(n, 1);
    int* cell = pl->tmp_cell.as<int>();
    int* counts = pl->tmp_count.as<int>();
    int* cursor = counts + (nseg + 1);
    B200DA_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * 2 * (size_t)(nseg + 1), st));
    if (n > 0) {
        k_cell_keys<<<grid1d(n, 256), 256, 0, st>>>(g, pos, n, morton, cell, keys, counts);
        B200DA_LAUNCH_CHECK();
    }
    k_exclusive_scan<int, int, 0><<<1, kScanThreads, 0, st>>>(counts, start_out, nseg);
    B200DA_LAUNCH_CHECK();
    if (n > 0) {
        k_fill_cells<<<grid1d(n, 256), 256, 0, st>>>(cell, keys, n, start_out, cursor, sorted);
        B200DA_LAUNCH_CHECK();
        const int nblk = (int)std::min<int64_t>(g.ncell, 148 * 64);
        k_segmented_sort<int><<<nblk, kSegSortThreads, 0, st>>>(sorted, start_out, g.ncell);
        B200DA_LAUNCH_CHECK();
    }
    *sorted_out = sorted;
    return B200DA_OK;
}

// Choose the cell grid from the bounding box of the grid points (bin space).
inline int choose_cells(b200da_plan* pl, const double bb[7]) {
    Geometry& g = pl->geom;
    if (bb[6] != 0.0) return B200DA_ERR_INVALID;            // non-finite coordinates
    const double cut = g.cut_bin * (1.0 + 1e-9) + 1e-300;
    double ext[3];
    int active = 0;
    for (int d = 0; d < 3; ++d) {
        const bool used = d >= 3 - g.nd;
        if (!used) { g.org[d] = -1.0; g.h[d] = 2.0; g.nc[d] = 1; ext[d] = 0.0; continue; }
        ++active;
        g.org[d] = bb[d] - cut;
        ext[d] = (bb[3 + d] + cut) - g.org[d];
        if (!(ext[d] > 0.0)) ext[d] = 1.0;
    }
    if (g.periodic) {
        if (bb[2] < 0.0 || bb[5] > g.period) return B200DA_ERR_INVALID;
        g.org[2] = 0.0; ext[2] = g.period;
    }
    double h = cut / kCellsPerCutoff;
    for (int iter = 0; iter < 64; ++iter) {
        double total = 1.0;
        for (int d = 3 - g.nd; d < 3; ++d) total *= std::max(1.0, std::ceil(ext[d] / h));
        if (total <= (double)kMaxCells) break;
        h *= std::pow(total / (double)kMaxCells, 1.0 / active) * 1.01;
    }
    int64_t ncell = 1;
    for (int d = 3 - g.nd; d < 3; ++d) {
        int nc = (int)std::max(1.0, std::ceil(ext[d] / h));
        if (g.periodic && d == 2) {
            nc = (int)std::max(1.0, std::floor(ext[d] / h));
            g.h[d] = ext[d] / nc;
        } else {
            g.h[d] = (nc == 1) ? ext[d] * (1.0 + 1e-12) : h;
        }
        g.nc[d] = nc;
        ncell *= nc;
    }
    g.ncell = (int)ncell;
    return B200DA_OK;
}

inline int set_grid_impl(b200da_plan* pl, const double* grid_coord, int64_t n, cudaStream_t st) {
    if (!pl || !grid_coord || n <= 0 || n > 0x7fffffffLL) return B200DA_ERR_INVALID;
    int rc;
    pl->have_grid = false; pl->have_obs = false;
    pl->n_grid = n;
    if ((rc = pl->tmp_pos.ensure(sizeof(Pos4) * (size_t)n))) return rc;
    if ((rc = pl->tmp_a.ensure(sizeof(int) * (size_t)(n + 2)))) return rc;
    if ((rc = pl->tmp_b.ensure(sizeof(int) * (size_t)(n + 2)))) return rc;
    if ((rc = pl->gpos.ensure(sizeof(Pos4) * (size_t)n))) return rc;
    if ((rc = pl->gorder.ensure(sizeof(int) * (size_t)n))) return rc;
    if ((rc = pl->block_off.ensure(sizeof(int) * (size_t)(n + 2)))) return rc;
    Pos4* pos = pl->tmp_pos.as<Pos4>();
    k_positions<<<grid1d(n, 256), 256, 0, st>>>(pl->geom, grid_coord, n, pos);
    B200DA_LAUNCH_CHECK();
    double* bb_dev = reinterpret_cast<double*>(pl->tmp_a.p);
    k_bbox<<<1, 1024, 0, st>>>(pos, n, bb_dev);
    B200DA_LAUNCH_CHECK();
    double bb[7];
    B200DA_CUDA(cudaMemcpyAsync(bb, bb_dev, sizeof(bb), cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    if ((rc = choose_cells(pl, bb))) return rc;
    const Geometry& g = pl->geom;
    if ((rc = pl->cell_start.ensure(sizeof(int) * (size_t)(g.ncell + 3)))) return rc;
    unsigned long long* sorted = nullptr;
    int* gstart = pl->cell_start.as<int>();                  // reused for the obs later
    if ((rc = sort_points_by_cell(pl, pos, n, /*morton=*/1, gstart, &sorted, st))) return rc;
    int* cell_sorted = pl->tmp_cell.as<int>();               // free again after the fill
    k_finalize_points<<<grid1d(n, 256), 256, 0, st>>>(g, sorted, pos, n, pl->gpos.as<Pos4>(), pl->gorder.as<int>(),
                                                      cell_sorted);
    B200DA_LAUNCH_CHECK();
    if (g.n_ext > 0) {
        if ((rc = pl->gext.ensure(sizeof(double) * (size_t)n * g.n_ext))) return rc;
        k_gather_ext<<<grid1d(n, 256), 256, 0, st>>>(sorted, grid_coord + (size_t)g.n_coord * n, n, g.n_ext, pl->gext.as<double>());
        B200DA_LAUNCH_CHECK();
    }
    int* fa = pl->tmp_a.as<int>();
    int* fb = pl->tmp_b.as<int>();
    k_segment_flags<<<grid1d(n, 256), 256, 0, st>>>(g, cell_sorted, n, fa);
    B200DA_LAUNCH_CHECK();
    k_exclusive_scan<int, int, 1><<<1, kScanThreads, 0, st>>>(fa, fb, n);          // fb[s+1] = seg start + 1
    B200DA_LAUNCH_CHECK();
    k_block_flags<<<grid1d(n, 256), 256, 0, st>>>(fb, n, pl->gpb, fa);            // fa = block-start flags
    B200DA_LAUNCH_CHECK();
    k_exclusive_scan<int, int, 0><<<1, kScanThreads, 0, st>>>(fa, fb, n);          // fb = block index per slot
    B200DA_LAUNCH_CHECK();
    k_block_scatter<<<grid1d(n, 256), 256, 0, st>>>(fa, fb, n, pl->block_off.as<int>());
    B200DA_LAUNCH_CHECK();
    int nb = 0;
    B200DA_CUDA(cudaMemcpyAsync(&nb, fb + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    pl->n_blocks = nb;
    pl->block_off_host.resize((size_t)nb + 1);
    B200DA_CUDA(cudaMemcpyAsync(pl->block_off_host.data(), pl->block_off.p, sizeof(int) * ((size_t)nb + 1),
                                cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    pl->have_grid = true;
    return B200DA_OK;
}

template <typename T>
inline int bin_obs_impl(b200da_plan* pl, const double* obs_coord, const T* yn, const T* d, int64_t m, cudaStream_t st) {
    if (!pl || m < 0 || m > 0x7fffffffLL) return B200DA_ERR_INVALID;
    if (m > 0 && (!obs_coord || !yn || !d)) return B200DA_ERR_INVALID;
    if (!pl->have_grid) return B200DA_ERR_STATE;
    int rc;
    const Geometry& g = pl->geom;
    pl->have_obs = false;
    pl->n_obs = m;
    const size_t mm = (size_t)std::max<int64_t>(m, 1);
    if ((rc = pl->tmp_pos.ensure(sizeof(Pos4) * mm))) return rc;
    if ((rc = pl->opos.ensure(sizeof(Pos4) * mm))) return rc;
    if ((rc = pl->ys.ensure(sizeof(T) * mm * (size_t)pl->kp))) return rc;
    Pos4* pos = pl->tmp_pos.as<Pos4>();
    if (m > 0) {
        k_positions<<<grid1d(m, 256), 256, 0, st>>>(g, obs_coord, m, pos);
        B200DA_LAUNCH_CHECK();
    }
    unsigned long long* sorted = nullptr;
    if ((rc = sort_points_by_cell(pl, pos, m, /*morton=*/0, pl->cell_start.as<int>(), &sorted, st))) return rc;
    if (m > 0) {
        dim3 blk(32, 8);
        k_gather_obs<T><<<grid1d(m, 8), blk, 0, st>>>(sorted, pos, m, yn, d, pl->k, pl->kp, pl->opos.as<Pos4>(),
                                                      pl->ys.as<T>());
        B200DA_LAUNCH_CHECK();
        if (g.n_ext > 0) {
            if ((rc = pl->oext.ensure(sizeof(double) * mm * g.n_ext))) return rc;
            k_gather_ext<<<grid1d(m, 256), 256, 0, st>>>(sorted, obs_coord + (size_t)g.n_coord * m, m, g.n_ext, pl->oext.as<double>());
            B200DA_LAUNCH_CHECK();
        }
    }
    pl->have_obs = true;
    return B200DA_OK;
}

}  // namespace b200da
