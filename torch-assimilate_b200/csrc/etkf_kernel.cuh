// Global ETKF (no localization): one k x k weight matrix from all observations, then a streaming update of
// the whole state.  Reference: pytassim/interface/etkf.py:99-120 -> core/etkf.py:79-103 (weights) and
// interface/base.py:257-278 (update).
#pragma once
#include "letkf_kernel.cuh"
#include "solve_kernel.cuh"

namespace b200da {

constexpr int kEtkfWarps = 8;

// Partial Gram of [Yn; d] over a chunk of observations per CTA; the lower-triangle tiles are split over the CTA's warps.
// ld = row stride of Yn (== m_obs for a whole array, the global M for a column shard of it).
// Yn is read in the reference layout (k, M) with coalesced row segments: a tile of kEtkfTileObs observations x (k + 1) rows
// goes through registers into shared memory once (the next tile's loads are in flight while the tensor work on the current
// one runs), and the eight warps take their DMMA fragments from there; every element is read from HBM exactly once.
constexpr int kEtkfTileObs = 64;
constexpr int kEtkfLd = kEtkfTileObs + 4;      // row stride of the staged tile in doubles: conflict-free fragment reads

// the tensor work of warp SUB on one staged tile: tile t of the lower triangle is owned by warp t % kEtkfWarps and is that
// warp's (t / kEtkfWarps)-th accumulator.  SUB is a template parameter so that only the warp's own DMMAs are in its
// instruction stream (a run-time ownership test turns the other 7/8 into predicated-off tensor instructions).
template <int KT, int SUB>
__device__ __forceinline__ void etkf_gram_tile(const double* __restrict__ tile, int nks, int lane,
                                               double (&acc)[(KT * (KT + 1) / 2 + kEtkfWarps - 1) / kEtkfWarps][2]) {
    for (int ks = 0; ks < nks; ++ks) {
        const double* fr = tile + (lane >> 2) * kEtkfLd + ks * 4 + (lane & 3);
        double f[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) f[t] = fr[t * 8 * kEtkfLd];
        int idx = 0;
#pragma unroll
        for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
            for (int nt = 0; nt <= mt; ++nt) {
                if (idx % kEtkfWarps == SUB) dmma884(acc[idx / kEtkfWarps][0], acc[idx / kEtkfWarps][1], f[mt], f[nt]);
                ++idx;
            }
        }
    }
}
template <typename T, int KT>
__global__ void __launch_bounds__(kEtkfWarps * 32) k_etkf_gram(const T* __restrict__ yn, const T* __restrict__ d,
                                                               int64_t m_obs, int64_t ld, int k, int64_t chunk,
                                                               double* __restrict__ partial) {
    constexpr int NTILES = KT * (KT + 1) / 2;
    constexpr int ACC = (NTILES + kEtkfWarps - 1) / kEtkfWarps;
    constexpr int KP = KT * 8;
    constexpr int NT = kEtkfWarps * 32;
    constexpr int PRE = (KP * kEtkfTileObs + NT - 1) / NT;      // staged elements per thread and tile
    extern __shared__ double etkf_tile[];                       // [KP][kEtkfLd]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[ACC][2];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    const int64_t j0 = (int64_t)blockIdx.x * chunk;
    const int64_t j1 = min(j0 + chunk, m_obs);
    T pre[PRE];
    auto fetch = [&](int64_t jb) {          // element e = row * 64 + obs: consecutive threads read consecutive observations
#pragma unroll
        for (int i = 0; i < PRE; ++i) {
            const int e = tid + i * NT;
            const int row = e / kEtkfTileObs, o = e - row * kEtkfTileObs;
            const int64_t j = jb + o;
            T v = (T)0;
            if (e < KP * kEtkfTileObs && j < j1) {
                if (row < k) v = yn[(int64_t)row * ld + j];
                else if (row == k) v = d[j];
            }
            pre[i] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < PRE; ++i) {
            const int e = tid + i * NT;
            const int row = e / kEtkfTileObs, o = e - row * kEtkfTileObs;
            if (e < KP * kEtkfTileObs) etkf_tile[row * kEtkfLd + o] = (double)pre[i];
        }
    };
    if (j0 < j1) fetch(j0);
    for (int64_t jb = j0; jb < j1; jb += kEtkfTileObs) {
        __syncthreads();                    // the previous tile has been consumed
        stash();
        __syncthreads();
        if (jb + kEtkfTileObs < j1) fetch(jb + kEtkfTileObs);
        const int nks = (int)min((int64_t)kEtkfTileObs, j1 - jb + 3) / 4;
        switch (warp) {
            case 0: etkf_gram_tile<KT, 0>(etkf_tile, nks, lane, acc); break;
            case 1: etkf_gram_tile<KT, 1>(etkf_tile, nks, lane, acc); break;
            case 2: etkf_gram_tile<KT, 2>(etkf_tile, nks, lane, acc); break;
            case 3: etkf_gram_tile<KT, 3>(etkf_tile, nks, lane, acc); break;
            case 4: etkf_gram_tile<KT, 4>(etkf_tile, nks, lane, acc); break;
            case 5: etkf_gram_tile<KT, 5>(etkf_tile, nks, lane, acc); break;
            case 6: etkf_gram_tile<KT, 6>(etkf_tile, nks, lane, acc); break;
            default: etkf_gram_tile<KT, 7>(etkf_tile, nks, lane, acc); break;
        }
    }
    double* out = partial + (size_t)blockIdx.x * KP * KP;
    int idx = 0;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            if (idx % kEtkfWarps == warp) {
                const int r = mt * 8 + (lane >> 2);
                const int c = nt * 8 + (lane & 3) * 2;
                out[r * KP + c] = acc[idx / kEtkfWarps][0];
                out[r * KP + c + 1] = acc[idx / kEtkfWarps][1];
            }
            ++idx;
        }
    }
}

// Sum the partial Grams in a fixed order into one tile-packed augmented Gram slot (common.cuh): the input format of the
// ensemble-space solve kernels.  One thread per entry (r, c), c <= r <= k; the slot is zeroed beforehand.
__global__ void k_etkf_reduce(const double* __restrict__ partial, int n_partial, int kp, int k, double* __restrict__ slot) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (k + 1) * kp) return;
    const int r = e / kp, c = e - r * kp;
    if (c > r || (c >= k && r != k)) return;                 // (k, k) = d d^T is kept for the kernelised ETKF (kernelise.cuh)
    double s = 0.0;
    for (int p = 0; p < n_partial; ++p) s += partial[(size_t)p * kp * kp + e];
    slot[sym_off(r, c)] = s;
}

// Sum the partial Grams in a fixed order into a dense (k+1) x (k+1) row-major lower triangle (row k = b, element (k, k) =
// d d^T, the upper triangle zero): the all-reduce payload of the observation-sharded global ETKF.
__global__ void k_etkf_reduce_dense(const double* __restrict__ partial, int n_partial, int kp, int k, double* __restrict__ gram) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (k + 1) * (k + 1)) return;
    const int r = e / (k + 1), c = e - r * (k + 1);
    double s = 0.0;
    if (c <= r)
        for (int p = 0; p < n_partial; ++p) s += partial[(size_t)p * kp * kp + r * kp + c];
    gram[e] = s;
}

// dense (k+1) x (k+1) lower triangle -> one kp x kp "partial" (the input format of the solve stage)
__global__ void k_etkf_dense_to_partial(const double* __restrict__ gram, int kp, int k, double* __restrict__ partial) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kp * kp) return;
    const int r = e / kp, c = e - r * kp;
    partial[e] = (r <= k && c <= r) ? gram[r * (k + 1) + c] : 0.0;
}

// Sum the partial Grams in a fixed order, eigendecompose, transform; W (k x k) row-major to global memory.
__global__ void __launch_bounds__(512, 1) k_etkf_solve(const double* __restrict__ partial, int n_partial, int kp, int k,
                                                    double rho, void* __restrict__ w_out, int io_f32) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const SolveSmem S = carve_solve_smem(smem_raw, k);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int x = tid; x < S.ne * S.lda; x += nt) S.A[x] = 0.0;
    if (tid < S.ne) S.bvec[tid] = 0.0;
    __syncthreads();
    for (int x = tid; x < (k + 1) * k; x += nt) {
        const int r = x / k, c = x % k;
        if (r < k && c > r) continue;
        double s = 0.0;
        for (int p = 0; p < n_partial; ++p) s += partial[(size_t)p * kp * kp + r * kp + c];
        if (r < k) { S.A[r * S.lda + c] = s; S.A[c * S.lda + r] = s; }
        else S.bvec[c] = s;
    }
    __syncthreads();
    jacobi_evd<2, 4, 4>(S.A, S.Vt, S.ne, S.lda, S.ldv, (double)(k - 1) / rho, tid, nt, 0);
    etkf_transform(S.A, S.Vt, S.bvec, S.vec, k, S.ne, S.lda, S.ldv, rho, tid, nt, 0);
    for (int x = tid; x < k * k; x += nt) st_io(w_out, x, S.A[(x / k) * S.lda + (x % k)], io_f32);
}

// Xa = mean + (X - mean) W.  One thread per grid point (coalesced along the grid axis), 8 output members per
// pass; W is staged in shared memory when it is global (per_grid = 0).
template <typename T, int JB>
__global__ void __launch_bounds__(128) k_apply_weights(const T* __restrict__ x, const T* __restrict__ w,
                                                       int per_grid, int k, int n_rows, int64_t n_grid, int64_t ld,
                                                       T* __restrict__ xa) {
    extern __shared__ double wsm_raw[];
    T* wsm = reinterpret_cast<T*>(wsm_raw);
    if (!per_grid) {
        for (int i = threadIdx.x; i < k * k; i += blockDim.x) wsm[i] = w[i];
        __syncthreads();
    }
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_grid) return;
    const T* wg = per_grid ? (w + gi * (int64_t)k * k) : wsm;
    for (int s = 0; s < n_rows; ++s) {
        const T* xs = x + (int64_t)s * k * ld + gi;
        T* xo = xa + (int64_t)s * k * ld + gi;
        double mean = 0.0;
        for (int i = 0; i < k; ++i) mean += (double)xs[(int64_t)i * ld];
        mean /= (double)k;
        for (int j0 = 0; j0 < k; j0 += JB) {
            double acc[JB];
#pragma unroll
            for (int j = 0; j < JB; ++j) acc[j] = 0.0;
            for (int i = 0; i < k; ++i) {
                const double p = (double)xs[(int64_t)i * ld] - mean;
                const T* wr = wg + i * k + j0;
#pragma unroll
                for (int j = 0; j < JB; ++j)
                    if (j0 + j < k) acc[j] = fma(p, (double)wr[j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < JB; ++j)
                if (j0 + j < k) xo[(int64_t)(j0 + j) * ld] = (T)(mean + acc[j]);
        }
    }
}

// Xa = mean + (X - mean) W with ONE global W (k x k): a streaming GEMM on the FP64 tensor pipe.
//   x_a[j, g] = sum_i x[i, g] W'[i][j],   W'[i][j] = W[i][j] + (1 - sum_i' W[i'][j]) / k
// (the ensemble mean folds into the weights: mean (1 - colsum_j) = sum_i x_i (1 - colsum_j) / k), i.e.
// Xa (k x N) = W'^T (k x k) X (k x N) for every state slice.  DMMA m8n8k4: A = W'^T fragments from shared memory (rows j,
// columns i), B = x fragments read straight from global memory in the reference layout (4 members x 8 consecutive grid
// points = four 64-byte segments), C = 8 members x 8 grid points.  One warp owns NTW * 8 consecutive grid points and all MT
// row tiles; every x element is read once and every x_a element written once (algorithmic HBM traffic), the FP64 work is
// 2 k^2 per state element.  n_grid columns are processed; ld is the row stride of x / xa (== n_grid for a whole state,
// the global N for a column shard of it); vec_ok: rows are 16-byte aligned, so the 2-element stores may be vectorised.
constexpr int kApplyWarps = 8;
// 8-point column tiles per warp: two while the accumulators (MT x NTW x 2 doubles) leave room for >= 2 CTAs per SM, one for
// large ensembles (k > 64: with two tiles the kernel needs 165 registers, one 256-thread CTA per SM, and the loads of X are
// not hidden: ncu long-scoreboard 4.1 per issue, DMMA pipe 67 %)
// (measured at cfg4: FP64 8.42 -> 8.26 ms with one tile and 2 CTAs per SM; FP32 plans keep two tiles: 8.6 vs 9.4 ms)
template <typename T> __host__ __device__ constexpr int apply_ntw(int mt) { return (sizeof(T) == 8 && mt > 8) ? 1 : 2; }
template <typename T> __host__ __device__ constexpr int apply_min_ctas(int mt) { return (apply_ntw<T>(mt) == 2 && mt > 8) ? 0 : 2; }   // 0: no occupancy request
__host__ __device__ inline int apply_lda(int k) {          // leading dimension of W'^T in shared memory: == 4 (mod 16) doubles
    int lda = (k + 3) & ~3;
    while ((lda & 15) != 4) lda += 4;
    return lda;
}
// Further destinations of the analysed columns: the same array in the memory of other GPUs of the node (peer pointers, already
// offset to this rank's first column).  The update kernel then IS the all-gather of the state-sharded global ETKF: every tile
// it produces goes to the local analysis and over NVLink to every peer, overlapped with the DMMA work of the next tile.
struct ApplyPeers { void* p[7]; int n; };
template <typename T, int MT>
__global__ void __launch_bounds__(kApplyWarps * 32, apply_min_ctas<T>(MT)) k_apply_global(const T* __restrict__ x, const T* __restrict__ w, int k, int n_rows,
                                                                 int64_t n_grid, int64_t ld, int vec_ok, T* __restrict__ xa,
                                                                 const ApplyPeers peers) {
    constexpr int kApplyNTW = apply_ntw<T>(MT);
    extern __shared__ double wsm_raw[];
    double* At = wsm_raw;                                   // [MT * 8][lda]: At[j][i] = W'[i][j]
    __shared__ double colsum[MT * 8];
    const int lda = apply_lda(k);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < MT * 8 * lda; e += blockDim.x) At[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < k * k; e += blockDim.x) {
        const int i = e / k, j = e - i * k;
        At[j * lda + i] = (double)w[e];
    }
    __syncthreads();
    for (int j = tid; j < k; j += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < k; ++i) s += At[j * lda + i];
        colsum[j] = (1.0 - s) / (double)k;
    }
    __syncthreads();
    for (int e = tid; e < k * k; e += blockDim.x) {
        const int j = e / k, i = e - j * k;
        At[j * lda + i] += colsum[j];
    }
    __syncthreads();
    const int ksteps = (k + 3) >> 2;
    const int r = lane >> 2, q = lane & 3;
    const int64_t n_chunks = (n_grid + kApplyNTW * 8 - 1) / (kApplyNTW * 8);
    for (int srow = 0; srow < n_rows; ++srow) {
        const T* xs = x + (int64_t)srow * k * ld;
        for (int64_t ch = (int64_t)blockIdx.x * kApplyWarps + warp; ch < n_chunks; ch += (int64_t)gridDim.x * kApplyWarps) {
            const int64_t g0 = ch * (kApplyNTW * 8);
            double acc[MT][kApplyNTW][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < kApplyNTW; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
#pragma unroll 4
            for (int ks = 0; ks < ksteps; ++ks) {
                const int i = ks * 4 + q;
                double b[kApplyNTW];
#pragma unroll
                for (int nt = 0; nt < kApplyNTW; ++nt) {
                    const int64_t g = g0 + nt * 8 + r;
                    b[nt] = (i < k && g < n_grid) ? (double)xs[(int64_t)i * ld + g] : 0.0;
                }
                const double* ap = At + r * lda + i;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const double a = ap[mt * 8 * lda];
#pragma unroll
                    for (int nt = 0; nt < kApplyNTW; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a, b[nt]);
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int j = mt * 8 + r;
#pragma unroll
                for (int nt = 0; nt < kApplyNTW; ++nt) {
                    const int64_t g = g0 + nt * 8 + q * 2;
                    if (j < k) {
                        const int64_t off = (int64_t)srow * k * ld + (int64_t)j * ld + g;
#pragma unroll
                        for (int pd = -1; pd < 7; ++pd) {     // unrolled: static indices into the kernel parameter (no local copy)
                            if (pd >= peers.n) break;
                            T* dst = (pd < 0 ? xa : static_cast<T*>(peers.p[pd < 0 ? 0 : pd])) + off;
                            if (g + 1 < n_grid && vec_ok) {
                                if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(dst) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                                else *reinterpret_cast<float2*>(dst) = make_float2((float)acc[mt][nt][0], (float)acc[mt][nt][1]);
                            } else {
                                if (g < n_grid) dst[0] = (T)acc[mt][nt][0];
                                if (g + 1 < n_grid) dst[1] = (T)acc[mt][nt][1];
                            }
                        }
                    }
                }
            }
        }
    }
}

// (n_rows, N) <-> dense (n_rows, n_cols) in block-sorted order: the all-gather payload of the grid-sharded run
template <typename T>
__global__ void k_pack_columns(const T* __restrict__ xa, const int* __restrict__ order, int64_t slot0, int64_t n_cols,
                               int n_rows, int64_t n_grid, T* __restrict__ packed, int64_t ld, int unpack) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= n_cols || r >= n_rows) return;
    const int64_t gi = order[slot0 + c];
    if (unpack) const_cast<T*>(xa)[(int64_t)r * n_grid + gi] = packed[(int64_t)r * ld + c];
    else packed[(int64_t)r * ld + c] = xa[(int64_t)r * n_grid + gi];
}

}  // namespace b200da
