// K2-K4 — the fused LETKF kernel: one CTA analyses one block of up to G neighbouring grid points.
//
//   candidate sweep   cell runs around the block -> quick reject against the block's bounding sphere
//   staging           cp.async of the surviving observations' rows of the cell-sorted, obs-major [Yn; d]
//                     copy into a 3-stage shared-memory ring; Gaspari-Cohn weights w_gj evaluated per
//                     (grid point, observation) pair in FP64 while the copies are in flight
//   Gram              C_g = sum_j w_gj y_j y_j^T (lower-triangle 8x8 tiles) and b_g = sum_j w_gj d_j y_j
//                     (the d row rides along as row k of the augmented matrix) with DMMA m8n8k4,
//                     accumulators in registers, WPG warps per grid point
//   EVD               parallel cyclic Jacobi of C_g in shared memory
//   transform         w_mean = U L^-1 U^T b,  W_p = U ((k-1) L^-1)^(1/2) U^T,  W = w_mean 1^T + W_p
//   update            x_a = mean + (x - mean) W for every state slice, streamed back to HBM
//
// Reference semantics (paths relative to /root/reference):
//   localize_obs + sqrt(w) gather     pytassim/localization/gaspari_cohn.py:97-136, interface/wrapper.py:86-98
//   Gram / EVD / transform            pytassim/core/etkf.py:57-103, core/utils.py:26-93
//   update                            pytassim/interface/base.py:257-278
// (w y)(y)^T equals (sqrt(w) y)(sqrt(w) y)^T of the reference up to rounding.
#pragma once
#include "plan.cuh"

namespace b200da {

constexpr int kTileObs = 64;      // observations per staged tile
constexpr int kStages = 3;
constexpr int kRing = 2048;       // survivor ring (sorted obs slots), power of two
constexpr int kMaxSweeps = 40;

struct LetkfParams {
    Geometry g;
    const Pos4* gpos;          // block-sorted grid positions (id = original index)
    const int* block_off;
    const Pos4* opos;          // cell-sorted obs positions
    const int* cell_start;
    const double* ys;          // [M][KP]
    const double* x;           // (n_slices, k, N)
    double* xa;
    double* w_out;             // (N, k, k) or null
    unsigned long long* n_ambiguous;   // or null
    unsigned long long* stats;         // or null: [0] gram cycles [1] evd cycles [2] sweeps [3] evds [4] setup cycles [5] tiles
    int64_t n_grid;
    int64_t n_obs;
    int block_begin;
    int k;
    int n_slices;
    int evd_conc;              // EVDs resident in shared memory at once (power of two, <= G)
    double rho;
    double cut_pad;            // padded cutoff in bin space
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// ---- Gram tile: acc += (w .* Y_tile) Y_tile^T for the tiles owned by warp SUB of the grid point ------------------
template <int KT, int WPG, int SUB>
__device__ __forceinline__ void gram_tile(const double* __restrict__ ytile, const double* __restrict__ wrow,
                                          double (&acc)[(KT * (KT + 1) / 2 + WPG - 1) / WPG][2], int lane) {
    constexpr int LDY = KT * 8 + 4;
#pragma unroll 2
    for (int ks = 0; ks < kTileObs / 4; ++ks) {
        const int j = ks * 4 + (lane & 3);
        const double w = wrow[j];
        if (__all_sync(0xffffffffu, w == 0.0)) continue;
        const double* yr = ytile + j * LDY + (lane >> 2);
        double f[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) f[t] = yr[t * 8];
        int idx = 0, n = 0;
#pragma unroll
        for (int mt = 0; mt < KT; ++mt) {
            const double fw = f[mt] * w;
#pragma unroll
            for (int nt = 0; nt <= mt; ++nt) {
                if (idx % WPG == SUB) { dmma884(acc[n][0], acc[n][1], fw, f[nt]); ++n; }
                ++idx;
            }
        }
    }
}

// accumulators -> symmetric A (k x k, leading dimension lda) and b (row k of the augmented Gram)
template <int KT, int WPG, int SUB>
__device__ __forceinline__ void dump_tiles(const double (&acc)[(KT * (KT + 1) / 2 + WPG - 1) / WPG][2],
                                           double* __restrict__ A, int lda, double* __restrict__ bvec, int k, int lane) {
    int idx = 0, n = 0;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            if (idx % WPG == SUB) {
                const int r = mt * 8 + (lane >> 2);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = nt * 8 + (lane & 3) * 2 + e;
                    const double v = acc[n][e];
                    if (r < k && c <= r) { A[r * lda + c] = v; A[c * lda + r] = v; }
                    else if (r == k && c < k) bvec[c] = v;
                }
                ++n;
            }
            ++idx;
        }
    }
}

// ---- parallel cyclic Jacobi (two-sided, round-robin ordering) on a group of threads ---------------------------------
// A (k x k symmetric, full storage) is diagonalised in place, V accumulates the rotations (columns = eigenvectors).
// Rotations are skipped when |a_pq| <= tol * sqrt((a_pp + shift)(a_qq + shift)): the criterion of a relative-accuracy
// Jacobi on A + shift*I, which is the matrix whose functions the transform needs (core/utils.py:58-60).
struct JacobiScratch {
    double* cs;     // [n2][2]
    int* pq;        // [n2][2]
    int* flag;      // [2]
};

__device__ int jacobi_evd(double* __restrict__ A, double* __restrict__ V, int k, int lda, double shift,
                          const JacobiScratch sc, int gtid, int gthreads, int bar_id) {
    const int ne = (k + 1) & ~1;
    const int n2 = ne >> 1;
    const double tol = 1e-15;
    for (int i = gtid; i < k * k; i += gthreads) {
        const int r = i / k, c = i % k;
        V[r * lda + c] = (r == c) ? 1.0 : 0.0;
    }
    if (gtid == 0) { sc.flag[0] = 0; sc.flag[1] = 0; }
    group_barrier(bar_id, gthreads);
    if (k < 2) return 0;
    int sweep = 0;
    for (; sweep < kMaxSweeps; ++sweep) {
        for (int step = 0; step < ne - 1; ++step) {
            // phase 1: rotation parameters of the n2 disjoint pairs
            if (gtid < n2) {
                int a, b;
                if (gtid == 0) { a = ne - 1; b = step; }
                else { a = (step + gtid) % (ne - 1); b = (step - gtid + (ne - 1)) % (ne - 1); }
                const int p = min(a, b), q = max(a, b);
                double c = 1.0, s = 0.0;
                if (q < k) {
                    const double apq = A[p * lda + q];
                    const double app = A[p * lda + p], aqq = A[q * lda + q];
                    if (fabs(apq) > tol * sqrt(fabs((app + shift) * (aqq + shift)))) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
                        c = rsqrt(fma(t, t, 1.0));
                        s = t * c;
                        sc.flag[sweep & 1] = 1;
                    }
                }
                sc.pq[2 * gtid] = p; sc.pq[2 * gtid + 1] = (q < k) ? q : -1;
                sc.cs[2 * gtid] = c; sc.cs[2 * gtid + 1] = s;
            }
            group_barrier(bar_id, gthreads);
            // phase 2: A <- J^T A J on independent 2x2 blocks (lower triangle of pair-pairs, mirrored), V <- V J
            const int nblk = n2 * (n2 + 1) / 2;
            for (int x = gtid; x < nblk; x += gthreads) {
                int ti = (int)((sqrtf(8.0f * (float)x + 1.0f) - 1.0f) * 0.5f);
                while (ti * (ti + 1) / 2 > x) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= x) ++ti;
                const int tj = x - ti * (ti + 1) / 2;
                const int pi = sc.pq[2 * ti], qi = sc.pq[2 * ti + 1];
                const int pj = sc.pq[2 * tj], qj = sc.pq[2 * tj + 1];
                const double ci = sc.cs[2 * ti], si = sc.cs[2 * ti + 1];
                const double cj = sc.cs[2 * tj], sj = sc.cs[2 * tj + 1];
                if (qi < 0 && qj < 0) continue;
                // rows (pi, qi) x cols (pj, qj); a dummy partner (q < 0) has identity rotation and no storage
                const double b00 = A[pi * lda + pj];
                const double b01 = qj >= 0 ? A[pi * lda + qj] : 0.0;
                const double b10 = qi >= 0 ? A[qi * lda + pj] : 0.0;
                const double b11 = (qi >= 0 && qj >= 0) ? A[qi * lda + qj] : 0.0;
                // left: rows' = J_i^T rows
                const double r00 = ci * b00 - si * b10, r01 = ci * b01 - si * b11;
                const double r10 = si * b00 + ci * b10, r11 = si * b01 + ci * b11;
                // right: cols' = cols J_j
                double n00 = cj * r00 - sj * r01, n01 = sj * r00 + cj * r01;
                double n10 = cj * r10 - sj * r11, n11 = sj * r10 + cj * r11;
                if (ti == tj) { n01 = 0.0; n10 = 0.0; }
                A[pi * lda + pj] = n00; A[pj * lda + pi] = n00;
                if (qj >= 0) { A[pi * lda + qj] = n01; A[qj * lda + pi] = n01; }
                if (qi >= 0) { A[qi * lda + pj] = n10; A[pj * lda + qi] = n10; }
                if (qi >= 0 && qj >= 0) { A[qi * lda + qj] = n11; A[qj * lda + qi] = n11; }
            }
            for (int x = gtid; x < n2 * k; x += gthreads) {
                const int t = x / k, r = x % k;
                const int p = sc.pq[2 * t], q = sc.pq[2 * t + 1];
                if (q < 0) continue;
                const double c = sc.cs[2 * t], s = sc.cs[2 * t + 1];
                const double vp = V[r * lda + p], vq = V[r * lda + q];
                V[r * lda + p] = c * vp - s * vq;
                V[r * lda + q] = s * vp + c * vq;
            }
            group_barrier(bar_id, gthreads);
        }
        const int rotated = sc.flag[sweep & 1];
        if (gtid == 0) sc.flag[(sweep + 1) & 1] = 0;
        group_barrier(bar_id, gthreads);
        if (!rotated) break;
    }
    return sweep + 1;
}

// ---- transform: A (diagonalised), V, b -> W = w_mean 1^T + W_p written over V ---------------------------------------
// core/utils.py:58-60 (clamp, + (k-1)/rho, reciprocal), core/etkf.py:70-77,102.
__device__ void etkf_transform(double* __restrict__ A, double* __restrict__ V, const double* __restrict__ bvec,
                               double* __restrict__ vec, int k, int lda, double rho, int gtid, int gthreads, int bar_id) {
    double* inv = vec;            // [k] 1 / (max(lambda, 0) + (k-1)/rho)
    double* z = vec + k;          // [k]
    double* wbar = vec + 2 * k;   // [k]
    const double reg = (double)(k - 1) / rho;
    for (int m = gtid; m < k; m += gthreads) {
        const double ev = fmax(A[m * lda + m], 0.0) + reg;
        const double iv = 1.0 / ev;
        inv[m] = iv;
        double acc = 0.0;                                   // z = L^-1 U^T b
        for (int i = 0; i < k; ++i) acc = fma(V[i * lda + m], bvec[i], acc);
        z[m] = acc * iv;
    }
    group_barrier(bar_id, gthreads);
    for (int i = gtid; i < k; i += gthreads) {              // w_mean = U z
        double acc = 0.0;
        for (int m = 0; m < k; ++m) acc = fma(V[i * lda + m], z[m], acc);
        wbar[i] = acc;
    }
    // B = U diag(((k-1) inv)^(1/4)) so that W_p = B B^T; B overwrites A
    for (int x = gtid; x < k * k; x += gthreads) {
        const int i = x / k, m = x % k;
        A[i * lda + m] = V[i * lda + m] * sqrt(sqrt((double)(k - 1) * inv[m]));
    }
    group_barrier(bar_id, gthreads);
    for (int x = gtid; x < k * (k + 1) / 2; x += gthreads) {
        int i = (int)((sqrtf(8.0f * (float)x + 1.0f) - 1.0f) * 0.5f);
        while (i * (i + 1) / 2 > x) --i;
        while ((i + 1) * (i + 2) / 2 <= x) ++i;
        const int j = x - i * (i + 1) / 2;
        double acc = 0.0;
        for (int m = 0; m < k; ++m) acc = fma(A[i * lda + m], A[j * lda + m], acc);
        V[i * lda + j] = acc + wbar[i];                     // W[i][j] = w_mean[i] + W_p[i][j]  (core/etkf.py:102)
        if (i != j) V[j * lda + i] = acc + wbar[j];
    }
    group_barrier(bar_id, gthreads);
}

// ---- update: x_a[s, j, g] = mean + sum_i (x[s, i, g] - mean) W[i][j]  (interface/base.py:257-278) ------------------
__device__ void apply_point(const double* __restrict__ W, int lda, int k, int n_slices, int64_t n_grid, int64_t gi,
                            const double* __restrict__ x, double* __restrict__ xa, double* __restrict__ w_out,
                            double* __restrict__ xbuf, int gtid, int gthreads, int bar_id) {
    if (w_out) {
        double* dst = w_out + gi * (int64_t)k * k;
        for (int i = gtid; i < k * k; i += gthreads) dst[i] = W[(i / k) * lda + (i % k)];
    }
    for (int s = 0; s < n_slices; ++s) {
        const double* xs = x + (int64_t)s * k * n_grid + gi;
        double* xo = xa + (int64_t)s * k * n_grid + gi;
        for (int i = gtid; i < k; i += gthreads) xbuf[i] = xs[(int64_t)i * n_grid];
        group_barrier(bar_id, gthreads);
        double mean = 0.0;
        for (int i = 0; i < k; ++i) mean += xbuf[i];       // same order for every thread
        mean /= (double)k;
        for (int j = gtid; j < k; j += gthreads) {
            double acc = 0.0;
            for (int i = 0; i < k; ++i) acc = fma(xbuf[i] - mean, W[i * lda + j], acc);
            xo[(int64_t)j * n_grid] = mean + acc;
        }
        group_barrier(bar_id, gthreads);
    }
}

// ---- block header kept in shared memory for the whole kernel ---------------------------------------------------------
template <int G>
struct BlockHeader {
    Pos4 gp[G];
    double cx, cy, cz, rb;
    int ng, n_runs, cand_total, pad;
    int warp_counts[32];
    int run_start[kMaxRuns];
    int run_pref[kMaxRuns + 1];
    int ring[kRing];
};

// Load the block's grid points, its bounding sphere and the list of candidate cell runs into the header.
template <int G>
__device__ void setup_block(BlockHeader<G>& H, const Geometry& g, const Pos4* __restrict__ gpos,
                            const int* __restrict__ block_off, const int* __restrict__ cell_start, int64_t n_obs,
                            double cut_pad, int blk) {
    const int tid = threadIdx.x;
    const int slot0 = block_off[blk];
    const int ng = block_off[blk + 1] - slot0;
    if (tid < G) {
        Pos4 p; p.x = 0; p.y = 0; p.z = 0; p.id = -1;
        if (tid < ng) p = gpos[slot0 + tid];
        H.gp[tid] = p;
    }
    __syncthreads();
    if (tid == 0) {
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = 0; i < ng; ++i) {
            const double v[3] = {H.gp[i].x, H.gp[i].y, H.gp[i].z};
            for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], v[d]); mx[d] = fmax(mx[d], v[d]); }
        }
        const double cx = 0.5 * (mn[0] + mx[0]), cy = 0.5 * (mn[1] + mx[1]), cz = 0.5 * (mn[2] + mx[2]);
        double rb = 0.0;
        for (int i = 0; i < ng; ++i) rb = fmax(rb, bin_distance(g, cx, cy, cz, H.gp[i].x, H.gp[i].y, H.gp[i].z));
        H.cx = cx; H.cy = cy; H.cz = cz; H.rb = rb * (1.0 + 1e-12);
        H.ng = ng;
        // candidate cell columns
        int lo[3], hi[3];
        bool empty = (n_obs == 0);
        for (int d = 0; d < 3; ++d) {
            if (g.periodic && d == 2) { lo[d] = 0; hi[d] = 0; continue; }
            const double a = floor((mn[d] - cut_pad - g.org[d]) / g.h[d]);
            const double b = floor((mx[d] + cut_pad - g.org[d]) / g.h[d]);
            lo[d] = (int)fmax(a, 0.0);
            hi[d] = (int)fmin(b, (double)(g.nc[d] - 1));
            if (g.nc[d] == 1) { lo[d] = 0; hi[d] = 0; }
            if (lo[d] > hi[d]) empty = true;
        }
        int zr[2][2]; int nz = 1;
        if (g.periodic) {
            const double a = floor((mn[2] - cut_pad) / g.h[2]);
            const double b = floor((mx[2] + cut_pad) / g.h[2]);
            if (b - a + 1.0 >= (double)g.nc[2]) { zr[0][0] = 0; zr[0][1] = g.nc[2] - 1; }
            else {
                int z0 = (int)fmod(a, (double)g.nc[2]); if (z0 < 0) z0 += g.nc[2];
                int z1 = (int)fmod(b, (double)g.nc[2]); if (z1 < 0) z1 += g.nc[2];
                if (z0 <= z1) { zr[0][0] = z0; zr[0][1] = z1; }
                else { zr[0][0] = z0; zr[0][1] = g.nc[2] - 1; zr[1][0] = 0; zr[1][1] = z1; nz = 2; }
            }
        } else { zr[0][0] = lo[2]; zr[0][1] = hi[2]; }
        int nr = 0, total = 0;
        H.run_pref[0] = 0;
        if (!empty) {
            for (int cx_ = lo[0]; cx_ <= hi[0]; ++cx_)
                for (int cy_ = lo[1]; cy_ <= hi[1]; ++cy_)
                    for (int q = 0; q < nz; ++q) {
                        if (nr >= kMaxRuns) break;
                        const int base = (cx_ * g.nc[1] + cy_) * g.nc[2];
                        const int s = cell_start[base + zr[q][0]];
                        const int e = cell_start[base + zr[q][1] + 1];
                        if (e > s) {
                            H.run_start[nr] = s; total += e - s; ++nr; H.run_pref[nr] = total;
                        }
                    }
        }
        H.n_runs = nr; H.cand_total = total;
    }
    __syncthreads();

}

template <int KT, int G, int WPG>
constexpr size_t gram_smem_bytes() {
    return sizeof(double) * ((size_t)kStages * kTileObs * (KT * 8 + 4) + (size_t)kStages * G * kTileObs);
}
__host__ __device__ inline size_t evd_smem_bytes_per_matrix(int k) {
    const int lda = k | 1, n2 = (k + 1) / 2;
    return sizeof(double) * ((size_t)2 * k * lda + 5 * (size_t)k + 2 * (size_t)n2 + 2) + sizeof(int) * (2 * (size_t)n2 + 4);
}

template <int KT, int G, int WPG>
__global__ void __launch_bounds__(G * WPG * 32, 1) k_letkf_fused(const LetkfParams P) {
    constexpr int NT = G * WPG * 32;
    constexpr int KP = KT * 8, LDY = KP + 4;
    constexpr int NTILES = KT * (KT + 1) / 2;
    constexpr int ACC = (NTILES + WPG - 1) / WPG;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    BlockHeader<G>& H = *reinterpret_cast<BlockHeader<G>*>(smem_raw);
    unsigned char* work = smem_raw + ((sizeof(BlockHeader<G>) + 31) & ~size_t(31));
    double* ybuf = reinterpret_cast<double*>(work);                       // [S][TS][LDY]
    double* wbuf = ybuf + (size_t)kStages * kTileObs * LDY;               // [S][G][TS]

    const Geometry& g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int my_g = warp / WPG, my_sub = warp % WPG;
    const int blk = P.block_begin + blockIdx.x;
    const long long t_start = clock64();
    setup_block<G>(H, g, P.gpos, P.block_off, P.cell_start, P.n_obs, P.cut_pad, blk);
    const int ng = H.ng;
    const long long t_setup = clock64();

    double acc[ACC][2];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }

    // ------------------------------------------------------------------------------------------------------------
    // Gram phase
    // ------------------------------------------------------------------------------------------------------------
    const int cand_total = H.cand_total, n_runs = H.n_runs;
    const double bcx = H.cx, bcy = H.cy, bcz = H.cz;
    const double reach = (P.cut_pad + H.rb) * (1.0 + 1e-12);
    int cand_pos = 0, ring_head = 0, ring_tail = 0;
    int produced = 0, consumed = 0;
    unsigned long long my_amb = 0;

    auto produce = [&]() -> bool {      // stage one more tile if any observation is left; uniform across the CTA
        // 1. refill the survivor ring
        while (ring_tail - ring_head < kTileObs && cand_pos < cand_total) {
            const int c = cand_pos + tid;
            bool keep = false;
            int s = 0;
            if (c < cand_total) {
                int lo_ = 0, hi_ = n_runs;            // run_pref[lo_] <= c < run_pref[hi_]
                while (hi_ - lo_ > 1) {
                    const int mid = (lo_ + hi_) >> 1;
                    if (H.run_pref[mid] <= c) lo_ = mid; else hi_ = mid;
                }
                s = H.run_start[lo_] + (c - H.run_pref[lo_]);
                const Pos4 po = P.opos[s];
                keep = bin_distance(g, bcx, bcy, bcz, po.x, po.y, po.z) <= reach;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) H.warp_counts[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                const int cnt = H.warp_counts[w];
                if (w < warp) before += cnt;
                total += cnt;
            }
            if (keep) H.ring[(ring_tail + before + __popc(bal & ((1u << lane) - 1u))) & (kRing - 1)] = s;
            __syncthreads();
            ring_tail += total;
            cand_pos += NT;
        }
        const int n_tile = min(kTileObs, ring_tail - ring_head);
        if (n_tile <= 0) return false;
        const int stage = produced % kStages;
        // 2. localization weights of the tile: one (slot, grid point) pair per thread and pass
        double* wst = wbuf + (size_t)stage * G * kTileObs;
        for (int q = tid; q < G * kTileObs; q += NT) {
            const int slot = q % kTileObs, gi = q / kTileObs;
            double w = 0.0;
            if (slot < n_tile && gi < ng) {
                const Pos4 po = P.opos[H.ring[(ring_head + slot) & (kRing - 1)]];
                bool amb;
                w = pair_weight(g, H.gp[gi].x, H.gp[gi].y, H.gp[gi].z, po.x, po.y, po.z, amb);
                if (amb) ++my_amb;
            }
            wst[q] = w;
        }
        // 3. asynchronous copy of the observation rows; padded slots re-read the first row with weight 0
        double* yst = ybuf + (size_t)stage * kTileObs * LDY;
        constexpr int CH = KP / 2;                   // 16-byte chunks per row
        for (int q = tid; q < kTileObs * CH; q += NT) {
            const int slot = q / CH, ch = q % CH;
            const int s = H.ring[(ring_head + (slot < n_tile ? slot : 0)) & (kRing - 1)];
            cp_async16(yst + slot * LDY + ch * 2, P.ys + (size_t)s * KP + ch * 2);
        }
        ring_head += n_tile;
        ++produced;
        return true;
    };

    if (cand_total > 0) {
        if (produce()) { cp_async_commit(); if (produce()) {} cp_async_commit(); }
        else { cp_async_commit(); cp_async_commit(); }
        while (consumed < produced) {
            cp_async_wait<1>();
            __syncthreads();
            produce();
            cp_async_commit();                        // one group per iteration, possibly empty
            if (my_g < ng) {
                const int stage = consumed % kStages;
                const double* yst = ybuf + (size_t)stage * kTileObs * LDY;
                const double* wrow = wbuf + ((size_t)stage * G + my_g) * kTileObs;
                if constexpr (WPG == 1) gram_tile<KT, WPG, 0>(yst, wrow, acc, lane);
                else if constexpr (WPG == 2) {
                    if (my_sub == 0) gram_tile<KT, WPG, 0>(yst, wrow, acc, lane);
                    else gram_tile<KT, WPG, 1>(yst, wrow, acc, lane);
                } else if constexpr (WPG == 4) {
                    switch (my_sub) {
                        case 0: gram_tile<KT, WPG, 0>(yst, wrow, acc, lane); break;
                        case 1: gram_tile<KT, WPG, 1>(yst, wrow, acc, lane); break;
                        case 2: gram_tile<KT, WPG, 2>(yst, wrow, acc, lane); break;
                        default: gram_tile<KT, WPG, 3>(yst, wrow, acc, lane); break;
                    }
                } else {
                    switch (my_sub) {
                        case 0: gram_tile<KT, WPG, 0>(yst, wrow, acc, lane); break;
                        case 1: gram_tile<KT, WPG, 1>(yst, wrow, acc, lane); break;
                        case 2: gram_tile<KT, WPG, 2>(yst, wrow, acc, lane); break;
                        case 3: gram_tile<KT, WPG, 3>(yst, wrow, acc, lane); break;
                        case 4: gram_tile<KT, WPG, 4>(yst, wrow, acc, lane); break;
                        case 5: gram_tile<KT, WPG, 5>(yst, wrow, acc, lane); break;
                        case 6: gram_tile<KT, WPG, 6>(yst, wrow, acc, lane); break;
                        default: gram_tile<KT, WPG, 7>(yst, wrow, acc, lane); break;
                    }
                }
            }
            ++consumed;
        }
        cp_async_wait<0>();
        if (P.n_ambiguous && my_amb) atomicAdd(P.n_ambiguous, my_amb);
    }
    __syncthreads();
    const long long t_gram = clock64();

    // ------------------------------------------------------------------------------------------------------------
    // EVD + transform + update, evd_conc grid points at a time (the shared memory of the Gram phase is reused)
    // ------------------------------------------------------------------------------------------------------------
    const int k = P.k, lda = k | 1, n2 = (k + 1) / 2;
    const int E = P.evd_conc;
    const int gthreads = NT / E;
    const int grp = tid / gthreads, gtid = tid % gthreads;
    const size_t per = (evd_smem_bytes_per_matrix(k) + 31) & ~size_t(31);
    unsigned char* mine = work + (size_t)grp * per;
    double* A = reinterpret_cast<double*>(mine);
    double* V = A + (size_t)k * lda;
    double* bvec = V + (size_t)k * lda;          // [k]
    double* vec = bvec + k;                      // [3k] inv, z, wbar
    double* xbuf = vec + 3 * k;                  // [k]
    JacobiScratch sc;
    sc.cs = xbuf + k;                            // [2 n2]
    sc.pq = reinterpret_cast<int*>(sc.cs + 2 * n2 + 2);
    sc.flag = sc.pq + 2 * n2;
    const double shift = (double)(k - 1) / P.rho;

    for (int round = 0; round * E < ng; ++round) {
        // warps owning a grid point of this round dump their tiles into that group's A / b
        if (my_g >= round * E && my_g < (round + 1) * E && my_g < ng) {
            unsigned char* dst = work + (size_t)(my_g - round * E) * per;
            double* Ad = reinterpret_cast<double*>(dst);
            double* bd = Ad + (size_t)2 * k * lda;
            if constexpr (WPG == 1) dump_tiles<KT, WPG, 0>(acc, Ad, lda, bd, k, lane);
            else if constexpr (WPG == 2) {
                if (my_sub == 0) dump_tiles<KT, WPG, 0>(acc, Ad, lda, bd, k, lane);
                else dump_tiles<KT, WPG, 1>(acc, Ad, lda, bd, k, lane);
            } else if constexpr (WPG == 4) {
                switch (my_sub) {
                    case 0: dump_tiles<KT, WPG, 0>(acc, Ad, lda, bd, k, lane); break;
                    case 1: dump_tiles<KT, WPG, 1>(acc, Ad, lda, bd, k, lane); break;
                    case 2: dump_tiles<KT, WPG, 2>(acc, Ad, lda, bd, k, lane); break;
                    default: dump_tiles<KT, WPG, 3>(acc, Ad, lda, bd, k, lane); break;
                }
            } else {
                switch (my_sub) {
                    case 0: dump_tiles<KT, WPG, 0>(acc, Ad, lda, bd, k, lane); break;
                    case 1: dump_tiles<KT, WPG, 1>(acc, Ad, lda, bd, k, lane); break;
                    case 2: dump_tiles<KT, WPG, 2>(acc, Ad, lda, bd, k, lane); break;
                    case 3: dump_tiles<KT, WPG, 3>(acc, Ad, lda, bd, k, lane); break;
                    case 4: dump_tiles<KT, WPG, 4>(acc, Ad, lda, bd, k, lane); break;
                    case 5: dump_tiles<KT, WPG, 5>(acc, Ad, lda, bd, k, lane); break;
                    case 6: dump_tiles<KT, WPG, 6>(acc, Ad, lda, bd, k, lane); break;
                    default: dump_tiles<KT, WPG, 7>(acc, Ad, lda, bd, k, lane); break;
                }
            }
        }
        __syncthreads();
        const int gp_idx = round * E + grp;
        if (gp_idx < ng) {
            const int bar_id = 1 + grp;
            const int nsw = jacobi_evd(A, V, k, lda, shift, sc, gtid, gthreads, bar_id);
            if (P.stats && gtid == 0) { atomicAdd(P.stats + 2, (unsigned long long)nsw); atomicAdd(P.stats + 3, 1ull); }
            etkf_transform(A, V, bvec, vec, k, lda, P.rho, gtid, gthreads, bar_id);
            apply_point(V, lda, k, P.n_slices, P.n_grid, H.gp[gp_idx].id, P.x, P.xa, P.w_out, xbuf, gtid, gthreads,
                        bar_id);
        }
        __syncthreads();
    }
    if (P.stats && tid == 0) {
        const long long t_end = clock64();
        atomicAdd(P.stats + 0, (unsigned long long)(t_gram - t_setup));
        atomicAdd(P.stats + 1, (unsigned long long)(t_end - t_gram));
        atomicAdd(P.stats + 4, (unsigned long long)(t_setup - t_start));
        atomicAdd(P.stats + 5, (unsigned long long)produced);
    }
}

}  // namespace b200da
