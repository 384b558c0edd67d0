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

#ifndef B200DA_GRAM_UNROLL
#define B200DA_GRAM_UNROLL 2
#endif
constexpr int kTileObs = 64;      // observations per staged tile
// ring depth of the staged tiles: 3, or 2 when three FP64 tiles of a large ensemble would not fit in shared memory
template <typename T, int KPT> __host__ __device__ constexpr int gram_stages() { return (sizeof(T) == 8 && KPT > 14) ? 2 : 3; }
constexpr int kRing = 2048;       // survivor ring (sorted obs slots), power of two

struct LetkfParams {
    Geometry g;
    const Pos4* gpos;          // block-sorted grid positions (id = original index)
    const int* block_off;
    const Pos4* opos;          // cell-sorted obs positions
    const double* gext;        // [N][g.n_ext] extra coordinates of the block-sorted grid points (null when n_ext = 0)
    const double* oext;        // [M][g.n_ext] extra coordinates of the cell-sorted observations
    const int* cell_start;
    const void* ys;            // [M][KP] of the plan dtype
    const void* x;             // (n_slices, k, N) of the plan dtype
    void* xa;
    void* w_out;               // (N, k, k) or null
    unsigned long long* n_ambiguous;   // or null
    unsigned long long* stats;         // or null: [0] gram cycles [1] evd cycles [2] sweeps [3] evds [4] setup cycles [5] tiles
    double* cmat;                      // scratch: per grid slot the tile-packed augmented Gram (rows 0..k-1 = C, row k = b)
    int64_t slot_base;                 // first grid slot of the chunk held in cmat
    int64_t n_grid;
    int64_t n_obs;
    int block_begin;
    int k;
    int n_slices;
    double rho;
    double cut_pad;            // padded cutoff in bin space
    PlanStatus* status;        // plan-owned: error flags, number of recorded ambiguous pairs
    PairRec* amb_list;         // [kAmbCapacity] pairs inside the ambiguity band met by this launch (FP64 taper path)
    const PairRec* over;       // [n_over] host decisions that replace the device weight of a pair
    int n_over;
    TaperTab tt;               // tabulated taper (common.cuh); tt.coef null: direct evaluation
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

// ---- Gram tile: acc += (w .* Y_tile) Y_tile^T for the tiles owned by warp SUB of the grid point ------------------
// padded row length of the staged tile: conflict-free fragment reads (4 observations x 8 members per warp load)
template <typename T, int KT>
__host__ __device__ constexpr int tile_ld() {
    return sizeof(T) == 8 ? KT * 8 + 4 : KT * 8 + ((8 - (KT * 8) % 32 + 32) % 32);
}

// Ownership of the KT (KT + 1) / 2 lower-triangle tiles (row-major index idx) and of the KT column tiles of the extra rows
// among the WPG warps of a grid point: contiguous ranges, so a warp touches as few tile rows (A fragments f * w) as possible.
template <int KT, int WPG, int SUB>
__host__ __device__ constexpr bool gram_tile_owned(int idx) {
    return idx >= (KT * (KT + 1) / 2) * SUB / WPG && idx < (KT * (KT + 1) / 2) * (SUB + 1) / WPG;
}
template <int KT, int WPG, int SUB>
__host__ __device__ constexpr bool gram_col_owned(int t) { return t >= KT * SUB / WPG && t < KT * (SUB + 1) / WPG; }
template <int KT, int WPG> __host__ __device__ constexpr int gram_acc() { return (KT * (KT + 1) / 2 + WPG - 1) / WPG; }
template <int KT, int WPG> __host__ __device__ constexpr int gram_ecols() { return (KT + WPG - 1) / WPG; }

// ER ("extra rows"): the augmented matrix [Yn; d] has k + 1 = 8 KT + ER rows.  Rows 0 .. 8 KT - 1 are ensemble members and run
// on the DMMA pipe as KT (KT + 1) / 2 lower-triangle 8 x 8 tiles; the last ER rows (ER - 1 members and the innovation row d)
// would cost a whole extra row of KT + 1 tiles for ER of its 8 rows, so their ER (8 KT) + ER (ER + 1) / 2 dot products are
// accumulated with plain DFMAs instead (k = 50: 21 tiles + 3 rows instead of 28 tiles).  ER = 0: every row sits in a tile
// (k + 1 <= 8 KT; kernelised plans, which need d.d from the tiles' diagonal, and remainders of 4..7 rows).
//   eacc[e][n]  partial sums (over this lane's observations) of row 8 KT + e against column n-th owned column tile * 8 + lane / 4
//   pacc[p]     the pairs among the extra rows, p = e (e + 1) / 2 + e2, e2 <= e (warp SUB 0 only)
template <typename T, int KT, int WPG, int SUB, int ER>
__device__ __forceinline__ void gram_tile(const T* __restrict__ ytile, const double* __restrict__ wrow,
                                          double (&acc)[gram_acc<KT, WPG>()][2],
                                          double (&eacc)[ER > 0 ? ER : 1][gram_ecols<KT, WPG>()],
                                          double (&pacc)[ER > 0 ? ER * (ER + 1) / 2 : 1], int lane) {
    constexpr int LDY = tile_ld<T, KT + (ER > 0 ? 1 : 0)>();
    constexpr int kUnroll = B200DA_GRAM_UNROLL;
#pragma unroll kUnroll
    for (int ks = 0; ks < kTileObs / 4; ++ks) {
        const int j = ks * 4 + (lane & 3);
        const double w = wrow[j];
        if (__all_sync(0xffffffffu, ((__double2hiint(w) << 1) | __double2loint(w)) == 0)) continue;    // w == 0 on the integer pipe (the FP64 pipe is the busy one)
        const T* yr = ytile + j * LDY + (lane >> 2);
        double f[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) f[t] = (double)yr[t * 8];
        if constexpr (ER > 0) {
            double x[ER], wx[ER];
#pragma unroll
            for (int e = 0; e < ER; ++e) { x[e] = (double)ytile[j * LDY + KT * 8 + e]; wx[e] = w * x[e]; }
#pragma unroll
            for (int e = 0; e < ER; ++e) {
                int nb = 0;
#pragma unroll
                for (int t = 0; t < KT; ++t)
                    if (gram_col_owned<KT, WPG, SUB>(t)) { eacc[e][nb] = fma(f[t], wx[e], eacc[e][nb]); ++nb; }
            }
            if constexpr (SUB == 0) {
                int p = 0;
#pragma unroll
                for (int e = 0; e < ER; ++e)
#pragma unroll
                    for (int e2 = 0; e2 <= e; ++e2) { pacc[p] = fma(wx[e], x[e2], pacc[p]); ++p; }
            }
        }
        int idx = 0, n = 0;
#pragma unroll
        for (int mt = 0; mt < KT; ++mt) {
            const double fw = f[mt] * w;
#pragma unroll
            for (int nt = 0; nt <= mt; ++nt) {
                if (gram_tile_owned<KT, WPG, SUB>(idx)) { dmma884(acc[n][0], acc[n][1], fw, f[nt]); ++n; }
                ++idx;
            }
        }
    }
}

// accumulators -> global scratch in the tile-packed layout (common.cuh): one 512-byte tile per accumulator pair set,
// rows 0..k-1 = C (diagonal tiles hold the full 8x8 product), row k = b; the extra rows land in tile row KT
template <int KT, int WPG, int SUB, int ER>
__device__ __forceinline__ void dump_tiles_global(const double (&acc)[gram_acc<KT, WPG>()][2],
                                                  double (&eacc)[ER > 0 ? ER : 1][gram_ecols<KT, WPG>()],
                                                  double (&pacc)[ER > 0 ? ER * (ER + 1) / 2 : 1], double* __restrict__ C, int lane) {
    const int r = lane >> 2, c = (lane & 3) * 2;
    if constexpr (ER > 0) {                     // rows KT * 8 + e of the augmented matrix: rows e of tile row KT
#pragma unroll
        for (int e = 0; e < ER; ++e) {
            int nb = 0;
#pragma unroll
            for (int t = 0; t < KT; ++t) {
                if (gram_col_owned<KT, WPG, SUB>(t)) {
                    double v = eacc[e][nb];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if ((lane & 3) == 0) C[tile_off(KT, t) + tile_elem(e, r)] = v;
                    ++nb;
                }
            }
        }
        if constexpr (SUB == 0) {
            int p = 0;
#pragma unroll
            for (int e = 0; e < ER; ++e)
#pragma unroll
                for (int e2 = 0; e2 <= e; ++e2) {
                    double v = pacc[p];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (lane == 0) C[tile_off(KT, KT) + tile_elem(e, e2)] = v;
                    ++p;
                }
        }
    }
    int idx = 0, n = 0;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            if (gram_tile_owned<KT, WPG, SUB>(idx)) {
                *reinterpret_cast<double2*>(C + tile_off(mt, nt) + tile_elem(r, c)) = make_double2(acc[n][0], acc[n][1]);
                ++n;
            }
            ++idx;
        }
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
                            double cut_pad, int blk, PlanStatus* status) {
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
                        if (nr >= kMaxRuns) {         // never silently drop candidates: the host turns this into an error
                            if (status) atomicOr(&status->error, (unsigned)kStatusRunOverflow);
                            break;
                        }
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

template <typename T, int KT, int G, int WPG, int ER>
__host__ __device__ constexpr size_t gram_smem_bytes() {
    constexpr int KPT = KT + (ER > 0 ? 1 : 0), S = gram_stages<T, KPT>();
    return sizeof(double) * ((size_t)S * G * kTileObs) + sizeof(T) * ((size_t)S * kTileObs * tile_ld<T, KPT>());
}

// run-time warp index -> compile-time SUB (the accumulator arrays stay in registers only if every access is static)
#define B200DA_FOR_SUB(WPG_, sub_, CALL)                                                                            \
    do {                                                                                                             \
        if constexpr ((WPG_) == 1) { CALL(0); }                                                                      \
        else if constexpr ((WPG_) == 2) { if ((sub_) == 0) { CALL(0); } else { CALL(1); } }                          \
        else if constexpr ((WPG_) == 4) {                                                                            \
            switch (sub_) { case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; default: CALL(3); break; } \
        } else {                                                                                                     \
            switch (sub_) { case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; \
                            case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; default: CALL(7); break; } \
        }                                                                                                            \
    } while (0)

template <typename T, int KT, int G, int WPG, int ER>
__global__ void __launch_bounds__(G * WPG * 32, (KT <= 4 && G * WPG <= 8) ? 2 : 1) k_letkf_gram(const LetkfParams P) {
    constexpr int NT = G * WPG * 32;
    constexpr int KPT = KT + (ER > 0 ? 1 : 0);                // 8-row tiles of the staged [Yn; d] rows
    constexpr int kStages = gram_stages<T, KPT>();
    constexpr int KP = KPT * 8, LDY = tile_ld<T, KPT>();
    constexpr int ACC = gram_acc<KT, WPG>(), ECOLS = gram_ecols<KT, WPG>(), NP = ER > 0 ? ER * (ER + 1) / 2 : 1;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    BlockHeader<G>& H = *reinterpret_cast<BlockHeader<G>*>(smem_raw);
    unsigned char* work = smem_raw + ((sizeof(BlockHeader<G>) + 31) & ~size_t(31));
    double* wbuf = reinterpret_cast<double*>(work);                       // [S][G][TS]
    T* ybuf = reinterpret_cast<T*>(wbuf + (size_t)kStages * G * kTileObs);   // [S][TS][LDY]
    double* ttab = reinterpret_cast<double*>(work + gram_smem_bytes<T, KT, G, WPG, ER>());   // [nseg * nint][6] taper table
    const T* ys = reinterpret_cast<const T*>(P.ys);
    const bool use_tab = P.tt.coef != nullptr;
    if (use_tab) {                                   // made visible by the barriers of setup_block
        const int n2 = (int)(taper_tab_doubles(P.tt) >> 1);
        for (int i = threadIdx.x; i < n2; i += G * WPG * 32)
            reinterpret_cast<double2*>(ttab)[i] = reinterpret_cast<const double2*>(P.tt.coef)[i];
    }

    const Geometry& g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int my_g = warp / WPG, my_sub = warp % WPG;
    const int blk = P.block_begin + blockIdx.x;
    const long long t_start = clock64();
    setup_block<G>(H, g, P.gpos, P.block_off, P.cell_start, P.n_obs, P.cut_pad, blk, P.status);
    const int ng = H.ng;
    const long long t_setup = clock64();

    double acc[ACC][2];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double eacc[ER > 0 ? ER : 1][ECOLS];
#pragma unroll
    for (int e = 0; e < (ER > 0 ? ER : 1); ++e)
#pragma unroll
        for (int i = 0; i < ECOLS; ++i) eacc[e][i] = 0.0;
    double pacc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) pacc[i] = 0.0;

    // ------------------------------------------------------------------------------------------------------------
    // Gram phase
    // ------------------------------------------------------------------------------------------------------------
    const int cand_total = H.cand_total, n_runs = H.n_runs;
    const double bcx = H.cx, bcy = H.cy, bcz = H.cz;
    const double reach = (P.cut_pad + H.rb) * (1.0 + 1e-12);
    int cand_pos = 0, ring_head = 0, ring_tail = 0;
    int produced = 0, consumed = 0;
    unsigned long long my_amb = 0;

    // 1. refill the survivor ring (CTA-collective: two barriers per pass of NT candidates; needed once every few tiles)
    // The position of this thread's candidate of the NEXT refill pass is fetched one pass ahead (s_next, px / py / pz): the
    // L2 latency of that load (the one long wait of a pass: all warps sit in the pass together, the DMMA pipe idles) is hidden
    // behind the tiles computed in between.
    int s_next = -1;
    double px = 0.0, py = 0.0, pz = 0.0;
    auto fetch_next = [&]() {
        const int c = cand_pos + tid;
        s_next = -1;
        if (c < cand_total) {
            int lo_ = 0, hi_ = n_runs;            // run_pref[lo_] <= c < run_pref[hi_]
            while (hi_ - lo_ > 1) {
                const int mid = (lo_ + hi_) >> 1;
                if (H.run_pref[mid] <= c) lo_ = mid; else hi_ = mid;
            }
            s_next = H.run_start[lo_] + (c - H.run_pref[lo_]);
            const Pos4 po = P.opos[s_next];
            px = po.x; py = po.y; pz = po.z;
        }
    };
    fetch_next();
    auto refill = [&]() {
        while (ring_tail - ring_head < kTileObs && cand_pos < cand_total) {
            const int s = s_next;
            const bool keep = s >= 0 && bin_distance(g, bcx, bcy, bcz, px, py, pz) <= reach;
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
            fetch_next();
        }
    };
    // 2. + 3. this thread's share of the next tile (no barrier inside): localization weights, one (slot, grid point) pair per
    // thread and pass, and the asynchronous copy of the observation rows; padded slots re-read the first row with weight 0
    auto stage_tile = [&](int n_tile) {
        const int stage = produced % kStages;
        double* wst = wbuf + (size_t)stage * G * kTileObs;
        for (int q = tid; q < G * kTileObs; q += NT) {
            const int slot = q % kTileObs, gi = q / kTileObs;
            double w = 0.0;
            if (slot < n_tile && gi < ng) {
                const int so = H.ring[(ring_head + slot) & (kRing - 1)];
                const Pos4 po = P.opos[so];
                bool amb;
                const double* ge = P.gext + (size_t)(P.block_off[blk] + gi) * g.n_ext;
                const double* oe = P.oext + (size_t)so * g.n_ext;
                if (use_tab) w = pair_weight_tab(g, P.tt, ttab, H.gp[gi].x, H.gp[gi].y, H.gp[gi].z, po.x, po.y, po.z, amb);
                else w = pair_weight(g, H.gp[gi].x, H.gp[gi].y, H.gp[gi].z, po.x, po.y, po.z, ge, oe, amb);
                if (amb) {                                 // rare: hand the pair to the host for the reference's own decision
                    ++my_amb;
                    const unsigned long long at = atomicAdd(&P.status->amb_found, 1ull);
                    if (at < (unsigned long long)kAmbCapacity) {
                        PairRec rec; rec.gid = H.gp[gi].id; rec.oid = po.id;
                        rec.w = pair_weight_raw(g, H.gp[gi].x, H.gp[gi].y, H.gp[gi].z, po.x, po.y, po.z, ge, oe);
                        P.amb_list[at] = rec;
                    }
                }
                if (P.n_over > 0) w = apply_override(P.over, P.n_over, H.gp[gi].id, po.id, w);
            }
            wst[q] = w;
        }
        T* yst = ybuf + (size_t)stage * kTileObs * LDY;
        constexpr int EPC = 16 / sizeof(T);          // elements per 16-byte chunk
        constexpr int CH = KP / EPC;                 // 16-byte chunks per row
        for (int q = tid; q < kTileObs * CH; q += NT) {
            const int slot = q / CH, ch = q % CH;
            const int s = H.ring[(ring_head + (slot < n_tile ? slot : 0)) & (kRing - 1)];
            cp_async16(yst + slot * LDY + ch * EPC, ys + (size_t)s * KP + ch * EPC);
        }
        ring_head += n_tile;
        ++produced;
    };
    auto consume_tile = [&]() {
        if (my_g < ng) {
            const int stage = consumed % kStages;
            const T* yst = ybuf + (size_t)stage * kTileObs * LDY;
            const double* wrow = wbuf + ((size_t)stage * G + my_g) * kTileObs;
#define B200DA_GT(S) gram_tile<T, KT, WPG, S, ER>(yst, wrow, acc, eacc, pacc, lane)
            B200DA_FOR_SUB(WPG, my_sub, B200DA_GT);
#undef B200DA_GT
        }
    };
    // Half of the warps of every scheduler (warp id bit 2) stage the tile after next BEFORE their share of the tensor work
    // on the current tile, the other half AFTER it: right behind the per-tile barrier there is always a warp with DMMA work on
    // every scheduler while the others sit in the latency chain of the FP64 taper (global load -> asin -> polynomial).
    const bool stage_first = ((warp >> 2) & 1) == 0;

    if (cand_total > 0) {
#pragma unroll
        for (int i = 0; i < kStages - 1; ++i) {          // prologue: one cp.async group per tile, possibly empty
            refill();
            const int n_tile = min(kTileObs, ring_tail - ring_head);
            if (n_tile > 0) stage_tile(n_tile);
            cp_async_commit();
        }
        while (consumed < produced) {
            cp_async_wait<kStages - 2>();
            __syncthreads();
            refill();
            const int n_tile = min(kTileObs, ring_tail - ring_head);
            if (stage_first) {
                if (n_tile > 0) stage_tile(n_tile);
                cp_async_commit();                        // one group per iteration, possibly empty
                consume_tile();
            } else {
                consume_tile();
                if (n_tile > 0) stage_tile(n_tile);
                cp_async_commit();
            }
            ++consumed;
        }
        cp_async_wait<0>();
        if (P.n_ambiguous && my_amb) atomicAdd(P.n_ambiguous, my_amb);
    }
    __syncthreads();
    const long long t_gram = clock64();

    // ------------------------------------------------------------------------------------------------------------
    // hand the augmented Gram matrices to the solve kernel through the (L2-resident) scratch
    // ------------------------------------------------------------------------------------------------------------
    if (my_g < ng) {
        double* C = P.cmat + (size_t)((int64_t)P.block_off[blk] + my_g - P.slot_base) * (size_t)(tri_tiles(KPT) * 64);
#define B200DA_DT(S) dump_tiles_global<KT, WPG, S, ER>(acc, eacc, pacc, C, lane)
        B200DA_FOR_SUB(WPG, my_sub, B200DA_DT);
#undef B200DA_DT
    }
    if (P.stats && tid == 0) {
        const long long t_end = clock64();
        atomicAdd(P.stats + 0, (unsigned long long)(t_gram - t_setup));
        atomicAdd(P.stats + 6, (unsigned long long)(t_end - t_gram));
        atomicAdd(P.stats + 4, (unsigned long long)(t_setup - t_start));
        atomicAdd(P.stats + 5, (unsigned long long)produced);
    }
}

}  // namespace b200da
