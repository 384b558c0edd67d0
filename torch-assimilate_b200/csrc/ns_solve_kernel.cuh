// K3b — ETKF transform by a tensor-core matrix iteration, then the state update.
//
// The reference obtains P~a = (C + aI)^-1 and W_p = ((k-1) (C + aI)^-1)^(1/2), a = (k-1)/rho, from a symmetric
// eigendecomposition of C (pytassim/core/utils.py:26-93, core/etkf.py:57-77).  Only these two symmetric functions
// of C are used, so this kernel computes Z = (C + aI)^(-1/2) directly with the coupled Newton-Schulz iteration
//
//     M = Z Y,  T = (3 I - g M) / 2,  Z <- sqrt(g) T Z,  Y <- sqrt(g) Y T        (Y0 = A / s, Z0 = I, A = C + aI)
//
// (Higham, Functions of Matrices, eq. 6.35, with the scaling g chosen from a tracked spectral interval), which is
// nothing but k x k x k matrix products: they run on the FP64 tensor pipe (DMMA m8n8k4) instead of the
// shared-memory-bandwidth-bound rotations of a Jacobi sweep.  sqrt(g) is folded into the stored T, so the two products
// Z' = T' Z, Y' = Y T' are stored as they come out of the accumulators.  Then
//
//     w_mean = Z (Z b) / s,  W_p = sqrt((k-1) / s) Z,  W = w_mean 1^T + W_p     (core/etkf.py:72-77,102)
//     x_a = mean + (x - mean) W = mean + xp . w_mean + sqrt((k-1) / s) Z xp      (interface/base.py:257-278)
//
// W itself is only formed where it is exported (w_out).
//
// Spectrum.  C is a Gram matrix, so eig(A) lies in [a, s] with s = min(||A||_F, ||A||_inf); the reference's
// clamp(min=0) (core/utils.py:58) only removes rounding noise of the same size as this kernel's own rounding.
// With mu = eig(g M) in [g lo, g], g = 3 / (1 + sqrt(lo) + lo) equalises f(g lo) = f(g), f(mu) = mu (3 - mu)^2 / 4,
// which is the largest lower bound reachable in one step; lo <- f(g lo).  The iteration count follows from the
// bound alone (no data-dependent test): it stops when 1 - lo < 2e-8 (FP32 plans: 3e-4), i.e. when the next step leaves an error
// of order (1 - lo)^2 below the FP64 rounding level.
//
// Layout.  Every matrix of the iteration is a polynomial in A, hence symmetric: only the lower-triangle 8x8
// tiles are stored (tile (mt, nt), nt <= mt, at ((mt (mt+1))/2 + nt) * 64 doubles; element (r, c) of a tile at
// r * 8 + (c ^ ((r & 2) << 1)): the XOR makes both the direct and the transposed DMMA fragment reads
// bank-conflict free).  The Gram kernel writes its accumulators to global memory in exactly this layout, so
// loading a matrix is a flat copy.  One group of WPM warps owns one matrix (WPM = 1 up to k = 40: no CTA-wide
// barrier anywhere in the iteration; ns_launch.cu: NsPick); a CTA holds several groups, each fetching grid slots from an
// atomic counter.  A group has three matrix buffers; the roles (Z, Y, T) alternate from one matrix to the next, because the
// next matrix is copied (cp.async) into the buffer that is dead once the dense result of the current one is complete.
#pragma once
#include "plan.cuh"

namespace b200da {

__device__ __forceinline__ double sym_get(const double* __restrict__ S, int i, int j) {
    return i >= j ? S[sym_off(i, j)] : S[sym_off(j, i)];
}

__device__ __forceinline__ void dmma884_ns(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct FragOff { int d0, d1, t0, t1; };      // per-lane offsets inside a tile: direct / transposed, k-step 0 / 1
__device__ __forceinline__ FragOff make_frag_off(int lane) {
    const int r = lane >> 2, q = lane & 3;
    FragOff f;
    f.d0 = tile_elem(r, q); f.d1 = tile_elem(r, 4 + q);
    f.t0 = tile_elem(q, r); f.t1 = tile_elem(4 + q, r);
    return f;
}

template <int KT, int WPM>
struct NsCfg {
    static constexpr int NT = tri_tiles(KT);
    static constexpr int OWN = (NT + WPM - 1) / WPM;           // accumulator tiles per warp
    static constexpr int KP = KT * 8;
    static constexpr int MAT = NT * 64;                        // doubles per matrix
    static constexpr int LDW = KP + 1;                         // dense W overlay (odd: conflict-free rows and columns)
    static constexpr int VEC = 5 * KP + 8;                     // b (two buffers: the next matrix is prefetched), u, w_mean, xbuf (+ reduction scratch)
    static constexpr size_t GROUP_BYTES = sizeof(double) * (3 * (size_t)MAT + VEC);
    static_assert(2 * MAT >= KP * LDW, "dense overlay must fit in two tile-packed matrices");
};

template <int KT, int WPM, int SUB>
__device__ __forceinline__ constexpr bool tile_owned(int idx) {
    return idx >= tri_tiles(KT) * SUB / WPM && idx < tri_tiles(KT) * (SUB + 1) / WPM;
}

// rows [t*8, t*8+8) x columns [kt*8 + ks*4, +4) of the symmetric matrix S as a DMMA A-fragment (equally the
// B-fragment of S used as right operand, because S is symmetric): read directly from tile (t, kt) when kt <= t,
// transposed from tile (kt, t) otherwise.  t is a compile-time constant after unrolling, kt a loop variable.
template <int KS>
__device__ __forceinline__ double sym_frag(const double* __restrict__ S, int t, int kt, const FragOff& fo) {
    const int off = kt <= t ? tile_off(t, kt) + (KS ? fo.d1 : fo.d0) : tile_off(kt, t) + (KS ? fo.t1 : fo.t0);
    return S[off];
}

// Operand storage of the products below:
//   kOpSym    symmetric, tile-packed lower triangle, in shared memory (every matrix of the iteration itself)
//   kOpSymG   the same in the per-group global scratch (written earlier by this group: read through L2, never the
//             non-coherent path)
//   kOpFullG  general matrix as a full KT x KT grid of 8x8 tiles (tile (mt, nt) at (mt * KT + nt) * 64) in the global
//             scratch, used as RIGHT operand: fragment = rows [kt*8 + ks*4, +4) x columns [t*8, +8)
enum { kOpSym = 0, kOpSymG = 1, kOpFullG = 2 };
template <int MODE, int KS, int KT>
__device__ __forceinline__ double op_frag(const double* S, int t, int kt, const FragOff& fo) {
    if constexpr (MODE == kOpSym) return sym_frag<KS>(S, t, kt, fo);
    else if constexpr (MODE == kOpSymG) {
        const int off = kt <= t ? tile_off(t, kt) + (KS ? fo.d1 : fo.d0) : tile_off(kt, t) + (KS ? fo.t1 : fo.t0);
        return __ldcg(S + off);
    } else {
        return __ldcg(S + (kt * KT + t) * 64 + (KS ? fo.t1 : fo.t0));
    }
}

template <int KT, int WPM, int SUB, int KS, int LM = kOpSym, int RM = kOpSym>
__device__ __forceinline__ void gemm_kstep(const double* L, const double* R, int kt,
                                           const FragOff& fo, double (&acc)[NsCfg<KT, WPM>::OWN][2]) {
    double fl[KT], fr[KT];
#pragma unroll
    for (int t = 0; t < KT; ++t) {            // fragments of rows / columns this warp does not own are dead code
        fl[t] = op_frag<LM, KS, KT>(L, t, kt, fo);
        fr[t] = op_frag<RM, KS, KT>(R, t, kt, fo);
    }
    int idx = 0, n = 0;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            if (tile_owned<KT, WPM, SUB>(idx)) { dmma884_ns(acc[n][0], acc[n][1], fl[mt], fr[nt]); ++n; }
            ++idx;
        }
    }
}

// acc = lower-triangle tiles (owned by warp SUB of the group, diagonal tiles complete) of L R
template <int KT, int WPM, int SUB, int LM = kOpSym, int RM = kOpSym>
__device__ __forceinline__ void sym_gemm_sub(const double* L, const double* R, const FragOff& fo,
                                             double (&acc)[NsCfg<KT, WPM>::OWN][2]) {
#pragma unroll
    for (int i = 0; i < NsCfg<KT, WPM>::OWN; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    // small matrices: fully unrolled, every fragment offset a compile-time constant (the run-time selection between a tile and
    // its transpose was 14 % of the instructions at k = 40); larger ones keep the loop (code size, registers)
    constexpr int kUnrollKt = KT <= 6 ? KT : 1;
#pragma unroll kUnrollKt
    for (int kt = 0; kt < KT; ++kt) {
        gemm_kstep<KT, WPM, SUB, 0, LM, RM>(L, R, kt, fo, acc);
        gemm_kstep<KT, WPM, SUB, 1, LM, RM>(L, R, kt, fo, acc);
    }
}

template <int KT, int WPM, int LM = kOpSym, int RM = kOpSym>
__device__ __forceinline__ void sym_gemm(int sub, const double* L, const double* R,
                                         const FragOff& fo, double (&acc)[NsCfg<KT, WPM>::OWN][2]) {
    if constexpr (WPM == 1) sym_gemm_sub<KT, 1, 0, LM, RM>(L, R, fo, acc);
    else if constexpr (WPM == 2) {
        if (sub == 0) sym_gemm_sub<KT, 2, 0, LM, RM>(L, R, fo, acc); else sym_gemm_sub<KT, 2, 1, LM, RM>(L, R, fo, acc);
    } else {
        switch (sub) {
            case 0: sym_gemm_sub<KT, 4, 0, LM, RM>(L, R, fo, acc); break;
            case 1: sym_gemm_sub<KT, 4, 1, LM, RM>(L, R, fo, acc); break;
            case 2: sym_gemm_sub<KT, 4, 2, LM, RM>(L, R, fo, acc); break;
            default: sym_gemm_sub<KT, 4, 3, LM, RM>(L, R, fo, acc); break;
        }
    }
}

// The warp's lower-triangle accumulator tiles into a FULL tile grid in global memory: directly (tile (mt, nt), nt <= mt), or
// transposed into the strictly upper tiles (element (r, c) of tile (mt, nt) -> element (c, r) of tile (nt, mt), nt < mt).
// lower(L R) stored directly plus lower(R L) stored transposed is the complete product L R of two symmetric matrices.
template <int KT, int WPM, int SUB, bool TRANSPOSED>
__device__ __forceinline__ void store_full_sub(double* F, const double (&acc)[NsCfg<KT, WPM>::OWN][2], int lane) {
    const int r = lane >> 2, c = (lane & 3) * 2;
    int idx = 0, n = 0;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            if (tile_owned<KT, WPM, SUB>(idx)) {
                if constexpr (!TRANSPOSED) {
                    *reinterpret_cast<double2*>(F + (mt * KT + nt) * 64 + tile_elem(r, c)) = make_double2(acc[n][0], acc[n][1]);
                } else if (mt != nt) {
                    double* tp = F + (nt * KT + mt) * 64;
                    tp[tile_elem(c, r)] = acc[n][0];
                    tp[tile_elem(c + 1, r)] = acc[n][1];
                }
                ++n;
            }
            ++idx;
        }
    }
}
template <int KT, int WPM, bool TRANSPOSED>
__device__ __forceinline__ void store_full(int sub, double* F, const double (&acc)[NsCfg<KT, WPM>::OWN][2], int lane) {
    if constexpr (WPM == 1) store_full_sub<KT, 1, 0, TRANSPOSED>(F, acc, lane);
    else if constexpr (WPM == 2) {
        if (sub == 0) store_full_sub<KT, 2, 0, TRANSPOSED>(F, acc, lane); else store_full_sub<KT, 2, 1, TRANSPOSED>(F, acc, lane);
    } else {
        switch (sub) {
            case 0: store_full_sub<KT, 4, 0, TRANSPOSED>(F, acc, lane); break;
            case 1: store_full_sub<KT, 4, 1, TRANSPOSED>(F, acc, lane); break;
            case 2: store_full_sub<KT, 4, 2, TRANSPOSED>(F, acc, lane); break;
            default: store_full_sub<KT, 4, 3, TRANSPOSED>(F, acc, lane); break;
        }
    }
}

// S <- scale * acc + diag * I over the warp's tiles; diagonal tiles are written symmetrically from their lower half
// PLAIN: scale = 1, diag = 0 (no FP64 instruction: the DMULs of the scaled store share the datapath with the other warps' DMMAs)
template <int KT, int WPM, int SUB, bool PLAIN = false>
__device__ __forceinline__ void store_tiles_sub(double* __restrict__ S, const double (&acc)[NsCfg<KT, WPM>::OWN][2],
                                            double scale, double diag, int lane) {
    const int r = lane >> 2, c = (lane & 3) * 2;
    int idx = 0, n = 0;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            if (tile_owned<KT, WPM, SUB>(idx)) {
                double* tp = S + tile_off(mt, nt);
                const double v0 = PLAIN ? acc[n][0] : acc[n][0] * scale, v1 = PLAIN ? acc[n][1] : acc[n][1] * scale;
                if (mt != nt) {
                    *reinterpret_cast<double2*>(tp + tile_elem(r, c)) = make_double2(v0, v1);
                } else {                       // four predicated stores, no divergent branches
                    const double d0 = (!PLAIN && c == r) ? v0 + diag : v0, d1 = (!PLAIN && c + 1 == r) ? v1 + diag : v1;
                    if (c <= r) tp[tile_elem(r, c)] = d0;
                    if (c < r) tp[tile_elem(c, r)] = v0;
                    if (c + 1 <= r) tp[tile_elem(r, c + 1)] = d1;
                    if (c + 1 < r) tp[tile_elem(c + 1, r)] = v1;
                }
                ++n;
            }
            ++idx;
        }
    }
}

template <int KT, int WPM>
__device__ __forceinline__ void store_tiles(int sub, double* __restrict__ S, const double (&acc)[NsCfg<KT, WPM>::OWN][2],
                                            double scale, double diag, int lane) {
    if constexpr (WPM == 1) store_tiles_sub<KT, 1, 0>(S, acc, scale, diag, lane);
    else if constexpr (WPM == 2) {
        if (sub == 0) store_tiles_sub<KT, 2, 0>(S, acc, scale, diag, lane); else store_tiles_sub<KT, 2, 1>(S, acc, scale, diag, lane);
    } else {
        switch (sub) {
            case 0: store_tiles_sub<KT, 4, 0>(S, acc, scale, diag, lane); break;
            case 1: store_tiles_sub<KT, 4, 1>(S, acc, scale, diag, lane); break;
            case 2: store_tiles_sub<KT, 4, 2>(S, acc, scale, diag, lane); break;
            default: store_tiles_sub<KT, 4, 3>(S, acc, scale, diag, lane); break;
        }
    }
}

// f(mt, nt, e) for every element slot e in [0, 64) of the lower-triangle tiles of warp SUB of the group (tile index % WPM == SUB):
// 32 lanes x 2 passes per tile, (mt, nt) uniform over the warp and compile-time constants where the loops unroll (KT <= 6), so an
// element pass over a tile-packed matrix is a handful of instructions with constant offsets instead of an index reconstruction
// per element (the flat loops with their (i, j) stepping were 20 % of the samples of the k = 40 solve).
template <int KT, int WPM, int SUB, typename F>
__device__ __forceinline__ void for_tiles(int lane, F&& f) {
    constexpr int kU = KT <= 6 ? KT : 1;
    int idx = 0;
#pragma unroll kU
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll kU
        for (int nt = 0; nt <= mt; ++nt) {
            if (idx % WPM == SUB) { f(mt, nt, lane); f(mt, nt, lane + 32); }
            ++idx;
        }
    }
}

// sum_j row[j * stride] * v[j], j < k, in four independent chains (a single chain of k dependent DFMAs is latency bound with the
// two warps per scheduler this kernel runs at)
__device__ __forceinline__ double dot4(const double* __restrict__ row, int stride, const double* __restrict__ v, int k) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 3 < k; j += 4) {
        a0 = fma(row[j * stride], v[j], a0);
        a1 = fma(row[(j + 1) * stride], v[j + 1], a1);
        a2 = fma(row[(j + 2) * stride], v[j + 2], a2);
        a3 = fma(row[(j + 3) * stride], v[j + 3], a3);
    }
    for (; j < k; ++j) a0 = fma(row[j * stride], v[j], a0);
    return (a0 + a1) + (a2 + a3);
}

// the same value in every lane of every warp that calls it with the same arguments: sum_i f(i), i < k, lanes stride over i,
// xor-butterfly (commutative at every level, so all lanes end with identical bits)
template <typename F>
__device__ __forceinline__ double warp_sum_k(int lane, int k, F&& f) {
    double a = 0.0;
    for (int i = lane; i < k; i += 32) a += f(i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}

struct NsParams {
    const double* cmat;        // [n_slots][slot_stride]: tile-packed augmented Gram (rows 0..k-1 = C, row k = b)
    const Pos4* gpos;          // block-sorted grid positions (id = original index)
    const void* x;             // state / analysis / exported weights in the plan dtype (io_f32)
    void* xa;
    void* w_out;
    int io_f32;
    unsigned int* counter;     // zeroed before the launch: next slot to solve
    unsigned long long* stats; // or null: [1] solve cycles [2] iterations [3] solves
    int64_t slot_base;
    int64_t n_slots;
    int64_t n_grid;
    int64_t slot_stride;       // doubles per slot in cmat
    int k;
    int n_slices;
    double rho;
    double* scratch;           // per group (blockIdx.x * GROUPS + group) ns_scratch_doubles(KT) doubles: stiff matrices only
    double stiff;              // spectral-bound ratio s / a above which the two-level solve is used
    double conv;               // the iteration takes its last step when 1 - lo < conv (error afterwards ~ conv^2): 2e-8, FP32 plans 3e-4
};

// doubles of global scratch per group: fixed-up A and B (tile-packed symmetric) + one full tile grid
__host__ __device__ constexpr size_t ns_scratch_doubles(int kt) { return 2 * (size_t)tri_tiles(kt) * 64 + (size_t)kt * kt * 64; }

__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int WPM>
__device__ __forceinline__ void ns_sync(int bar_id) {
    if constexpr (WPM == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "n"(WPM * 32) : "memory");
}

// SUB (the warp's index inside its group) is a template parameter: with a run-time index every product call is a branch over
// WPM instantiations that all write the accumulator array, which the compiler then keeps in local memory (measured: ~700
// LDL/STL.128 in the two-warp kernel).
// Asynchronous load of one augmented Gram: flat copy of the lower-triangle tiles into `Yd`, b (row k) into `bd[0..k)`; the caller
// waits (cp_async_wait_all + group barrier) before it reads them.
template <int KT, int WPM>
__device__ __forceinline__ void ns_issue_load(const NsParams& P, int64_t slot, double* Yd, double* bd, int gtid) {
    constexpr int GT = WPM * 32, MAT = NsCfg<KT, WPM>::MAT;
    const double* gC = P.cmat + (size_t)slot * (size_t)P.slot_stride;
    const int k = P.k;
    for (int e = gtid * 2; e < MAT; e += GT * 2) cp_async16(Yd + e, gC + e);
    for (int c = gtid; c < k; c += GT) cp_async8(bd + c, gC + sym_off(k, c));
}

// One matrix: everything between "augmented Gram in shared memory (Y, bvec)" and "analysis columns written".  Buffers: the
// iteration runs in Z, Y, T; the dense result D overlays the two adjacent buffers that are not Z; once D is complete Z is dead, and
// the NEXT matrix of this group is fetched from the slot counter and copied into it (and its b into `bnext`) behind the rest of
// this one (matrix-vector products, refinement, update).  Returns that next slot.
template <int KT, int WPM, int SUB>
__device__ long long ns_solve_one(const NsParams& P, int64_t slot, double* Z, double* Y, double* T, double* D,
                                  double* bvec, double* bnext, double* vec, double* scratch, long long* slot_sh,
                                  int gtid, int sub, int lane, int bar_id) {
    using Cfg = NsCfg<KT, WPM>;
    constexpr int GT = WPM * 32, KP = Cfg::KP, MAT = Cfg::MAT, LDW = Cfg::LDW;
    const int k = P.k;
    double* uvec = vec + 2 * KP;        // [KP]   (vec + 0, vec + KP: the two b buffers)
    double* wbar = vec + 3 * KP;        // [KP]
    double* xbuf = vec + 4 * KP;        // [KP]
    double* red = vec + 5 * KP;         // [8] cross-warp reduction scratch
    const double* gC = P.cmat + (size_t)slot * (size_t)P.slot_stride;
    const double alpha = (double)(k - 1) / P.rho;
    const FragOff fo = make_frag_off(lane);

    // ---- fix-up: zero the padding (and the b row if it shares the last tile row), mirror diagonal tiles, add a I -------
    if ((k & 7) != 0) {
        constexpr int MT = KT - 1;
        for (int e = gtid; e < KT * 64; e += GT) {
            const int nt = e >> 6, r = (e >> 3) & 7, c = e & 7;
            if (MT * 8 + r >= k || nt * 8 + c >= k) Y[tile_off(MT, nt) + tile_elem(r, c)] = 0.0;
        }
        ns_sync<WPM>(bar_id);
    }
    for (int e = gtid; e < KT * 64; e += GT) {
        const int mt = e >> 6, r = (e >> 3) & 7, c = e & 7;
        double* tp = Y + tile_off(mt, mt);
        if (c > r) tp[tile_elem(r, c)] = tp[tile_elem(c, r)];
        else if (c == r && mt * 8 + r < k) tp[tile_elem(r, c)] += alpha;
    }
    ns_sync<WPM>(bar_id);
    // ---- s = min(||A||_F, ||A||_inf) -------------------------------------------------------------------------------
    double rs_max = 0.0, sq = 0.0;
    for (int i = gtid; i < k; i += GT) {            // row i over all KT tile columns (the padding is zero): tile (mt, nt) read directly
        const int mt = i >> 3, r = i & 7;           // for nt <= mt, tile (nt, mt) transposed beyond the diagonal
        const int xr = (r & 2) << 1, r4 = r ^ 4;
        double rs0 = 0.0, rs1 = 0.0, sq1 = 0.0;
        for (int nt = 0; nt < KT; ++nt) {
            const bool dir = nt <= mt;
            const double* tp = Y + (dir ? tile_off(mt, nt) + r * 8 : tile_off(nt, mt));
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const double a0 = tp[dir ? (c ^ xr) : c * 8 + ((c & 2) ? r4 : r)];
                const double a1 = tp[dir ? ((c + 1) ^ xr) : (c + 1) * 8 + ((c & 2) ? r4 : r)];
                rs0 += fabs(a0); sq = fma(a0, a0, sq);
                rs1 += fabs(a1); sq1 = fma(a1, a1, sq1);
            }
        }
        sq += sq1;
        rs_max = fmax(rs_max, rs0 + rs1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        rs_max = fmax(rs_max, __shfl_xor_sync(0xffffffffu, rs_max, off));
        sq += __shfl_xor_sync(0xffffffffu, sq, off);
    }
    if constexpr (WPM > 1) {
        if (lane == 0) { red[sub] = rs_max; red[4 + sub] = sq; }
        ns_sync<WPM>(bar_id);
        rs_max = red[0]; sq = red[4];
#pragma unroll
        for (int w = 1; w < WPM; ++w) { rs_max = fmax(rs_max, red[w]); sq += red[4 + w]; }
    }
    const double s = fmin(sqrt(sq), rs_max) * (1.0 + 1e-12);
    double acc[Cfg::OWN][2];
    // ---- Z <- ((Y + shift I) / sc)^(-1/2) for a symmetric Y whose shifted spectrum lies in [lo_abs, sc]; Y, T destroyed -------
    auto inv_sqrt = [&](const double shift, const double sc, const double lo_abs) -> int {
        const double inv_sc = 1.0 / sc;
        // iteration 0 (Z0 = I): T = (3 I - g Y0) / 2, Z1 = sqrt(g) T, Y1 = sqrt(g) Y0 T = Y0 Z1.  sqrt(g) is folded into the stored
        // T of every iteration (T' = sqrt(g) T), so that the products Z' = T' Z and Y' = Y T' are stored as they come
        double lo = fmin(lo_abs * inv_sc, 1.0);
        double g = 3.0 / (1.0 + sqrt(lo) + lo);
        double sg = sqrt(g);
        for_tiles<KT, WPM, SUB>(lane, [&](int mt, int nt, int e) {
            const int o = tile_off(mt, nt) + e;
            const int r = e >> 3, c = (e & 7) ^ ((r & 2) << 1);
            const bool on_diag = (mt == nt) && (r == c);
            double y = (Y[o] + (on_diag ? shift : 0.0)) * inv_sc;
            if (on_diag && mt * 8 + r >= k) y = 1.0;                 // padding: decoupled unit eigenvalues
            const double t = fma(-0.5 * g, y, on_diag ? 1.5 : 0.0);
            Y[o] = y; Z[o] = sg * t;
        });
        ns_sync<WPM>(bar_id);
        int iters = 1;
        // the scaling of the NEXT iteration depends on the tracked bound only: it is computed ahead of the products of the current
        // one, so that its divisions and square roots (dependent FP64 chains) interleave with the DMMA stream
        auto next_scaling = [&]() {
            const double m = g * lo;
            lo = fmin(1.0, 0.25 * m * (3.0 - m) * (3.0 - m));
            g = 3.0 / (1.0 + sqrt(lo) + lo);
            sg = sqrt(g);
        };
        {
            next_scaling();
            sym_gemm_sub<KT, WPM, SUB>(Y, Z, fo, acc);
            ns_sync<WPM>(bar_id);
            store_tiles_sub<KT, WPM, SUB, true>(Y, acc, 1.0, 0.0, lane);
            ns_sync<WPM>(bar_id);
        }
        for (; iters < 64; ++iters) {
            const bool last = (1.0 - lo) < P.conv;
            const double gc = g, sgc = sg;
            next_scaling();
            sym_gemm_sub<KT, WPM, SUB>(Z, Y, fo, acc);                    // M = Z Y
            store_tiles_sub<KT, WPM, SUB>(T, acc, -0.5 * gc * sgc, 1.5 * sgc, lane);   // T' = sqrt(g) (3 I - g M) / 2
            ns_sync<WPM>(bar_id);
            sym_gemm_sub<KT, WPM, SUB>(T, Z, fo, acc);                    // Z' = T' Z
            ns_sync<WPM>(bar_id);
            store_tiles_sub<KT, WPM, SUB, true>(Z, acc, 1.0, 0.0, lane);
            if (last) break;
            sym_gemm_sub<KT, WPM, SUB>(Y, T, fo, acc);                    // Y' = Y T'
            ns_sync<WPM>(bar_id);
            store_tiles_sub<KT, WPM, SUB, true>(Y, acc, 1.0, 0.0, lane);
            ns_sync<WPM>(bar_id);
        }
        ns_sync<WPM>(bar_id);
        return iters + 1;
    };
    // D = dense A^(-1/2) / dscale, over the two buffers that are not Z
    double dscale = 1.0;
    int iters;
    unsigned int next_u = 0;
    if constexpr (WPM == 1) { if (lane == 0) next_u = atomicAdd(P.counter, 1u); }     // the next slot: its latency hides behind the D pass
    else { if (gtid == 0) *slot_sh = (long long)atomicAdd(P.counter, 1u); }           // (visible after the barrier below)
    const bool stiff_path = !(scratch == nullptr || s <= P.stiff * alpha);
    if (!stiff_path) {
        // ---- one level: D = A^(-1/2) = Z / sqrt(s) ---------------------------------------------------------------------
        iters = inv_sqrt(0.0, s, alpha);
        dscale = sqrt(1.0 / s);
        for_tiles<KT, WPM, SUB>(lane, [&](int mt, int nt, int e) {       // tile-packed -> dense, both halves (padding included)
            const int r = e >> 3, c = (e & 7) ^ ((r & 2) << 1);
            const double v = Z[tile_off(mt, nt) + e];
            const int i = mt * 8 + r, j = nt * 8 + c;
            D[i * LDW + j] = v;
            if (mt != nt) D[j * LDW + i] = v;
        });
    } else {
        // ---- stiff matrix (s / a above P.stiff): two levels.  Every product of the iteration is stored as a symmetric matrix
        // (lower-triangle tiles); the rounding-level commutators this discards are amplified from one iteration to the next,
        // which is harmless up to ~8 iterations (s / a of a few thousand: error ~1e-12) and ruinous beyond (1e-7 at 1e4, no
        // correct digit at 1e5).  So the spectrum is split with an intermediate shift a1, a < a1 < s:
        //     B = (C + a1 I)^(-1/2)              spectrum ratio s / a1
        //     E = B A B                          a congruence: symmetric, the error of B enters multiplicatively, spectrum
        //                                        (l + a) / (l + a1) in [a / a1, 1)
        //     A^(-1/2) = sym(E^(-1/2) B)         B and E are functions of C, so the product is symmetric up to rounding
        // The error of B enters E relative to B's own conditioning, sqrt(s / a1), so the first level gets the smaller share of the
        // ratio: s / a1 = max(30, sqrt(s / a) / 2), a1 / a = the rest.  Measured error of A^(-1/2): < 1e-11 up to s / a = 1e5,
        // ~1e-10 at 5e6 (where the eigendecomposition route itself is no better).  A, B and one full-grid temporary live in
        // global scratch.
        double* gA = scratch;
        double* gB = gA + MAT;
        double* gF = gB + MAT;
        for (int e = gtid; e < MAT; e += GT) gA[e] = Y[e];
        ns_sync<WPM>(bar_id);            // the iteration's first pass rewrites Y tile-wise (another element-to-thread mapping)
        const double a1 = s / fmax(30.0, 0.5 * sqrt(s / alpha));
        const double shift = a1 - alpha;
        const double s1 = (s + shift) * (1.0 + 1e-12);
        iters = inv_sqrt(shift, s1, a1);
        const double zs1 = sqrt(1.0 / s1);
        for (int e = gtid; e < MAT; e += GT) gB[e] = Z[e] * zs1;
        ns_sync<WPM>(bar_id);
        sym_gemm_sub<KT, WPM, SUB, kOpSymG, kOpSymG>(gA, gB, fo, acc);    // W = A B: lower tiles, then the upper ones as lower(B A)^T
        store_full_sub<KT, WPM, SUB, false>(gF, acc, lane);
        sym_gemm_sub<KT, WPM, SUB, kOpSymG, kOpSymG>(gB, gA, fo, acc);
        store_full_sub<KT, WPM, SUB, true>(gF, acc, lane);
        ns_sync<WPM>(bar_id);
        sym_gemm_sub<KT, WPM, SUB, kOpSymG, kOpFullG>(gB, gF, fo, acc);   // E = B W (padding rows and columns are exactly 0)
        store_tiles_sub<KT, WPM, SUB>(Y, acc, 1.0, 0.0, lane);
        ns_sync<WPM>(bar_id);
        const double s2 = 1.0 + 1e-9;
        iters += inv_sqrt(0.0, s2, (alpha / a1) * (1.0 - 1e-9));
        sym_gemm_sub<KT, WPM, SUB, kOpSym, kOpSymG>(Z, gB, fo, acc);      // F = E^(-1/2) B, full, over W
        store_full_sub<KT, WPM, SUB, false>(gF, acc, lane);
        sym_gemm_sub<KT, WPM, SUB, kOpSymG, kOpSym>(gB, Z, fo, acc);
        store_full_sub<KT, WPM, SUB, true>(gF, acc, lane);
        ns_sync<WPM>(bar_id);
        const double zs = 0.5 * sqrt(1.0 / s2);
        for (int e = gtid, i = gtid / k, j = gtid - (gtid / k) * k; e < k * k; e += GT) {
            const double fij = __ldcg(gF + ((i >> 3) * KT + (j >> 3)) * 64 + tile_elem(i & 7, j & 7));
            const double fji = __ldcg(gF + ((j >> 3) * KT + (i >> 3)) * 64 + tile_elem(j & 7, i & 7));
            D[i * LDW + j] = (fij + fji) * zs;
            for (j += GT; j >= k; j -= k) ++i;
        }
    }
    if (P.stats && gtid == 0) { atomicAdd(P.stats + 2, (unsigned long long)iters); atomicAdd(P.stats + 3, 1ull); }
    ns_sync<WPM>(bar_id);
    // ---- Z is dead: the next matrix of this group starts to arrive in it ------------------------------------------------
    long long next;
    if constexpr (WPM == 1) next = (long long)__shfl_sync(0xffffffffu, next_u, 0);
    else next = *slot_sh;
    if (next < P.n_slots) ns_issue_load<KT, WPM>(P, next, Z, bnext, gtid);
    const double ds2 = dscale * dscale;
    for (int i = gtid; i < k; i += GT) uvec[i] = dot4(D + i * LDW, 1, bvec, k);         // u = D b
    ns_sync<WPM>(bar_id);
    for (int i = gtid; i < k; i += GT) wbar[i] = ds2 * dot4(D + i * LDW, 1, uvec, k);   // w_mean = D u = A^-1 b     core/etkf.py:72-73
    ns_sync<WPM>(bar_id);
    // ---- iterative refinement of w_mean against the Gram in global memory.  D carries an unstructured error e |D|; through
    // w_mean = D D b with |b| ~ lambda_max it becomes e * (lambda_max / a) in w_mean.  One residual step with the exact A
    // removes it (contraction e * s / a per step); two steps on the two-level path, none when s / a is small anyway.
    const int n_refine = stiff_path ? 2 : (s > 64.0 * alpha ? 1 : 0);
    for (int it = 0; it < n_refine; ++it) {
        for (int i = gtid; i < k; i += GT) {                      // r = b - A w_mean
            double a = fma(-alpha, wbar[i], bvec[i]);
            for (int j = 0; j < k; ++j) a = fma(-sym_get(gC, i, j), wbar[j], a);
            xbuf[i] = a;
        }
        ns_sync<WPM>(bar_id);
        for (int i = gtid; i < k; i += GT) uvec[i] = dot4(D + i * LDW, 1, xbuf, k);
        ns_sync<WPM>(bar_id);
        for (int i = gtid; i < k; i += GT) wbar[i] = fma(ds2, dot4(D + i * LDW, 1, uvec, k), wbar[i]);
        ns_sync<WPM>(bar_id);
    }
    // ---- W = w_mean 1^T + sqrt(k-1) D (core/etkf.py:75-76,102) is only formed where it is exported; the update applies its two
    // terms separately:  x_a[s, j, g] = mean + sum_i xp_i W[i][j] = mean + xp . w_mean + sqrt(k-1) (D xp)[j],  xp = x - mean
    // (interface/base.py:257-278)
    const double sk = sqrt((double)(k - 1)) * dscale;
    const int64_t gi = P.gpos[P.slot_base + slot].id;
    const int f32 = P.io_f32;
    if (P.w_out) {
        for (int i = 0; i < k; ++i) {
            const double wi = wbar[i];
            const int64_t row = (gi * (int64_t)k + i) * k;
            for (int j = gtid; j < k; j += GT) st_io(P.w_out, row + j, fma(sk, D[i * LDW + j], wi), f32);
        }
    }
    for (int sl = 0; sl < P.n_slices; ++sl) {
        const int64_t base = (int64_t)sl * k * P.n_grid + gi;
        for (int i = gtid; i < k; i += GT) xbuf[i] = ld_io(P.x, base + (int64_t)i * P.n_grid, f32);
        ns_sync<WPM>(bar_id);
        const double mean = warp_sum_k(lane, k, [&](int i) { return xbuf[i]; }) / (double)k;      // identical bits in every thread
        ns_sync<WPM>(bar_id);
        for (int i = gtid; i < k; i += GT) xbuf[i] -= mean;
        ns_sync<WPM>(bar_id);
        const double c0 = warp_sum_k(lane, k, [&](int i) { return xbuf[i] * wbar[i]; });
        for (int j = gtid; j < k; j += GT)
            st_io(P.xa, base + (int64_t)j * P.n_grid, mean + fma(sk, dot4(D + j, LDW, xbuf, k), c0), f32);
        ns_sync<WPM>(bar_id);
    }
    return next;
}

template <int KT, int WPM, int GROUPS>
__global__ void __launch_bounds__(GROUPS * WPM * 32, 1) k_letkf_solve_ns(const NsParams P) {
    using Cfg = NsCfg<KT, WPM>;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    __shared__ long long next_slot[GROUPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = warp / WPM, sub = warp % WPM, gtid = tid - group * WPM * 32;
    double* base = reinterpret_cast<double*>(smem_raw + (size_t)group * Cfg::GROUP_BYTES);
    double* b0 = base;                                   // three matrix buffers; the roles (Z, Y, T) alternate, see below
    double* b1 = base + Cfg::MAT;
    double* b2 = base + 2 * Cfg::MAT;
    double* vec = base + 3 * Cfg::MAT;
    double* bA = vec;
    double* bB = vec + Cfg::KP;
    const int bar_id = 1 + group;
    double* scratch = P.scratch ? P.scratch + ((size_t)blockIdx.x * GROUPS + group) * ns_scratch_doubles(KT) : nullptr;
    const long long t0 = clock64();
    for (int c = P.k + gtid; c < Cfg::KP; c += WPM * 32) { bA[c] = 0.0; bB[c] = 0.0; }      // the copies fill [0, k) only
    long long slot;
    if constexpr (WPM == 1) {
        unsigned int v = 0;
        if (lane == 0) v = atomicAdd(P.counter, 1u);
        slot = (long long)__shfl_sync(0xffffffffu, v, 0);
    } else {
        if (gtid == 0) next_slot[group] = (long long)atomicAdd(P.counter, 1u);
        ns_sync<WPM>(bar_id);
        slot = next_slot[group];
        ns_sync<WPM>(bar_id);
    }
    if (slot < P.n_slots) ns_issue_load<KT, WPM>(P, slot, b2, bA, gtid);
    // The matrix of a slot arrives in an END buffer (b2, then b0, b2, ...): Y = that buffer, Z = the other end, T = b1; the dense
    // result overlays Y and T, and the next matrix is copied into Z as soon as that result is complete.
    int odd = 0;
    while (slot < P.n_slots) {
        double* Y = odd ? b0 : b2;
        double* Z = odd ? b2 : b0;
        double* D = odd ? b0 : b1;
        double* bcur = odd ? bB : bA;
        double* bnext = odd ? bA : bB;
        cp_async_wait_all();
        ns_sync<WPM>(bar_id);
        if constexpr (WPM == 1) slot = ns_solve_one<KT, 1, 0>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id);
        else if constexpr (WPM == 2) {
            if (sub == 0) slot = ns_solve_one<KT, 2, 0>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id);
            else slot = ns_solve_one<KT, 2, 1>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id);
        } else {
            switch (sub) {
                case 0: slot = ns_solve_one<KT, 4, 0>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id); break;
                case 1: slot = ns_solve_one<KT, 4, 1>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id); break;
                case 2: slot = ns_solve_one<KT, 4, 2>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id); break;
                default: slot = ns_solve_one<KT, 4, 3>(P, slot, Z, Y, b1, D, bcur, bnext, vec, scratch, next_slot + group, gtid, sub, lane, bar_id); break;
            }
        }
        odd ^= 1;
    }
    if (P.stats && tid == 0) atomicAdd(P.stats + 1, (unsigned long long)(clock64() - t0));
}

}  // namespace b200da
