// Iterative ensemble Kalman smoother (IEnKS) weight update on top of the augmented Gram: transform and bundle variants.
// Included by b200da.cu only.
//
// Reference semantics restated here (paths relative to /root/reference):
//   IEnKSTransformModule._update_weights / forward     pytassim/core/ienks.py:117-151
//   IEnKSBundleModule._get_dh_dw                        pytassim/core/ienks.py:168-174
//   svd / rev_svd / matrix_product / diagonal_add       pytassim/core/utils.py:96-199
//   localized call with incoming weights (args_to_skip) pytassim/interface/lienks.py:68-118, interface/wrapper.py:86-98
//
// With the incoming weights W (k x k) of a grid point, w = rowmean(W - I) (ienks.py:50-53), Wp = W - w 1^T, T = Wp^-1
// (:62-65: the SVD there is only used for the inverse and for (Wp Wp^T)^-1 = T^T T), S = T (transform, :74-75) or I / eps
// (bundle, :173), and C = Yn Yn^T, b = Yn d^T from the Gram kernel:
//     dh_dw dh_dw^T = S C S^T,   dh_dw d^T = S b                                                 (:74-75, :84-88, :96)
//     grad = (k - 1) w - S b                                                                     (:84-88)
//     P    = (1 - tau) (k - 1) T^T T + tau (S C S^T + (k - 1) I)                                 (:96-98)
//     W'   = w - tau P^-1 grad + ((k - 1) P^-1)^(1/2)                                            (:99-103, :130-131, :149)
// P = A' + a' I with A' = (1 - tau)(k - 1) T^T T + tau S C S^T (positive semi-definite) and a' = tau (k - 1): exactly the
// problem the ensemble-space solve kernels handle with inflation 1 / tau, and w - tau P^-1 grad = P^-1 b' with
// b' = P w - tau grad.  k_ienks_pre rewrites a Gram slot in place into (A', b'); the solve kernels then run unchanged
// (W' = P^-1 b' 1^T + sqrt(k - 1) P^-1/2, then the update x_mean + X' W').  Grid points without local observations keep
// their incoming weights (ienks.py:143: the update only runs for p > 0): k_ienks_pre flags them and k_ienks_keep restores
// W and the state columns after the solve.
#pragma once
#include "common.cuh"

namespace b200da {

struct IenksParams {
    double* cmat;              // [n_slots][slot_stride] tile-packed augmented Gram, rewritten in place
    int64_t n_slots;
    int64_t slot_stride;
    const Pos4* gpos;          // block-sorted grid positions (id = original index)
    int64_t slot_base;
    const void* w_in;          // incoming weights in the plan dtype: (N, k, k) when per_grid else (k, k)
    int per_grid;
    int io_f32;
    int k;
    double tau;                // learning rate in (0, 1]
    double eps;                // > 0: bundle variant (S = I / eps); <= 0: transform variant (S = T)
    int* empty;                // [n_slots] out: 1 when the grid point has no local observation
    // k_ienks_keep only
    const void* x;
    void* xa;
    void* w_out;
    int n_slices;
    int64_t n_grid;
};

__host__ __device__ inline size_t ienks_smem_bytes(int k) {
    return sizeof(double) * (3 * (size_t)k * (k + 1) + 7 * (size_t)k) + sizeof(int) * (size_t)k + 64;
}

// One CTA per slot (grid-stride), 256 threads = 8 warps: warp `ty` walks rows ty, ty + 8, ..., lane `tx` walks columns tx,
// tx + 32, ... (no integer division in the inner loops).  Shared memory: three k x (k + 1) matrices, seven k-vectors and the
// pivot list; the odd leading dimension keeps column accesses conflict-free.
__global__ void __launch_bounds__(256) k_ienks_pre(const IenksParams P) {
    extern __shared__ double ism[];
    const int k = P.k, ld = k + 1;
    double* A = ism;                   // Wp, inverted in place: T = Wp^-1
    double* G = A + (size_t)k * ld;    // Gram, then S C S^T
    double* M = G + (size_t)k * ld;    // temporary, then A'
    double* wbar = M + (size_t)k * ld;
    double* bv = wbar + k;
    double* colv = bv + k;
    double* rowa = colv + k;
    double* rowc = rowa + k;
    double* sb = rowc + k;
    double* bout = sb + k;
    int* pivs = reinterpret_cast<int*>(bout + k);
    const int tid = threadIdx.x, nt = blockDim.x, tx = tid & 31, ty = tid >> 5, nw = nt >> 5;
    const bool bundle = P.eps > 0.0;
    const double km1 = (double)(k - 1);
    for (int64_t s = blockIdx.x; s < P.n_slots; s += gridDim.x) {
        double* C = P.cmat + (size_t)s * (size_t)P.slot_stride;
        int nonzero = 0;
        for (int i = tid; i < k; i += nt) nonzero |= (C[sym_off(i, i)] != 0.0) | (C[sym_off(k, i)] != 0.0);
        nonzero = __syncthreads_or(nonzero);
        if (tid == 0) P.empty[s] = nonzero ? 0 : 1;
        if (!nonzero) continue;
        const int64_t gi = P.gpos[P.slot_base + s].id;
        const int64_t wbase = P.per_grid ? gi * (int64_t)k * k : 0;
        for (int r = ty; r < k; r += nw)
            for (int j = tx; j < k; j += 32) {
                A[r * ld + j] = ld_io(P.w_in, wbase + r * k + j, P.io_f32);
                G[r * ld + j] = j <= r ? C[sym_off(r, j)] : C[sym_off(j, r)];
            }
        for (int i = tid; i < k; i += nt) bv[i] = C[sym_off(k, i)];
        __syncthreads();
        for (int i = tid; i < k; i += nt) {                          // w = mean over columns of (W - I)        ienks.py:50-51
            double sum = 0.0;
            for (int j = 0; j < k; ++j) sum += A[i * ld + j] - (i == j ? 1.0 : 0.0);
            wbar[i] = sum / (double)k;
        }
        __syncthreads();
        for (int r = ty; r < k; r += nw) {                           // Wp = W - w 1^T                           ienks.py:52
            const double wr = wbar[r];
            for (int j = tx; j < k; j += 32) A[r * ld + j] -= wr;
        }
        __syncthreads();
        // ---- T = Wp^-1: in-place Gauss-Jordan elimination with partial (row) pivoting; the row swaps are undone as column
        // swaps in reverse order at the end ((Q Wp)^-1 = Wp^-1 Q^T).  Every warp finds the pivot itself (same data, same
        // result: no barrier and no idle warps), lanes own columns tx, tx + 32, tx + 64.                          ienks.py:62-65
        const bool h0 = tx < k, h1 = tx + 32 < k, h2 = tx + 64 < k;
        const bool any1 = k > 32, any2 = k > 64;                     // uniform: skip whole column groups
        for (int c = 0; c < k; ++c) {
            // pivot: the largest |A[r][c]|, r >= c, compared on the upper 32 bits of the doubles (monotone for non-negative
            // values; a pivot within 2^-20 of the largest is as good), one redux per warp instead of a shuffle ladder
            unsigned int key = 0u; int pr = c;
            for (int r = c + tx; r < k; r += 32) {
                const unsigned int kr = (unsigned int)__double2hiint(fabs(A[r * ld + c]));
                if (kr > key || r == c + tx) { key = kr; pr = r; }
            }
            const unsigned int kmax = __reduce_max_sync(0xffffffffu, key);
            const unsigned int who = __ballot_sync(0xffffffffu, key == kmax && c + tx < k);
            pr = __shfl_sync(0xffffffffu, pr, __ffs(who) - 1);
            if (tid < k) {                                           // k <= 96 < 256: one pass, the other warps skip the division
                const int j = tid;
                const double pinv = 1.0 / A[pr * ld + c];
                rowa[j] = (j == c ? 1.0 : A[pr * ld + j]) * pinv;   // the new row c: the scaled pivot row, unit column folded in
                rowc[j] = A[c * ld + j];                             // the old row c moves to the pivot's place
                colv[j] = A[j * ld + c];                             // column c before the elimination
                if (j == 0) pivs[c] = pr;
            }
            __syncthreads();
            const double ra0 = h0 ? rowa[tx] : 0.0, ra1 = h1 ? rowa[tx + 32] : 0.0, ra2 = h2 ? rowa[tx + 64] : 0.0;
            for (int r = ty; r < k; r += nw) {
                double* Ar = A + r * ld;
                if (r == c) {
                    if (h0) Ar[tx] = ra0;
                    if (h1) Ar[tx + 32] = ra1;
                    if (h2) Ar[tx + 64] = ra2;
                    continue;
                }
                const bool moved = (r == pr);                        // pr > c here: this row now holds the old row c
                const double* src = moved ? rowc : Ar;
                const double f = moved ? rowc[c] : colv[r];
                if (h0) Ar[tx] = fma(-f, ra0, tx == c ? 0.0 : src[tx]);
                if (any1 && h1) Ar[tx + 32] = fma(-f, ra1, tx + 32 == c ? 0.0 : src[tx + 32]);
                if (any2 && h2) Ar[tx + 64] = fma(-f, ra2, tx + 64 == c ? 0.0 : src[tx + 64]);
            }
            __syncthreads();
        }
        for (int r = tid; r < k; r += nt)                            // every thread un-permutes the columns of its own row
            for (int c = k - 1; c >= 0; --c) {
                const int p = pivs[c];
                if (p != c) { const double t = A[r * ld + c]; A[r * ld + c] = A[r * ld + p]; A[r * ld + p] = t; }
            }
        __syncthreads();
        // k x k products: a warp takes two rows at a time, a lane up to three columns: five shared-memory loads for six FMAs
        //   out[r][j] = sum_l L(r, l) * R(l, j)
#define B200DA_IENKS_PRODUCT(LEXPR0, LEXPR1, REXPR, STORE)                                                        \
        for (int r0 = ty; r0 < k; r0 += 2 * nw) {                                                                     \
            const int r1 = r0 + nw;                                                                                   \
            const bool two = r1 < k;                                                                                  \
            const int r1c = two ? r1 : r0;                                                                            \
            double a00 = 0.0, a01 = 0.0, a02 = 0.0, a10 = 0.0, a11 = 0.0, a12 = 0.0;                                  \
            for (int l = 0; l < k; ++l) {                                                                             \
                const double x0 = LEXPR0, x1 = LEXPR1;                                                                \
                if (h0) { const int j = tx; const double y = REXPR; a00 = fma(x0, y, a00); a10 = fma(x1, y, a10); }       \
                if (h1) { const int j = tx + 32; const double y = REXPR; a01 = fma(x0, y, a01); a11 = fma(x1, y, a11); }  \
                if (h2) { const int j = tx + 64; const double y = REXPR; a02 = fma(x0, y, a02); a12 = fma(x1, y, a12); }  \
            }                                                                                                         \
            if (h0) { const int j = tx; { const int r = r0; const double v = a00; STORE; } if (two) { const int r = r1; const double v = a10; STORE; } }        \
            if (h1) { const int j = tx + 32; { const int r = r0; const double v = a01; STORE; } if (two) { const int r = r1; const double v = a11; STORE; } }   \
            if (h2) { const int j = tx + 64; { const int r = r0; const double v = a02; STORE; } if (two) { const int r = r1; const double v = a12; STORE; } }   \
        }
        if (bundle) {                                                // dh_dw = Yn / eps                                   :173
            const double ie = 1.0 / P.eps;
            for (int r = ty; r < k; r += nw)
                for (int j = tx; j < k; j += 32) G[r * ld + j] = (G[r * ld + j] * ie) * ie;
            for (int i = tid; i < k; i += nt) sb[i] = bv[i] * ie;
            __syncthreads();
        } else {                                                     // dh_dw = T Yn: S C S^T = (T C) T^T                   :74-75
            B200DA_IENKS_PRODUCT(A[r0 * ld + l], A[r1c * ld + l], G[l * ld + j], M[r * ld + j] = v)
            for (int i = tid; i < k; i += nt) {
                double acc = 0.0;
                for (int l = 0; l < k; ++l) acc = fma(A[i * ld + l], bv[l], acc);
                sb[i] = acc;
            }
            __syncthreads();
            B200DA_IENKS_PRODUCT(M[r0 * ld + l], M[r1c * ld + l], A[j * ld + l], G[r * ld + j] = v)
            __syncthreads();
        }
        // ---- A' = (1 - tau)(k - 1) T^T T + tau S C S^T                                                         :62-66, :96-98
        const double c_old = (1.0 - P.tau) * km1;
        if (c_old != 0.0) {
            B200DA_IENKS_PRODUCT(A[l * ld + r0], A[l * ld + r1c], A[l * ld + j], M[r * ld + j] = fma(c_old, v, P.tau * G[r * ld + j]))
        } else {
            for (int r = ty; r < k; r += nw)
                for (int j = tx; j < k; j += 32) M[r * ld + j] = P.tau * G[r * ld + j];
        }
#undef B200DA_IENKS_PRODUCT
        __syncthreads();
        // ---- b' = P w - tau grad,  grad = (k - 1) w - S b,  P = A' + tau (k - 1) I                            :84-88, :130-131
        for (int i = tid; i < k; i += nt) {
            double acc = P.tau * km1 * wbar[i];
            for (int j = 0; j < k; ++j) acc = fma(0.5 * (M[i * ld + j] + M[j * ld + i]), wbar[j], acc);
            bout[i] = acc - P.tau * (km1 * wbar[i] - sb[i]);
        }
        for (int r = ty; r < k; r += nw)
            for (int j = tx; j <= r; j += 32) C[sym_off(r, j)] = 0.5 * (M[r * ld + j] + M[j * ld + r]);
        __syncthreads();
        for (int i = tid; i < k; i += nt) C[sym_off(k, i)] = bout[i];
        __syncthreads();
    }
}

// Grid points flagged by k_ienks_pre keep their incoming weights: W' = W (ienks.py:143-150), x_a = x_mean + X' W.
__global__ void __launch_bounds__(128) k_ienks_keep(const IenksParams P) {
    extern __shared__ double ksm2[];
    const int k = P.k;
    double* W = ksm2;                  // [k][k]
    double* xbuf = W + (size_t)k * k;  // [k]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t s = blockIdx.x; s < P.n_slots; s += gridDim.x) {
        if (!P.empty[s]) continue;     // uniform over the CTA
        const int64_t gi = P.gpos[P.slot_base + s].id;
        const int64_t wbase = P.per_grid ? gi * (int64_t)k * k : 0;
        for (int e = tid; e < k * k; e += nt) {
            const double w = ld_io(P.w_in, wbase + e, P.io_f32);
            W[e] = w;
            if (P.w_out) st_io(P.w_out, gi * (int64_t)k * k + e, w, P.io_f32);
        }
        __syncthreads();
        for (int sl = 0; sl < P.n_slices; ++sl) {
            const int64_t base = (int64_t)sl * k * P.n_grid + gi;
            for (int i = tid; i < k; i += nt) xbuf[i] = ld_io(P.x, base + (int64_t)i * P.n_grid, P.io_f32);
            __syncthreads();
            double mean = 0.0;
            for (int i = 0; i < k; ++i) mean += xbuf[i];
            mean /= (double)k;
            for (int j = tid; j < k; j += nt) {
                double acc = 0.0;
                for (int i = 0; i < k; ++i) acc = fma(xbuf[i] - mean, W[i * k + j], acc);
                st_io(P.xa, base + (int64_t)j * P.n_grid, mean + acc, P.io_f32);
            }
            __syncthreads();
        }
    }
}

}  // namespace b200da
