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

__host__ __device__ inline size_t ienks_smem_bytes(int k) { return sizeof(double) * (3 * (size_t)k * (k + 1) + 6 * (size_t)k) + 64; }

// One CTA per slot (grid-stride), 256 threads.  Shared memory: three k x (k + 1) matrices and six k-vectors.
__global__ void __launch_bounds__(256) k_ienks_pre(const IenksParams P) {
    extern __shared__ double ism[];
    const int k = P.k, ld = k + 1;
    double* A = ism;                   // Wp, later the Gram / S C S^T / A'
    double* B = A + (size_t)k * ld;    // T = Wp^-1
    double* M = B + (size_t)k * ld;    // temporary
    double* wbar = M + (size_t)k * ld;
    double* bv = wbar + k;
    double* colv = bv + k;
    double* rowa = colv + k;
    double* rowb = rowa + k;
    double* sb = rowb + k;
    __shared__ int piv_row;
    const int tid = threadIdx.x, nt = blockDim.x;
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
        for (int e = tid; e < k * k; e += nt) {
            const int i = e / k, j = e - i * k;
            A[i * ld + j] = ld_io(P.w_in, wbase + e, P.io_f32);
            B[i * ld + j] = i == j ? 1.0 : 0.0;
        }
        __syncthreads();
        for (int i = tid; i < k; i += nt) {                          // w = mean over columns of (W - I)        ienks.py:50-51
            double sum = 0.0;
            for (int j = 0; j < k; ++j) sum += A[i * ld + j] - (i == j ? 1.0 : 0.0);
            wbar[i] = sum / (double)k;
        }
        __syncthreads();
        for (int e = tid; e < k * k; e += nt) { const int i = e / k, j = e - i * k; A[i * ld + j] -= wbar[i]; }   // :52
        __syncthreads();
        // ---- T = Wp^-1 by Gauss-Jordan elimination with partial pivoting ([A | B] -> [I | T])                  :62-64
        for (int c = 0; c < k; ++c) {
            if (tid < 32) {
                double best = -1.0; int br = c;
                for (int r = c + tid; r < k; r += 32) { const double v = fabs(A[r * ld + c]); if (v > best) { best = v; br = r; } }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                    const int orr = __shfl_xor_sync(0xffffffffu, br, off);
                    if (ob > best || (ob == best && orr < br)) { best = ob; br = orr; }
                }
                if (tid == 0) piv_row = br;
            }
            __syncthreads();
            const int pr = piv_row;
            const double pinv = 1.0 / A[pr * ld + c];
            for (int j = tid; j < k; j += nt) {                      // pivot row, scaled; column c of every row
                rowa[j] = A[pr * ld + j] * pinv;
                rowb[j] = B[pr * ld + j] * pinv;
                colv[j] = A[j * ld + c];
            }
            __syncthreads();
            if (pr != c) {                                           // move row c into the pivot's place
                for (int j = tid; j < k; j += nt) { A[pr * ld + j] = A[c * ld + j]; B[pr * ld + j] = B[c * ld + j]; }
                __syncthreads();
                if (tid == 0) colv[pr] = colv[c];
                __syncthreads();
            }
            for (int e = tid; e < k * k; e += nt) {
                const int r = e / k, j = e - r * k;
                if (r == c) { A[r * ld + j] = rowa[j]; B[r * ld + j] = rowb[j]; }
                else { const double f = colv[r]; A[r * ld + j] = fma(-f, rowa[j], A[r * ld + j]); B[r * ld + j] = fma(-f, rowb[j], B[r * ld + j]); }
            }
            __syncthreads();
        }
        // ---- Gram of the grid point
        for (int e = tid; e < k * k; e += nt) {
            const int i = e / k, j = e - i * k;
            A[i * ld + j] = j <= i ? C[sym_off(i, j)] : C[sym_off(j, i)];
        }
        for (int i = tid; i < k; i += nt) bv[i] = C[sym_off(k, i)];
        __syncthreads();
        if (bundle) {                                                // dh_dw = Yn / eps                                   :173
            const double ie = 1.0 / P.eps;
            for (int e = tid; e < k * k; e += nt) { const int i = e / k, j = e - i * k; A[i * ld + j] = (A[i * ld + j] * ie) * ie; }
            for (int i = tid; i < k; i += nt) sb[i] = bv[i] * ie;
            __syncthreads();
        } else {                                                     // dh_dw = T Yn                                       :74-75
            for (int e = tid; e < k * k; e += nt) {
                const int i = e / k, j = e - i * k;
                double acc = 0.0;
                for (int l = 0; l < k; ++l) acc = fma(B[i * ld + l], A[l * ld + j], acc);
                M[i * ld + j] = acc;
            }
            for (int i = tid; i < k; i += nt) {
                double acc = 0.0;
                for (int l = 0; l < k; ++l) acc = fma(B[i * ld + l], bv[l], acc);
                sb[i] = acc;
            }
            __syncthreads();
            for (int e = tid; e < k * k; e += nt) {
                const int i = e / k, j = e - i * k;
                double acc = 0.0;
                for (int l = 0; l < k; ++l) acc = fma(M[i * ld + l], B[j * ld + l], acc);
                A[i * ld + j] = acc;
            }
            __syncthreads();
        }
        // ---- A' = (1 - tau)(k - 1) T^T T + tau S C S^T                                                         :62-66, :96-98
        const double c_old = (1.0 - P.tau) * km1;
        for (int e = tid; e < k * k; e += nt) {
            const int i = e / k, j = e - i * k;
            double acc = 0.0;
            if (c_old != 0.0)
                for (int l = 0; l < k; ++l) acc = fma(B[l * ld + i], B[l * ld + j], acc);
            M[i * ld + j] = fma(c_old, acc, P.tau * A[i * ld + j]);
        }
        __syncthreads();
        // ---- b' = P w - tau grad,  grad = (k - 1) w - S b,  P = A' + tau (k - 1) I                            :84-88, :130-131
        for (int i = tid; i < k; i += nt) {
            double acc = P.tau * km1 * wbar[i];
            for (int j = 0; j < k; ++j) acc = fma(0.5 * (M[i * ld + j] + M[j * ld + i]), wbar[j], acc);
            colv[i] = acc - P.tau * (km1 * wbar[i] - sb[i]);
        }
        __syncthreads();
        for (int e = tid; e < k * k; e += nt) {
            const int i = e / k, j = e - i * k;
            if (j <= i) C[sym_off(i, j)] = 0.5 * (M[i * ld + j] + M[j * ld + i]);
        }
        for (int i = tid; i < k; i += nt) C[sym_off(k, i)] = colv[i];
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
