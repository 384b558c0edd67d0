// Device side of the kernelised ETKF: evaluation of a kernel program and the in-place rewrite of a Gram slot (see kernelise.cuh
// for the reference semantics).  Included by b200da.cu only (k_kernelise is not a template).
#pragma once
#include "kernelise.cuh"

namespace b200da {

// xy = x . y, xx = |x|^2, yy = |y|^2; same_set: x and y come from the same sample set (K(perts, perts)); diag: same_set and i == j.
__device__ inline double kernel_eval(const KernelProgram& K, double xy, double xx, double yy, bool diag, bool same_set) {
    double st[kKernelStack];
    int sp = 0;
    for (int t = 0; t < K.n; ++t) {
        const double a = K.p0[t], b = K.p1[t];
        switch (K.op[t]) {
            case B200DA_KOP_LINEAR:                                  // kernels/linear.py:62-63
                st[sp++] = xy; break;
            case B200DA_KOP_GAUSS: {                                 // kernels/rbf.py forward: exp(-|x/l - y/l|^2 / 2); a = l
                const double d2 = fmax((xx + yy - 2.0 * xy) / (a * a), 0.0);
                st[sp++] = exp(-(d2 / 2.0)); break;
            }
            case B200DA_KOP_POLY:                                    // kernels/polynomial.py: (x.y + c)^p; a = p, b = c
                st[sp++] = pow(xy + b, a); break;
            case B200DA_KOP_TANH:                                    // kernels/tanh.py: tanh(alpha x.y + c); a = alpha, b = c
                st[sp++] = tanh(fma(a, xy, b)); break;
            case B200DA_KOP_RATIONAL: {                              // kernels/rational.py: (1 + |x/l - y/l|^2 / (2 w))^(-w); a = l, b = w
                const double d2 = fmax((xx + yy - 2.0 * xy) / (a * a), 0.0);
                st[sp++] = pow(1.0 + d2 / (2.0 * b), -b); break;
            }
            case B200DA_KOP_SCALE:                                   // kernels/scale.py: constant
                st[sp++] = a; break;
            case B200DA_KOP_DIAG:                                    // kernels/diag.py: c I for equal sample counts, else 0
                st[sp++] = (same_set && diag) ? a : 0.0; break;
            case B200DA_KOP_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;          // base_kernels.py:89-90
            case B200DA_KOP_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;          // :119-120
            default:             --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;      // :160-161 (B200DA_KOP_POW)
        }
    }
    return st[0];
}

// One CTA per slot (grid-stride).  Slots whose augmented Gram has a zero diagonal (no local observation) are left alone:
// the reference returns the inflated prior weights there (core/etkf.py:91-95) and the zero Gram gives exactly that.
// Every kernel value is evaluated once (lower triangle, mirrored into a k x (k + 1) shared-memory copy); row means and the
// centring then work from shared memory.  Dynamic shared memory: kernelise_smem_bytes(k).
__host__ __device__ inline size_t kernelise_smem_bytes(int k) { return sizeof(double) * ((size_t)k * (k + 1) + 3 * (size_t)(k + 1)); }

__global__ void __launch_bounds__(256) k_kernelise(double* __restrict__ cmat, int64_t n_slots, int64_t slot_stride, int k,
                                                   const KernelProgram K) {
    extern __shared__ double ksm[];
    const int ld = k + 1;
    double* Ks = ksm;                 // [k][k + 1] kernel matrix
    double* dg = Ks + (size_t)k * ld; // [k + 1] diagonal of G
    double* rmean = dg + (k + 1);     // [k] row means of the kernel matrix                          ketkf.py:81
    double* kobs = rmean + (k + 1);   // [k] K(x_i, d)                                               ketkf.py:90
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t s = blockIdx.x; s < n_slots; s += gridDim.x) {
        double* C = cmat + (size_t)s * (size_t)slot_stride;
        int nonzero = 0;
        for (int i = tid; i <= k; i += nt) { const double v = C[sym_off(i, i)]; dg[i] = v; nonzero |= (v != 0.0); }
        if (!__syncthreads_or(nonzero)) continue;
        for (int e = tid; e < k * k; e += nt) {                      // K[i][j] = K[j][i], j <= i                   ketkf.py:80
            const int i = e / k, j = e - i * k;
            if (j > i) continue;
            const double v = kernel_eval(K, C[sym_off(i, j)], dg[i], dg[j], i == j, true);
            Ks[i * ld + j] = v;
            Ks[j * ld + i] = v;
        }
        // one observation vector: diag.py:66-67 gives zeros unless k == 1
        for (int i = tid; i < k; i += nt) kobs[i] = kernel_eval(K, C[sym_off(k, i)], dg[i], dg[k], k == 1, k == 1);
        __syncthreads();
        for (int i = tid; i < k; i += nt) {
            double sum = 0.0;
            for (int j = 0; j < k; ++j) sum += Ks[i * ld + j];
            rmean[i] = sum / (double)k;
        }
        __syncthreads();
        double mu = 0.0, ko = 0.0;                                   // same order in every thread
        for (int i = 0; i < k; ++i) { mu += rmean[i]; ko += kobs[i]; }
        mu /= (double)k; ko /= (double)k;
        // K_c[i, j] = K[i, j] - mean_i K[:, j] - (mean_j K[i, :] - mu)                             ketkf.py:81-85
        for (int e = tid; e < k * k; e += nt) {
            const int i = e / k, j = e - i * k;
            if (j <= i) C[sym_off(i, j)] = Ks[i * ld + j] - rmean[j] - (rmean[i] - mu);
        }
        // k_obs_c[i] = k_obs[i] - mean(k_obs) - (mean_j K[i, :] - mu)                              ketkf.py:91-92
        for (int i = tid; i < k; i += nt) C[sym_off(k, i)] = (kobs[i] - ko) - (rmean[i] - mu);
        __syncthreads();
    }
}

}  // namespace b200da
