// Kernelised ensemble-space problem (KETKF / LKETKF with a non-linear kernel) on top of the augmented Gram.
//
// Reference semantics restated here (paths relative to /root/reference):
//   KETKFModule._estimate_weights      pytassim/core/ketkf.py:69-100   (kernel matrix, double centring, centred k_obs)
//   kernels                             pytassim/kernels/{linear,rbf,polynomial,tanh,rational,scale,diag}.py
//   compositions (+, *, **)             pytassim/kernels/base_kernels.py:40-161
//
// Every kernel above is an element-wise function of x_i . x_j, |x_i|^2 and |x_j|^2, i.e. of entries of the augmented Gram
// G = [Yn; d] [Yn; d]^T that the Gram kernels already leave in the tile-packed slot of a grid point (common.cuh): rows
// 0..k-1 = C, row k = b = Yn d^T, element (k, k) = d d^T.  k_kernelise rewrites a slot IN PLACE into the centred kernel
// matrix (rows 0..k-1) and the centred kernel column of the observations (row k); the ensemble-space solve kernels then run
// unchanged, because core/ketkf.py:87-96 is core/etkf.py:67-76 on those two quantities.  Kernels that need the L1 distance
// (Ornstein-Uhlenbeck, periodic) are not functions of the Gram and are rejected on the host.
//
// This header only declares the program (it is part of the plan); the device code is in kernelise_kernel.cuh, which only
// b200da.cu includes.
#pragma once
#include "common.cuh"

namespace b200da {

constexpr int kMaxKernelOps = 16;
constexpr int kKernelStack = 8;

// Postfix program of a (composed) kernel: leaves push K(x, y), the three compositions pop two values.
struct KernelProgram {
    int n;                          // 0: plain ETKF (no kernelise pass)
    int op[kMaxKernelOps];
    double p0[kMaxKernelOps];
    double p1[kMaxKernelOps];
};

}  // namespace b200da
