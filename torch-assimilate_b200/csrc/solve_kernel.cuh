// K3b — symmetric eigendecomposition (parallel cyclic Jacobi, Brent-Luk data movement), ETKF transform and
// state update.
//
// Reference semantics (paths relative to /root/reference):
//   evd: symeig, clamp(min=0), + (k-1)/rho, reciprocal           pytassim/core/utils.py:26-61
//   rev_evd: U diag(f) U^T                                        pytassim/core/utils.py:64-93
//   w_mean = P~a Y d,  W_p = U ((k-1) L^-1)^(1/2) U^T, W = w_mean + W_p     pytassim/core/etkf.py:57-103
//   x_a = mean + (x - mean) W                                      pytassim/interface/base.py:257-278
//
// Jacobi layout.  The ne = 2*n2 (k rounded up to even) indices live in n2 pair slots: slot t = positions
// (2t, 2t+1) of the matrix.  A step rotates every slot pair, A <- J^T A J, which decomposes into n2 x n2
// independent 2x2 blocks B_ij <- R_i^T B_ij R_j, and then MOVES the data with the fixed tournament permutation
// (position 0 stays; 2 -> 4 -> ... -> 2(n2-1) -> 2(n2-1)+1 -> ... -> 3 -> 1 -> 2), so that after ne-1 steps every
// pair of indices has met once.  Because the pairing is expressed by where the data sits, every thread reads and
// writes the same addresses in every step: no index arithmetic in the loop.  Warps own block rows, lanes own block
// columns; lane t of every warp evaluates rotation t (redundantly per warp — no hand-off through shared memory, no
// serial phase) and a row's rotation is fetched with a shuffle.  A thread keeps its blocks in registers between
// the "all reads" and the "all writes" barrier, which also gives the ILP needed to cover the ~45-cycle FP64 latency.
// V is stored transposed (row = eigenvector slot) so that its update is conflict-free and contiguous.
// Rotations are skipped when a_pq^2 <= tol^2 (a_pp + shift)(a_qq + shift): the criterion of a relative-accuracy
// Jacobi on A + shift*I, the matrix whose functions the transform needs.  The tangent is evaluated in FP32 on
// operands pre-scaled by a power of two and (c, s) is renormalised in FP64, so each rotation is orthogonal to
// working precision (an inexact angle only costs a little convergence).
#pragma once
#include "plan.cuh"

namespace b200da {

constexpr int kMaxSweeps = 40;

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ int bl_dest(int c, int n2) {      // where the content of position c goes after a step
    if (n2 == 1) return c;
    const int t = c >> 1;
    if (c & 1) return t >= 1 ? 2 * (t - 1) + 1 : 2;
    if (t == 0) return 0;
    return t <= n2 - 2 ? 2 * (t + 1) : 2 * (n2 - 1) + 1;
}

__device__ __forceinline__ float scaled_float(double v, int e) {     // v * 2^(127 - e) as float, flushing small values
    const int hi = __double2hiint(v), lo = __double2loint(v);
    const int ev = ((hi >> 20) & 0x7ff) - e + 127;
    if (ev <= 0) return 0.f;
    const unsigned bits = ((unsigned)hi & 0x80000000u) | ((unsigned)ev << 23) | (((unsigned)hi & 0xfffffu) << 3) |
                          ((unsigned)lo >> 29);
    return __uint_as_float(bits);
}

__device__ __forceinline__ bool jacobi_rotation(double app, double aqq, double apq, double shift, double& c, double& s) {
    const double tol2 = 1e-30;
    const bool rot = apq * apq > tol2 * fabs((app + shift) * (aqq + shift));
    const double d = aqq - app;
    const int ed = (__double2hiint(d) >> 20) & 0x7ff, eb = ((__double2hiint(apq) >> 20) & 0x7ff) + 1;
    const int e = max(ed, eb);
    const float fd = scaled_float(d, e), fb = 2.f * scaled_float(apq, e);
    const float h = sqrtf(fmaf(fd, fd, fb * fb));
    const float tf = __fdividef(fd >= 0.f ? fb : -fb, fabsf(fd) + h);   // |t| <= 1: the inner rotation
    const float c0 = rsqrtf(fmaf(tf, tf, 1.f));
    const float s0 = tf * c0;
    const double cd = (double)c0, sd = (double)s0;
    const double err = fma(cd, cd, fma(sd, sd, -1.0));                  // c0^2 + s0^2 - 1  (~1e-7)
    const double f = fma(err, fma(err, 0.375, -0.5), 1.0);              // (1 + err)^(-1/2) to O(err^3)
    c = rot ? cd * f : 1.0;
    s = rot ? sd * f : 0.0;
    return rot;
}

// A: [ne][lda] (lda even), Vt: [ne][ldv]; returns the number of sweeps.  NR: block columns per lane, RW: block rows
// per warp, NV: 32-row chunks of Vt per lane.  Requires n2 <= 32*NR, n2 <= RW*(nthreads/32), ne <= 32*NV.
template <int NR, int RW, int NV>
__device__ int jacobi_evd(double* __restrict__ A, double* __restrict__ Vt, int ne, int lda, int ldv, double shift,
                          int tid, int nthreads, int bar_id) {
    const int lane = tid & 31, gw = tid >> 5, GW = nthreads >> 5;
    const int n2 = ne >> 1;
    for (int i = tid; i < ne * ldv; i += nthreads) Vt[i] = 0.0;
    group_barrier(bar_id, nthreads);
    for (int i = tid; i < ne; i += nthreads) Vt[i * ldv + i] = 1.0;
    // loop-invariant addressing
    int dc0[NR], dc1[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        const int tj = min(lane + 32 * u, n2 - 1);
        dc0[u] = bl_dest(2 * tj, n2); dc1[u] = bl_dest(2 * tj + 1, n2);
    }
    int dr0[RW], dr1[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) {
        const int ti = min(gw + GW * i, n2 - 1);
        dr0[i] = bl_dest(2 * ti, n2); dr1[i] = bl_dest(2 * ti + 1, n2);
    }
    const int rows_mine = (n2 - gw + GW - 1) / GW;             // block rows / V pairs this warp owns (may be <= 0)
    group_barrier(bar_id, nthreads);
    const int nsteps = n2 == 1 ? 1 : ne - 1;
    int sweep = 0;
    for (; sweep < kMaxSweeps; ++sweep) {
        bool any_rot = false;
        for (int step = 0; step < nsteps; ++step) {
            // ---- phase A: everything this thread will need, into registers --------------------------------------
            double app[NR], aqq[NR], apq[NR];
#pragma unroll
            for (int u = 0; u < NR; ++u) {
                const int t = lane + 32 * u;
                app[u] = 1.0; aqq[u] = 1.0; apq[u] = 0.0;
                if (t < n2) {
                    const double2 top = *reinterpret_cast<const double2*>(A + (2 * t) * lda + 2 * t);
                    app[u] = top.x; apq[u] = top.y;
                    aqq[u] = A[(2 * t + 1) * lda + 2 * t + 1];
                }
            }
            double2 bt[RW][NR], bb[RW][NR];
            double vp[RW][NV], vq[RW][NV];
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                if (i < rows_mine) {
                    const int ti = gw + GW * i;
                    const double* rt = A + (2 * ti) * lda;
#pragma unroll
                    for (int u = 0; u < NR; ++u) {
                        const int tj = lane + 32 * u;
                        if (tj < n2) {
                            bt[i][u] = *reinterpret_cast<const double2*>(rt + 2 * tj);
                            bb[i][u] = *reinterpret_cast<const double2*>(rt + lda + 2 * tj);
                        }
                    }
                    const double* vt = Vt + (2 * ti) * ldv;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const int r = lane + 32 * j;
                        if (r < ne) { vp[i][j] = vt[r]; vq[i][j] = vt[ldv + r]; }
                    }
                }
            }
            group_barrier(bar_id, nthreads);
            // ---- phase B: rotations, 2x2 block updates, writes to the permuted positions -------------------------
            double cj[NR], sj[NR];
            bool rot = false;
#pragma unroll
            for (int u = 0; u < NR; ++u) rot |= jacobi_rotation(app[u], aqq[u], apq[u], shift, cj[u], sj[u]);
            any_rot |= (__any_sync(0xffffffffu, rot) != 0);
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                if (i < rows_mine) {
                    const int ti = gw + GW * i;
                    double ci = __shfl_sync(0xffffffffu, cj[0], ti & 31), si = __shfl_sync(0xffffffffu, sj[0], ti & 31);
                    if (NR > 1) {
                        const double c1 = __shfl_sync(0xffffffffu, cj[NR - 1], ti & 31);
                        const double s1 = __shfl_sync(0xffffffffu, sj[NR - 1], ti & 31);
                        if (ti >= 32) { ci = c1; si = s1; }
                    }
                    double* o0 = A + dr0[i] * lda;
                    double* o1 = A + dr1[i] * lda;
#pragma unroll
                    for (int u = 0; u < NR; ++u) {
                        if (lane + 32 * u < n2) {
                            const double b00 = bt[i][u].x, b01 = bt[i][u].y, b10 = bb[i][u].x, b11 = bb[i][u].y;
                            const double r00 = ci * b00 - si * b10, r01 = ci * b01 - si * b11;   // rows' = R_i^T rows
                            const double r10 = si * b00 + ci * b10, r11 = si * b01 + ci * b11;
                            o0[dc0[u]] = cj[u] * r00 - sj[u] * r01;                               // cols' = cols R_j
                            o0[dc1[u]] = sj[u] * r00 + cj[u] * r01;
                            o1[dc0[u]] = cj[u] * r10 - sj[u] * r11;
                            o1[dc1[u]] = sj[u] * r10 + cj[u] * r11;
                        }
                    }
                    double* v0 = Vt + dr0[i] * ldv;
                    double* v1 = Vt + dr1[i] * ldv;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const int r = lane + 32 * j;
                        if (r < ne) {
                            v0[r] = ci * vp[i][j] - si * vq[i][j];
                            v1[r] = si * vp[i][j] + ci * vq[i][j];
                        }
                    }
                }
            }
            group_barrier(bar_id, nthreads);
        }
        if (!any_rot) break;                                   // identical in every warp: all see the same rotations
    }
    return sweep + 1;
}

// A (diagonal = eigenvalues per slot), Vt (row m = eigenvector of slot m), b -> W = w_mean 1^T + W_p as
// [k][ldw] row-major written over A.  Sums run over all ne slots; the dummy slot of an odd ensemble size is a
// decoupled eigenpair (0, e_dummy) that never touches the first k rows.
__device__ inline void etkf_transform(double* __restrict__ A, double* __restrict__ Vt, const double* __restrict__ bvec,
                               double* __restrict__ vec, int k, int ne, int lda, int ldv, double rho, int tid,
                               int nthreads, int bar_id) {
    double* inv = vec;             // [ne] 1 / (max(lambda, 0) + (k-1)/rho)        core/utils.py:58-60
    double* z = vec + ne;          // [ne] L^-1 U^T b
    double* wbar = vec + 2 * ne;   // [ne] w_mean                                  core/etkf.py:72-73
    const double reg = (double)(k - 1) / rho;
    for (int m = tid; m < ne; m += nthreads) {
        const double iv = 1.0 / (fmax(A[m * lda + m], 0.0) + reg);
        inv[m] = iv;
        double acc = 0.0;
        for (int i = 0; i < k; ++i) acc = fma(Vt[m * ldv + i], bvec[i], acc);
        z[m] = acc * iv;
    }
    group_barrier(bar_id, nthreads);
    for (int i = tid; i < k; i += nthreads) {
        double acc = 0.0;
        for (int m = 0; m < ne; ++m) acc = fma(Vt[m * ldv + i], z[m], acc);
        wbar[i] = acc;
    }
    group_barrier(bar_id, nthreads);            // w_mean reads the unscaled Vt: finish before any thread scales it
    // scale row m of Vt by ((k-1) inv_m)^(1/4): W_p = B^T B with B = that matrix            core/etkf.py:75-76
    for (int x = tid; x < ne * ne; x += nthreads) {
        const int m = x / ne, i = x - m * ne;
        Vt[m * ldv + i] *= sqrt(sqrt((double)(k - 1) * inv[m]));
    }
    group_barrier(bar_id, nthreads);
    for (int x = tid; x < k * (k + 1) / 2; x += nthreads) {
        int i = (int)((sqrtf(8.0f * (float)x + 1.0f) - 1.0f) * 0.5f);
        while (i * (i + 1) / 2 > x) --i;
        while ((i + 1) * (i + 2) / 2 <= x) ++i;
        const int j = x - i * (i + 1) / 2;
        double acc = 0.0;
        for (int m = 0; m < ne; ++m) acc = fma(Vt[m * ldv + i], Vt[m * ldv + j], acc);
        A[i * lda + j] = acc + wbar[i];                      // W[i][j] = w_mean[i] + W_p[i][j]   core/etkf.py:102
        if (i != j) A[j * lda + i] = acc + wbar[j];
    }
    group_barrier(bar_id, nthreads);
}

// x_a[s, j, g] = mean + sum_i (x[s, i, g] - mean) W[i][j]                                     interface/base.py:257-278
__device__ inline void apply_point(const double* __restrict__ W, int ldw, int k, int n_slices, int64_t n_grid, int64_t gi,
                            const void* __restrict__ x, void* __restrict__ xa, void* __restrict__ w_out, int f32,
                            double* __restrict__ xbuf, int tid, int nthreads, int bar_id) {
    if (w_out) {
        for (int i = tid; i < k * k; i += nthreads) st_io(w_out, gi * (int64_t)k * k + i, W[(i / k) * ldw + (i % k)], f32);
    }
    for (int s = 0; s < n_slices; ++s) {
        const int64_t base = (int64_t)s * k * n_grid + gi;
        for (int i = tid; i < k; i += nthreads) xbuf[i] = ld_io(x, base + (int64_t)i * n_grid, f32);
        group_barrier(bar_id, nthreads);
        double mean = 0.0;
        for (int i = 0; i < k; ++i) mean += xbuf[i];       // same order for every thread
        mean /= (double)k;
        for (int j = tid; j < k; j += nthreads) {
            double acc = 0.0;
            for (int i = 0; i < k; ++i) acc = fma(xbuf[i] - mean, W[i * ldw + j], acc);
            st_io(xa, base + (int64_t)j * n_grid, mean + acc, f32);
        }
        group_barrier(bar_id, nthreads);
    }
}

// shared-memory carve-up of one solve: A [ne][lda], Vt [ne][ldv], b [ne], vec [3 ne], xbuf [ne]
struct SolveSmem {
    double *A, *Vt, *bvec, *vec, *xbuf;
    int ne, lda, ldv;
};
__host__ __device__ inline int solve_ne(int k) { return (k + 1) & ~1; }
__host__ __device__ inline size_t solve_smem_bytes(int k) {
    const int ne = solve_ne(k);
    return sizeof(double) * ((size_t)ne * ne + (size_t)ne * (ne + 1) + 5 * (size_t)ne) + 32;
}
__device__ inline SolveSmem carve_solve_smem(unsigned char* base, int k) {
    SolveSmem S;
    S.ne = solve_ne(k); S.lda = S.ne; S.ldv = S.ne + 1;
    S.A = reinterpret_cast<double*>(base);
    S.Vt = S.A + (size_t)S.ne * S.lda;
    S.bvec = S.Vt + (size_t)S.ne * S.ldv;
    S.vec = S.bvec + S.ne;
    S.xbuf = S.vec + 3 * S.ne;
    return S;
}

struct SolveParams {
    const double* cmat;        // [n_slots][slot_stride]: tile-packed augmented Gram (common.cuh), row k = b
    int64_t slot_stride;
    const Pos4* gpos;          // block-sorted grid positions (id = original index)
    const void* x;             // state / analysis / exported weights in the plan dtype (io_f32)
    void* xa;
    void* w_out;
    int io_f32;
    unsigned long long* stats;
    int64_t slot_base;
    int64_t n_slots;
    int64_t n_grid;
    int k;
    int n_slices;
    double rho;
};

// One CTA per grid point (grid-stride over the chunk's slots).
template <int THREADS, int MINB, int NR, int RW, int NV>
__global__ void __launch_bounds__(THREADS, MINB) k_letkf_solve(const SolveParams P) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int k = P.k;
    const SolveSmem S = carve_solve_smem(smem_raw, k);
    const int tid = threadIdx.x;
    const long long t0 = clock64();
    for (int64_t s = blockIdx.x; s < P.n_slots; s += gridDim.x) {
        const double* C = P.cmat + (size_t)s * (size_t)P.slot_stride;
        for (int x = tid; x < S.ne * S.lda; x += THREADS) S.A[x] = 0.0;
        if (tid < S.ne) S.bvec[tid] = 0.0;
        __syncthreads();
        for (int x = tid; x < (k + 1) * k; x += THREADS) {
            const int r = x / k, c = x - r * k;
            if (r < k) { if (c <= r) { const double v = C[sym_off(r, c)]; S.A[r * S.lda + c] = v; S.A[c * S.lda + r] = v; } }
            else S.bvec[c] = C[sym_off(k, c)];
        }
        __syncthreads();
        const int nsw = jacobi_evd<NR, RW, NV>(S.A, S.Vt, S.ne, S.lda, S.ldv, (double)(k - 1) / P.rho, tid, THREADS, 0);
        if (P.stats && tid == 0) { atomicAdd(P.stats + 2, (unsigned long long)nsw); atomicAdd(P.stats + 3, 1ull); }
        etkf_transform(S.A, S.Vt, S.bvec, S.vec, k, S.ne, S.lda, S.ldv, P.rho, tid, THREADS, 0);
        apply_point(S.A, S.lda, k, P.n_slices, P.n_grid, P.gpos[P.slot_base + s].id, P.x, P.xa, P.w_out, P.io_f32, S.xbuf, tid,
                    THREADS, 0);
        __syncthreads();
    }
    if (P.stats && tid == 0) atomicAdd(P.stats + 1, (unsigned long long)(clock64() - t0));
}

}  // namespace b200da
