// Micro-benchmarks for the FP64 roofline denominators on B200 (sm_100a):
// DFMA, DMMA m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16, FFMA, and a DMMA loop fed from
// shared memory. Prints one JSON object. Timing: CUDA events, best of 5.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
    double c[NACC][4]; double av[2] = {a, a + 1};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1684(c[i], av, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[NACC][4]; double av[4] = {a, a + 1, a + 2, a + 3}; double bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[NACC][4]; double av[8]; double bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// DMMA m8n8k4 fed from shared memory like the Gram kernel: 7 fragment loads + 1 weight load + 7 DMUL
// per k-step of 4 obs, 14 DMMAs per warp (half of the 28 lower-triangle tiles of a 56x56 Gram).
template <int HALF>
__device__ __forceinline__ void gramlike_body(const double* sm, const double* wt, int iters, int lane, double (&c)[14][2]) {
    const int S = 60, T = 64;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int ks = 0; ks < T / 4; ++ks) {
            const int j = ks * 4 + (lane & 3);
            const double w = wt[j];
            double f[7];
#pragma unroll
            for (int t = 0; t < 7; ++t) f[t] = sm[j * S + t * 8 + (lane >> 2)];
            int idx = 0, n = 0;
#pragma unroll
            for (int mt = 0; mt < 7; ++mt) {
#pragma unroll
                for (int nt = 0; nt <= mt; ++nt) {
                    if ((idx & 1) == HALF) { dmma884(c[n][0], c[n][1], f[mt] * w, f[nt]); ++n; }
                    ++idx;
                }
            }
        }
    }
}
__global__ void __launch_bounds__(512) k_gramlike(double* out, int iters) {
    extern __shared__ double sm[];
    const int S = 60, T = 64;
    for (int i = threadIdx.x; i < T * S + 8 * T; i += blockDim.x) sm[i] = (i % 97) * 1e-3;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp >> 1;
    double c[14][2];
#pragma unroll
    for (int i = 0; i < 14; ++i) { c[i][0] = 0; c[i][1] = 0; }
    const double* wt = sm + T * S + g * T;
    if (warp & 1) gramlike_body<1>(sm, wt, iters, lane, c); else gramlike_body<0>(sm, wt, iters, lane, c);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 14; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// Do DFMA and DMMA co-issue?  ND independent DMMA chains + NF independent DFMA chains per iteration, interleaved: if the two
// instruction classes had their own datapaths the mixed loop would take max(t_dmma, t_dfma), if they share the FP64 units it
// takes t_dmma + t_dfma.
template <int ND, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
    double c[ND > 0 ? ND : 1][2];
    double f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < (ND > 0 ? ND : 1); ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
            if (i < ND) dmma884(c[i][0], c[i][1], a, b);
            if (i < NF) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (ND > 0 ? ND : 1); ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

// The k = 50 Gram inner loop with the last three rows of [Yn; d] on the DFMA pipe (letkf_kernel.cuh, ER = 3): 6 tile rows =
// 21 lower-triangle tiles split 10 / 11 over the two warps of a grid point + 3 x 3 column FMAs per warp + the 6 pair terms.
template <int HALF>
__device__ __forceinline__ void gramlike_er_body(const double* sm, const double* wt, int iters, int lane, double (&c)[11][2],
                                                 double (&e)[3][3], double (&p)[6]) {
    const int S = 60, T = 64;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int ks = 0; ks < T / 4; ++ks) {
            const int j = ks * 4 + (lane & 3);
            const double w = wt[j];
            double f[6], x[3], wx[3];
#pragma unroll
            for (int t = 0; t < 6; ++t) f[t] = sm[j * S + t * 8 + (lane >> 2)];
#pragma unroll
            for (int q = 0; q < 3; ++q) { x[q] = sm[j * S + 48 + q]; wx[q] = w * x[q]; }
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int t = 0; t < 3; ++t) e[q][t] = fma(f[HALF * 3 + t], wx[q], e[q][t]);
            if (HALF == 0) {
                int n = 0;
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int q2 = 0; q2 <= q; ++q2) { p[n] = fma(wx[q], x[q2], p[n]); ++n; }
            }
            int idx = 0, n = 0;
#pragma unroll
            for (int mt = 0; mt < 6; ++mt) {
#pragma unroll
                for (int nt = 0; nt <= mt; ++nt) {
                    if ((idx >= 10) == (HALF == 1)) { dmma884(c[n][0], c[n][1], f[mt] * w, f[nt]); ++n; }
                    ++idx;
                }
            }
        }
    }
}
__global__ void __launch_bounds__(512) k_gramlike_er(double* out, int iters) {
    extern __shared__ double sm[];
    const int S = 60, T = 64;
    for (int i = threadIdx.x; i < T * S + 8 * T; i += blockDim.x) sm[i] = (i % 97) * 1e-3;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp >> 1;
    double c[11][2], e[3][3], p[6];
#pragma unroll
    for (int i = 0; i < 11; ++i) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { e[i][0] = 0; e[i][1] = 0; e[i][2] = 0; }
#pragma unroll
    for (int i = 0; i < 6; ++i) p[i] = 0;
    const double* wt = sm + T * S + g * T;
    if (warp & 1) gramlike_er_body<1>(sm, wt, iters, lane, c, e, p); else gramlike_er_body<0>(sm, wt, iters, lane, c, e, p);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 11; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 3; ++i) s += e[i][0] + e[i][1] + e[i][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) s += p[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1024));
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int occ = 1; occ <= 8; occ *= 2) {
        const int grid = sms * occ, block = 256;
        const double thr = (double)grid * block;
        double ms;
        ms = time_ms([&] { k_dfma<16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dfma_tflops_occ%d\": %.2f", occ, thr * iters * 16 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_ffma<16><<<grid, block>>>((float*)out, iters, 1.0000001f, 1e-9f); });
        printf(", \"ffma_tflops_occ%d\": %.2f", occ, thr * iters * 16 * 2 / ms * 1e-9);
        const double warps = thr / 32;
        ms = time_ms([&] { k_dmma884<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmma884_tflops_occ%d\": %.2f", occ, warps * iters * 8 * (8 * 8 * 4) * 2 / ms * 1e-9);
        ms = time_ms([&] { k_dmma1684<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmma1684_tflops_occ%d\": %.2f", occ, warps * iters * 8 * (16 * 8 * 4) * 2 / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmma1688_tflops_occ%d\": %.2f", occ, warps * iters * 8 * (16 * 8 * 8) * 2 / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dmma16816_tflops_occ%d\": %.2f", occ, warps * iters * 8 * (16 * 8 * 16) * 2 / ms * 1e-9);
    }
    {
        const int smem = (64 * 60 + 8 * 64) * 8;
        CK(cudaFuncSetAttribute(k_gramlike, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int it2 = 256;
        double ms = time_ms([&] { k_gramlike<<<sms, 512, smem>>>(out, it2); });
        // 16 warps x 16 k-steps x 14 DMMA x 256 FMA x 2
        printf(", \"gramlike_tflops\": %.2f", (double)sms * 16 * it2 * 16 * 14 * 256 * 2 / ms * 1e-9);
    }
    {
        // co-issue probe at 4 CTAs of 256 threads per SM (8 warps per scheduler)
        const int grid = sms * 4, block = 256;
        const double t_d = time_ms([&] { k_mix<8, 0><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        const double t_f8 = time_ms([&] { k_mix<0, 8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        const double t_m8 = time_ms([&] { k_mix<8, 8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        const double t_f16 = time_ms([&] { k_mix<0, 16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        const double t_m16 = time_ms([&] { k_mix<8, 16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"coissue\": {\"dmma8_ms\": %.3f, \"dfma8_ms\": %.3f, \"dmma8_dfma8_ms\": %.3f, \"dfma16_ms\": %.3f, "
               "\"dmma8_dfma16_ms\": %.3f, \"overlap_8\": %.3f, \"overlap_16\": %.3f, \"note\": \"overlap = (t_dmma + t_dfma - "
               "t_mixed) / min(t_dmma, t_dfma): 1 = separate datapaths, 0 = one shared FP64 datapath\"}",
               t_d, t_f8, t_m8, t_f16, t_m16, (t_d + t_f8 - t_m8) / (t_d < t_f8 ? t_d : t_f8),
               (t_d + t_f16 - t_m16) / (t_d < t_f16 ? t_d : t_f16));
    }
    {
        const int smem = (64 * 60 + 8 * 64) * 8;
        CK(cudaFuncSetAttribute(k_gramlike_er, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int it2 = 256;
        double ms = time_ms([&] { k_gramlike_er<<<sms, 512, smem>>>(out, it2); });
        // algorithmic Gram FLOPs of k = 50 per observation and grid point: 2 * 51 * 51 (full square incl. the d row)
        printf(", \"gramlike_er3_ms\": %.3f, \"gramlike_er3_alg_tflops\": %.2f", ms,
               (double)sms * 8 * it2 * 64 * (2.0 * 50 * 50 + 2.0 * 50) / ms * 1e-9);
        const int smem0 = (64 * 60 + 8 * 64) * 8;
        double ms0 = time_ms([&] { k_gramlike<<<sms, 512, smem0>>>(out, it2); });
        printf(", \"gramlike_kt7_ms\": %.3f, \"gramlike_kt7_alg_tflops\": %.2f", ms0,
               (double)sms * 8 * it2 * 64 * (2.0 * 50 * 50 + 2.0 * 50) / ms0 * 1e-9);
    }
    printf("}\n");
    return 0;
}
