// K2-K3a for FP32 plans with k >= 32 — the localization-weighted Gram matrices of 128 neighbouring grid points at once on
// the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory).
//
// Reference semantics (paths relative to /root/reference): the same as letkf_kernel.cuh —
//   localize_obs + sqrt(w) gather     pytassim/localization/gaspari_cohn.py:97-136, interface/wrapper.py:86-98
//   C = Y~ Y~^T, b = Y~ d~^T           pytassim/core/etkf.py:68,72 (core/utils.py:153-173)
//
// Formulation.  C_g[a][b] = sum_j w_gj y_aj y_bj is, for a whole block of grid points, ONE matrix product
//
//     D[g, (a,b)] = sum_j W[g, j] Z[(a,b), j],      W[g, j] = w_gj,   Z[(a,b), j] = y_aj y_bj,   b <= a <= k  (y_k = d~)
//
// with M = 128 grid points, N = the (k+1)(k+2)/2 lower-triangle pairs of the augmented matrix (split into chunks of
// <= 512 columns = one CTA each, because the FP32 accumulator D[128][N] lives in the 512 columns of tensor memory) and
// K = the candidate observations of the block.  Z is shared by all 128 grid points of the CTA, W is the taper:
// neither operand exists in memory, both are generated tile by tile (32 observations) in shared memory by the CUDA
// cores while the tensor core consumes the previous tile.  The symmetric half comes for free (only pairs b <= a are
// columns) and there is no padding of k to a multiple of 8.
//
// Precision.  Operands are split into two bfloat16 terms (x = hi + lo, 16 significant bits) and the product is
// accumulated as hi*hi + hi*lo + lo*hi in FP32 (3 kind::f16 MMAs per 16 observations; the dropped lo*lo term is 2^-18
// relative).  The FP32 accumulation of the tensor core truncates: over the ~3 500 accumulation steps of a cfg3 grid point
// (p ~ 56 000) a sum of same-sign terms comes out 4e-4 too small (measured, tools/diag_tc_gram.py: every diagonal entry
// biased by -4.1e-4 relative, off-diagonal entries by the same fraction of their own magnitude).  The accumulators are
// therefore kept small: every pair column is centred, Z[(a,b), j] = y_aj y_bj - c_ab with c_ab the mean of y_a y_b over
// ALL observations (b200da.cu: tc_centre_constants), and the epilogue adds c_ab * sum_j w_gj back in FP64, the sum of the
// weights being accumulated by the generator threads on the CUDA cores (FP64).  The taper is read from a per-CTA table of the weight as a function of the SQUARED bin-space distance
// (2048 intervals, linear interpolation, error < 1e-6; built in FP64 from the same taper code as the FP64 path, mask
// w > eps included), evaluated on FP32 positions relative to the block centre: a pair costs 3 subtractions, 3 FMAs, one
// table read and one FMA instead of a square root, an arc sine and two polynomials.  The k x k solve and the update stay
// in FP64 (ns_solve_kernel.cuh).
//
// Shared-memory operand tiles use the no-swizzle K-major canonical layout of the UMMA shared-memory descriptor: core
// matrices of 8 rows x 16 bytes, element (row, kk) at  (kk / 8) * ROWS * 16 + row * 16 + (kk % 8) * 2  bytes, i.e.
// leading (K) byte offset = ROWS * 16, stride (M/N) byte offset = 128.
#pragma once
#include <cuda_bf16.h>
#include "letkf_kernel.cuh"

namespace b200da {

constexpr int kTcM = 128;          // grid points per CTA (rows of D)
constexpr int kTcObs = 32;         // observations per staged tile (two K = 16 MMA steps)
constexpr int kTcGenWarps = 16;    // operand-generating warps (also the epilogue warps)
constexpr int kTcGenThreads = kTcGenWarps * 32;
constexpr int kTcMmaWarp = kTcGenWarps;        // issues the tcgen05.mma stream, owns the tensor-memory allocation
constexpr int kTcLoadWarp = kTcGenWarps + 1;   // first of the loader warps: candidate filter + observation-tile loads
constexpr int kTcYStages = 6;      // loader warps = observation-tile buffers (one each) between loaders and generators
constexpr int kTcThreads = (kTcGenWarps + 1 + kTcYStages) * 32;
constexpr int kTcYLd = 36;         // row length of the member-major observation tile: 32 obs + 4 (conflict-free LDS.128)
constexpr bool kTcATmem = true;    // W operand (A) in tensor memory (written with tcgen05.st) instead of shared memory
constexpr int kTcTmemCols = 512;   // tensor-memory allocation
constexpr int kTcACols = 64;       // A operand in tensor memory: 2 stages x (hi, lo) x 2 K steps x 8 columns (16 bf16 each)
constexpr int kTcMaxCols = kTcATmem ? kTcTmemCols - kTcACols : kTcTmemCols;   // accumulator columns per CTA
constexpr float kTcFar = 1.0e18f;  // coordinate of padding observations (+) and padding grid points (-)
constexpr int kTcFilterUnroll = 4; // candidates per lane and filter pass (independent loads in flight)

struct TcParams {
    LetkfParams L;
    int n_cols;        // (k + 1)(k + 2) / 2
    int n_chunks;      // column chunks = CTAs per grid-point block
    int nc;            // columns per chunk, multiple of 32, <= 512
    int kp;            // row length of the staging copy ys (floats)
    int n_load;        // active loader warps = observation-tile buffers (2..kTcYStages, limited by shared memory)
    float q_max;       // table domain: squared bin-space distance beyond which the weight is 0
    float period;
    int w_closed;      // 1: closed-form FP32 weights (no table reads), 0: table in squared distance
    int asin_poly;     // closed form, haversine: chord / 2 stays below 0.3, asin by its series
    float r_scale;     // closed form, distance -> r: 1 / radius (haversine: 2 R / radius, applied to asin(chord / 2))
    float eps;
    const float* centre;   // [n_cols] centring constant of every pair column (see "Precision")
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {     // release at CTA scope
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {      // arrives on `bar` when all earlier MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] B[smem]^T, bf16 x bf16 -> f32, M = 128, N from the instruction descriptor, K = 16
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (SM100 descriptor version 1)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128
__host__ __device__ constexpr uint32_t tc_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x = first argument = lower 16 bits
    return *reinterpret_cast<const uint32_t*>(&v);
}
// x[0..7] -> eight bf16 "hi" terms and eight bf16 "lo" terms (x = hi + lo to 16 significant bits)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = pack_bf16x2(x[2 * i] - h0, x[2 * i + 1] - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- localization weights from the table ---------------------------------------------------------------------------------
constexpr int kTcTab = 2048;       // intervals of the weight table over [0, q_max]; entry kTcTab is the zero beyond

// distance kinds (uniform per launch): squared distance in three coordinates (haversine through the chord, Euclidean),
// |dz|, periodic |dz|
enum { kTcSq3 = 0, kTcAbs = 1, kTcPeriodic = 2 };

// Weight of a pair from positions relative to the block centre (bin space).  Padding observations and the unused
// grid-point rows of a partial block sit at +kTcFar / -kTcFar in all three coordinates: their squared distance to every
// real partner is huge, the table index saturates at the zero entry.  The 1-D metrics, whose x coordinate is otherwise 0,
// add |dx| to the distance for that purpose.
template <int DIST>
__device__ __forceinline__ float pair_weight_f32(const float2* __restrict__ tab, float xs, float period, float gx, float gy,
                                                 float gz, float4 o) {
    const float dx = o.x - gx, dy = o.y - gy, dz = o.z - gz;
    float q;
    if (DIST == kTcSq3) {
        q = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    } else {
        float d = fabsf(dz);
        if (DIST == kTcPeriodic) d = fminf(d, fabsf(period - d));
        d += fabsf(dx);
        q = d * d;
    }
    const float x = fminf(q * xs, (float)kTcTab);                 // table coordinate, saturated at the zero entry
    const int i = __float2int_rz(x);
    const float2 e = tab[i];
    return fmaf(x - (float)i, e.y, e.x);                           // linear interpolation: e = (w_i, w_{i+1} - w_i)
}

// taper weights of one grid point for the 8 observations ot[0..7] -> bf16 hi / lo
template <int DIST>
__device__ __forceinline__ void w_chunk(const float4* __restrict__ ot, const float2* __restrict__ tab, float xs, float period,
                                        float gx, float gy, float gz, uint4& hi, uint4& lo, float& wsum) {
    float w[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) w[jj] = pair_weight_f32<DIST>(tab, xs, period, gx, gy, gz, ot[jj]);
    wsum = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
    split8(w, hi, lo);
}

// ---- the same weights by closed forms in FP32 (no shared-memory traffic, more instructions) ---------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// gaspari_cohn.py:78-95, branch-free; the outer branch in s = 2 - r:  f2(r) = s^4 (5/8 - s/2 + s^2/12) / r  (no cancellation,
// the reference form sums terms of magnitude 10 to get 1e-5)
__device__ __forceinline__ float taper_gc_f32(float r) {
    const float r2 = r * r;
    const float inner = fmaf(r2, fmaf(r, fmaf(r, fmaf(r, -0.25f, 0.5f), 0.625f), -5.0f / 3.0f), 1.0f);
    const float s = fmaxf(2.0f - r, 0.0f), s2 = s * s;
    const float outer = s2 * s2 * fmaf(s, fmaf(s, 1.0f / 12.0f, -0.5f), 0.625f) * rcp_approx(r);
    return r < 1.0f ? inner : outer;
}
// gaspari_cohn.py:172-210; the last branch in s = 2 - r:  f4(r) = s^4 (10/11 - 8 s/11 + 4 s^2/33) / r
__device__ __forceinline__ float taper_gcinf_f32(float r) {
    const float r2 = r * r, ri = rcp_approx(r);
    const float p1 = fmaf(r2, fmaf(r, fmaf(r, fmaf(r, -28.0f / 33.0f, 8.0f / 11.0f), 20.0f / 11.0f), -80.0f / 33.0f), 1.0f);
    const float p2 = fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, 20.0f / 33.0f, -16.0f / 11.0f), 0.0f), 100.0f / 33.0f),
                                  -45.0f / 11.0f), 51.0f / 22.0f) - (7.0f / 44.0f) * ri;
    const float p3 = fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, -4.0f / 11.0f, 16.0f / 11.0f), -10.0f / 11.0f),
                                         -100.0f / 33.0f), 5.0f), -61.0f / 22.0f) + (115.0f / 132.0f) * ri;
    const float s = fmaxf(2.0f - r, 0.0f), s2 = s * s;
    const float p4 = s2 * s2 * fmaf(s, fmaf(s, 4.0f / 33.0f, -8.0f / 11.0f), 10.0f / 11.0f) * ri;
    return r < 0.5f ? p1 : (r < 1.0f ? p2 : (r < 1.5f ? p3 : p4));
}
// closed-form kinds: haversine with asin by its series (chord / 2 <= 0.3), haversine with asinf, Euclidean, |dz|, periodic
enum { kTcHavPoly = 0, kTcHavAsin = 1, kTcEuclid = 2, kTcAbsF = 3, kTcPeriodicF = 4 };
template <int DIST, int TAPER>
__device__ __forceinline__ float pair_weight_closed(float r_scale, float eps, float period, float gx, float gy, float gz, float4 o) {
    const float dx = o.x - gx, dy = o.y - gy, dz = o.z - gz;
    float r;
    if (DIST == kTcHavPoly) {          // asin(h) = h (1 + h^2/6 + 3 h^4/40 + 15 h^6/336 + 35 h^8/1152 + ...)
        const float h2 = 0.25f * fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        r = (r_scale * sqrt_approx(h2)) *
            fmaf(h2, fmaf(h2, fmaf(h2, fmaf(h2, 35.0f / 1152.0f, 15.0f / 336.0f), 3.0f / 40.0f), 1.0f / 6.0f), 1.0f);
    } else if (DIST == kTcHavAsin) {
        const float h = 0.5f * sqrt_approx(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        r = fmaf(fmaxf(h - 1.0f, 0.0f), kTcFar, r_scale * asinf(fminf(h, 1.0f)));      // h > 1 only for padding
    } else if (DIST == kTcEuclid) {
        r = r_scale * sqrt_approx(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    } else {
        float d = fabsf(dz);
        if (DIST == kTcPeriodicF) d = fminf(d, fabsf(period - d));
        r = r_scale * (d + fabsf(dx));
    }
    const float w = TAPER == B200DA_TAPER_GCINF ? taper_gcinf_f32(r) : taper_gc_f32(r);
    return w > eps ? w : 0.0f;                                     // gaspari_cohn.py:135
}
template <int DIST, int TAPER>
__device__ __forceinline__ void w_chunk_closed(const float4* __restrict__ ot, float r_scale, float eps, float period, float gx,
                                               float gy, float gz, uint4& hi, uint4& lo, float& wsum) {
    float w[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) w[jj] = pair_weight_closed<DIST, TAPER>(r_scale, eps, period, gx, gy, gz, ot[jj]);
    wsum = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
    split8(w, hi, lo);
}

// table entry i: weight at squared bin-space distance q_i = i q_max / kTcTab, evaluated in FP64 with the reference's
// selection logic and mask (gaspari_cohn.py:126-135)
__device__ inline double tab_weight(const Geometry& g, double q) {
    double dist;
    if (g.metric == B200DA_METRIC_HAVERSINE) dist = 2.0 * g.sphere_r * asin(fmin(0.5 * sqrt(q), 1.0));
    else dist = sqrt(q);
    const double w = taper_eval(g.taper, dist / g.radius);
    return w > g.eps ? w : 0.0;
}

// ---- shared-memory carve-up ------------------------------------------------------------------------------------------------
// barriers: [0..1] operand stage full (16 generator warps), [2..3] operand stage free (tcgen05.commit),
//           [4..7] observation tile full (loader), [8..11] observation tile free (16 generator warps), [12] all MMAs done
struct TcSmem {
    BlockHeader<kTcM>* H;
    uint64_t* bars;
    uint32_t* tmem_slot;
    int* ymeta;              // [kTcYStages] valid observations of the tile, 0 = end of the candidate stream
    int* op_last;            // [2] set by the generators when the operand stage carries the end marker instead of a tile
    float2* wtab;            // [kTcTab + 1] weight table (w_i, w_{i+1} - w_i)
    float4* otile;           // [kTcYStages][kTcObs] observation positions relative to the block centre, .w = 1 valid / 0 padding
    float* ytile;            // [kTcYStages][kp][kTcYLd] member-major [Yn; d] of the tile
    unsigned char* a_hi;     // [2][kTcM * 64]
    unsigned char* a_lo;
    unsigned char* b_hi;     // [2][nc * 64]
    unsigned char* b_lo;
};
__host__ __device__ constexpr size_t tc_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t tc_smem_bytes(int kp, int nc, int n_load) {
    size_t o = tc_align(sizeof(BlockHeader<kTcM>), 128);
    o += 256;                                             // barriers, tensor-memory address, tile meta data
    o += tc_align(sizeof(float2) * (kTcTab + 1), 128);
    o += sizeof(float4) * kTcYStages * kTcObs;
    o = tc_align(o + sizeof(float) * n_load * (size_t)kp * kTcYLd, 128);
    o += kTcATmem ? 0 : 2 * 2 * (size_t)kTcM * 64;
    o += 2 * 2 * (size_t)nc * 64;
    return o;
}
__device__ __forceinline__ TcSmem tc_carve(unsigned char* base, int kp, int nc, int n_load) {
    TcSmem S;
    size_t o = tc_align(sizeof(BlockHeader<kTcM>), 128);
    S.H = reinterpret_cast<BlockHeader<kTcM>*>(base);
    S.bars = reinterpret_cast<uint64_t*>(base + o);
    S.tmem_slot = reinterpret_cast<uint32_t*>(base + o + 128);
    S.ymeta = reinterpret_cast<int*>(base + o + 144);
    S.op_last = reinterpret_cast<int*>(base + o + 176);
    o += 256;
    S.wtab = reinterpret_cast<float2*>(base + o);
    o += tc_align(sizeof(float2) * (kTcTab + 1), 128);
    S.otile = reinterpret_cast<float4*>(base + o);
    o += sizeof(float4) * kTcYStages * kTcObs;
    S.ytile = reinterpret_cast<float*>(base + o);
    o = tc_align(o + sizeof(float) * n_load * (size_t)kp * kTcYLd, 128);
    S.a_hi = base + o; o += kTcATmem ? 0 : 2 * (size_t)kTcM * 64;
    S.a_lo = base + o; o += kTcATmem ? 0 : 2 * (size_t)kTcM * 64;
    S.b_hi = base + o; o += 2 * (size_t)nc * 64;
    S.b_lo = base + o;
    return S;
}

// pair column c = a (a + 1) / 2 + b, b <= a
__device__ __forceinline__ void col_to_pair(int c, int& a, int& b) {
    a = (int)((sqrtf(8.0f * (float)c + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > c) --a;
    while ((a + 1) * (a + 2) / 2 <= c) ++a;
    b = c - a * (a + 1) / 2;
}

// Warp roles.  16 generator warps turn observation tiles into operand tiles (W: taper weights of 128 grid points x 32
// observations, Z: the pair products of this CTA's columns x 32 observations; bf16 hi / lo); one warp issues the MMAs; four
// loader warps each filter an interleaved quarter of the candidate observations of the block (the sum over observations
// does not depend on their order) and load the surviving rows into their own tile buffer.  The pipelines are coupled by
// mbarriers only: there is no CTA-wide barrier between set-up and the epilogue.
__global__ void __launch_bounds__(kTcThreads, 1) k_tc_gram(const TcParams P) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    const int kp = P.kp, nc = P.nc;
    const int n_load = P.n_load;
    const TcSmem S = tc_carve(smem_dyn, kp, nc, n_load);
    BlockHeader<kTcM>& H = *S.H;
    const LetkfParams& L = P.L;
    const Geometry& g = L.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blk = L.block_begin + (int)(blockIdx.x / P.n_chunks);
    const int chunk = (int)(blockIdx.x % P.n_chunks);
    const int k1 = L.k + 1;                                       // rows of the augmented [Yn; d]
    uint64_t* op_full = S.bars;
    uint64_t* op_free = S.bars + 2;
    uint64_t* y_full = S.bars + 4;
    uint64_t* y_free = S.bars + 4 + kTcYStages;
    uint64_t* all_done = S.bars + 4 + 2 * kTcYStages;

    // ---- one-time set-up: candidate runs of the block, barriers, tensor memory ----------------------------------------
    setup_block<kTcM>(H, g, L.gpos, L.block_off, L.cell_start, L.n_obs, L.cut_pad, blk, L.status);
    if (tid == 0) {
        mbar_init(&op_full[0], kTcGenWarps); mbar_init(&op_full[1], kTcGenWarps);
        mbar_init(&op_free[0], 1); mbar_init(&op_free[1], 1);
        for (int i = 0; i < kTcYStages; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_free[i], kTcGenWarps); }
        mbar_init(all_done, 1);
        S.op_last[0] = 0; S.op_last[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i <= kTcTab; i += kTcThreads) {          // weight table (FP64 evaluation, once per CTA)
        const double dq = (double)P.q_max / kTcTab;
        const double w0 = i < kTcTab ? tab_weight(g, i * dq) : 0.0;
        const double w1 = i + 1 < kTcTab ? tab_weight(g, (i + 1) * dq) : 0.0;
        S.wtab[i] = make_float2((float)w0, (float)(w1 - w0));
    }
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(S.tmem_slot)), "r"((uint32_t)kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *S.tmem_slot;
    const int ng = H.ng;
    const uint32_t a_lbo = kTcM * 16, b_lbo = (uint32_t)nc * 16;
    const float far_rel = kTcFar;                                   // padding coordinate, see pair_weight_f32
    int n_tiles = 0;

    if (warp >= kTcLoadWarp + n_load) {
        // spare loader warp (fewer tile buffers than loader warps fit in shared memory): nothing to do
    } else if (warp >= kTcLoadWarp) {
        // =================================================================================================================
        // loaders: candidate filter (bounding sphere of the block) -> private ring of surviving observation slots -> tiles
        // =================================================================================================================
        const int lw = warp - kTcLoadWarp;                        // loader index = tile buffer
        const float* __restrict__ ys = reinterpret_cast<const float*>(L.ys);
        const int cand_total = H.cand_total, n_runs = H.n_runs;
        const double bcx = H.cx, bcy = H.cy, bcz = H.cz;
        const double reach = (L.cut_pad + H.rb) * (1.0 + 1e-12);
        const int chr = kp >> 2;                                  // 16-byte chunks per staged row
        constexpr int kPass = kTcFilterUnroll * 32, kRingW = 256;  // private ring: power of two >= kTcObs - 1 + kPass
        static_assert(kRingW * kTcYStages <= kRing && kTcObs - 1 + kPass <= kRingW, "loader rings do not fit");
        int* ring = H.ring + lw * kRingW;
        int cand_pos = lw * kPass, head = 0, tail = 0;
        for (int u = 0;; ++u) {
            while (tail - head < kTcObs && cand_pos < cand_total) {
                int s[kTcFilterUnroll];
                Pos4 po[kTcFilterUnroll];
#pragma unroll
                for (int q = 0; q < kTcFilterUnroll; ++q) {
                    const int c = cand_pos + q * 32 + lane;
                    s[q] = -1;
                    if (c < cand_total) {
                        int lo_ = 0, hi_ = n_runs;
                        while (hi_ - lo_ > 1) {
                            const int mid = (lo_ + hi_) >> 1;
                            if (H.run_pref[mid] <= c) lo_ = mid; else hi_ = mid;
                        }
                        s[q] = H.run_start[lo_] + (c - H.run_pref[lo_]);
                    }
                }
#pragma unroll
                for (int q = 0; q < kTcFilterUnroll; ++q) po[q] = L.opos[max(s[q], 0)];
#pragma unroll
                for (int q = 0; q < kTcFilterUnroll; ++q) {
                    const bool keep = s[q] >= 0 && bin_distance(g, bcx, bcy, bcz, po[q].x, po[q].y, po[q].z) <= reach;
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) ring[(tail + __popc(bal & ((1u << lane) - 1u))) & (kRingW - 1)] = s[q];
                    tail += __popc(bal);
                }
                cand_pos += n_load * kPass;
                __syncwarp();
            }
            const int n = min(kTcObs, tail - head);
            if (u >= 1) mbar_wait(&y_free[lw], (uint32_t)((u - 1) & 1));
            if (n == 0) {
                if (lane == 0) { S.ymeta[lw] = 0; mbar_arrive(&y_full[lw]); }
                break;
            }
            const int slot = ring[(head + (lane < n ? lane : 0)) & (kRingW - 1)];
            const float* row = ys + (size_t)slot * kp;
            float* yt = S.ytile + (size_t)lw * kp * kTcYLd + lane;
            const Pos4 p = L.opos[slot];
            for (int q0 = 0; q0 < chr; q0 += 16) {                 // 16 independent 16-byte loads in flight per lane
                float4 v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q0 + q < chr) v[q] = *reinterpret_cast<const float4*>(row + (q0 + q) * 4);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q0 + q < chr) {
                        float* dst = yt + (size_t)(q0 + q) * 4 * kTcYLd;
                        dst[0] = v[q].x; dst[kTcYLd] = v[q].y; dst[2 * kTcYLd] = v[q].z; dst[3 * kTcYLd] = v[q].w;
                    }
            }
            {
                float4 o = make_float4(far_rel, far_rel, far_rel, 0.f);   // padding: out of reach of every grid point -> weight 0
                if (lane < n) o = make_float4((float)(p.x - bcx), (float)(p.y - bcy), (float)(p.z - bcz), 1.0f);
                S.otile[lw * kTcObs + lane] = o;
            }
            head += n;
            __syncwarp();
            if (lane == 0) { S.ymeta[lw] = n; mbar_arrive(&y_full[lw]); }
        }
    } else if (warp == kTcMmaWarp) {
        // =================================================================================================================
        // MMA issuer: per tile 2 K steps x 2 column halves x (hi hi + hi lo + lo hi)
        // =================================================================================================================
        const uint32_t idesc = tc_idesc_bf16(nc >> 1);
        const uint64_t da_hi0 = tc_smem_desc(smem_u32(S.a_hi), a_lbo, 128), da_lo0 = tc_smem_desc(smem_u32(S.a_lo), a_lbo, 128);
        const uint64_t db_hi0 = tc_smem_desc(smem_u32(S.b_hi), b_lbo, 128), db_lo0 = tc_smem_desc(smem_u32(S.b_lo), b_lbo, 128);
        int t = 0;
        for (;; ++t) {
            const int st = t & 1;
            mbar_wait(&op_full[st], (uint32_t)((t >> 1) & 1));
            if (S.op_last[st]) break;
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int ks = 0; ks < kTcObs / 16; ++ks) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // descriptor start addresses advance in 16-byte units inside the 14-bit address field
                        const uint32_t aoff = ((uint32_t)st * kTcM * 64 + (uint32_t)ks * 2 * a_lbo) >> 4;
                        const uint32_t boff = ((uint32_t)st * (uint32_t)nc * 64 + (uint32_t)ks * 2 * b_lbo + (uint32_t)h * (uint32_t)(nc >> 1) * 16) >> 4;
                        const uint32_t d = tmem + (uint32_t)h * (uint32_t)(nc >> 1);
                        if constexpr (kTcATmem) {
                            // A columns: stage * 32 + {hi: 0, lo: 16} + K step * 8 behind the accumulator columns
                            const uint32_t a_hi = tmem + (uint32_t)kTcMaxCols + (uint32_t)st * 32 + (uint32_t)ks * 8;
                            tc_mma_bf16_ts(d, a_hi, db_hi0 + boff, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                            tc_mma_bf16_ts(d, a_hi, db_lo0 + boff, idesc, 1u);
                            tc_mma_bf16_ts(d, a_hi + 16, db_hi0 + boff, idesc, 1u);
                        } else {
                            tc_mma_bf16(d, da_hi0 + aoff, db_hi0 + boff, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                            tc_mma_bf16(d, da_hi0 + aoff, db_lo0 + boff, idesc, 1u);
                            tc_mma_bf16(d, da_lo0 + aoff, db_hi0 + boff, idesc, 1u);
                        }
                    }
                }
                tc_commit(&op_free[st]);
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (t > 0) tc_commit(all_done); else mbar_arrive(all_done);
        }
    } else {
        // =================================================================================================================
        // generators
        // =================================================================================================================
        const int my_g = tid & (kTcM - 1), my_kc = tid >> 7;      // 128 grid points x 4 chunks of 8 observations
        float gxr = -far_rel, gyr = -far_rel, gzr = -far_rel;
        const bool g_ok = my_g < ng;
        if (g_ok) {
            gxr = (float)(H.gp[my_g].x - H.cx); gyr = (float)(H.gp[my_g].y - H.cy); gzr = (float)(H.gp[my_g].z - H.cz);
        }
        const int my_col = chunk * nc + tid;
        const bool c_ok = tid < nc && my_col < P.n_cols;
        int ca = 0, cb = 0;
        if (c_ok) col_to_pair(my_col, ca, cb);
        const float ncz = c_ok ? -P.centre[my_col] : 0.f;            // minus the centring constant of this thread's pair column
        double wsum_acc = 0.0;                                      // sum of this thread's taper weights (grid point my_g)
        const int dist_kind = g.metric == B200DA_METRIC_PERIODIC1D ? kTcPeriodic
                            : g.metric == B200DA_METRIC_ABS1D ? kTcAbs : kTcSq3;
        const float xs = (float)kTcTab / P.q_max, period = P.period, r_scale = P.r_scale, eps = P.eps;
        const float2* wtab = S.wtab;
        int wmode = dist_kind;
        if (P.w_closed) {
            const int ck = g.metric == B200DA_METRIC_HAVERSINE ? (P.asin_poly ? kTcHavPoly : kTcHavAsin)
                         : g.metric == B200DA_METRIC_EUCLID ? kTcEuclid
                         : g.metric == B200DA_METRIC_PERIODIC1D ? kTcPeriodicF : kTcAbsF;
            wmode = 8 + ck * 2 + (g.taper == B200DA_TAPER_GCINF ? 1 : 0);
        }
        unsigned alive = (1u << n_load) - 1u, par = 0u;       // loaders still producing; phase parity of their buffers
        int t = 0;                                                  // operand tiles produced so far
        for (int yst = -1; alive != 0u;) {
            if (++yst == n_load) yst = 0;
            if (!((alive >> yst) & 1u)) continue;
            mbar_wait(&y_full[yst], (par >> yst) & 1u);
            if (S.ymeta[yst] == 0) { alive &= ~(1u << yst); continue; }      // this loader has run out of candidates
            par ^= 1u << yst;
            const int st = t & 1;
            if (t >= 2) mbar_wait(&op_free[st], (uint32_t)(((t >> 1) - 1) & 1));
            // ---- W tile: taper weights of this thread's grid point for 8 observations -----------------------------------
            {
                const float4* ot = S.otile + yst * kTcObs + my_kc * 8;
                uint4 hi, lo;
                float ws8;
                switch (wmode) {                                   // uniform over the launch
                    case 0: w_chunk<kTcSq3>(ot, wtab, xs, period, gxr, gyr, gzr, hi, lo, ws8); break;
                    case 1: w_chunk<kTcAbs>(ot, wtab, xs, period, gxr, gyr, gzr, hi, lo, ws8); break;
                    case 2: w_chunk<kTcPeriodic>(ot, wtab, xs, period, gxr, gyr, gzr, hi, lo, ws8); break;
#define B200DA_TC_W(D, T) case 8 + (D) * 2 + (T): w_chunk_closed<D, T>(ot, r_scale, eps, period, gxr, gyr, gzr, hi, lo, ws8); break;
                    B200DA_TC_W(kTcHavPoly, 0) B200DA_TC_W(kTcHavPoly, 1) B200DA_TC_W(kTcHavAsin, 0) B200DA_TC_W(kTcHavAsin, 1)
                    B200DA_TC_W(kTcEuclid, 0) B200DA_TC_W(kTcEuclid, 1) B200DA_TC_W(kTcAbsF, 0) B200DA_TC_W(kTcAbsF, 1)
                    B200DA_TC_W(kTcPeriodicF, 0)
                    default: w_chunk_closed<kTcPeriodicF, 1>(ot, r_scale, eps, period, gxr, gyr, gzr, hi, lo, ws8); break;
#undef B200DA_TC_W
                }
                wsum_acc += (double)ws8;
                if constexpr (kTcATmem) {
                    // this thread's row (grid point) = its tensor-memory lane; 8 bf16 = 4 columns at K offset my_kc * 8
                    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)kTcMaxCols + (uint32_t)st * 32 +
                                           (uint32_t)(my_kc >> 1) * 8 + (uint32_t)(my_kc & 1) * 4;
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                                 :: "r"(taddr), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                                 :: "r"(taddr + 16), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
                } else {
                    const size_t off = (size_t)st * kTcM * 64 + (size_t)my_kc * a_lbo + (size_t)my_g * 16;
                    *reinterpret_cast<uint4*>(S.a_hi + off) = hi;
                    *reinterpret_cast<uint4*>(S.a_lo + off) = lo;
                }
            }
            // ---- Z tile: products y_a y_b of this thread's pair column (loads of chunk kc + 1 ahead of the stores of kc) --------
            if (c_ok) {
                const float* yt = S.ytile + (size_t)yst * kp * kTcYLd;
                const float4* ya = reinterpret_cast<const float4*>(yt + ca * kTcYLd);
                const float4* yb = reinterpret_cast<const float4*>(yt + cb * kTcYLd);
                unsigned char* zh = S.b_hi + (size_t)st * nc * 64 + (size_t)tid * 16;
                unsigned char* zl = S.b_lo + (size_t)st * nc * 64 + (size_t)tid * 16;
                float4 a0 = ya[0], a1 = ya[1], b0 = yb[0], b1 = yb[1];
#pragma unroll
                for (int kc = 0; kc < kTcObs / 8; ++kc) {
                    const float z[8] = {fmaf(a0.x, b0.x, ncz), fmaf(a0.y, b0.y, ncz), fmaf(a0.z, b0.z, ncz), fmaf(a0.w, b0.w, ncz),
                                        fmaf(a1.x, b1.x, ncz), fmaf(a1.y, b1.y, ncz), fmaf(a1.z, b1.z, ncz), fmaf(a1.w, b1.w, ncz)};
                    if (kc + 1 < kTcObs / 8) { a0 = ya[2 * kc + 2]; a1 = ya[2 * kc + 3]; b0 = yb[2 * kc + 2]; b1 = yb[2 * kc + 3]; }
                    uint4 hi, lo;
                    split8(z, hi, lo);
                    *reinterpret_cast<uint4*>(zh + (size_t)kc * b_lbo) = hi;
                    *reinterpret_cast<uint4*>(zl + (size_t)kc * b_lbo) = lo;
                }
            }
            // ---- publish the operand stage to the tensor core, release the observation tile -----------------------------------
            fence_async_smem();
            if constexpr (kTcATmem) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&op_full[st]); mbar_arrive(&y_free[yst]); }
            ++t;
        }
        {   // end marker for the MMA warp in the next operand stage
            const int st = t & 1;
            if (t >= 2) mbar_wait(&op_free[st], (uint32_t)(((t >> 1) - 1) & 1));
            if (tid == 0) S.op_last[st] = 1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&op_full[st]);
            n_tiles = t;
        }

        // ---- epilogue: tensor memory -> FP64 tile-packed Gram scratch of the solve kernel ------------------------------------
        mbar_wait(all_done, 0);
        tc_fence_after();
        // sum of the weights of every grid point: four partial sums (observation chunks) through the idle operand stage
        double* wpart = reinterpret_cast<double*>(S.b_hi);
        wpart[my_kc * kTcM + my_g] = wsum_acc;
        asm volatile("bar.sync 1, %0;" :: "n"(kTcGenThreads) : "memory");
        const int lq = warp & 3, cw = warp >> 2;                  // TMEM lane quarter of this warp, column group
        const int gi = lq * 32 + lane;
        const double wsum = (wpart[gi] + wpart[kTcM + gi]) + (wpart[2 * kTcM + gi] + wpart[3 * kTcM + gi]);
        const int64_t slot = (int64_t)L.block_off[blk] + gi - L.slot_base;
        const int kt = (k1 + 7) >> 3;
        double* C = L.cmat + (size_t)(gi < ng ? slot : 0) * (size_t)(tri_tiles(kt) * 64);
        for (int cb16 = cw; cb16 < (nc >> 4); cb16 += kTcGenThreads / 128) {
            uint32_t v[16];
            if (n_tiles > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(cb16 * 16);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0u;
            }
            const int col = chunk * nc + cb16 * 16;
            if (gi < ng && col < P.n_cols) {
                int a, b;
                col_to_pair(col, a, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (col + i < P.n_cols) C[sym_off(a, b)] = fma((double)__ldg(P.centre + col + i), wsum, (double)__uint_as_float(v[i]));
                    if (++b > a) { ++a; b = 0; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kTcTmemCols) : "memory");
    }
}

}  // namespace b200da
