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
// relative).  The taper is evaluated in FP32 on positions relative to the block centre, with the outer Gaspari-Cohn
// branch rewritten in s = 2 - r, w = s^4 (5/8 - s/2 + s^2/12) / r, which has no cancellation (the reference form sums
// terms of magnitude 10 to get 1e-5).  The k x k solve and the update stay in FP64 (ns_solve_kernel.cuh).
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
constexpr int kTcThreads = 512;
constexpr int kTcYLd = 36;         // row length of the member-major observation tile: 32 obs + 4 (conflict-free LDS.128)
constexpr int kTcMaxCols = 512;    // tensor-memory columns = accumulator columns per CTA
constexpr int kTcMaxYItems = 3;    // 16-byte chunks of the observation tile per thread (k + 1 <= 136)

struct TcParams {
    LetkfParams L;
    int n_cols;        // (k + 1)(k + 2) / 2
    int n_chunks;      // column chunks = CTAs per grid-point block
    int nc;            // columns per chunk, multiple of 32, <= 512
    int kp;            // row length of the staging copy ys (floats)
    float r_scale;     // distance -> r:  1 / radius  (haversine: 2 R / radius, applied to asin(chord / 2))
    float eps;
    float period;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
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
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = __bfloat162float(__float2bfloat16_rn(x[i]));
        l[i] = x[i] - h[i];
    }
    hi = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    lo = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
}

// ---- FP32 tapers ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float taper_gc_f32(float r) {          // gaspari_cohn.py:78-95, outer branch in s = 2 - r
    if (r < 1.0f) {
        const float r2 = r * r;
        return fmaf(r2, fmaf(r, fmaf(r, fmaf(r, -0.25f, 0.5f), 0.625f), -5.0f / 3.0f), 1.0f);
    }
    if (r < 2.0f) {
        const float s = 2.0f - r, s2 = s * s;
        return __fdividef(s2 * s2 * fmaf(s, fmaf(s, 1.0f / 12.0f, -0.5f), 0.625f), r);
    }
    return 0.0f;
}
__device__ __forceinline__ float taper_gcinf_f32(float r) {       // gaspari_cohn.py:172-210, last branch in s = 2 - r
    if (r < 0.5f) {
        const float r2 = r * r;
        return fmaf(r2, fmaf(r, fmaf(r, fmaf(r, -28.0f / 33.0f, 8.0f / 11.0f), 20.0f / 11.0f), -80.0f / 33.0f), 1.0f);
    }
    if (r < 1.0f) {
        const float poly = fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, 20.0f / 33.0f, -16.0f / 11.0f), 0.0f), 100.0f / 33.0f),
                                        -45.0f / 11.0f), 51.0f / 22.0f);
        return poly - __fdividef(7.0f / 44.0f, r);
    }
    if (r < 1.5f) {
        const float poly = fmaf(r, fmaf(r, fmaf(r, fmaf(r, fmaf(r, -4.0f / 11.0f, 16.0f / 11.0f), -10.0f / 11.0f),
                                               -100.0f / 33.0f), 5.0f), -61.0f / 22.0f);
        return poly + __fdividef(115.0f / 132.0f, r);
    }
    if (r < 2.0f) {
        const float s = 2.0f - r, s2 = s * s;
        return __fdividef(s2 * s2 * fmaf(s, fmaf(s, 4.0f / 33.0f, -8.0f / 11.0f), 10.0f / 11.0f), r);
    }
    return 0.0f;
}

// localization weight of a pair from positions relative to the block centre (bin space)
__device__ __forceinline__ float pair_weight_f32(int metric, int taper, float r_scale, float eps, float period,
                                                 float gx, float gy, float gz, float ox, float oy, float oz) {
    const float dx = ox - gx, dy = oy - gy, dz = oz - gz;
    float r;
    if (metric == B200DA_METRIC_HAVERSINE) {
        const float h = fminf(0.5f * sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz))), 1.0f);
        r = r_scale * asinf(h);
    } else if (metric == B200DA_METRIC_EUCLID) {
        r = r_scale * sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    } else {
        float d = fabsf(dz);
        if (metric == B200DA_METRIC_PERIODIC1D) d = fminf(d, period - d);
        r = r_scale * d;
    }
    const float w = taper == B200DA_TAPER_GCINF ? taper_gcinf_f32(r) : taper_gc_f32(r);
    return w > eps ? w : 0.0f;                                     // gaspari_cohn.py:135
}

// ---- shared-memory carve-up ------------------------------------------------------------------------------------------------
struct TcSmem {
    BlockHeader<kTcM>* H;
    uint64_t* bars;          // [0], [1]: operand stage free; [2]: all MMAs done
    uint32_t* tmem_slot;
    float4* otile;           // [2][kTcObs] observation positions relative to the block centre, .w = 1 valid / 0 padding
    float* ytile;            // [2][kp][kTcYLd] member-major [Yn; d] of the tile
    unsigned char* a_hi;     // [2][kTcM * 64]
    unsigned char* a_lo;
    unsigned char* b_hi;     // [2][nc * 64]
    unsigned char* b_lo;
};
__host__ __device__ inline size_t tc_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t tc_smem_bytes(int kp, int nc) {
    size_t o = tc_align(sizeof(BlockHeader<kTcM>), 128);
    o += 128;                                             // barriers + tensor-memory address
    o += sizeof(float4) * 2 * kTcObs;
    o = tc_align(o + sizeof(float) * 2 * (size_t)kp * kTcYLd, 128);
    o += 2 * 2 * (size_t)kTcM * 64;
    o += 2 * 2 * (size_t)nc * 64;
    return o + 128;                                       // slack for the manual 128-byte alignment of the base
}
__device__ inline TcSmem tc_carve(unsigned char* base, int kp, int nc) {
    TcSmem S;
    size_t o = 0;
    S.H = reinterpret_cast<BlockHeader<kTcM>*>(base);
    o = tc_align(sizeof(BlockHeader<kTcM>), 128);
    S.bars = reinterpret_cast<uint64_t*>(base + o);
    S.tmem_slot = reinterpret_cast<uint32_t*>(base + o + 64);
    o += 128;
    S.otile = reinterpret_cast<float4*>(base + o);
    o += sizeof(float4) * 2 * kTcObs;
    S.ytile = reinterpret_cast<float*>(base + o);
    o = tc_align(o + sizeof(float) * 2 * (size_t)kp * kTcYLd, 128);
    S.a_hi = base + o; o += 2 * (size_t)kTcM * 64;
    S.a_lo = base + o; o += 2 * (size_t)kTcM * 64;
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

__global__ void __launch_bounds__(kTcThreads, 1) k_tc_gram(const TcParams P) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 127) & ~(uintptr_t)127);
    const int kp = P.kp, nc = P.nc;
    const TcSmem S = tc_carve(base, kp, nc);
    BlockHeader<kTcM>& H = *S.H;
    const LetkfParams& L = P.L;
    const Geometry& g = L.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blk = L.block_begin + (int)(blockIdx.x / P.n_chunks);
    const int chunk = (int)(blockIdx.x % P.n_chunks);
    const int k1 = L.k + 1;                                       // rows of the augmented [Yn; d]
    const float* __restrict__ ys = reinterpret_cast<const float*>(L.ys);

    // ---- one-time set-up: candidate runs of the block, barriers, tensor memory ----------------------------------------
    setup_block<kTcM>(H, g, L.gpos, L.block_off, L.cell_start, L.n_obs, L.cut_pad, blk);
    if (tid == 0) {
        mbar_init(&S.bars[0], 1); mbar_init(&S.bars[1], 1); mbar_init(&S.bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(S.tmem_slot)), "r"((uint32_t)kTcMaxCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *S.tmem_slot;
    const int ng = H.ng;
    const int slot0 = L.block_off[blk];

    // this thread's grid point (W generation) and pair column (Z generation)
    const int my_g = tid & (kTcM - 1), my_kc = tid >> 7;          // 128 grid points x 4 chunks of 8 observations
    float gxr = 0.f, gyr = 0.f, gzr = 0.f;
    const bool g_ok = my_g < ng;
    if (g_ok) {
        gxr = (float)(H.gp[my_g].x - H.cx); gyr = (float)(H.gp[my_g].y - H.cy); gzr = (float)(H.gp[my_g].z - H.cz);
    }
    const int my_col = chunk * nc + tid;
    const bool c_ok = tid < nc && my_col < P.n_cols;
    int ca = 0, cb = 0;
    if (c_ok) col_to_pair(my_col, ca, cb);

    const int cand_total = H.cand_total, n_runs = H.n_runs;
    const double bcx = H.cx, bcy = H.cy, bcz = H.cz;
    const double reach = (L.cut_pad + H.rb) * (1.0 + 1e-12);
    int cand_pos = 0, ring_head = 0, ring_tail = 0;

    auto refill = [&]() {                 // keep at least one tile of surviving candidates in the ring; CTA-uniform
        while (ring_tail - ring_head < kTcObs && cand_pos < cand_total) {
            const int c = cand_pos + tid;
            bool keep = false;
            int s = 0;
            if (c < cand_total) {
                int lo_ = 0, hi_ = n_runs;
                while (hi_ - lo_ > 1) {
                    const int mid = (lo_ + hi_) >> 1;
                    if (H.run_pref[mid] <= c) lo_ = mid; else hi_ = mid;
                }
                s = H.run_start[lo_] + (c - H.run_pref[lo_]);
                const Pos4 po = L.opos[s];
                keep = bin_distance(g, bcx, bcy, bcz, po.x, po.y, po.z) <= reach;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) H.warp_counts[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kTcThreads / 32; ++w) {
                const int cnt = H.warp_counts[w];
                if (w < warp) before += cnt;
                total += cnt;
            }
            if (keep) H.ring[(ring_tail + before + __popc(bal & ((1u << lane) - 1u))) & (kRing - 1)] = s;
            __syncthreads();
            ring_tail += total;
            cand_pos += kTcThreads;
        }
    };

    // global -> registers of one tile (positions by the first 32 threads, observation rows by everybody)
    const int chr = kp >> 2;                                      // 16-byte chunks per staged row
    const int y_items = kTcObs * chr;
    float4 pre_y[kTcMaxYItems];
    float4 pre_o = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int head, int n_tile) {
#pragma unroll
        for (int i = 0; i < kTcMaxYItems; ++i) {
            const int item = tid + i * kTcThreads;
            if (item < y_items) {
                const int j = item & (kTcObs - 1), q = item >> 5;
                const int s = H.ring[(head + (j < n_tile ? j : 0)) & (kRing - 1)];
                pre_y[i] = *reinterpret_cast<const float4*>(ys + (size_t)s * kp + q * 4);
            }
        }
        if (tid < kTcObs) {
            pre_o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < n_tile) {
                const Pos4 po = L.opos[H.ring[(head + tid) & (kRing - 1)]];
                pre_o = make_float4((float)(po.x - bcx), (float)(po.y - bcy), (float)(po.z - bcz), 1.0f);
            }
        }
    };
    auto stash = [&](int buf) {           // registers -> shared memory, observation rows transposed to member-major
        float* yt = S.ytile + (size_t)buf * kp * kTcYLd;
#pragma unroll
        for (int i = 0; i < kTcMaxYItems; ++i) {
            const int item = tid + i * kTcThreads;
            if (item < y_items) {
                const int j = item & (kTcObs - 1), q = item >> 5;
                yt[(q * 4 + 0) * kTcYLd + j] = pre_y[i].x; yt[(q * 4 + 1) * kTcYLd + j] = pre_y[i].y;
                yt[(q * 4 + 2) * kTcYLd + j] = pre_y[i].z; yt[(q * 4 + 3) * kTcYLd + j] = pre_y[i].w;
            }
        }
        if (tid < kTcObs) S.otile[buf * kTcObs + tid] = pre_o;
    };

    const uint32_t idesc = tc_idesc_bf16(nc >> 1);                // two MMAs of N = nc / 2 per K step
    const uint32_t a_lbo = kTcM * 16, b_lbo = (uint32_t)nc * 16;
    int n_tiles = 0;

    refill();
    int n_tile = min(kTcObs, ring_tail - ring_head);
    if (n_tile > 0) {
        fetch(ring_head, n_tile);
        ring_head += n_tile;
        stash(0);
    }
    __syncthreads();

    while (n_tile > 0) {
        const int t = n_tiles, st = t & 1;
        // ---- prefetch the next tile into registers -------------------------------------------------------------------
        refill();
        const int n_next = min(kTcObs, ring_tail - ring_head);
        if (n_next > 0) { fetch(ring_head, n_next); ring_head += n_next; }
        // ---- operand stage free?  (MMAs of tile t - 2 have finished reading it) ----------------------------------------
        if (t >= 2) mbar_wait(&S.bars[st], (uint32_t)(((t >> 1) - 1) & 1));
        // ---- W tile: taper weights of 128 grid points x 32 observations ------------------------------------------------
        {
            float w[8];
            const float4* ot = S.otile + st * kTcObs + my_kc * 8;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 o = ot[jj];
                float v = 0.0f;
                if (g_ok && o.w != 0.0f)
                    v = pair_weight_f32(g.metric, g.taper, P.r_scale, P.eps, P.period, gxr, gyr, gzr, o.x, o.y, o.z);
                w[jj] = v;
            }
            uint4 hi, lo;
            split8(w, hi, lo);
            const size_t off = (size_t)st * kTcM * 64 + (size_t)my_kc * a_lbo + (size_t)my_g * 16;
            *reinterpret_cast<uint4*>(S.a_hi + off) = hi;
            *reinterpret_cast<uint4*>(S.a_lo + off) = lo;
        }
        // ---- Z tile: products y_a y_b of this thread's pair column -------------------------------------------------------
        if (c_ok) {
            const float* yt = S.ytile + (size_t)st * kp * kTcYLd;
            const float* ya = yt + ca * kTcYLd;
            const float* yb = yt + cb * kTcYLd;
#pragma unroll
            for (int kc = 0; kc < kTcObs / 8; ++kc) {
                const float4 a0 = *reinterpret_cast<const float4*>(ya + kc * 8), a1 = *reinterpret_cast<const float4*>(ya + kc * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(yb + kc * 8), b1 = *reinterpret_cast<const float4*>(yb + kc * 8 + 4);
                const float z[8] = {a0.x * b0.x, a0.y * b0.y, a0.z * b0.z, a0.w * b0.w, a1.x * b1.x, a1.y * b1.y, a1.z * b1.z, a1.w * b1.w};
                uint4 hi, lo;
                split8(z, hi, lo);
                const size_t off = (size_t)st * nc * 64 + (size_t)kc * b_lbo + (size_t)tid * 16;
                *reinterpret_cast<uint4*>(S.b_hi + off) = hi;
                *reinterpret_cast<uint4*>(S.b_lo + off) = lo;
            }
        }
        // ---- hand the prefetched tile to the other buffer, publish the operands to the tensor core -----------------------
        if (n_next > 0) stash(st ^ 1);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t ahi = smem_u32(S.a_hi + (size_t)st * kTcM * 64), alo = smem_u32(S.a_lo + (size_t)st * kTcM * 64);
            const uint32_t bhi = smem_u32(S.b_hi + (size_t)st * nc * 64), blo = smem_u32(S.b_lo + (size_t)st * nc * 64);
#pragma unroll
            for (int ks = 0; ks < kTcObs / 16; ++ks) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t boff = (uint32_t)ks * 2 * b_lbo + (uint32_t)h * (uint32_t)(nc >> 1) * 16;
                    const uint64_t da_hi = tc_smem_desc(ahi + ks * 2 * a_lbo, a_lbo, 128), da_lo = tc_smem_desc(alo + ks * 2 * a_lbo, a_lbo, 128);
                    const uint64_t db_hi = tc_smem_desc(bhi + boff, b_lbo, 128), db_lo = tc_smem_desc(blo + boff, b_lbo, 128);
                    const uint32_t d = tmem + (uint32_t)h * (uint32_t)(nc >> 1);
                    tc_mma_bf16(d, da_hi, db_hi, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                    tc_mma_bf16(d, da_hi, db_lo, idesc, 1u);
                    tc_mma_bf16(d, da_lo, db_hi, idesc, 1u);
                }
            }
            tc_commit(&S.bars[st]);
        }
        ++n_tiles;
        n_tile = n_next;
    }

    // ---- epilogue: tensor memory -> FP64 tile-packed Gram scratch of the solve kernel ----------------------------------------
    if (n_tiles > 0) {
        if (tid == 0) tc_commit(&S.bars[2]);
        mbar_wait(&S.bars[2], 0);
        tc_fence_after();
    }
    {
        const int lq = warp & 3, cw = warp >> 2;                  // TMEM lane quarter of this warp, column group
        const int gi = lq * 32 + lane;
        const int64_t slot = (int64_t)slot0 + gi - L.slot_base;
        const int kt = (k1 + 7) >> 3;
        double* C = L.cmat + (size_t)(gi < ng ? slot : 0) * (size_t)(tri_tiles(kt) * 64);
        for (int cb16 = cw; cb16 < (nc >> 4); cb16 += kTcThreads / 128) {
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
            int col = chunk * nc + cb16 * 16;
            if (gi < ng && col < P.n_cols) {
                int a, b;
                col_to_pair(col, a, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (col + i < P.n_cols) C[sym_off(a, b)] = (double)__uint_as_float(v[i]);
                    if (++b > a) { ++a; b = 0; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kTcMaxCols) : "memory");
    }
}

}  // namespace b200da
