// C ABI of libb200da.so (see include/b200da.h): plan management, kernel dispatch, host-buffer convenience path.
#include <cmath>
#include <mutex>
#include "binning.cuh"
#include "etkf_kernel.cuh"
#include "letkf_kernel.cuh"
#include "neighbour_kernel.cuh"
#include "ns_solve_kernel.cuh"
#include "kernelise_kernel.cuh"
#include "ienks_kernel.cuh"
#include "launch.cuh"

namespace b200da {
thread_local std::string g_last_cuda_error;
int64_t g_launch_count = 0;
int g_debug_sync = [] { const char* e = getenv("B200DA_DEBUG_SYNC"); return e ? atoi(e) : 0; }();

// host copies of the tapers (only used to locate the cutoff radius r with w(r) = eps)
static double host_taper(int taper, double r) {
    if (taper == B200DA_TAPER_GCINF) {
        if (r < 0.5) return -28 * std::pow(r, 5) / 33 + 8 * std::pow(r, 4) / 11 + 20 * std::pow(r, 3) / 11 - 80 * r * r / 33 + 1;
        if (r < 1.0) return 20 * std::pow(r, 5) / 33 - 16 * std::pow(r, 4) / 11 + 100 * r * r / 33 - 45 * r / 11 + 51.0 / 22 - 7 / (44 * r);
        if (r < 1.5) return -4 * std::pow(r, 5) / 11 + 16 * std::pow(r, 4) / 11 - 10 * std::pow(r, 3) / 11 - 100 * r * r / 33 + 5 * r - 61.0 / 22 + 115 / (132 * r);
        if (r < 2.0) return 4 * std::pow(r, 5) / 33 - 8 * std::pow(r, 4) / 11 + 10 * std::pow(r, 3) / 11 + 80 * r * r / 33 - 80 * r / 11 + 64.0 / 11 - 32 / (33 * r);
        return 0.0;
    }
    if (r < 1.0) return -0.25 * std::pow(r, 5) + 0.5 * std::pow(r, 4) + 0.625 * std::pow(r, 3) - 5.0 / 3 * r * r + 1;
    if (r < 2.0) return std::pow(r, 5) / 12 - 0.5 * std::pow(r, 4) + 0.625 * std::pow(r, 3) + 5.0 / 3 * r * r - 5 * r + 4 - 2.0 / 3 / r;
    return 0.0;
}

// ---- tabulated taper of the Gram kernels (common.cuh: TaperTab) ---------------------------------------------------------------
// piece `seg` of the taper as an analytic function of r (each piece is evaluated with its own formula also slightly outside
// its interval, so that the interpolation nodes next to a breakpoint never pick the neighbouring piece)
static long double taper_piece(int taper, int seg, long double r) {
    const long double r2 = r * r, r3 = r2 * r, r4 = r3 * r, r5 = r4 * r;
    if (taper == B200DA_TAPER_GCINF) {
        switch (seg) {
            case 0: return -28.0L * r5 / 33 + 8.0L * r4 / 11 + 20.0L * r3 / 11 - 80.0L * r2 / 33 + 1;
            case 1: return 20.0L * r5 / 33 - 16.0L * r4 / 11 + 100.0L * r2 / 33 - 45.0L * r / 11 + 51.0L / 22 - 7.0L / (44 * r);
            case 2: return -4.0L * r5 / 11 + 16.0L * r4 / 11 - 10.0L * r3 / 11 - 100.0L * r2 / 33 + 5 * r - 61.0L / 22 + 115.0L / (132 * r);
            default: return 4.0L * r5 / 33 - 8.0L * r4 / 11 + 10.0L * r3 / 11 + 80.0L * r2 / 33 - 80.0L * r / 11 + 64.0L / 11 - 32.0L / (33 * r);
        }
    }
    if (seg == 0) return -0.25L * r5 + 0.5L * r4 + 0.625L * r3 - 5.0L / 3 * r2 + 1;
    return r5 / 12 - 0.5L * r4 + 0.625L * r3 + 5.0L / 3 * r2 - 5 * r + 4 - 2.0L / 3 / r;
}
// Builds the table for the plan's geometry; leaves pl->tt.coef null (direct evaluation) when the table does not apply: a
// haversine support that reaches beyond a third of the sphere (the chord is a poor coordinate near the antipode), a degenerate
// radius, or B200DA_TAPER_TABLE=0 in the environment.
static int build_taper_table(b200da_plan* pl) {
    const Geometry& g = pl->geom;
    pl->tt = TaperTab{};
    if (const char* e = getenv("B200DA_TAPER_TABLE")) { if (!atoi(e)) return B200DA_OK; }
    if (!(g.radius > 0.0) || !std::isfinite(g.radius)) return B200DA_OK;
    const bool hav = g.metric == B200DA_METRIC_HAVERSINE;
    const int nseg = g.taper == B200DA_TAPER_GCINF ? 4 : 2;
    const long double dr = 2.0L / nseg;
    if (hav && 2.0 * g.radius / g.sphere_r > 2.0 * M_PI / 3.0) return B200DA_OK;
    auto u_of_r = [&](long double r) -> long double {            // bin-space distance at taper argument r
        return hav ? 2.0L * sinl(0.5L * r * (long double)g.radius / (long double)g.sphere_r) : r * (long double)g.radius;
    };
    auto r_of_u = [&](long double u) -> long double {
        return hav ? 2.0L * (long double)g.sphere_r * asinl(0.5L * u) / (long double)g.radius : u / (long double)g.radius;
    };
    TaperTab tt{};
    tt.nseg = nseg;
    tt.nint = 256 / nseg;                                         // 128 intervals per unit of r
    std::vector<double> coef((size_t)nseg * tt.nint * 6);
    for (int sg = 0; sg <= nseg; ++sg) tt.ub[sg] = (double)u_of_r(dr * sg);
    static const long double kPi = 3.14159265358979323846264338327950288L;
    for (int sg = 0; sg < nseg; ++sg) {
        const long double ua = tt.ub[sg], ub = tt.ub[sg + 1];      // the device locates the segment with these rounded values
        if (!(ub > ua)) return B200DA_OK;
        const long double h = (ub - ua) / tt.nint;
        tt.scale[sg] = (double)(1.0L / h);
        for (int i = 0; i < tt.nint; ++i) {
            // the device computes x = (u - ub[s]) * scale[s], i = floor(x), t = 2 (x - i) - 1: interval i is [ua + i / scale, ...)
            const long double inv = 1.0L / (long double)tt.scale[sg];
            const long double mid = ua + (i + 0.5L) * inv, half = 0.5L * inv;
            long double f[6], a[6];
            for (int j = 0; j < 6; ++j) {
                const long double t = cosl((2 * j + 1) * kPi / 12);
                long double r = r_of_u(mid + half * t);
                if (r < 1e-30L) r = 1e-30L;
                f[j] = taper_piece(g.taper, sg, r);
            }
            for (int kk = 0; kk < 6; ++kk) {                      // Chebyshev coefficients of the interpolant
                long double acc = 0;
                for (int j = 0; j < 6; ++j) acc += f[j] * cosl(kk * (2 * j + 1) * kPi / 12);
                a[kk] = acc * (kk == 0 ? 1.0L / 6 : 2.0L / 6);
            }
            // T0..T5 -> monomials in t
            double* c = &coef[((size_t)sg * tt.nint + i) * 6];
            c[0] = (double)(a[0] - a[2] + a[4]);
            c[1] = (double)(a[1] - 3 * a[3] + 5 * a[5]);
            c[2] = (double)(2 * a[2] - 8 * a[4]);
            c[3] = (double)(4 * a[3] - 20 * a[5]);
            c[4] = (double)(8 * a[4]);
            c[5] = (double)(16 * a[5]);
        }
    }
    int rc = pl->taper_tab.ensure(sizeof(double) * coef.size());
    if (rc) return rc;
    B200DA_CUDA(cudaMemcpy(pl->taper_tab.p, coef.data(), sizeof(double) * coef.size(), cudaMemcpyHostToDevice));
    tt.coef = pl->taper_tab.as<double>();
    pl->tt = tt;
    return B200DA_OK;
}

// smallest r (padded) beyond which the taper never exceeds eps: the tapers decrease monotonically on (0, 2)
static double cutoff_radius(int taper, double eps) {
    if (!(eps > 0.0)) return 2.0;
    if (eps >= 1.0) return 1e-6;
    double lo = 0.0, hi = 2.0;
    for (int i = 0; i < 200; ++i) {
        const double mid = 0.5 * (lo + hi);
        if (host_taper(taper, mid) > eps) lo = mid; else hi = mid;
    }
    return std::min(2.0, hi * (1.0 + 1e-7) + 1e-9);
}

// tile-packed augmented Gram of the chunk's grid slots -> dense (N, k+1, k+1) in original grid order (lower triangle)
__global__ void k_unpack_gram(const double* __restrict__ cmat, const Pos4* __restrict__ gpos, int64_t slot_base, int64_t n_slots,
                              int64_t slot_stride, int k1, double* __restrict__ out) {
    for (int64_t s = blockIdx.x; s < n_slots; s += gridDim.x) {
        const double* C = cmat + (size_t)s * (size_t)slot_stride;
        double* o = out + (size_t)gpos[slot_base + s].id * (size_t)k1 * k1;
        for (int e = threadIdx.x; e < k1 * k1; e += blockDim.x) {
            const int r = e / k1, c = e - r * k1;
            o[e] = (c <= r && !(r == k1 - 1 && c == k1 - 1)) ? C[sym_off(r, c)] : 0.0;
        }
    }
}

struct KernelConfig { int g, wpg; };
static KernelConfig config_for_kt(int kt) {
    if (kt <= 5) return {8, 1};
    if (kt <= 7) return {8, 2};
    if (kt <= 10) return {4, 4};
    return {2, 8};
}
constexpr int kMaxKt = 17;

// name of the DMMA Gram variant that will run for the plan (dispatch_fused) and its number of DFMA rows
static void name_dmma_kernel(b200da_plan* pl) {
    const KernelConfig cfg = config_for_kt(pl->kt);
    const int er = gram_extra_rows(pl);
    pl->gram_er = er;
    pl->kernel_name = std::string("letkf_gram_") + (pl->dtype == B200DA_F32 ? "f32in_f64dmma" : "f64") +
                      (er > 0 ? "_er" + std::to_string(er) : std::string()) + "_kt" + std::to_string(er > 0 ? pl->kt - 1 : pl->kt) +
                      "_g" + std::to_string(cfg.g) + "_w" + std::to_string(cfg.wpg) + (pl->kprog.n > 0 ? "+kernelise" : "");
}

template <int KT>
static int launch_etkf_gram(int f32, const void* yn, const void* d, int64_t m, int64_t ld, int k, int ncta, int64_t chunk,
                            double* partial, cudaStream_t st) {
    const size_t smem = sizeof(double) * (size_t)KT * 8 * kEtkfLd;
    if (f32) {
        B200DA_CUDA(cudaFuncSetAttribute(k_etkf_gram<float, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_etkf_gram<float, KT><<<ncta, kEtkfWarps * 32, smem, st>>>((const float*)yn, (const float*)d, m, ld, k, chunk, partial);
    } else {
        B200DA_CUDA(cudaFuncSetAttribute(k_etkf_gram<double, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_etkf_gram<double, KT><<<ncta, kEtkfWarps * 32, smem, st>>>((const double*)yn, (const double*)d, m, ld, k, chunk, partial);
    }
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}
#define B200DA_EG_CASE(KT) case KT: return launch_etkf_gram<KT>(f32, yn, d, m, ld, k, ncta, chunk, partial, st);
static int dispatch_etkf_gram(int kt, int f32, const void* yn, const void* d, int64_t m, int64_t ld, int k, int ncta,
                              int64_t chunk, double* partial, cudaStream_t st) {
    switch (kt) {
        B200DA_EG_CASE(1) B200DA_EG_CASE(2) B200DA_EG_CASE(3) B200DA_EG_CASE(4) B200DA_EG_CASE(5) B200DA_EG_CASE(6)
        B200DA_EG_CASE(7) B200DA_EG_CASE(8) B200DA_EG_CASE(9) B200DA_EG_CASE(10) B200DA_EG_CASE(11) B200DA_EG_CASE(12)
        B200DA_EG_CASE(13) B200DA_EG_CASE(14) B200DA_EG_CASE(15) B200DA_EG_CASE(16) B200DA_EG_CASE(17)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

template <int MT>
static int launch_apply_global(int f32, const void* x, const void* w, int k, int n_rows, int64_t n_grid, int64_t ld, void* xa,
                               const ApplyPeers& peers, cudaStream_t st) {
    const size_t esz = f32 ? 4 : 8;
    const int vec_ok = ((ld * (int64_t)esz) % (2 * esz) == 0 && (reinterpret_cast<uintptr_t>(xa) % (2 * esz)) == 0) ? 1 : 0;
    const size_t smem = sizeof(double) * (size_t)MT * 8 * apply_lda(k);
    if (smem > kMaxSmem - 2048) return B200DA_ERR_UNSUPPORTED;
    const int ntw = f32 ? apply_ntw<float>(MT) : apply_ntw<double>(MT);
    const int64_t n_chunks = (n_grid + ntw * 8 - 1) / (ntw * 8);
    int occ = 1;
    if (f32) {
        B200DA_CUDA(cudaFuncSetAttribute(k_apply_global<float, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B200DA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_apply_global<float, MT>, kApplyWarps * 32, smem));
    } else {
        B200DA_CUDA(cudaFuncSetAttribute(k_apply_global<double, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B200DA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_apply_global<double, MT>, kApplyWarps * 32, smem));
    }
    const int grid = (int)std::min<int64_t>((n_chunks + kApplyWarps - 1) / kApplyWarps, 148 * (int64_t)std::max(occ, 1));
    if (f32) {
        k_apply_global<float, MT><<<grid, kApplyWarps * 32, smem, st>>>((const float*)x, (const float*)w, k, n_rows, n_grid, ld, vec_ok, (float*)xa, peers);
    } else {
        k_apply_global<double, MT><<<grid, kApplyWarps * 32, smem, st>>>((const double*)x, (const double*)w, k, n_rows, n_grid, ld, vec_ok, (double*)xa, peers);
    }
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}
#define B200DA_AG_CASE(MT) case MT: return launch_apply_global<MT>(f32, x, w, k, n_rows, n_grid, ld, xa, peers, st);
static int dispatch_apply_global(int f32, const void* x, const void* w, int k, int n_rows, int64_t n_grid, int64_t ld, void* xa,
                                 const ApplyPeers& peers, cudaStream_t st) {
    switch ((k + 7) / 8) {
        B200DA_AG_CASE(1) B200DA_AG_CASE(2) B200DA_AG_CASE(3) B200DA_AG_CASE(4) B200DA_AG_CASE(5) B200DA_AG_CASE(6)
        B200DA_AG_CASE(7) B200DA_AG_CASE(8) B200DA_AG_CASE(9) B200DA_AG_CASE(10) B200DA_AG_CASE(11) B200DA_AG_CASE(12)
        B200DA_AG_CASE(13) B200DA_AG_CASE(14) B200DA_AG_CASE(15) B200DA_AG_CASE(16)
        default: return B200DA_ERR_UNSUPPORTED;
    }
}

// spectral-bound ratio s / a above which the solve kernel switches to its two-level path (ns_solve_kernel.cuh): the
// single-level error is ~1e-12 at 2e3 (FP64 tolerance 1e-10) and ~1e-8 at 2e4 (FP32 tolerance 1e-4)
static double ns_stiff_ratio(const b200da_plan* pl) {
    if (pl->ns_stiff > 0.0) return pl->ns_stiff;
    return pl->dtype == B200DA_F32 ? 2.0e4 : 2.0e3;
}

static int check_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return B200DA_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return B200DA_ERR_NO_DEVICE; }
    if (prop.major != 10) return B200DA_ERR_NO_DEVICE;
    return B200DA_OK;
}

static NeighbourParams neighbour_params(const b200da_plan* pl) {
    NeighbourParams P{};
    P.g = pl->geom;
    P.gpos = pl->gpos.as<Pos4>();
    P.block_off = pl->block_off.as<int>();
    P.opos = pl->opos.as<Pos4>();
    P.gext = pl->gext.as<double>(); P.oext = pl->oext.as<double>();
    P.cell_start = pl->cell_start.as<int>();
    P.n_obs = pl->n_obs;
    P.cut_pad = pl->geom.cut_bin * (1.0 + 1e-9) + 1e-300;
    P.status = pl->devstat.as<PlanStatus>();
    P.over = pl->over_list.as<PairRec>(); P.n_over = pl->n_over;
    return P;
}

template <int MODE>
static int launch_neighbours(const b200da_plan* pl, const NeighbourParams& P, cudaStream_t st) {
    if (pl->n_blocks == 0) return B200DA_OK;
    switch (pl->gpb) {
        case 128: k_neighbours<128, MODE><<<(int)pl->n_blocks, 256, 0, st>>>(P); break;
        case 8: k_neighbours<8, MODE><<<(int)pl->n_blocks, 256, 0, st>>>(P); break;
        case 4: k_neighbours<4, MODE><<<(int)pl->n_blocks, 256, 0, st>>>(P); break;
        default: k_neighbours<2, MODE><<<(int)pl->n_blocks, 256, 0, st>>>(P); break;
    }
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

}  // namespace b200da

using namespace b200da;

extern "C" {

int b200da_version(void) { return 100; }
int64_t b200da_launch_count(void) { return g_launch_count; }
const char* b200da_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

const char* b200da_strerror(int status) {
    switch (status) {
        case B200DA_OK: return "ok";
        case B200DA_ERR_INVALID: return "invalid argument";
        case B200DA_ERR_SIZE: return "observational size of ensemble perturbations and observations do not match";
        case B200DA_ERR_UNSUPPORTED: return "metric / taper / dtype / ensemble size not supported by the B200 engine";
        case B200DA_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
        case B200DA_ERR_CUDA: return "CUDA runtime error";
        case B200DA_ERR_STATE: return "call order: set_grid and bin_obs must precede this call";
        case B200DA_ERR_NOMEM: return "out of device memory";
        case B200DA_ERR_OVERFLOW: return "a grid-point block needs more candidate cell columns than the kernels hold: results are invalid";
        default: return "unknown status";
    }
}

int b200da_plan_create(b200da_plan** plan, int k, int n_slices, int n_coord, int metric, const double* metric_params,
                       int n_metric_params, const double* radius, int n_radius, double epsilon, double inf_factor,
                       int dtype, int taper) {
    if (!plan) return B200DA_ERR_INVALID;
    *plan = nullptr;
    if (k < 2 || n_slices < 1 || !radius || n_radius < 1 || !(radius[0] > 0.0) || !(inf_factor > 0.0))
        return B200DA_ERR_INVALID;
    if (dtype != B200DA_F64 && dtype != B200DA_F32) return B200DA_ERR_UNSUPPORTED;
    if (taper != B200DA_TAPER_GC && taper != B200DA_TAPER_GCINF) return B200DA_ERR_UNSUPPORTED;
    const int kt = (k + 1 + 7) / 8;
    if (kt > kMaxKt) return B200DA_ERR_UNSUPPORTED;
    int rc = check_device();
    if (rc) return rc;
    b200da_plan* pl = new (std::nothrow) b200da_plan();
    if (!pl) return B200DA_ERR_NOMEM;
    pl->k = k; pl->n_slices = n_slices; pl->n_coord = n_coord; pl->dtype = dtype; pl->rho = inf_factor;
    pl->kt = kt; pl->kp = kt * 8;
    const KernelConfig cfg = config_for_kt(kt);
    pl->gpb = cfg.g;
    pl->use_tc = (dtype == B200DA_F32 && k >= 8);
    name_dmma_kernel(pl);
    if (pl->use_tc) {
        pl->gram_er = 0;
        int n_cols, n_chunks, nc;
        tc_chunking(k, &n_cols, &n_chunks, &nc);
        pl->gpb = 128;
        pl->kernel_name = "letkf_gram_tcgen05_bf16x3_k" + std::to_string(k) + "_m128_n" + std::to_string(nc) + "x" + std::to_string(n_chunks);
    }
    Geometry& g = pl->geom;
    g.metric = metric; g.taper = taper; g.n_coord = n_coord; g.periodic = 0;
    g.radius = radius[0]; g.eps = epsilon; g.period = 0.0; g.sphere_r = 1.0;
    g.rcut = cutoff_radius(taper, epsilon);
    switch (metric) {
        case B200DA_METRIC_ABS1D:
            if (n_coord != 1) { delete pl; return B200DA_ERR_INVALID; }
            g.nd = 1; g.cut_bin = g.rcut * g.radius; break;
        case B200DA_METRIC_PERIODIC1D:
            if (n_coord != 1 || n_metric_params < 1 || !metric_params || !(metric_params[0] > 0.0)) { delete pl; return B200DA_ERR_INVALID; }
            g.nd = 1; g.periodic = 1; g.period = metric_params[0]; g.cut_bin = g.rcut * g.radius; break;
        case B200DA_METRIC_EUCLID:
            if (n_coord < 1 || n_coord > 3) { delete pl; return B200DA_ERR_INVALID; }
            g.nd = n_coord; g.cut_bin = g.rcut * g.radius; break;
        case B200DA_METRIC_HAVERSINE: {
            if (n_coord != 2 || n_metric_params < 1 || !metric_params || !(metric_params[0] > 0.0)) { delete pl; return B200DA_ERR_INVALID; }
            g.nd = 3; g.sphere_r = metric_params[0];
            const double ang = std::min(g.rcut * g.radius / g.sphere_r, M_PI);
            g.cut_bin = 2.0 * std::sin(0.5 * ang) * (1.0 + 1e-9);
            break;
        }
        default: delete pl; return B200DA_ERR_UNSUPPORTED;
    }
    if (cudaEventCreate(&pl->ev0) != cudaSuccess || cudaEventCreate(&pl->ev1) != cudaSuccess) {
        delete pl; return B200DA_ERR_CUDA;
    }
    if (int rc = build_taper_table(pl)) { b200da_plan_destroy(pl); return rc; }
    if (pl->devstat.ensure(sizeof(PlanStatus)) || pl->amb_list.ensure(sizeof(PairRec) * kAmbCapacity) ||
        pl->over_list.ensure(sizeof(PairRec) * 16) || cudaMemset(pl->devstat.p, 0, sizeof(PlanStatus)) != cudaSuccess) {
        b200da_plan_destroy(pl); return B200DA_ERR_NOMEM;
    }
    *plan = pl;
    return B200DA_OK;
}

void b200da_plan_destroy(b200da_plan* pl) {
    if (!pl) return;
    DevBuf* bufs[] = {&pl->gpos, &pl->gorder, &pl->block_off, &pl->opos, &pl->cell_start, &pl->ys, &pl->tmp_keys,
                      &pl->tmp_cell, &pl->tmp_count, &pl->tmp_a, &pl->tmp_b, &pl->tmp_pos, &pl->host_stage_obs,
                      &pl->host_stage_y, &pl->host_stage_d, &pl->host_stage_x, &pl->host_stage_xa, &pl->etkf_partial,
                      &pl->etkf_w, &pl->stats, &pl->cmat, &pl->counter, &pl->ns_scratch, &pl->gext, &pl->oext,
                      &pl->devstat, &pl->amb_list, &pl->over_list, &pl->tc_centre, &pl->taper_tab};
    for (cudaEvent_t ev : pl->ev_pool) cudaEventDestroy(ev);
    for (DevBuf* b : bufs) b->release();
    if (pl->ev0) cudaEventDestroy(pl->ev0);
    if (pl->ev1) cudaEventDestroy(pl->ev1);
    delete pl;
}

int b200da_plan_set_extra(b200da_plan* pl, int n_extra, const double* extra_radius) {
    if (!pl || n_extra < 0 || n_extra > 2 || (n_extra > 0 && !extra_radius)) return B200DA_ERR_INVALID;
    // GaspariCohnInf.localize_obs evaluates a single distance (localization/gaspari_cohn.py:216-254)
    if (n_extra > 0 && pl->geom.taper != B200DA_TAPER_GC) return B200DA_ERR_UNSUPPORTED;
    for (int e = 0; e < n_extra; ++e) if (!(extra_radius[e] > 0.0)) return B200DA_ERR_INVALID;
    pl->geom.n_ext = n_extra;
    for (int e = 0; e < 2; ++e) pl->geom.ext_radius[e] = e < n_extra ? extra_radius[e] : 1.0;
    if (n_extra > 0 && pl->use_tc) {
        // the tcgen05 Gram evaluates its FP32 weights in closed form for one distance: plans with extra components use the
        // DMMA Gram (FP32 tiles converted on load)
        const KernelConfig cfg = config_for_kt(pl->kt);
        pl->use_tc = false; pl->gpb = cfg.g;
        name_dmma_kernel(pl);
    }
    pl->have_grid = false; pl->have_obs = false;
    return B200DA_OK;
}

int b200da_plan_set_kernel(b200da_plan* pl, int n_ops, const int* ops, const double* p0, const double* p1) {
    if (!pl || n_ops < 0 || n_ops > kMaxKernelOps || (n_ops > 0 && (!ops || !p0 || !p1))) return B200DA_ERR_INVALID;
    int depth = 0;                                   // the program must be a well-formed postfix expression
    for (int i = 0; i < n_ops; ++i) {
        if (ops[i] < B200DA_KOP_LINEAR || ops[i] > B200DA_KOP_POW) return B200DA_ERR_UNSUPPORTED;
        if (ops[i] >= B200DA_KOP_ADD) { if (depth < 2) return B200DA_ERR_INVALID; --depth; }
        else if (++depth > kKernelStack) return B200DA_ERR_UNSUPPORTED;
        if ((ops[i] == B200DA_KOP_GAUSS || ops[i] == B200DA_KOP_RATIONAL) && !(p0[i] != 0.0)) return B200DA_ERR_INVALID;
    }
    if (n_ops > 0 && depth != 1) return B200DA_ERR_INVALID;
    // multiples of 8 run the Gram with one more tile row (innovation row inside the tiles): instantiated up to 16 rows
    if (n_ops > 0 && pl->k % 8 == 0 && pl->kt > 16) return B200DA_ERR_UNSUPPORTED;
    pl->kprog = KernelProgram{};
    pl->kprog.n = n_ops;
    for (int i = 0; i < n_ops; ++i) { pl->kprog.op[i] = ops[i]; pl->kprog.p0[i] = p0[i]; pl->kprog.p1[i] = p1[i]; }
    if (n_ops > 0 && pl->use_tc) {
        // the kernel functions are evaluated on FP64 Gram entries: FP32 plans use the DMMA Gram (FP32 tiles converted on load)
        const KernelConfig cfg = config_for_kt(pl->kt);
        pl->use_tc = false; pl->gpb = cfg.g;
        pl->have_grid = false; pl->have_obs = false;
    }
    if (!pl->use_tc) name_dmma_kernel(pl);
    return B200DA_OK;
}

// Gram slots -> centred kernel matrix + centred kernel column of the observations, in place (kernelise.cuh)
static int launch_kernelise(b200da_plan* pl, double* cmat, int64_t n_slots, int64_t slot_stride, cudaStream_t st) {
    const size_t smem = kernelise_smem_bytes(pl->k);
    if (smem > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    B200DA_CUDA(cudaFuncSetAttribute(k_kernelise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_kernelise<<<(int)std::min<int64_t>(n_slots, 148 * 16), 256, smem, st>>>(cmat, n_slots, slot_stride, pl->k, pl->kprog);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_set_grid(b200da_plan* plan, const double* grid_coord, int64_t n_grid, void* stream) {
    if (plan) { plan->slot_of_host.clear(); plan->n_over = 0; }
    return set_grid_impl(plan, grid_coord, n_grid, (cudaStream_t)stream);
}

static int etkf_partial_grams(b200da_plan* pl, const void* Yn, const void* d, int64_t m, int64_t ld, int* ncta_out, cudaStream_t st);

// Centring constants of the tcgen05 Gram (tc_gram_kernel.cuh): c[(a, b)] = mean over ALL observations of y_a y_b, i.e. the
// unlocalized augmented Gram (the global-ETKF Gram kernel, etkf_kernel.cuh) divided by M, in pair-column order.
__global__ void k_tc_centre(const double* __restrict__ partial, int n_partial, int kp, int k, int64_t m, int n_cols,
                            float* __restrict__ centre) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n_cols) return;
    int a = (int)((sqrtf(8.0f * (float)col + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > col) --a;
    while ((a + 1) * (a + 2) / 2 <= col) ++a;
    const int b = col - a * (a + 1) / 2;
    double s = 0.0;
    for (int p = 0; p < n_partial; ++p) s += partial[(size_t)p * kp * kp + a * kp + b];
    centre[col] = m > 0 ? (float)(s / (double)m) : 0.f;
}
static int tc_centre_constants(b200da_plan* pl, const void* Yn, const void* d, int64_t m, cudaStream_t st) {
    const int n_cols = (pl->k + 1) * (pl->k + 2) / 2;
    int rc, n_partial = 0;
    if ((rc = pl->tc_centre.ensure(sizeof(float) * (size_t)(n_cols + 512)))) return rc;
    if (m <= 0) { B200DA_CUDA(cudaMemsetAsync(pl->tc_centre.p, 0, sizeof(float) * (size_t)(n_cols + 512), st)); return B200DA_OK; }
    if ((rc = etkf_partial_grams(pl, Yn, d, m, m, &n_partial, st))) return rc;
    B200DA_CUDA(cudaMemsetAsync(pl->tc_centre.p, 0, sizeof(float) * (size_t)(n_cols + 512), st));
    k_tc_centre<<<(n_cols + 127) / 128, 128, 0, st>>>(pl->etkf_partial.as<double>(), n_partial, pl->kp, pl->k, m, n_cols,
                                                       pl->tc_centre.as<float>());
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_bin_obs(b200da_plan* plan, const double* obs_coord, const void* Yn, const void* d, int64_t n_obs, void* stream) {
    if (!plan) return B200DA_ERR_INVALID;
    // new observations: decisions and records of the ambiguity protocol refer to the previous set
    plan->n_over = 0;
    B200DA_CUDA(cudaMemsetAsync(&plan->devstat.as<PlanStatus>()->amb_found, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    if (plan->dtype == B200DA_F32) {
        int rc = bin_obs_impl<float>(plan, obs_coord, (const float*)Yn, (const float*)d, n_obs, (cudaStream_t)stream);
        if (rc == B200DA_OK && plan->use_tc) rc = tc_centre_constants(plan, Yn, d, n_obs, (cudaStream_t)stream);
        return rc;
    }
    return bin_obs_impl<double>(plan, obs_coord, (const double*)Yn, (const double*)d, n_obs, (cudaStream_t)stream);
}

int b200da_obs_prep(b200da_plan* pl, const void* HX, const void* y, const void* variance, int64_t m, void* Yn, void* d,
                    void* stream) {
    if (!pl || m < 0 || (m > 0 && (!HX || !y || !variance || !Yn || !d))) return B200DA_ERR_INVALID;
    if (m == 0) return B200DA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->dtype == B200DA_F32)
        k_obs_prep<float><<<grid1d(m, 256), 256, 0, st>>>((const float*)HX, (const float*)y, (const float*)variance, pl->k, m,
                                                          (float*)Yn, (float*)d);
    else
        k_obs_prep<double><<<grid1d(m, 256), 256, 0, st>>>((const double*)HX, (const double*)y, (const double*)variance, pl->k, m,
                                                           (double*)Yn, (double*)d);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_obs_gather_prep(b200da_plan* pl, const void* Xp, const int64_t* src_offset, int64_t member_stride, const void* y,
                           const void* variance, int64_t m, void* Yn, void* d, void* stream) {
    if (!pl || m < 0 || member_stride <= 0 || (m > 0 && (!Xp || !src_offset || !y || !variance || !Yn || !d))) return B200DA_ERR_INVALID;
    if (m == 0) return B200DA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->dtype == B200DA_F32)
        k_obs_gather_prep<float><<<grid1d(m, 256), 256, 0, st>>>((const float*)Xp, (const long long*)src_offset, member_stride,
                                                                 (const float*)y, (const float*)variance, pl->k, m, (float*)Yn, (float*)d);
    else
        k_obs_gather_prep<double><<<grid1d(m, 256), 256, 0, st>>>((const double*)Xp, (const long long*)src_offset, member_stride,
                                                                  (const double*)y, (const double*)variance, pl->k, m, (double*)Yn, (double*)d);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int64_t b200da_num_blocks(const b200da_plan* plan) { return plan ? plan->n_blocks : 0; }
int64_t b200da_num_grid(const b200da_plan* plan) { return plan ? plan->n_grid : 0; }
int64_t b200da_num_obs(const b200da_plan* plan) { return plan ? plan->n_obs : 0; }
int64_t b200da_block_offset(const b200da_plan* plan, int64_t block) {
    if (!plan || !plan->have_grid || block < 0 || block > plan->n_blocks) return -1;
    return plan->block_off_host[(size_t)block];
}
int b200da_block_offsets(const b200da_plan* plan, int64_t* offsets_host) {
    if (!plan || !offsets_host) return B200DA_ERR_INVALID;
    if (!plan->have_grid) return B200DA_ERR_STATE;
    for (size_t b = 0; b <= (size_t)plan->n_blocks; ++b) offsets_host[b] = plan->block_off_host[b];
    return B200DA_OK;
}
int b200da_gram_extra_rows(const b200da_plan* plan) { return plan ? plan->gram_er : 0; }
int b200da_grid_order(const b200da_plan* plan, int32_t* order_out, void* stream) {
    if (!plan || !order_out) return B200DA_ERR_INVALID;
    if (!plan->have_grid) return B200DA_ERR_STATE;
    B200DA_CUDA(cudaMemcpyAsync(order_out, plan->gorder.p, sizeof(int) * (size_t)plan->n_grid, cudaMemcpyDeviceToDevice,
                                (cudaStream_t)stream));
    return B200DA_OK;
}

// incoming weights and parameters of one IEnKS iteration (ienks_kernel.cuh); null: plain (L)ETKF
struct IenksArgs {
    const void* w_in;
    int per_grid;
    double tau;
    double eps;     // > 0: bundle variant
};

static int ienks_check(const b200da_plan* pl, const IenksArgs& ie) {
    if (!ie.w_in || !(ie.tau > 0.0) || ie.tau > 1.0) return B200DA_ERR_INVALID;
    if (pl->kprog.n > 0) return B200DA_ERR_UNSUPPORTED;           // the reference has no kernelised IEnKS either
    if (ienks_smem_bytes(pl->k) > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    return B200DA_OK;
}

static IenksParams ienks_params(const b200da_plan* pl, const IenksArgs& ie, double* cmat, int64_t n_slots, int64_t slot_stride,
                                const Pos4* gpos, int64_t slot_base, int* empty) {
    IenksParams I{};
    I.cmat = cmat; I.n_slots = n_slots; I.slot_stride = slot_stride; I.gpos = gpos; I.slot_base = slot_base;
    I.w_in = ie.w_in; I.per_grid = ie.per_grid; I.io_f32 = pl->dtype == B200DA_F32 ? 1 : 0; I.k = pl->k;
    I.tau = ie.tau; I.eps = ie.eps; I.empty = empty;
    return I;
}

static int launch_ienks_pre(const b200da_plan* pl, const IenksParams& I, cudaStream_t st) {
    const size_t smem = ienks_smem_bytes(pl->k);
    B200DA_CUDA(cudaFuncSetAttribute(k_ienks_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ienks_pre<<<(int)std::min<int64_t>(I.n_slots, 148 * 8), 256, smem, st>>>(I);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

static int launch_ienks_keep(const b200da_plan* pl, const IenksParams& I, cudaStream_t st) {
    const size_t smem = sizeof(double) * ((size_t)pl->k * pl->k + pl->k);
    B200DA_CUDA(cudaFuncSetAttribute(k_ienks_keep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ienks_keep<<<(int)std::min<int64_t>(I.n_slots, 148 * 16), 128, smem, st>>>(I);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

static int letkf_impl(b200da_plan* pl, const void* X, void* Xa, void* W_opt, int64_t block_begin, int64_t block_end,
                      int64_t* n_ambiguous_opt, double* gram_out, void* stream, const IenksArgs* ie = nullptr) {
    if (!pl || (!gram_out && (!X || !Xa))) return B200DA_ERR_INVALID;
    if (!pl->have_grid || !pl->have_obs) return B200DA_ERR_STATE;
    if (block_begin < 0 || block_end > pl->n_blocks || block_begin > block_end) return B200DA_ERR_INVALID;
    if (block_begin == block_end) return B200DA_OK;
    LetkfParams P{};
    P.g = pl->geom;
    P.gpos = pl->gpos.as<Pos4>();
    P.block_off = pl->block_off.as<int>();
    P.opos = pl->opos.as<Pos4>();
    P.gext = pl->gext.as<double>(); P.oext = pl->oext.as<double>();
    P.cell_start = pl->cell_start.as<int>();
    P.ys = pl->ys.p;
    P.x = X; P.xa = Xa; P.w_out = W_opt;
    const int f32 = pl->dtype == B200DA_F32 ? 1 : 0;
    P.n_ambiguous = (unsigned long long*)n_ambiguous_opt;
    P.n_grid = pl->n_grid; P.n_obs = pl->n_obs;
    P.block_begin = (int)block_begin;
    P.k = pl->k; P.n_slices = pl->n_slices; P.rho = pl->rho;
    P.cut_pad = pl->geom.cut_bin * (1.0 + 1e-9) + 1e-300;
    P.status = pl->devstat.as<PlanStatus>();
    P.amb_list = pl->amb_list.as<PairRec>();
    P.over = pl->over_list.as<PairRec>(); P.n_over = pl->n_over;
    P.tt = pl->geom.n_ext == 0 ? pl->tt : TaperTab{};              // several distance rows: direct evaluation, row by row
    P.stats = nullptr;
    if (pl->collect_stats) {
        int rc = pl->stats.ensure(sizeof(unsigned long long) * 16);
        if (rc) return rc;
        B200DA_CUDA(cudaMemsetAsync(pl->stats.p, 0, sizeof(unsigned long long) * 16, (cudaStream_t)stream));
        P.stats = pl->stats.as<unsigned long long>();
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int k = pl->k;
    const int64_t slot_stride = (int64_t)tri_tiles(pl->kt) * 64;           // doubles per grid slot of the Gram scratch
    const size_t per_slot = sizeof(double) * (size_t)slot_stride;
    const bool jacobi = pl->solver == B200DA_SOLVER_JACOBI;
    const size_t smem_solve = solve_smem_bytes(k);
    const bool big = k > 64;
    if (jacobi) {
        if (smem_solve > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
        if (big) B200DA_CUDA(cudaFuncSetAttribute(k_letkf_solve<512, 1, 2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
        else B200DA_CUDA(cudaFuncSetAttribute(k_letkf_solve<256, 2, 1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
    } else {
        int rc = pl->counter.ensure(sizeof(unsigned int) * 4);
        if (rc) return rc;
    }
    // chunk the block range so that the Gram scratch stays below ~2 GiB
    const int64_t max_slots = std::max<int64_t>(pl->gpb, (int64_t)((size_t(2) << 30) / per_slot));
    if (pl->timing) { B200DA_CUDA(cudaEventRecord(pl->ev0, st)); pl->gram_ms = 0.f; pl->solve_ms = 0.f; pl->n_ev = 0; }
    int64_t b = block_begin;
    while (b < block_end) {
        int64_t e = b;
        const int64_t s0 = pl->block_off_host[(size_t)b];
        while (e < block_end && pl->block_off_host[(size_t)e + 1] - s0 <= max_slots) ++e;
        if (e == b) ++e;
        const int64_t n_slots = pl->block_off_host[(size_t)e] - s0;
        int rc = pl->cmat.ensure(per_slot * (size_t)n_slots);
        if (rc) return rc;
        P.cmat = pl->cmat.as<double>();
        P.slot_base = s0;
        P.block_begin = (int)b;
        cudaEvent_t ea = nullptr, eb = nullptr, ec = nullptr;
        if (pl->timing) {
            if ((rc = pl->next_events(&ea, &eb, &ec))) return rc;
            B200DA_CUDA(cudaEventRecord(ea, st));
        }
        if (pl->use_tc) { if ((rc = launch_tc_gram(pl, P, (int)(e - b), st))) return rc; }
        else if ((rc = dispatch_fused(pl, P, (int)(e - b), st))) return rc;
        if (pl->timing) B200DA_CUDA(cudaEventRecord(eb, st));
        if (pl->kprog.n > 0 && (rc = launch_kernelise(pl, pl->cmat.as<double>(), n_slots, slot_stride, st))) return rc;
        IenksParams I{};
        if (ie) {
            if ((rc = pl->tmp_a.ensure(sizeof(int) * (size_t)n_slots))) return rc;
            I = ienks_params(pl, *ie, pl->cmat.as<double>(), n_slots, slot_stride, P.gpos, s0, pl->tmp_a.as<int>());
            I.x = P.x; I.xa = P.xa; I.w_out = P.w_out; I.n_slices = pl->n_slices; I.n_grid = pl->n_grid;
            if ((rc = launch_ienks_pre(pl, I, st))) return rc;
        }
        const double rho_solve = ie ? 1.0 / ie->tau : pl->rho;      // P = A' + tau (k - 1) I  (ienks_kernel.cuh)
        if (gram_out) {
            k_unpack_gram<<<(int)std::min<int64_t>(n_slots, 148 * 16), 128, 0, st>>>(P.cmat, P.gpos, s0, n_slots, slot_stride, k + 1, gram_out);
            B200DA_LAUNCH_CHECK();
        } else if (jacobi) {
            SolveParams S{};
            S.cmat = P.cmat; S.slot_stride = slot_stride; S.gpos = P.gpos; S.x = P.x; S.xa = P.xa; S.w_out = P.w_out; S.stats = P.stats;
            S.io_f32 = f32;
            S.slot_base = s0; S.n_slots = n_slots; S.n_grid = pl->n_grid; S.k = k; S.n_slices = pl->n_slices; S.rho = rho_solve;
            const int grid = (int)std::min<int64_t>(n_slots, 148 * 64);
            if (big) k_letkf_solve<512, 1, 2, 4, 4><<<grid, 512, smem_solve, st>>>(S);
            else k_letkf_solve<256, 2, 1, 4, 2><<<grid, 256, smem_solve, st>>>(S);
            B200DA_LAUNCH_CHECK();
        } else {
            B200DA_CUDA(cudaMemsetAsync(pl->counter.p, 0, sizeof(unsigned int) * 4, st));
            NsParams S{};
            S.cmat = P.cmat; S.slot_stride = slot_stride; S.gpos = P.gpos; S.x = P.x; S.xa = P.xa; S.w_out = P.w_out; S.stats = P.stats;
            S.counter = pl->counter.as<unsigned int>();
            S.io_f32 = f32;
            S.slot_base = s0; S.n_slots = n_slots; S.n_grid = pl->n_grid; S.k = k; S.n_slices = pl->n_slices; S.rho = rho_solve;
            if ((rc = pl->ns_scratch.ensure(ns_scratch_bytes((k + 7) / 8)))) return rc;
            S.scratch = pl->ns_scratch.as<double>(); S.stiff = ns_stiff_ratio(pl); S.conv = pl->dtype == B200DA_F32 ? 3e-4 : 2e-8;
            if ((rc = dispatch_ns((k + 7) / 8, S, st))) return rc;
        }
        if (ie && !gram_out && (rc = launch_ienks_keep(pl, I, st))) return rc;
        if (pl->timing) B200DA_CUDA(cudaEventRecord(ec, st));
        b = e;
    }
    if (pl->timing) B200DA_CUDA(cudaEventRecord(pl->ev1, st));
    return B200DA_OK;
}

int b200da_letkf(b200da_plan* pl, const void* X, void* Xa, void* W_opt, int64_t block_begin, int64_t block_end,
                 int64_t* n_ambiguous_opt, void* stream) {
    return letkf_impl(pl, X, Xa, W_opt, block_begin, block_end, n_ambiguous_opt, nullptr, stream);
}

int b200da_letkf_ienks(b200da_plan* pl, const void* X, void* Xa, const void* W_in, int w_per_grid, void* W_out, double tau,
                       double epsilon, int64_t block_begin, int64_t block_end, void* stream) {
    if (!pl) return B200DA_ERR_INVALID;
    IenksArgs ie{W_in, w_per_grid ? 1 : 0, tau, epsilon};
    int rc = ienks_check(pl, ie);
    if (rc) return rc;
    return letkf_impl(pl, X, Xa, W_out, block_begin, block_end, nullptr, nullptr, stream, &ie);
}

int b200da_letkf_gram(b200da_plan* pl, double* gram_out, int64_t block_begin, int64_t block_end, void* stream) {
    if (!gram_out) return B200DA_ERR_INVALID;
    return letkf_impl(pl, nullptr, nullptr, nullptr, block_begin, block_end, nullptr, gram_out, stream);
}

int b200da_letkf_host(b200da_plan* pl, const double* obs_coord_host, const void* Yn_host, const void* d_host, int64_t m,
                      const void* X_host, void* Xa_host, void* stream) {
    if (!pl || !X_host || !Xa_host) return B200DA_ERR_INVALID;
    if (!pl->have_grid) return B200DA_ERR_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const size_t mm = (size_t)std::max<int64_t>(m, 1);
    const size_t es = pl->dtype == B200DA_F32 ? sizeof(float) : sizeof(double);
    const size_t xbytes = es * (size_t)pl->n_slices * pl->k * (size_t)pl->n_grid;
    const size_t n_rows_coord = (size_t)pl->n_coord + pl->geom.n_ext;
    if ((rc = pl->host_stage_obs.ensure(sizeof(double) * mm * n_rows_coord))) return rc;
    if ((rc = pl->host_stage_y.ensure(es * mm * pl->k))) return rc;
    if ((rc = pl->host_stage_d.ensure(es * mm))) return rc;
    if ((rc = pl->host_stage_x.ensure(xbytes))) return rc;
    if ((rc = pl->host_stage_xa.ensure(xbytes))) return rc;
    if (m > 0) {
        B200DA_CUDA(cudaMemcpyAsync(pl->host_stage_obs.p, obs_coord_host, sizeof(double) * (size_t)m * n_rows_coord, cudaMemcpyHostToDevice, st));
        B200DA_CUDA(cudaMemcpyAsync(pl->host_stage_y.p, Yn_host, es * (size_t)m * pl->k, cudaMemcpyHostToDevice, st));
        B200DA_CUDA(cudaMemcpyAsync(pl->host_stage_d.p, d_host, es * (size_t)m, cudaMemcpyHostToDevice, st));
    }
    B200DA_CUDA(cudaMemcpyAsync(pl->host_stage_x.p, X_host, xbytes, cudaMemcpyHostToDevice, st));
    if ((rc = b200da_bin_obs(pl, pl->host_stage_obs.as<double>(), pl->host_stage_y.p, pl->host_stage_d.p, m, stream))) return rc;
    if ((rc = b200da_letkf(pl, pl->host_stage_x.p, pl->host_stage_xa.p, nullptr, 0, pl->n_blocks, nullptr, stream))) return rc;
    B200DA_CUDA(cudaMemcpyAsync(Xa_host, pl->host_stage_xa.p, xbytes, cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    return B200DA_OK;
}

int b200da_letkf_host_blocks(b200da_plan* pl, void* Xa_host, int64_t block_begin, int64_t block_end, void* stream) {
    if (!pl || !Xa_host) return B200DA_ERR_INVALID;
    if (!pl->have_grid || !pl->have_obs || !pl->host_stage_x.p || !pl->host_stage_xa.p) return B200DA_ERR_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = pl->dtype == B200DA_F32 ? sizeof(float) : sizeof(double);
    const size_t xbytes = es * (size_t)pl->n_slices * pl->k * (size_t)pl->n_grid;
    int rc;
    if ((rc = b200da_letkf(pl, pl->host_stage_x.p, pl->host_stage_xa.p, nullptr, block_begin, block_end, nullptr, stream))) return rc;
    B200DA_CUDA(cudaMemcpyAsync(Xa_host, pl->host_stage_xa.p, xbytes, cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    return B200DA_OK;
}

int b200da_neighbour_count(b200da_plan* pl, int64_t* counts, int64_t* n_ambiguous_opt, void* stream) {
    if (!pl || !counts) return B200DA_ERR_INVALID;
    if (!pl->have_grid || !pl->have_obs) return B200DA_ERR_STATE;
    NeighbourParams P = neighbour_params(pl);
    P.counts = (long long*)counts;
    P.n_ambiguous = (unsigned long long*)n_ambiguous_opt;
    return launch_neighbours<0>(pl, P, (cudaStream_t)stream);
}

int b200da_neighbour_fill(b200da_plan* pl, const int64_t* offsets, int32_t* idx, double* w_opt, uint8_t* ambiguous_opt,
                          void* stream) {
    if (!pl || !offsets || !idx) return B200DA_ERR_INVALID;
    if (!pl->have_grid || !pl->have_obs) return B200DA_ERR_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nnz = 0;
    B200DA_CUDA(cudaMemcpyAsync(&nnz, offsets + pl->n_grid, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    if (nnz <= 0) return B200DA_OK;
    int rc;
    if ((rc = pl->tmp_keys.ensure(sizeof(unsigned long long) * (size_t)nnz))) return rc;
    if ((rc = pl->tmp_a.ensure(sizeof(int) * (size_t)pl->n_grid))) return rc;
    B200DA_CUDA(cudaMemsetAsync(pl->tmp_a.p, 0, sizeof(int) * (size_t)pl->n_grid, st));
    NeighbourParams P = neighbour_params(pl);
    P.offsets = (const long long*)offsets;
    P.keys = pl->tmp_keys.as<unsigned long long>();
    P.cursor = pl->tmp_a.as<int>();
    if ((rc = launch_neighbours<1>(pl, P, st))) return rc;
    const int nblk = (int)std::min<int64_t>(pl->n_grid, 148 * 64);
    k_segmented_sort<long long><<<nblk, kSegSortThreads, 0, st>>>(P.keys, (const long long*)offsets, pl->n_grid);
    B200DA_LAUNCH_CHECK();
    k_neighbour_finalize<<<nblk, 128, 0, st>>>(pl->geom, pl->gpos.as<Pos4>(), pl->opos.as<Pos4>(), pl->gext.as<double>(),
                                                 pl->oext.as<double>(), (const long long*)offsets,
                                               P.keys, pl->n_grid, idx, w_opt, ambiguous_opt, P.over, P.n_over);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_neighbour_ambiguous(b200da_plan* pl, int64_t capacity, int64_t* grid_idx, int64_t* obs_idx, double* w,
                               int64_t* n_found, void* stream) {
    if (!pl || !n_found || capacity < 0 || (capacity > 0 && (!grid_idx || !obs_idx || !w))) return B200DA_ERR_INVALID;
    if (!pl->have_grid || !pl->have_obs) return B200DA_ERR_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    B200DA_CUDA(cudaMemsetAsync(n_found, 0, sizeof(int64_t), st));
    NeighbourParams P = neighbour_params(pl);
    P.capacity = capacity; P.amb_grid = (long long*)grid_idx; P.amb_obs = (long long*)obs_idx; P.amb_w = w;
    P.amb_found = (unsigned long long*)n_found;
    return launch_neighbours<2>(pl, P, st);
}

int b200da_pending_status(b200da_plan* pl, int64_t capacity, int64_t* grid_idx_host, int64_t* obs_idx_host, double* w_host,
                          int64_t* n_found_host, void* stream) {
    if (!pl || !n_found_host || capacity < 0 || (capacity > 0 && (!grid_idx_host || !obs_idx_host || !w_host))) return B200DA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    PlanStatus hs{};
    B200DA_CUDA(cudaMemcpyAsync(&hs, pl->devstat.p, sizeof(PlanStatus), cudaMemcpyDeviceToHost, st));
    B200DA_CUDA(cudaStreamSynchronize(st));
    *n_found_host = (int64_t)hs.amb_found;
    const int64_t n = std::min<int64_t>(std::min<int64_t>((int64_t)hs.amb_found, kAmbCapacity), capacity);
    if (n > 0) {
        std::vector<PairRec> rec((size_t)n);
        B200DA_CUDA(cudaMemcpyAsync(rec.data(), pl->amb_list.p, sizeof(PairRec) * (size_t)n, cudaMemcpyDeviceToHost, st));
        B200DA_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < n; ++i) { grid_idx_host[i] = rec[(size_t)i].gid; obs_idx_host[i] = rec[(size_t)i].oid; w_host[i] = rec[(size_t)i].w; }
    }
    if (hs.amb_found || hs.error) B200DA_CUDA(cudaMemsetAsync(pl->devstat.p, 0, sizeof(PlanStatus), st));
    if (hs.error & kStatusRunOverflow) return B200DA_ERR_OVERFLOW;
    return B200DA_OK;
}

int b200da_plan_set_overrides(b200da_plan* pl, int64_t n, const int64_t* grid_idx_host, const int64_t* obs_idx_host,
                              const double* w_host, void* stream) {
    if (!pl || n < 0 || n > kAmbCapacity || (n > 0 && (!grid_idx_host || !obs_idx_host || !w_host))) return B200DA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    pl->n_over = 0;
    if (n == 0) return B200DA_OK;
    int rc;
    std::vector<PairRec> rec((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (grid_idx_host[i] < 0 || grid_idx_host[i] >= pl->n_grid || obs_idx_host[i] < 0 || obs_idx_host[i] >= pl->n_obs ||
            !(w_host[i] >= 0.0)) return B200DA_ERR_INVALID;
        rec[(size_t)i].gid = grid_idx_host[i]; rec[(size_t)i].oid = obs_idx_host[i]; rec[(size_t)i].w = w_host[i];
    }
    B200DA_CUDA(cudaStreamSynchronize(st));                    // kernels in flight may still read the previous list
    if ((rc = pl->over_list.ensure(sizeof(PairRec) * (size_t)n))) return rc;
    B200DA_CUDA(cudaMemcpyAsync(pl->over_list.p, rec.data(), sizeof(PairRec) * (size_t)n, cudaMemcpyHostToDevice, st));
    B200DA_CUDA(cudaStreamSynchronize(st));                    // rec is a stack-lifetime staging buffer
    pl->n_over = (int)n;
    return B200DA_OK;
}

int b200da_blocks_of_grid(b200da_plan* pl, int64_t n, const int64_t* grid_idx_host, int64_t* block_host, void* stream) {
    if (!pl || n < 0 || (n > 0 && (!grid_idx_host || !block_host))) return B200DA_ERR_INVALID;
    if (!pl->have_grid) return B200DA_ERR_STATE;
    if (pl->slot_of_host.size() != (size_t)pl->n_grid) {
        std::vector<int32_t> order((size_t)pl->n_grid);
        B200DA_CUDA(cudaMemcpyAsync(order.data(), pl->gorder.p, sizeof(int32_t) * (size_t)pl->n_grid, cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream));
        B200DA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        pl->slot_of_host.assign((size_t)pl->n_grid, 0);
        for (int64_t s = 0; s < pl->n_grid; ++s) pl->slot_of_host[(size_t)order[(size_t)s]] = (int32_t)s;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (grid_idx_host[i] < 0 || grid_idx_host[i] >= pl->n_grid) return B200DA_ERR_INVALID;
        const int32_t slot = pl->slot_of_host[(size_t)grid_idx_host[i]];
        const auto it = std::upper_bound(pl->block_off_host.begin(), pl->block_off_host.end(), slot);
        block_host[i] = (int64_t)(it - pl->block_off_host.begin()) - 1;
    }
    return B200DA_OK;
}

// stage 1 of the global ETKF: per-CTA partial Grams of an observation range into pl->etkf_partial; returns the number of partials
static int etkf_partial_grams(b200da_plan* pl, const void* Yn, const void* d, int64_t m, int64_t ld, int* ncta_out, cudaStream_t st) {
    const int k = pl->k, kp = pl->kp;
    int ncta = 1;
    int64_t chunk = 4;
    if (m > 0) {
        ncta = (int)std::min<int64_t>(148 * 2, (m + 255) / 256);
        chunk = ((m + ncta - 1) / ncta + kEtkfTileObs - 1) / kEtkfTileObs * kEtkfTileObs;
        ncta = (int)((m + chunk - 1) / chunk);
    }
    int rc;
    if ((rc = pl->etkf_partial.ensure(sizeof(double) * (size_t)ncta * kp * kp))) return rc;
    if (m > 0) {
        if ((rc = dispatch_etkf_gram(pl->kt, pl->dtype == B200DA_F32, Yn, d, m, ld, k, ncta, chunk,
                                     pl->etkf_partial.as<double>(), st))) return rc;
    }
    *ncta_out = m > 0 ? ncta : 0;
    return B200DA_OK;
}

// stage 2: n_partial partial Grams in pl->etkf_partial -> W (k, k); n_partial = 0 gives sqrt(rho) I
static int etkf_solve_partials(b200da_plan* pl, int n_partial, void* W, cudaStream_t st, const IenksArgs* ie = nullptr) {
    const int k = pl->k, kp = pl->kp;
    const int f32 = pl->dtype == B200DA_F32 ? 1 : 0;
    int rc;
    const bool kernelised = pl->kprog.n > 0;
    const double rho_solve = ie ? 1.0 / ie->tau : pl->rho;
    if (ie && n_partial == 0) {                            // no observations: the weights stay as they are (ienks.py:143-150)
        B200DA_CUDA(cudaMemcpyAsync(W, ie->w_in, (f32 ? 4 : 8) * (size_t)k * k, cudaMemcpyDeviceToDevice, st));
        return B200DA_OK;
    }
    if (n_partial > 0 && (pl->solver == B200DA_SOLVER_NEWTON_SCHULZ || kernelised || ie)) {
        // partial Grams -> one tile-packed slot -> the tensor-core Newton-Schulz solve (one matrix, no state update)
        const size_t slot_doubles = (size_t)tri_tiles(pl->kt) * 64;
        if ((rc = pl->etkf_w.ensure(sizeof(Pos4) + sizeof(double) * slot_doubles))) return rc;
        if ((rc = pl->counter.ensure(sizeof(unsigned int) * 4))) return rc;
        B200DA_CUDA(cudaMemsetAsync(pl->etkf_w.p, 0, sizeof(Pos4) + sizeof(double) * slot_doubles, st));   // Pos4.id = 0
        B200DA_CUDA(cudaMemsetAsync(pl->counter.p, 0, sizeof(unsigned int) * 4, st));
        double* slot = reinterpret_cast<double*>(pl->etkf_w.as<unsigned char>() + sizeof(Pos4));
        k_etkf_reduce<<<grid1d((int64_t)(k + 1) * kp, 128), 128, 0, st>>>(pl->etkf_partial.as<double>(), n_partial, kp, k, slot);
        B200DA_LAUNCH_CHECK();
        if (kernelised && (rc = launch_kernelise(pl, slot, 1, (int64_t)slot_doubles, st))) return rc;
        if (ie) {
            if ((rc = pl->tmp_a.ensure(sizeof(int) * 4))) return rc;
            IenksParams I = ienks_params(pl, *ie, slot, 1, (int64_t)slot_doubles, pl->etkf_w.as<Pos4>(), 0, pl->tmp_a.as<int>());
            I.per_grid = 0;
            if ((rc = launch_ienks_pre(pl, I, st))) return rc;
        }
        if (pl->solver == B200DA_SOLVER_JACOBI) {         // kernelised / IEnKS only: the slot goes through the per-point Jacobi kernel
            const size_t smem_j = solve_smem_bytes(k);
            if (smem_j > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
            SolveParams J{};
            J.cmat = slot; J.slot_stride = (int64_t)slot_doubles; J.gpos = pl->etkf_w.as<Pos4>(); J.x = nullptr; J.xa = nullptr;
            J.w_out = W; J.io_f32 = f32; J.stats = nullptr; J.slot_base = 0; J.n_slots = 1; J.n_grid = 0; J.k = k; J.n_slices = 0;
            J.rho = rho_solve;
            if (k > 64) {
                B200DA_CUDA(cudaFuncSetAttribute(k_letkf_solve<512, 1, 2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_j));
                k_letkf_solve<512, 1, 2, 4, 4><<<1, 512, smem_j, st>>>(J);
            } else {
                B200DA_CUDA(cudaFuncSetAttribute(k_letkf_solve<256, 2, 1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_j));
                k_letkf_solve<256, 2, 1, 4, 2><<<1, 256, smem_j, st>>>(J);
            }
            B200DA_LAUNCH_CHECK();
            return B200DA_OK;
        }
        NsParams S{};
        S.cmat = slot; S.slot_stride = (int64_t)slot_doubles; S.gpos = pl->etkf_w.as<Pos4>(); S.x = nullptr; S.xa = nullptr;
        S.w_out = W; S.io_f32 = f32; S.stats = nullptr; S.counter = pl->counter.as<unsigned int>();
        S.slot_base = 0; S.n_slots = 1; S.n_grid = 0; S.k = k; S.n_slices = 0; S.rho = rho_solve;
        if ((rc = pl->ns_scratch.ensure(ns_scratch_bytes((k + 7) / 8)))) return rc;
        S.scratch = pl->ns_scratch.as<double>(); S.stiff = ns_stiff_ratio(pl); S.conv = pl->dtype == B200DA_F32 ? 3e-4 : 2e-8;
        return dispatch_ns((k + 7) / 8, S, st);
    }
    const size_t smem = solve_smem_bytes(k);
    if (smem > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    B200DA_CUDA(cudaFuncSetAttribute(k_etkf_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_etkf_solve<<<1, 512, smem, st>>>(pl->etkf_partial.as<double>(), n_partial, kp, k, pl->rho, W, f32);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_etkf_weights(b200da_plan* pl, const void* Yn, const void* d, int64_t m, void* W, void* stream) {
    if (!pl || !W || m < 0 || (m > 0 && (!Yn || !d))) return B200DA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc, n_partial = 0;
    if ((rc = etkf_partial_grams(pl, Yn, d, m, m, &n_partial, st))) return rc;
    return etkf_solve_partials(pl, n_partial, W, st);
}

int b200da_etkf_ienks_weights(b200da_plan* pl, const void* Yn, const void* d, int64_t m, const void* W_in, double tau,
                              double epsilon, void* W_out, void* stream) {
    if (!pl || !W_out || m < 0 || (m > 0 && (!Yn || !d))) return B200DA_ERR_INVALID;
    IenksArgs ie{W_in, 0, tau, epsilon};
    int rc = ienks_check(pl, ie);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int n_partial = 0;
    if ((rc = etkf_partial_grams(pl, Yn, d, m, m, &n_partial, st))) return rc;
    return etkf_solve_partials(pl, n_partial, W_out, st, &ie);
}

int b200da_etkf_gram(b200da_plan* pl, const void* Yn, const void* d, int64_t m, int64_t ld_obs, double* gram_out, void* stream) {
    if (!pl || !gram_out || m < 0 || ld_obs < m || (m > 0 && (!Yn || !d))) return B200DA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc, n_partial = 0;
    if ((rc = etkf_partial_grams(pl, Yn, d, m, ld_obs, &n_partial, st))) return rc;
    const int k = pl->k;
    k_etkf_reduce_dense<<<grid1d((int64_t)(k + 1) * (k + 1), 128), 128, 0, st>>>(pl->etkf_partial.as<double>(), n_partial, pl->kp, k, gram_out);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_etkf_weights_from_gram(b200da_plan* pl, const double* gram, int64_t n_obs_total, void* W, void* stream) {
    if (!pl || !gram || !W || n_obs_total < 0) return B200DA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = pl->kp;
    int rc;
    if ((rc = pl->etkf_partial.ensure(sizeof(double) * (size_t)kp * kp))) return rc;
    if (n_obs_total > 0) {
        k_etkf_dense_to_partial<<<grid1d((int64_t)kp * kp, 128), 128, 0, st>>>(gram, kp, pl->k, pl->etkf_partial.as<double>());
        B200DA_LAUNCH_CHECK();
    }
    return etkf_solve_partials(pl, n_obs_total > 0 ? 1 : 0, W, st);
}

static int apply_weights_impl(b200da_plan* pl, const void* X, const void* W, int per_grid, int64_t n_grid, int64_t ld, void* Xa,
                              cudaStream_t st, const ApplyPeers& peers = ApplyPeers{}) {
    const int k = pl->k;
    if (!per_grid) return dispatch_apply_global(pl->dtype == B200DA_F32, X, W, k, pl->n_slices, n_grid, ld, Xa, peers, st);
    if (peers.n > 0) return B200DA_ERR_UNSUPPORTED;
    const size_t smem = 0;
    if (pl->dtype == B200DA_F32) {
        B200DA_CUDA(cudaFuncSetAttribute(k_apply_weights<float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        k_apply_weights<float, 8><<<grid1d(n_grid, 128), 128, smem, st>>>((const float*)X, (const float*)W, per_grid, k,
                                                                        pl->n_slices, n_grid, ld, (float*)Xa);
    } else {
        B200DA_CUDA(cudaFuncSetAttribute(k_apply_weights<double, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        k_apply_weights<double, 8><<<grid1d(n_grid, 128), 128, smem, st>>>((const double*)X, (const double*)W, per_grid, k,
                                                                         pl->n_slices, n_grid, ld, (double*)Xa);
    }
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

int b200da_apply_weights(b200da_plan* pl, const void* X, const void* W, int per_grid, int64_t n_grid, void* Xa, void* stream) {
    if (!pl || !X || !W || !Xa || n_grid <= 0) return B200DA_ERR_INVALID;
    return apply_weights_impl(pl, X, W, per_grid, n_grid, n_grid, Xa, (cudaStream_t)stream);
}

int b200da_apply_weights_cols(b200da_plan* pl, const void* X, const void* W, int per_grid, int64_t col_begin, int64_t col_end,
                              int64_t n_grid, void* Xa, void* stream) {
    if (!pl || !X || !W || !Xa || n_grid <= 0 || col_begin < 0 || col_end > n_grid || col_begin > col_end) return B200DA_ERR_INVALID;
    if (col_begin == col_end) return B200DA_OK;
    const size_t esz = pl->dtype == B200DA_F32 ? 4 : 8;
    const unsigned char* x = static_cast<const unsigned char*>(X) + (size_t)col_begin * esz;
    unsigned char* xa = static_cast<unsigned char*>(Xa) + (size_t)col_begin * esz;
    const unsigned char* w = static_cast<const unsigned char*>(W) + (per_grid ? (size_t)col_begin * pl->k * pl->k * esz : 0);
    return apply_weights_impl(pl, x, w, per_grid, col_end - col_begin, n_grid, xa, (cudaStream_t)stream);
}

int b200da_apply_weights_cols_peers(b200da_plan* pl, const void* X, const void* W, int64_t col_begin, int64_t col_end,
                                    int64_t n_grid, void* Xa, int n_peers, void* const* Xa_peers, void* stream) {
    if (!pl || !X || !W || !Xa || n_grid <= 0 || col_begin < 0 || col_end > n_grid || col_begin > col_end || n_peers < 0 ||
        n_peers > 7 || (n_peers > 0 && !Xa_peers)) return B200DA_ERR_INVALID;
    if (col_begin == col_end) return B200DA_OK;
    const size_t esz = pl->dtype == B200DA_F32 ? 4 : 8;
    ApplyPeers peers{};
    peers.n = n_peers;
    for (int i = 0; i < n_peers; ++i) {
        if (!Xa_peers[i]) return B200DA_ERR_INVALID;
        peers.p[i] = static_cast<unsigned char*>(Xa_peers[i]) + (size_t)col_begin * esz;
    }
    const unsigned char* x = static_cast<const unsigned char*>(X) + (size_t)col_begin * esz;
    unsigned char* xa = static_cast<unsigned char*>(Xa) + (size_t)col_begin * esz;
    return apply_weights_impl(pl, x, W, 0, col_end - col_begin, n_grid, xa, (cudaStream_t)stream, peers);
}

int b200da_peer_copy_cols(void* dst, const void* src, int64_t rows, int64_t ld, int64_t col_begin, int64_t col_end,
                          int elem_bytes, void* stream) {
    if (!dst || !src || rows <= 0 || ld <= 0 || col_begin < 0 || col_end > ld || col_begin > col_end ||
        (elem_bytes != 4 && elem_bytes != 8)) return B200DA_ERR_INVALID;
    if (col_begin == col_end) return B200DA_OK;
    const size_t es = (size_t)elem_bytes;
    B200DA_CUDA(cudaMemcpy2DAsync(static_cast<unsigned char*>(dst) + (size_t)col_begin * es, (size_t)ld * es,
                                  static_cast<const unsigned char*>(src) + (size_t)col_begin * es, (size_t)ld * es,
                                  (size_t)(col_end - col_begin) * es, (size_t)rows, cudaMemcpyDefault, (cudaStream_t)stream));
    return B200DA_OK;
}

static int pack_impl(b200da_plan* pl, const void* xa, int64_t b0, int64_t b1, void* packed, int64_t ld, int unpack,
                     cudaStream_t st) {
    if (!pl || !xa || !packed) return B200DA_ERR_INVALID;
    if (!pl->have_grid) return B200DA_ERR_STATE;
    if (b0 < 0 || b1 > pl->n_blocks || b0 > b1) return B200DA_ERR_INVALID;
    const int64_t s0 = pl->block_off_host[(size_t)b0], s1 = pl->block_off_host[(size_t)b1];
    const int64_t ncols = s1 - s0;
    if (ncols == 0) return B200DA_OK;
    if (ld <= 0) ld = ncols;
    if (ld < ncols) return B200DA_ERR_INVALID;
    const int rows = pl->n_slices * pl->k;
    if (rows > 65535) return B200DA_ERR_UNSUPPORTED;
    dim3 grid((unsigned)grid1d(ncols, 256), (unsigned)rows);
    if (pl->dtype == B200DA_F32)
        k_pack_columns<float><<<grid, 256, 0, st>>>((const float*)xa, pl->gorder.as<int>(), s0, ncols, rows, pl->n_grid, (float*)packed, ld, unpack);
    else
        k_pack_columns<double><<<grid, 256, 0, st>>>((const double*)xa, pl->gorder.as<int>(), s0, ncols, rows, pl->n_grid, (double*)packed, ld, unpack);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}
int b200da_pack_columns(b200da_plan* pl, const void* Xa, int64_t b0, int64_t b1, void* packed, int64_t ld, void* stream) {
    return pack_impl(pl, Xa, b0, b1, packed, ld, 0, (cudaStream_t)stream);
}
int b200da_unpack_columns(b200da_plan* pl, const void* packed, int64_t b0, int64_t b1, void* Xa, int64_t ld, void* stream) {
    return pack_impl(pl, Xa, b0, b1, const_cast<void*>(packed), ld, 1, (cudaStream_t)stream);
}

int b200da_set_solver(b200da_plan* pl, int solver) {
    if (!pl) return B200DA_ERR_INVALID;
    if (solver != B200DA_SOLVER_NEWTON_SCHULZ && solver != B200DA_SOLVER_JACOBI) return B200DA_ERR_UNSUPPORTED;
    pl->solver = solver;
    return B200DA_OK;
}

int b200da_collect_stats(b200da_plan* pl, int on) { if (!pl) return B200DA_ERR_INVALID; pl->collect_stats = on != 0; return B200DA_OK; }
int b200da_get_stats(b200da_plan* pl, int64_t* out8) {
    if (!pl || !out8) return B200DA_ERR_INVALID;
    for (int i = 0; i < 16; ++i) out8[i] = 0;
    if (!pl->stats.p) return B200DA_OK;
    B200DA_CUDA(cudaDeviceSynchronize());
    B200DA_CUDA(cudaMemcpy(out8, pl->stats.p, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost));
    return B200DA_OK;
}

const char* b200da_kernel_name(const b200da_plan* pl) { return pl ? pl->kernel_name.c_str() : ""; }
int b200da_enable_timing(b200da_plan* pl, int on) { if (!pl) return B200DA_ERR_INVALID; pl->timing = on != 0; return B200DA_OK; }
float b200da_last_kernel_ms(b200da_plan* pl) {
    if (!pl || !pl->timing) return -1.f;
    if (cudaEventSynchronize(pl->ev1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, pl->ev0, pl->ev1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    pl->gram_ms = 0.f; pl->solve_ms = 0.f;
    for (int i = 0; i + 2 < pl->n_ev; i += 3) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, pl->ev_pool[i], pl->ev_pool[i + 1]);
        cudaEventElapsedTime(&b, pl->ev_pool[i + 1], pl->ev_pool[i + 2]);
        pl->gram_ms += a; pl->solve_ms += b;
    }
    return ms;
}
/* after b200da_last_kernel_ms: device time of the Gram kernels (which = 0) or the solve kernels (which = 1) */
float b200da_last_phase_ms(b200da_plan* pl, int which) {
    if (!pl) return -1.f;
    return which == 0 ? pl->gram_ms : pl->solve_ms;
}

}  // extern "C"
