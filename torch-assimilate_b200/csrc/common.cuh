// Shared definitions of the B200 LETKF engine: geometry of the cell grid, the distance functions and
// the Gaspari-Cohn tapers evaluated on the device.
//
// Reference semantics restated here (paths relative to /root/reference):
//   taper      pytassim/localization/gaspari_cohn.py:78-136 (GaspariCohn), :172-254 (GaspariCohnInf)
//   distance   user `dist_func` (gaspari_cohn.py:125); closed set documented in include/b200da.h
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/b200da.h"

namespace b200da {

constexpr double kAmbiguityBand = 1e-13;   // |w - eps| below this is reported as ambiguous
constexpr int kMaxRuns = 512;              // candidate cell columns per grid-point block (incl. periodic split)
constexpr int kCellsPerCutoff = 6;         // cell edge = cutoff / 6
constexpr int64_t kMaxCells = (1 << 22);

// Everything a kernel needs to map coordinates to cells and to evaluate the localization weight.
struct Geometry {
    int metric;
    int taper;
    int nd;            // dimensions of the bin space that are in use (1..3); unused leading dims are 0
    int n_coord;
    int periodic;      // last dimension wraps (PERIODIC1D)
    double radius;     // Gaspari-Cohn length scale c
    double eps;        // use_obs = w > eps
    double rcut;       // w(r) <= eps for every r >= rcut (conservative)
    double cut_bin;    // rcut * c expressed in bin-space units (chord length for HAVERSINE)
    double period;
    double sphere_r;
    double org[3];     // origin of the cell grid in bin space
    double h[3];       // cell edge per dimension
    int nc[3];         // cells per dimension
    int ncell;         // nc[0]*nc[1]*nc[2]; cell id ncell is the discard bin (outside the domain)
    int n_ext;         // extra distance components |x_g - x_o| on further coordinate columns (0..2), one taper factor each
    double ext_radius[2];
};

struct __align__(32) Pos4 {     // bin-space position + payload (original index, as raw bits)
    double x, y, z;
    long long id;
};

// One (grid point, observation) pair in ORIGINAL indices with a taper value: the records of the ambiguity protocol.  The Gram
// kernel appends every pair whose taper value lies inside the ambiguity band to the plan's list (the mask w > eps of
// gaspari_cohn.py:135 is then decided on the host with the reference's own numpy expression); the host hands its decisions
// back as overrides: `w` replaces the device's weight of that pair (0 = not a local observation).
struct PairRec { long long gid; long long oid; double w; };
constexpr int kAmbCapacity = 4096;          // recorded ambiguous pairs per plan between two b200da_pending_status calls

// device-side status words of a plan
enum { kStatusRunOverflow = 1 };            // a block needed more than kMaxRuns candidate cell columns: results invalid
struct PlanStatus { unsigned int error; unsigned int pad; unsigned long long amb_found; };

__device__ __forceinline__ double apply_override(const PairRec* __restrict__ over, int n_over, long long gid, long long oid,
                                                 double w) {
    for (int o = 0; o < n_over; ++o)
        if (over[o].oid == oid && over[o].gid == gid) w = over[o].w;
    return w;
}

__host__ __device__ inline double deg2rad(double d) { return d * 0.017453292519943295; }

// Coordinates (struct of arrays, n_coord rows of length n) -> bin-space position.
// 1-D metrics live on the last axis so that a candidate range is one contiguous run of cells.
__device__ inline void bin_position(const Geometry& g, const double* __restrict__ coord, int64_t n, int64_t i,
                                    double& x, double& y, double& z) {
    x = 0.0; y = 0.0; z = 0.0;
    if (g.metric == B200DA_METRIC_HAVERSINE) {
        const double phi = deg2rad(coord[i]);
        const double lam = deg2rad(coord[n + i]);
        double sp, cp, sl, cl;
        sincos(phi, &sp, &cp);
        sincos(lam, &sl, &cl);
        x = cp * cl; y = cp * sl; z = sp;
    } else if (g.nd == 1) {
        z = coord[i];
    } else if (g.nd == 2) {
        y = coord[i]; z = coord[n + i];
    } else {
        x = coord[i]; y = coord[n + i]; z = coord[2 * n + i];
    }
}

// Cell index along one dimension (monotone in v, so candidate ranges computed with the same formula
// are conservative).  Returns -1 / nc when outside.
__device__ inline int cell_coord(const Geometry& g, int dim, double v) {
    if (g.nc[dim] == 1 && !(g.periodic && dim == 2)) {
        // single cell: still reject points outside the domain slab
        const double t = (v - g.org[dim]) / g.h[dim];
        return (t < 0.0) ? -1 : (t >= 1.0 ? 1 : 0);
    }
    const double t = floor((v - g.org[dim]) / g.h[dim]);
    if (t < 0.0) return -1;
    if (t >= (double)g.nc[dim]) return g.nc[dim];
    return (int)t;
}

__device__ inline int cell_of(const Geometry& g, double x, double y, double z) {
    int cz = cell_coord(g, 2, z);
    if (g.periodic) {                       // coordinates are validated to lie in [0, period]
        if (cz >= g.nc[2]) cz = g.nc[2] - 1;
        if (cz < 0) cz = 0;
    }
    const int cx = cell_coord(g, 0, x), cy = cell_coord(g, 1, y);
    if (cx < 0 || cx >= g.nc[0] || cy < 0 || cy >= g.nc[1] || cz < 0 || cz >= g.nc[2]) return g.ncell;
    return (cx * g.nc[1] + cy) * g.nc[2] + cz;
}

// ---- tapers -------------------------------------------------------------------------------------------------

// GaspariCohn: gaspari_cohn.py:78-95, selection logic :126-134 (f2 where r < 2, overwritten by f1 where r < 1).
__device__ __forceinline__ double taper_gc(double r) {
    if (r < 1.0) {
        const double r2 = r * r;
        return fma(r2, fma(r, fma(r, fma(r, -0.25, 0.5), 0.625), -5.0 / 3.0), 1.0);
    }
    if (r < 2.0) {
        const double poly = fma(r, fma(r, fma(r, fma(r, fma(r, 1.0 / 12.0, -0.5), 0.625), 5.0 / 3.0), -5.0), 4.0);
        return poly - (2.0 / 3.0) / r;
    }
    return 0.0;                 // also for NaN distances: every comparison is false (gaspari_cohn.py:128)
}

// GaspariCohnInf: gaspari_cohn.py:172-210, selection :247-252 (thresholds 2, 1.5, 1, 0.5).
__device__ __forceinline__ double taper_gcinf(double r) {
    if (r < 0.5) {
        const double r2 = r * r;
        return fma(r2, fma(r, fma(r, fma(r, -28.0 / 33.0, 8.0 / 11.0), 20.0 / 11.0), -80.0 / 33.0), 1.0);
    }
    if (r < 1.0) {
        const double poly = fma(r, fma(r, fma(r, fma(r, fma(r, 20.0 / 33.0, -16.0 / 11.0), 0.0), 100.0 / 33.0),
                                       -45.0 / 11.0), 51.0 / 22.0);
        return poly - 7.0 / (44.0 * r);
    }
    if (r < 1.5) {
        const double poly = fma(r, fma(r, fma(r, fma(r, fma(r, -4.0 / 11.0, 16.0 / 11.0), -10.0 / 11.0),
                                              -100.0 / 33.0), 5.0), -61.0 / 22.0);
        return poly + 115.0 / (132.0 * r);
    }
    if (r < 2.0) {
        const double poly = fma(r, fma(r, fma(r, fma(r, fma(r, 4.0 / 33.0, -8.0 / 11.0), 10.0 / 11.0), 80.0 / 33.0),
                                       -80.0 / 11.0), 64.0 / 11.0);
        return poly - 32.0 / (33.0 * r);
    }
    return 0.0;
}

__device__ __forceinline__ double taper_eval(int taper, double r) {
    return taper == B200DA_TAPER_GCINF ? taper_gcinf(r) : taper_gc(r);
}

// ---- distances ------------------------------------------------------------------------------------------------

// Distance in the metric's own units between two bin-space positions.  ABS1D / PERIODIC1D / EUCLID follow the
// numpy expression order without FMA contraction, so the value is bit-identical to the numpy metric objects in
// pytassim_b200/localization/metrics.py; HAVERSINE goes through the chord of the unit vectors (equal to the
// haversine formula up to rounding).
__device__ __forceinline__ double metric_distance(const Geometry& g, double gx, double gy, double gz,
                                                  double ox, double oy, double oz) {
    switch (g.metric) {
        case B200DA_METRIC_ABS1D:
            return fabs(gz - oz);
        case B200DA_METRIC_PERIODIC1D: {
            const double d = fabs(gz - oz);
            return fmin(d, g.period - d);
        }
        case B200DA_METRIC_EUCLID: {
            const double dz = oz - gz;
            if (g.nd == 1) return sqrt(__dmul_rn(dz, dz));
            const double dy = oy - gy;
            if (g.nd == 2) return sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz)));
            const double dx = ox - gx;
            return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        }
        default: {   // HAVERSINE
            const double dx = ox - gx, dy = oy - gy, dz = oz - gz;
            double a = 0.25 * fma(dx, dx, fma(dy, dy, dz * dz));
            a = fmin(fmax(a, 0.0), 1.0);
            return 2.0 * g.sphere_r * asin(sqrt(a));
        }
    }
}

// Cheap, conservative proximity measure in bin space used to reject candidates against a whole block of grid
// points: Euclidean in bin space (chord for the sphere), wrapped for the periodic line.
__device__ __forceinline__ double bin_distance(const Geometry& g, double ax, double ay, double az,
                                               double bx, double by, double bz) {
    if (g.nd == 1 && g.metric != B200DA_METRIC_HAVERSINE) {
        double d = fabs(az - bz);
        if (g.periodic) d = fmin(d, g.period - d);
        return d;
    }
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return sqrt(fma(dx, dx, fma(dy, dy, dz * dz)));
}

// Localization weight of one (grid point, observation) pair: returns w if w > eps else 0
// (gaspari_cohn.py:126-135); `ambiguous` is set when |w - eps| is inside the ambiguity band.
// ge / oe: the extra coordinate values of the grid point / observation (g.n_ext each; may be null when n_ext = 0).  A
// dist_func that returns several rows gets one taper factor per row, each row divided by its own radius, multiplied in
// row order starting from 1 (gaspari_cohn.py:124-134); the candidate search only uses the first component, which is
// conservative because every factor is <= 1.
__device__ __forceinline__ double pair_weight(const Geometry& g, double gx, double gy, double gz,
                                              double ox, double oy, double oz, const double* __restrict__ ge,
                                              const double* __restrict__ oe, bool& ambiguous) {
    const double dist = metric_distance(g, gx, gy, gz, ox, oy, oz);
    const double r = dist / g.radius;                      // gaspari_cohn.py:127
    double w = taper_eval(g.taper, r);
    for (int e = 0; e < g.n_ext; ++e) {
        const double re = fabs(ge[e] - oe[e]) / g.ext_radius[e];
        w = __dmul_rn(w, taper_eval(g.taper, re));          // gaspari_cohn.py:134
    }
    ambiguous = fabs(w - g.eps) < kAmbiguityBand;
    return (w > g.eps) ? w : 0.0;                          // gaspari_cohn.py:135
}

// ---- tabulated taper of the Gram kernels -----------------------------------------------------------------------------------
// The direct evaluation above costs ~150 FP64 instructions per (grid point, observation) pair for the haversine metric (a
// square root, an arc sine, two divisions, the polynomial) on the datapath the DMMA instructions need.  The Gram kernels
// therefore read the weight as a function of the BIN-SPACE distance u (chord length for haversine, the metric's own distance
// otherwise) from a table: the pieces of the taper (breakpoints r = 1, 2 or 0.5, 1, 1.5, 2) map to u-segments, every segment
// is cut into `nint` equal intervals, and every interval carries the degree-5 interpolant of w(u) through 6 Chebyshev nodes
// (coefficients in t = 2 (u - u_left) / h - 1, computed on the host in long double; b200da.cu: build_taper_table).  The
// interpolation error is below 1e-16 for radii up to a third of the sphere, i.e. smaller than the rounding of the direct
// evaluation; the mask w > eps is applied to the tabulated value, and pairs inside the ambiguity band go to the host as before.
// The neighbour-list kernels keep the direct evaluation (their weights are compared with numpy's).
constexpr int kTaperSegMax = 4;
struct TaperTab {
    const double* coef;              // [nseg * nint][6], null: direct evaluation
    int nseg, nint;
    double ub[kTaperSegMax + 1];     // u at the breakpoints, ub[0] = 0, ub[nseg] = u where the taper reaches 0
    double scale[kTaperSegMax];      // nint / (ub[s + 1] - ub[s])
};
__host__ __device__ inline size_t taper_tab_doubles(const TaperTab& tt) { return tt.coef ? (size_t)tt.nseg * tt.nint * 6 : 0; }

// bin-space distance of the pair, the argument of the table
__device__ __forceinline__ double taper_tab_arg(const Geometry& g, double gx, double gy, double gz, double ox, double oy, double oz) {
    switch (g.metric) {
        case B200DA_METRIC_ABS1D: return fabs(gz - oz);
        case B200DA_METRIC_PERIODIC1D: { const double d = fabs(gz - oz); return fmin(d, g.period - d); }
        default: {                  // EUCLID (unused leading coordinates are 0 in bin space) and HAVERSINE (chord)
            const double dx = ox - gx, dy = oy - gy, dz = oz - gz;
            return sqrt(fma(dx, dx, fma(dy, dy, dz * dz)));
        }
    }
}
__device__ __forceinline__ double taper_tab_eval(const TaperTab& tt, const double* __restrict__ tab, double u) {
    if (!(u < tt.ub[tt.nseg])) return 0.0;                  // beyond the support (and NaN distances, gaspari_cohn.py:128)
    int s = 0;
#pragma unroll
    for (int i = 1; i < kTaperSegMax; ++i) if (i < tt.nseg && u >= tt.ub[i]) s = i;
    const double x = (u - tt.ub[s]) * tt.scale[s];
    const int i = min(__double2int_rd(x), tt.nint - 1);
    const double t = fma(2.0, x - (double)i, -1.0);
    const double2* c = reinterpret_cast<const double2*>(tab + (size_t)(s * tt.nint + i) * 6);
    const double2 c01 = c[0], c23 = c[1], c45 = c[2];
    return fma(t, fma(t, fma(t, fma(t, fma(t, c45.y, c45.x), c23.y), c23.x), c01.y), c01.x);
}
// pair_weight with the tabulated taper (single distance row only)
__device__ __forceinline__ double pair_weight_tab(const Geometry& g, const TaperTab& tt, const double* __restrict__ tab,
                                                  double gx, double gy, double gz, double ox, double oy, double oz, bool& ambiguous) {
    const double w = taper_tab_eval(tt, tab, taper_tab_arg(g, gx, gy, gz, ox, oy, oz));
    ambiguous = fabs(w - g.eps) < kAmbiguityBand;
    return (w > g.eps) ? w : 0.0;                          // gaspari_cohn.py:135
}

// the taper value itself (no mask): what the ambiguity records carry
__device__ __forceinline__ double pair_weight_raw(const Geometry& g, double gx, double gy, double gz,
                                                  double ox, double oy, double oz, const double* __restrict__ ge,
                                                  const double* __restrict__ oe) {
    double w = taper_eval(g.taper, metric_distance(g, gx, gy, gz, ox, oy, oz) / g.radius);
    for (int e = 0; e < g.n_ext; ++e) w = __dmul_rn(w, taper_eval(g.taper, fabs(ge[e] - oe[e]) / g.ext_radius[e]));
    return w;
}

// ---- state / weight I/O in the plan's dtype --------------------------------------------------------------------
// The ensemble-space algebra always runs in FP64; with an FP32 plan only the HBM-resident arrays (state, analysis,
// observation-space staging copy, exported weights) are FP32.  `f32` is uniform over a launch.
__device__ __forceinline__ double ld_io(const void* __restrict__ p, int64_t i, int f32) {
    return f32 ? (double)reinterpret_cast<const float*>(p)[i] : reinterpret_cast<const double*>(p)[i];
}
__device__ __forceinline__ void st_io(void* __restrict__ p, int64_t i, double v, int f32) {
    if (f32) reinterpret_cast<float*>(p)[i] = (float)v; else reinterpret_cast<double*>(p)[i] = v;
}

// ---- tile-packed symmetric matrices ----------------------------------------------------------------------------
// Lower-triangle 8x8 tiles, tile (mt, nt), nt <= mt, at ((mt (mt+1))/2 + nt) * 64 doubles; element (r, c) of a tile at
// r * 8 + (c ^ ((r & 2) << 1)).  The XOR makes both the direct and the transposed DMMA fragment reads of a tile
// bank-conflict free.  This is the hand-over format between the Gram kernel and the solve kernels.
__host__ __device__ constexpr int tri_tiles(int kt) { return kt * (kt + 1) / 2; }
__host__ __device__ constexpr int tile_off(int mt, int nt) { return (mt * (mt + 1) / 2 + nt) * 64; }   // nt <= mt
__host__ __device__ constexpr int tile_elem(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }
// offset of element (i, j), j <= i
__host__ __device__ constexpr int sym_off(int i, int j) { return tile_off(i >> 3, j >> 3) + tile_elem(i & 7, j & 7); }

}  // namespace b200da
