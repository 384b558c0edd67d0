// Launch of the tcgen05 Gram kernel (tc_gram_kernel.cuh): its own translation unit so that the library builds in parallel.
#include "launch.cuh"
#include "tc_gram_kernel.cuh"

namespace b200da {

// column chunking of the (k + 1)(k + 2) / 2 pair columns: as few chunks as fit 512 tensor-memory columns each and, with at
// least two loader warps, the shared memory of one SM
void tc_chunking(int k, int* n_cols, int* n_chunks, int* nc) {
    const int kp = (k + 1 + 7) / 8 * 8;
    *n_cols = (k + 1) * (k + 2) / 2;
    for (*n_chunks = (*n_cols + kTcMaxCols - 1) / kTcMaxCols;; ++*n_chunks) {
        const int per = (*n_cols + *n_chunks - 1) / *n_chunks;
        *nc = (per + 31) / 32 * 32;
        if (tc_smem_bytes(kp, *nc, 2) <= kMaxSmem || *nc <= 64) break;
    }
}

int launch_tc_gram(b200da_plan* pl, const LetkfParams& L, int nblocks, cudaStream_t st) {
    TcParams P{};
    P.L = L;
    tc_chunking(pl->k, &P.n_cols, &P.n_chunks, &P.nc);
    P.kp = pl->kp;
    const Geometry& g = pl->geom;
    // the weight vanishes beyond the padded cutoff (bin space: chord length for haversine)
    const double cut = g.cut_bin * (1.0 + 1e-6);
    P.q_max = (float)(cut * cut);
    P.period = (float)g.period;
    P.w_closed = 1;
    P.asin_poly = (g.metric == B200DA_METRIC_HAVERSINE && g.cut_bin <= 0.4) ? 1 : 0;
    P.r_scale = (float)(g.metric == B200DA_METRIC_HAVERSINE ? 2.0 * g.sphere_r / g.radius : 1.0 / g.radius);
    P.eps = (float)g.eps;
    P.centre = pl->tc_centre.as<float>();
    if (!P.centre) return B200DA_ERR_STATE;
    if (const char* e = getenv("B200DA_TC_WTABLE")) P.w_closed = atoi(e) ? 0 : 1;
    P.n_load = kTcYStages;
    while (P.n_load > 2 && tc_smem_bytes(P.kp, P.nc, P.n_load) > kMaxSmem) --P.n_load;
    const size_t smem = tc_smem_bytes(P.kp, P.nc, P.n_load);
    if (smem > kMaxSmem) return B200DA_ERR_UNSUPPORTED;
    B200DA_CUDA(cudaFuncSetAttribute(k_tc_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_gram<<<nblocks * P.n_chunks, kTcThreads, smem, st>>>(P);
    B200DA_LAUNCH_CHECK();
    return B200DA_OK;
}

}  // namespace b200da
