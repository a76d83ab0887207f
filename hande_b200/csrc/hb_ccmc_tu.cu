// hande_b200: CCMC kernels + the excitation-generator probe kernel for ONE W, selected with -DHB_TU_W=<1..4>.
#include "hb_ccmc.cuh"

#if !defined(HB_TU_W)
#error "compile with -DHB_TU_W=<1..4>"
#endif
#define HB_CAT2_(a, b) a##b
#define HB_CAT2(a, b) HB_CAT2_(a, b)
#define HB_FN(name) HB_CAT2(name, HB_TU_W)

int HB_FN(hb_ccmc_cluster_w)(hb200_engine* e, const Params& p, const CcmcLaunch& L) {
    constexpr int W = HB_TU_W;
    const int c = e->cur;
    k_ccmc_cluster<W><<<(unsigned)L.nblk, 256, 0, e->stream>>>(e->sys, p, L.a, e->d_states[c], e->d_pops[c], e->d_dat[c], L.cum,
                                                               e->d_spawn[0], e->d_head, e->block_size, e->d_proc_map,
                                                               L.partials, e->d_err);
    CK(cudaGetLastError());
    return 0;
}
int HB_FN(hb_ccmc_nc_w)(hb200_engine* e, const Params& p, const CcmcLaunch& L) {
    constexpr int W = HB_TU_W;
    const int c = e->cur;
    k_ccmc_nc<W><<<(unsigned)L.nblk, 256, 0, e->stream>>>(e->sys, p, L.a, e->d_states[c], e->d_pops[c], e->d_dat[c], e->d_spawn[0],
                                                          e->d_head, e->block_size, e->d_proc_map, L.partials, nullptr, e->d_err);
    CK(cudaGetLastError());
    return 0;
}
int HB_FN(hb_ccmc_redistribute_w)(hb200_engine* e, const Params& p) {
    constexpr int W = HB_TU_W;
    const int c = e->cur;
    const long long n = e->nstates;
    k_ccmc_redistribute<W><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->sys, p, e->d_states[c], e->d_pops[c], n,
                                                                               e->d_spawn[0], e->d_head, e->block_size,
                                                                               e->d_proc_map, e->d_err);
    CK(cudaGetLastError());
    return 0;
}
int HB_FN(hb_ccmc_find_det_w)(hb200_engine* e, const Params& p) {
    constexpr int W = HB_TU_W;
    const int c = e->cur;
    k_find_det<W><<<1, 32, 0, e->stream>>>(p, e->d_states[c], e->d_pops[c], e->nstates, e->d_ll);
    CK(cudaGetLastError());
    return 0;
}
int HB_FN(hb_gen_excit_batch_w)(hb200_engine* e, const Params& p, const uint64_t* d_f, const int64_t* d_p, const uint32_t* d_a,
                                long long n, int* d_io, double* d_do, int64_t* d_ns, const double* d_rn, int nrn, int* d_nused) {
    constexpr int W = HB_TU_W;
    k_gen_excit_batch<W><<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(e->sys, p, d_f, d_p, d_a, n, e->d_proc_map, d_io, d_do,
                                                                             d_ns, d_rn, nrn, d_nused);
    CK(cudaGetLastError());
    return 0;
}
