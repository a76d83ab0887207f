// hande_b200: k_spawn_death instantiations of ONE (W, generator group), selected with -DHB_TU_W=<1..4>
// -DHB_TU_GROUP=<0..4>; hande_b200/build.py compiles the 20 combinations in parallel.
#include "hb_spawn.cuh"

#if !defined(HB_TU_W) || !defined(HB_TU_GROUP)
#error "compile with -DHB_TU_W=<1..4> -DHB_TU_GROUP=<0..4>"
#endif
#define HB_CAT4_(a, b, c, d) a##b##c##d
#define HB_CAT4(a, b, c, d) HB_CAT4_(a, b, c, d)

template <int W, int GEN>
static int launch_spawn(hb200_engine* e, const Params& p, const SpawnLaunch& L) {
    const int eg = e->cfg.excit_gen;
    const int nsu = (eg == HB200_EXCIT_GEN_POWER_PITZER_ORDERN) ? e->sys.nel :
                    (e->sys.kind == SYS_READ_IN && eg != HB200_EXCIT_GEN_POWER_PITZER && eg != HB200_EXCIT_GEN_NO_RENORM && eg != HB200_EXCIT_GEN_NO_RENORM_SPIN &&
                     eg != HB200_EXCIT_GEN_HEAT_BATH &&
                     eg != HB200_EXCIT_GEN_HEAT_BATH_SINGLE)
                        ? 2 * e->sys.nsym_tot : 0;
    const bool hb = eg == HB200_EXCIT_GEN_HEAT_BATH;
    const size_t smem = SpawnSmem(e->W, e->sys.nel, nsu, e->sys.nbasis, hb, hb_uses_heat_bath_tables(e), !hb && p.ps_part != nullptr,
                                  p.qn != 0, GEN == EXCIT_GEN_HEAT_BATH_SINGLE).total;
    // the attribute is per device: one bit per device ordinal and instantiation
    static unsigned long long attr_set = 0ull;
    const unsigned long long bit = 1ull << (e->cfg.device & 63);
    if (!(attr_set & bit)) {
        CK(cudaFuncSetAttribute(k_spawn_death<W, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set |= bit;
    }
    const int c = e->cur;
    k_spawn_death<W, GEN><<<L.ntiles, TILE, smem, e->stream>>>(e->sys, p, e->d_states[c], e->d_pops[c], e->d_dat[c], L.n,
                                                                e->d_spawn[0], e->d_head, e->block_size, e->d_proc_map,
                                                                e->d_partials, e->d_err, L.tile0);
    CK(cudaGetLastError());
    e->npartials = L.tile0 + L.ntiles;
    return 0;
}

int HB_CAT4(hb_spawn_w, HB_TU_W, _g, HB_TU_GROUP)(hb200_engine* e, const Params& p, const SpawnLaunch& L) {
    constexpr int W = HB_TU_W;
    switch (L.gen) {
#if HB_TU_GROUP == 0
        case EXCIT_GEN_HEAT_BATH: return launch_spawn<W, EXCIT_GEN_HEAT_BATH>(e, p, L);
#elif HB_TU_GROUP == 1
        case EXCIT_GEN_HEAT_BATH_UNIFORM: return launch_spawn<W, EXCIT_GEN_HEAT_BATH_UNIFORM>(e, p, L);
        case EXCIT_GEN_HEAT_BATH_SINGLE: return launch_spawn<W, EXCIT_GEN_HEAT_BATH_SINGLE>(e, p, L);
#elif HB_TU_GROUP == 2
        case EXCIT_GEN_POWER_PITZER_OCC: return launch_spawn<W, EXCIT_GEN_POWER_PITZER_OCC>(e, p, L);
        case EXCIT_GEN_POWER_PITZER_OCC_IJ: return launch_spawn<W, EXCIT_GEN_POWER_PITZER_OCC_IJ>(e, p, L);
#elif HB_TU_GROUP == 3
        case EXCIT_GEN_RENORM: return launch_spawn<W, EXCIT_GEN_RENORM>(e, p, L);
        case EXCIT_GEN_RENORM_SPIN: return launch_spawn<W, EXCIT_GEN_RENORM_SPIN>(e, p, L);
        case EXCIT_GEN_NO_RENORM: return launch_spawn<W, EXCIT_GEN_NO_RENORM>(e, p, L);
        case EXCIT_GEN_NO_RENORM_SPIN: return launch_spawn<W, EXCIT_GEN_NO_RENORM_SPIN>(e, p, L);
#else
#if HB_TU_W <= 4      // molecular generators: reference-mapped tables hold byte orbital lists (nbasis <= 254)
        case EXCIT_GEN_POWER_PITZER: return launch_spawn<W, EXCIT_GEN_POWER_PITZER>(e, p, L);
        case EXCIT_GEN_POWER_PITZER_ORDERN: return launch_spawn<W, EXCIT_GEN_POWER_PITZER_ORDERN>(e, p, L);
#endif
        case GEN_UEG: return launch_spawn<W, GEN_UEG>(e, p, L);
        case GEN_UEG_PP: return launch_spawn<W, GEN_UEG_PP>(e, p, L);
#endif
        default: break;
    }
    FAIL("spawn_death: generator not in this object file's group");
}
