// hande_b200: k_spawn_death instantiations of ONE (W, generator group), selected with -DHB_TU_W=<1..4>
// -DHB_TU_GROUP=<0..4>; hande_b200/build.py compiles the 20 combinations in parallel.
#include "hb_spawn.cuh"
#if HB_TU_GROUP == 0
#include "hb_spawn_hb.cuh"
#include "hb_spawn_wf.cuh"
#endif

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

#if HB_TU_GROUP == 0
// excit_gen = heat_bath: the warp-synchronous wavefront kernel (hb_spawn_hb.cuh) on a persistent grid, followed by the
// launch that spreads the attempts of deferred (huge-population) determinants over the grid
template <int W, class Mask>
static int launch_spawn_hb(hb200_engine* e, const Params& p, const SpawnLaunch& L) {
    using namespace hbw;
    if (!e->d_heavy_count) {
        e->heavy_cap = (unsigned)std::max<long long>(1 << 16, e->cfg.walker_length / HEAVY + 1024);
        void* q = nullptr;
        CK(cudaMalloc(&q, (size_t)e->heavy_cap * sizeof(HeavyItem)));
        e->owned.push_back(q); e->d_heavy_items = q;
        CK(cudaMalloc(&q, 16));
        e->owned.push_back(q); e->d_heavy_count = (unsigned*)q;
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, e->cfg.device));
        e->num_sms = prop.multiProcessorCount;
    }
    const WarpSmem WS(W, e->sys.nel, p.qn != 0);
    const size_t smem = (size_t)((e->sys.nbasis * 8 + 15) & ~15) + (size_t)NWARP * WS.total;
    static unsigned long long attr_set = 0ull;
    const unsigned long long bit = 1ull << (e->cfg.device & 63);
    if (!(attr_set & bit)) {
        CK(cudaFuncSetAttribute(k_spawn_hb<W, Mask>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_spawn_hb_heavy<W, Mask>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set |= bit;
    }
    if (smem > 200 * 1024) FAIL("spawn_death: heat_bath kernel needs more shared memory than a block has (nel too large)");
    if (getenv("HB200_WF")) {     // development switch: the wavefront kernels (hb_spawn_wf.cuh)
        using namespace hbwf;
        const int nel = e->sys.nel;
        const SelSmem SS(W, nel);
        const size_t sm1 = (size_t)((e->sys.nbasis * 8 + 15) & ~15) + (size_t)K1_WARPS * SS.total;
        static bool once = false;
        if (!once) {
            CK(cudaFuncSetAttribute(k_wf_select<W, Mask>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once = true;
        }
        const size_t sA = rec_stride(sizeof(RecA), nel, W), sD = rec_stride(sizeof(RecD), nel, W), sS = rec_stride(sizeof(RecS), nel, W);
        if (!e->d_wf_cnt) {
            void* q = nullptr;
            e->wf_cap = (unsigned)std::min<long long>(e->cfg.walker_length + 1024, 160000000ll);
            CK(cudaMalloc(&q, (size_t)e->wf_cap * sA)); e->owned.push_back(q); e->d_wf_recA = q;
            CK(cudaMalloc(&q, (size_t)e->wf_cap * sD)); e->owned.push_back(q); e->d_wf_recD = q;
            CK(cudaMalloc(&q, (size_t)e->wf_cap * sS)); e->owned.push_back(q); e->d_wf_recS = q;
            CK(cudaMalloc(&q, 64)); e->owned.push_back(q); e->d_wf_cnt = q;
        }
        CK(cudaMemsetAsync(e->d_wf_cnt, 0, 64, e->stream));
        CK(cudaMemsetAsync(e->d_heavy_count, 0, sizeof(unsigned), e->stream));
        int bps1 = (int)((227 * 1024) / (sm1 + 1024));
        bps1 = std::max(1, std::min(bps1, 4));
        const long long nt = (L.n + 31) / 32;
        const int g1 = (int)std::max<long long>(1, std::min<long long>((long long)e->num_sms * bps1, (nt + K1_WARPS - 1) / K1_WARPS));
        HeavyQueue hq;
        hq.items = (HeavyItem*)e->d_heavy_items; hq.count = e->d_heavy_count; hq.cap = e->heavy_cap;
        const int c = e->cur;
        Counters* cnt = (Counters*)e->d_wf_cnt;
        k_wf_select<W, Mask><<<g1, K1_WARPS * 32, sm1, e->stream>>>(e->sys, p, e->d_states[c], e->d_pops[c], e->d_dat[c], 0, L.n,
                                                                   (unsigned char*)e->d_wf_recA, cnt, e->wf_cap, e->d_partials,
                                                                   e->d_err, hq);
        const int g2 = e->num_sms * 6;
        k_wf_coin<W><<<g2, 256, 0, e->stream>>>(e->sys, p, e->d_states[c], (const unsigned char*)e->d_wf_recA, cnt,
                                               (unsigned char*)e->d_wf_recD, (unsigned char*)e->d_wf_recS, e->wf_cap, e->d_err);
        k_wf_double<W><<<g2, 256, 0, e->stream>>>(e->sys, p, e->d_states[c], (const unsigned char*)e->d_wf_recD, cnt, e->wf_cap,
                                                 e->d_spawn[0], e->d_head, e->block_size, e->d_proc_map, e->d_err);
        k_wf_single<W><<<g2, 256, 0, e->stream>>>(e->sys, p, e->d_states[c], (const unsigned char*)e->d_wf_recS, cnt, e->wf_cap,
                                                 e->d_spawn[0], e->d_head, e->block_size, e->d_proc_map, e->d_err);
        // deferred (huge-population) determinants: the fused kernel's heavy path
        const size_t smh = (size_t)((e->sys.nbasis * 8 + 15) & ~15) + (size_t)NWARP * WS.total;
        int bpsh = std::max(1, std::min((int)((227 * 1024) / (smh + 1024)), 4));
        k_spawn_hb_heavy<W, Mask><<<e->num_sms * bpsh, NWARP * 32, smh, e->stream>>>(e->sys, p, e->d_states[c], e->d_spawn[0], e->d_head,
                                                                                    e->block_size, e->d_proc_map, e->d_err, hq);
        CK(cudaGetLastError());
        e->launches += 4;
        e->npartials = g1;
        return 0;
    }
    int bps = (int)((227 * 1024) / (smem + 1024));
    bps = std::max(1, std::min(bps, 4));
    const long long ntile = (L.n + SLOTS - 1) / SLOTS;
    const int grid = (int)std::max<long long>(1, std::min<long long>((long long)e->num_sms * bps, (ntile + NWARP - 1) / NWARP));
    if ((long long)grid > e->max_tiles) FAIL("spawn_death: partial-sum scratch too small");
    HeavyQueue hq;
    hq.items = (HeavyItem*)e->d_heavy_items; hq.count = e->d_heavy_count; hq.cap = e->heavy_cap;
    CK(cudaMemsetAsync(e->d_heavy_count, 0, sizeof(unsigned), e->stream));
    const int c = e->cur;
    k_spawn_hb<W, Mask><<<grid, NWARP * 32, smem, e->stream>>>(e->sys, p, e->d_states[c], e->d_pops[c], e->d_dat[c], L.n, e->d_spawn[0],
                                                              e->d_head, e->block_size, e->d_proc_map, e->d_partials, e->d_err, hq);
    CK(cudaGetLastError());
    k_spawn_hb_heavy<W, Mask><<<e->num_sms * bps, NWARP * 32, smem, e->stream>>>(e->sys, p, e->d_states[c], e->d_spawn[0], e->d_head,
                                                                                e->block_size, e->d_proc_map, e->d_err, hq);
    CK(cudaGetLastError());
    e->launches++;          // the caller counts the first launch
    e->npartials = grid;
    return 0;
}
#endif

int HB_CAT4(hb_spawn_w, HB_TU_W, _g, HB_TU_GROUP)(hb200_engine* e, const Params& p, const SpawnLaunch& L) {
    constexpr int W = HB_TU_W;
    switch (L.gen) {
#if HB_TU_GROUP == 0
        case EXCIT_GEN_HEAT_BATH:
            if (getenv("HB200_WF") || getenv("HB200_MEGA")) {      // development switches: the other two designs measured
                if (e->sys.nel <= 32) return launch_spawn_hb<W, uint32_t>(e, p, L);
                return launch_spawn_hb<W, uint64_t>(e, p, L);
            }
            return launch_spawn<W, EXCIT_GEN_HEAT_BATH>(e, p, L);
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
