// hande_b200: CUDA engine (sm_100a) for the FCIQMC propagation hot path + its C ABI.
//
// One engine = one GPU = one MPI rank of the reference.  Data layout in HBM:
//   main walker list (particle_t, src/qmc_data.f90:615-682), double-buffered:
//       states[N][W] uint64 (sorted ascending, bit_str_cmp order), pops[N] int64 (encoded), dat[N] double
//   spawn store (spawn_t, src/spawn_data.F90:35-145): two buffers of `spawned_walker_length` elements,
//       element = [f(0..W-1), population, flag] int64  (32 B for W=2), partitioned in nprocs blocks
//   system tables (integral store, symmetry tables, heat-bath alias tables): replicated per GPU, L2-resident
//       except the nb^4 heat-bath tables.
//
// Kernels (all HBM/L2-bound integer + fp64 scalar work; no tensor cores - nothing is a dense contraction):
//   k_spawn_death     fused: decode, initiator flag, projected energy, decide_nattempts, spawning attempts
//                     (load-balanced over a 256-state tile), stochastic death; warp-aggregated append
//   k_radix_*         LSD radix sort of the spawn list on the bit-string key (8-bit digits)
//   k_annihilate      segmented sum of equal keys + initiator flag algebra + binary search into the main list
//   k_round_count     stochastic rounding of main-list populations + per-tile survivor counts
//   k_merge           single pass merge of survivors and new determinants into the other main-list buffer
//   k_sc0             <D|H|D> - H00 for new determinants
#include <cuda_runtime.h>
#include <nccl.h>   // types only: the library is bound at run time (see NcclApi below)
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/hande_b200.h"
#include "hb_core.cuh"

using namespace hb;

static thread_local std::string g_err;
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + std::to_string(__LINE__); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)
// NCCL is resolved lazily with dlopen/dlsym instead of a link-time dependency: a host process that also uses
// torch (bench.py, the multi-GPU tests) must end up with ONE libnccl.so.2, and torch's bundled copy (2.28) is newer
// than the system one (2.27).  Order: a copy already loaded in the process, $HB200_NCCL_LIB, the system library.
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) { const char* p = getenv("HB200_NCCL_LIB"); if (p && *p) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define HB_SYM(field, name) field = (decltype(field))dlsym(h, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
        HB_SYM(GetUniqueId, "ncclGetUniqueId") HB_SYM(CommInitRank, "ncclCommInitRank") HB_SYM(CommDestroy, "ncclCommDestroy")
        HB_SYM(AllGather, "ncclAllGather") HB_SYM(Broadcast, "ncclBroadcast") HB_SYM(Send, "ncclSend") HB_SYM(Recv, "ncclRecv")
        HB_SYM(GroupStart, "ncclGroupStart") HB_SYM(GroupEnd, "ncclGroupEnd") HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
        return true;
    }
};
static NcclApi g_nccl;

#define NCK(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t _e = (call);                                                                  \
        if (_e != ncclSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + g_nccl.GetErrorString(_e) + " @" + std::to_string(__LINE__); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)
#define FAIL(msg)        \
    do {                 \
        g_err = (msg);   \
        return 1;        \
    } while (0)

constexpr int TILE = 256;  // states per block in the fused spawn kernel and in the merge passes

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// exclusive scan over a block of TILE threads; returns exclusive prefix, total in *total.  warp_sums: >= 8 ints smem
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) {
        int s = warp_sums[w];
        if (w < warp) off += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return off + incl - v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int W>
__device__ __forceinline__ void load_det(const uint64_t* p, uint64_t* f) {
    if (W == 2) {
        ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(p));   // streamed once per cycle
        f[0] = v.x; f[1] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) f[k] = p[k];
    }
}
template <int W>
__device__ __forceinline__ void store_det(uint64_t* p, const uint64_t* f) {
    if (W == 2) {
        *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(f[0], f[1]);
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) p[k] = f[k];
    }
}

// lower_bound in the sorted main list: first index with states[idx] >= key
template <int W>
__device__ __forceinline__ long long lower_bound_det(const uint64_t* states, long long n, const uint64_t* key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        uint64_t f[W];
        load_det<W>(states + mid * W, f);
        if (det_less<W>(f, key)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// Kernel: fused spawn + death + estimators  (src/fciqmc.f90:315-371, 635-769)
// ------------------------------------------------------------------------------------------------
struct SpawnPartials {  // one per block; reduced in fixed order by k_reduce_partials
    double pe, d0;
    long long ndeath, npart, nattempts;
};

constexpr int SINGLES_CHUNK = 64;  // single excitations whose pgen terms are staged in shared memory at a time

// Shared-memory carve-up of k_spawn_death (same arithmetic on host and device).  For the heat-bath generator the
// phase-A staging area (column i of hb_ij_w at the occupied orbitals, [q][thread]) shares its storage with the buffers
// that are only live in the later phases.
struct SpawnSmem {
    size_t sf, shash, ssign, sscan, swarp, sred, siw, sw, sh1, shm, ssp, spsum, sterm, sok, sq, scnt, sflag, slo, sperm, ssq,
        ssi, socc, ssu, sps, sdf, total;
    // heat_bath: the original heat-bath generator (phase buffers); hb_stage: any generator that selects i, j from the
    // heat-bath weights (needs the hb_i_w copy and the per-thread staging area of nel doubles)
    // ps: pattempt_update statistics are accumulated (per thread: two doubles and two counters)
    // qn: quasi-Newton propagator (fock_sum of each state of the tile)
    __host__ __device__ SpawnSmem(int W, int nel, int nsu, int nb, bool heat_bath, bool hb_stage, bool ps, bool qn) {
        size_t o = 0;
        sf = o;     o += (size_t)TILE * W * 8;
        sred = o;   o += 40 * 8;
        shash = o;  o += hb_stage ? 0 : (size_t)TILE * 8;         // stream selector per state (recomputed per attempt when
                                                                   // shared memory is scarce: it also buys L1 capacity)
        siw = o;    o += hb_stage ? (size_t)nb * 8 : 0;           // copy of hb_i_w
        // ---- union: phase A staging | phase B..F buffers
        const size_t u0 = o;
        sw = o;
        size_t v = u0;
        sh1 = v;    v += heat_bath ? (size_t)TILE * 8 : 0;        // signed slater_condon1(i,a) per attempt slot
        shm = v;    v += heat_bath ? (size_t)3 * TILE * 8 : 0;    // |slater_condon1| of the three other orderings
        ssp = v;    v += heat_bath ? (size_t)2 * TILE * 8 : 0;    // singles: hmod_ia, ij_tot
        spsum = v;  v += heat_bath ? (size_t)TILE * 8 : 0;        // singles: sum of pgen terms
        sterm = v;  v += heat_bath ? (size_t)SINGLES_CHUNK * nel * 8 : 0;  // singles: pgen terms of one chunk
        sok = v;    v += heat_bath ? (size_t)SINGLES_CHUNK * nel : 0;
        v = (v + 3) & ~(size_t)3;
        sq = v;     v += heat_bath ? (size_t)4 * TILE * 4 : 0;    // request queues: [TILE] phase B, [3*TILE] phase D
        const size_t stage = hb_stage ? (size_t)TILE * nel * 8 : 0;
        o = u0 + (stage > (v - u0) ? stage : (v - u0));
        o = (o + 7) & ~(size_t)7;
        // ---- end of union
        sscan = o;  o += (size_t)(TILE + 1) * 4;
        swarp = o;  o += 8 * 4;
        scnt = o;   o += 4 * 4;                                    // queue counters
        sflag = o;  o += TILE;
        ssign = o;  o += TILE;                                     // sign of the parent population (attempt_to_spawn)
        slo = o;    o += heat_bath ? TILE : 0;                     // tile-state index of each attempt slot
        sperm = o;  o += heat_bath ? TILE : 0;
        ssq = o;    o += heat_bath ? TILE : 0;                     // queue of single-excitation slots
        ssi = o;    o += heat_bath ? 2 * TILE : 0;                 // singles: i, a
        socc = o;   o += (size_t)TILE * nel;
        ssu = o;    o += (size_t)TILE * nsu;
        o = (o + 7) & ~(size_t)7;
        sps = o;    o += ps ? (size_t)TILE * 24 : 0;
        sdf = o;    o += qn ? (size_t)TILE * 8 : 0;
        total = (o + 15) & ~(size_t)15;
    }
};

// create_excited_det (src/excitations.F90:365-406)
template <int W>
__device__ __forceinline__ void make_child(const uint64_t* f, const Gen& g, uint64_t* child) {
#pragma unroll
    for (int k = 0; k < W; ++k) child[k] = f[k];
    child[(g.from1 - 1) >> 6] &= ~(1ull << ((g.from1 - 1) & 63));
    child[(g.to1 - 1) >> 6] |= (1ull << ((g.to1 - 1) & 63));
    if (g.nexcit == 2) {
        child[(g.from2 - 1) >> 6] &= ~(1ull << ((g.from2 - 1) & 63));
        child[(g.to2 - 1) >> 6] |= (1ull << ((g.to2 - 1) & 63));
    }
}

// GEN: compile-time generator of this instantiation (one kernel per generator keeps the code - and the instruction
// cache footprint - to what the run actually executes): EXCIT_GEN_* for read_in systems, GEN_UEG for the UEG.
enum { GEN_UEG = 100, GEN_UEG_PP = 101 };
template <int W, int GEN>
__global__ void __launch_bounds__(TILE, (GEN == EXCIT_GEN_HEAT_BATH || GEN == EXCIT_GEN_HEAT_BATH_UNIFORM) ? 3 : 4)
k_spawn_death(Sys s, Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ pops,
              const double* __restrict__ dat, long long nstates, int64_t* __restrict__ spawn,
              unsigned long long* __restrict__ head, long long block_size, const int* __restrict__ proc_map,
              SpawnPartials* __restrict__ partials, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nel = s.nel;
    constexpr bool heat_bath = (GEN == EXCIT_GEN_HEAT_BATH);
    constexpr bool hb_stage = heat_bath || (GEN == EXCIT_GEN_HEAT_BATH_UNIFORM) || (GEN == EXCIT_GEN_HEAT_BATH_SINGLE) ||
                              (GEN == EXCIT_GEN_POWER_PITZER_OCC_IJ);
    const int nsu = (GEN == EXCIT_GEN_POWER_PITZER_ORDERN) ? nel : (GEN == EXCIT_GEN_RENORM || GEN == EXCIT_GEN_RENORM_SPIN || GEN == EXCIT_GEN_HEAT_BATH_UNIFORM || GEN == EXCIT_GEN_POWER_PITZER_OCC ||
                     GEN == EXCIT_GEN_POWER_PITZER_OCC_IJ) ? 2 * s.nsym_tot : 0;
    const bool ps_on = !heat_bath && p.ps_part != nullptr;
    const SpawnSmem L(W, nel, nsu, s.nbasis, heat_bath, hb_stage, ps_on, p.qn != 0);
    double* sdf = reinterpret_cast<double*>(smem_raw + L.sdf);
    double* sps_h = reinterpret_cast<double*>(smem_raw + L.sps);                  // [2][TILE]: singles, doubles
    unsigned* sps_n = reinterpret_cast<unsigned*>(smem_raw + L.sps + 16 * TILE);  // [2][TILE]
    if (ps_on) {
        sps_h[threadIdx.x] = 0.0; sps_h[TILE + threadIdx.x] = 0.0;
        sps_n[threadIdx.x] = 0u; sps_n[TILE + threadIdx.x] = 0u;
    }
    uint64_t* sf = reinterpret_cast<uint64_t*>(smem_raw + L.sf);
    uint8_t* ssign = smem_raw + L.ssign;
    uint64_t* shash = reinterpret_cast<uint64_t*>(smem_raw + L.shash);
    double* sred = reinterpret_cast<double*>(smem_raw + L.sred);
    double* sh1 = reinterpret_cast<double*>(smem_raw + L.sh1);
    double* shm = reinterpret_cast<double*>(smem_raw + L.shm);
    double* ssp = reinterpret_cast<double*>(smem_raw + L.ssp);
    double* spsum = reinterpret_cast<double*>(smem_raw + L.spsum);
    double* siw = reinterpret_cast<double*>(smem_raw + L.siw);
    double* sw = reinterpret_cast<double*>(smem_raw + L.sw);
    double* sterm = reinterpret_cast<double*>(smem_raw + L.sterm);
    uint8_t* sok = smem_raw + L.sok;
    int* sscan = reinterpret_cast<int*>(smem_raw + L.sscan);
    int* swarp = reinterpret_cast<int*>(smem_raw + L.swarp);
    uint32_t* sq1 = reinterpret_cast<uint32_t*>(smem_raw + L.sq);
    uint32_t* sq2 = sq1 + TILE;
    int* scnt = reinterpret_cast<int*>(smem_raw + L.scnt);
    uint8_t* sflag = smem_raw + L.sflag;
    uint8_t* slo = smem_raw + L.slo;
    uint8_t* sperm = smem_raw + L.sperm;
    uint8_t* ssq = smem_raw + L.ssq;
    uint8_t* ssi = smem_raw + L.ssi;
    uint8_t* socc = smem_raw + L.socc;
    uint8_t* ssu = smem_raw + L.ssu;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const long long idx = (long long)blockIdx.x * TILE + tid;
    const int E = W + 2;

    double pe = 0.0, d0 = 0.0;
    long long ndeath = 0, npart = 0;
    int natt = 0;
    if (hb_stage)
        for (int k = tid; k < s.nbasis; k += TILE) siw[k] = s.hb_i_w[k];
    if (idx < nstates) {
        uint64_t f[W];
        load_det<W>(states + idx * W, f);
        const int64_t pop = __ldcs(pops + idx);
        const double Kii = __ldcs(dat + idx);
#pragma unroll
        for (int k = 0; k < W; ++k) sf[tid * W + k] = f[k];
        ssign[tid] = pop < 0;
        uint8_t* occ = socc + tid * nel;
        decode_det<W>(f, occ);
        if (GEN == EXCIT_GEN_POWER_PITZER_ORDERN) find_diff_ref_cdet<W>(s, p.f0, f, ssu + tid * nsu);
        else if (nsu) build_symunocc_masks<W>(s, f, ssu + tid * nsu);
        const uint64_t h = det_hash64<W>(f);
        if (!hb_stage) shash[tid] = h;
        const double real_pop = (double)pop / (double)p.real_factor;
        // set_parent_flag (src/ifciqmc.f90:13-57)
        sflag[tid] = (fabs(real_pop) > p.initiator_pop) ? 0 : 1;
        // update_proj_energy_mol (src/energy_evaluation.F90:906-986)
        bool is_ref;
        double hm = proj_energy_hmatel<W>(s, p, f, occ, is_ref);
        if (is_ref) d0 = real_pop; else pe = hm * real_pop;
        PhiloxStream rng;
        rng.begin(p.seed, p.cycle, RNG_NATTEMPTS, h, 0);
        natt = decide_nattempts(rng, real_pop);
        rng.begin(p.seed, p.cycle, RNG_DEATH, h, 0);
        int64_t kill_abs;
        double death_weight = 1.0;
        if (p.qn) {
            const double dfock = qn_fock_sum(s, p, occ);
            sdf[tid] = dfock;
            death_weight = qn_weighting(p, dfock);
        }
        const int64_t newpop = stochastic_death(rng, p, Kii, pop, kill_abs, death_weight);
        pops[idx] = newpop;
        ndeath = kill_abs;
        npart = newpop < 0 ? -newpop : newpop;
    }
    int T;
    const int excl = block_excl_scan(natt, swarp, &T);
    sscan[tid] = excl;
    if (tid == 0) sscan[TILE] = T;
    __syncthreads();

    for (int base = 0; base < T; base += TILE) {
        const int a = base + tid;
        const bool active = a < T;
        int lo = 0, att = 0;
        if (active) {
            int hi = TILE;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (sscan[mid] <= a) lo = mid; else hi = mid;
            }
            att = a - sscan[lo];
        }
        uint64_t f[W];
#pragma unroll
        for (int k = 0; k < W; ++k) f[k] = sf[lo * W + k];
        PhiloxStream rng;
        rng.begin(p.seed, p.cycle, RNG_SPAWN, hb_stage ? det_hash64<W>(f) : shash[lo], (uint32_t)att);
        if (!hb_stage) rng.prefetch();   // uniform generators draw inside divergent rejection loops
        Gen g;
        if (heat_bath) {
            // ---- phase A: i, j, a for every attempt of the round
            if (tid < 4) scnt[tid] = 0;
            __syncthreads();
            HbState st;
            st.allowed = false; st.need_ia = false; st.dbl = true; st.need_k = 0;
            if (active) hb_phase_a<W>(rng, s, f, socc + lo * nel, st, siw, sw + tid, TILE);
            __syncthreads();   // the request queues share their storage with the phase-A staging area
            if (active) {
                slo[tid] = (uint8_t)lo;
                if (st.allowed && st.need_ia) {
                    const int q = atomicAdd(&scnt[0], 1);
                    sq1[q] = (uint32_t)tid | ((uint32_t)st.i << 8) | ((uint32_t)st.a << 16);
                }
            }
            __syncthreads();
            // ---- phase B: the queued slater_condon1(i,a), one request per thread
            for (int r = tid; r < scnt[0]; r += TILE) {
                const uint32_t rq = sq1[r];
                const int slot = rq & 255, fr = (rq >> 8) & 255, to = (rq >> 16) & 255, l = slo[slot];
                uint64_t ff[W];
#pragma unroll
                for (int k = 0; k < W; ++k) ff[k] = sf[l * W + k];
                bool pm;
                sh1[slot] = hb_sc1<W>(s, ff, socc + l * nel, fr, to, pm);
                sperm[slot] = pm;
            }
            __syncthreads();
            // ---- phase C: single/double coin, b; queue the remaining slater_condon1 and the singles
            if (active) {
                if (st.allowed && st.need_ia) { st.h_ia = sh1[tid]; st.perm_ia = sperm[tid] != 0; }
                hb_phase_c<W>(rng, s, f, st);
                if (st.allowed) {
                    if (st.dbl) {
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (st.need_k & (1u << k)) {
                                int fr, to, ot;
                                hb_ordering(st, k, fr, to, ot);
                                const int q = atomicAdd(&scnt[1], 1);
                                sq2[q] = (uint32_t)tid | ((uint32_t)fr << 8) | ((uint32_t)to << 16) | ((uint32_t)k << 24);
                            }
                    } else {
                        const int q = atomicAdd(&scnt[2], 1);
                        ssq[q] = (uint8_t)tid;
                        ssi[tid] = (uint8_t)st.i; ssi[TILE + tid] = (uint8_t)st.a;
                        ssp[tid] = st.hmod_ia; ssp[TILE + tid] = st.ij_tot;
                    }
                }
            }
            __syncthreads();
            // ---- phase D: queued slater_condon1 of the other orderings (dense)
            for (int r = tid; r < scnt[1]; r += TILE) {
                const uint32_t rq = sq2[r];
                const int slot = rq & 255, fr = (rq >> 8) & 255, to = (rq >> 16) & 255, k = rq >> 24, l = slo[slot];
                uint64_t ff[W];
#pragma unroll
                for (int kk = 0; kk < W; ++kk) ff[kk] = sf[l * W + kk];
                bool pm;
                shm[k * TILE + slot] = fabs(hb_sc1<W>(s, ff, socc + l * nel, fr, to, pm));
            }
            // ---- phase E: singles.  The (division-heavy) pgen terms are evaluated one per thread over all
            //      (single, occupied orbital) pairs of a chunk, then summed in occ_list order, one single per thread.
            {
                const int ns = scnt[2];
                for (int c0 = 0; c0 < ns; c0 += SINGLES_CHUNK) {
                    const int nc = min(SINGLES_CHUNK, ns - c0);
                    for (int w = tid; w < nc * nel; w += TILE) {
                        const int r = w / nel, q = w - r * nel;
                        const int slot = ssq[c0 + r], l = slo[slot];
                        double term = 0.0;
                        const bool ok = hb_single_term(s, ssi[slot], ssi[TILE + slot], ssp[slot], ssp[TILE + slot],
                                                       socc[l * nel + q], term);
                        sterm[w] = term;
                        sok[w] = ok;
                    }
                    __syncthreads();
                    for (int r = tid; r < nc; r += TILE) {
                        double psum = 0.0;
                        for (int q = 0; q < nel; ++q)
                            if (sok[r * nel + q]) psum = psum + sterm[r * nel + q];
                        spsum[ssq[c0 + r]] = psum;
                    }
                    __syncthreads();
                }
            }
            __syncthreads();
            // ---- phase F: pgen, H_ij
            double hmk[3] = {0.0, 0.0, 0.0};
            double psum = 0.0;
            if (active && st.allowed) {
                if (st.dbl) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (st.need_k & (1u << k)) hmk[k] = shm[k * TILE + tid];
                } else {
                    psum = spsum[tid];
                }
            }
            hb_phase_f<W>(s, f, st, hmk, psum, g);
        } else if (active) {
            if (GEN == GEN_UEG) gen_excit_ueg_no_renorm<W>(rng, s, f, socc + lo * nel, g);
            else if (GEN == GEN_UEG_PP) gen_excit_ueg_power_pitzer<W>(rng, s, f, socc + lo * nel, g);
            else if (GEN == EXCIT_GEN_HEAT_BATH_UNIFORM)
                gen_excit_heat_bath_uniform<W>(rng, s, p, f, socc + lo * nel, ssu + lo * nsu, siw, sw + tid, TILE, g);
            else if (GEN == EXCIT_GEN_HEAT_BATH_SINGLE)
                gen_excit_heat_bath_uniform<W, true>(rng, s, p, f, socc + lo * nel, nullptr, siw, sw + tid, TILE, g);
            else if (GEN == EXCIT_GEN_POWER_PITZER_OCC)     // also cauchy_schwarz_occ (p.excit_gen picks the integral)
                gen_excit_power_pitzer_occ<W>(rng, s, p, f, socc + lo * nel, ssu + lo * nsu, nullptr, nullptr, 0, g);
            else if (GEN == EXCIT_GEN_POWER_PITZER_OCC_IJ)  // also cauchy_schwarz_occ_ij
                gen_excit_power_pitzer_occ<W>(rng, s, p, f, socc + lo * nel, ssu + lo * nsu, siw, sw + tid, TILE, g);
            else if (GEN == EXCIT_GEN_RENORM) gen_excit_renorm<W>(rng, s, p, f, socc + lo * nel, ssu + lo * nsu, g);
            else if (GEN == EXCIT_GEN_RENORM_SPIN) gen_excit_renorm<W, true>(rng, s, p, f, socc + lo * nel, ssu + lo * nsu, g);
            else if (GEN == EXCIT_GEN_POWER_PITZER) gen_excit_power_pitzer_ref<W>(rng, s, p, f, socc + lo * nel, g);
            else if (GEN == EXCIT_GEN_POWER_PITZER_ORDERN)     // ssu holds ref_cdet_occ_list of each state (nsu = nel)
                gen_excit_power_pitzer_orderN<W>(rng, s, p, f, socc + lo * nel, ssu + lo * nsu, g);
            else if (GEN == EXCIT_GEN_NO_RENORM_SPIN) gen_excit_no_renorm<W, true>(rng, s, p, f, socc + lo * nel, g);
            else gen_excit_no_renorm<W>(rng, s, p, f, socc + lo * nel, g);
        }
        int64_t nspawn = 0;
        uint64_t child[W];
        int dest = 0, pflag = 0;
        if (active) {
            double hmq = g.hmatel;
            if (p.qn && g.allowed) hmq = hmq * qn_spawned_weighting(p, sdf[lo], g);   // spawn_standard (src/spawning.F90:101-103)
            if (ps_on && g.allowed) {   // update_p_single_double_data (src/spawning.F90:104-109,2139-2215)
                const int k = (g.nexcit == 2) ? TILE : 0;
                sps_h[k + tid] = sps_h[k + tid] + (fabs(hmq) * (g.nexcit == 2 ? p.pattempt_double : p.pattempt_single)) / g.pgen;
                sps_n[k + tid] += 1u;
            }
            nspawn = attempt_to_spawn(rng, p, hmq, g.pgen, ssign[lo] ? (int64_t)-1 : (int64_t)1);
            if (nspawn != 0) {
                make_child<W>(f, g, child);
                // create_spawned_particle[_initiator]_truncated (src/spawning.F90:1186-1319)
                if (p.trunc_level >= 0 && excit_level<W>(child, p.f0) > p.trunc_level) {
                    nspawn = 0;
                } else {
                    // assign_particle_processor (src/spawning.F90:770-838)
                    dest = (p.nprocs > 1) ? proc_map[owner_slot(child, s.nbasis, p.hash_seed, p.nprocs, p.nslots)] : 0;
                    pflag = p.initiator ? sflag[lo] : 0;
                }
            }
        }
        __syncwarp();
        const unsigned has = __ballot_sync(0xffffffffu, nspawn != 0);
        if (nspawn != 0) {
            // add_[flagged_]spawned_particle (src/spawning.F90:907-1018): warp-aggregated pointer bump per destination
            const unsigned peers = (p.nprocs > 1) ? __match_any_sync(has, dest) : has;
            const int leader = __ffs(peers) - 1;
            const int rank = __popc(peers & ((1u << lane) - 1u));
            unsigned long long slot0 = 0;
            if (lane == leader) slot0 = atomicAdd(&head[dest], (unsigned long long)__popc(peers));
            slot0 = __shfl_sync(peers, slot0, leader);
            const long long slot = (long long)slot0 + rank;
            if (slot < block_size) {
                int64_t* dst = spawn + ((long long)dest * block_size + slot) * E;
                if (W == 2) {
                    reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(child[0], child[1]);
                    reinterpret_cast<longlong2*>(dst)[1] = make_longlong2((long long)nspawn, (long long)pflag);
                } else {
#pragma unroll
                    for (int k = 0; k < W; ++k) dst[k] = (int64_t)child[k];
                    dst[W] = nspawn;
                    dst[W + 1] = pflag;
                }
            } else {
                atomicOr(err, 1);  // spawn%error: no space left in the spawning array
            }
        }
    }

    if (ps_on) {
        const double a = warp_sum_d(sps_h[tid]), b = warp_sum_d(sps_h[TILE + tid]);
        const long long c = warp_sum_ll((long long)sps_n[tid]), d = warp_sum_ll((long long)sps_n[TILE + tid]);
        __syncthreads();
        if (lane == 0) {
            sred[warp] = a; sred[8 + warp] = b;
            reinterpret_cast<long long*>(sred)[16 + warp] = c;
            reinterpret_cast<long long*>(sred)[24 + warp] = d;
        }
        __syncthreads();
        if (tid == 0) {
            PsPartials out = {0.0, 0.0, 0, 0};
            for (int w = 0; w < TILE / 32; ++w) {
                out.h_pgen_singles_sum += sred[w]; out.h_pgen_doubles_sum += sred[8 + w];
                out.excit_gen_singles += reinterpret_cast<long long*>(sred)[16 + w];
                out.excit_gen_doubles += reinterpret_cast<long long*>(sred)[24 + w];
            }
            p.ps_part[blockIdx.x] = out;
        }
    }
    // deterministic block reduction of the estimators
    double r0 = warp_sum_d(pe), r1 = warp_sum_d(d0);
    long long r2 = warp_sum_ll(ndeath), r3 = warp_sum_ll(npart);
    __syncthreads();
    if (lane == 0) {
        sred[warp] = r0; sred[8 + warp] = r1;
        reinterpret_cast<long long*>(sred)[16 + warp] = r2;
        reinterpret_cast<long long*>(sred)[24 + warp] = r3;
    }
    __syncthreads();
    if (tid == 0) {
        SpawnPartials out;
        out.pe = 0.0; out.d0 = 0.0; out.ndeath = 0; out.npart = 0;
        for (int w = 0; w < TILE / 32; ++w) {
            out.pe += sred[w]; out.d0 += sred[8 + w];
            out.ndeath += reinterpret_cast<long long*>(sred)[16 + w];
            out.npart += reinterpret_cast<long long*>(sred)[24 + w];
        }
        out.nattempts = T;
        partials[blockIdx.x] = out;
    }
}

// ------------------------------------------------------------------------------------------------
// CCMC (src/ccmc.f90:603-896): one thread per cluster-selection attempt - select_cluster, do_ccmc_accumulation,
// spawner_ccmc and stochastic_ccmc_death; spawned and killed excips are appended to the spawn list and then go through
// the same sort / annihilation / merge kernels as FCIQMC (direct_annihilation).
// ------------------------------------------------------------------------------------------------
// add_spawned_particle (src/spawning.F90:907-1018): warp-aggregated pointer bump; every lane of the warp must call it
template <int W>
__device__ __forceinline__ void append_spawn_warp(const uint64_t* f, int64_t nspawn, int dest, int nprocs,
                                                  int64_t* __restrict__ spawn, unsigned long long* __restrict__ head,
                                                  long long block_size, int* __restrict__ err) {
    constexpr int E = W + 2;
    const int lane = threadIdx.x & 31;
    const unsigned has = __ballot_sync(0xffffffffu, nspawn != 0);
    if (nspawn != 0) {
        const unsigned peers = (nprocs > 1) ? __match_any_sync(has, dest) : has;
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        unsigned long long slot0 = 0;
        if (lane == leader) slot0 = atomicAdd(&head[dest], (unsigned long long)__popc(peers));
        slot0 = __shfl_sync(peers, slot0, leader);
        const long long slot = (long long)slot0 + rank;
        if (slot < block_size) {
            int64_t* dst = spawn + ((long long)dest * block_size + slot) * E;
#pragma unroll
            for (int k = 0; k < W; ++k) dst[k] = (int64_t)f[k];
            dst[W] = nspawn;
            dst[W + 1] = 0;
        } else {
            atomicOr(err, 1);
        }
    }
}

struct CcmcPartials { double pe, d0; long long ndeath, nattempts_spawn; };

// block sums of the pattempt_update statistics of a 256-thread CCMC block (all threads call it)
__device__ __forceinline__ void ps_block_reduce(PsPartials* out, double hs, double hd, int ns, int nd) {
    __shared__ double sh[2][8];
    __shared__ long long sn[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double a = warp_sum_d(hs), b = warp_sum_d(hd);
    const long long c = warp_sum_ll((long long)ns), d = warp_sum_ll((long long)nd);
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; sn[0][warp] = c; sn[1][warp] = d; }
    __syncthreads();
    if (tid == 0) {
        PsPartials o = {0.0, 0.0, 0, 0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            o.h_pgen_singles_sum += sh[0][w]; o.h_pgen_doubles_sum += sh[1][w];
            o.excit_gen_singles += sn[0][w]; o.excit_gen_doubles += sn[1][w];
        }
        out[blockIdx.x] = o;
    }
}
// running totals of the report loop: acc[0..3] += block sums in fixed order
__global__ void k_reduce_ps(const PsPartials* __restrict__ part, int n, double* __restrict__ acc) {
    __shared__ double sh[4][32];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        v[0] += part[i].h_pgen_singles_sum; v[1] += (double)part[i].excit_gen_singles;
        v[2] += part[i].h_pgen_doubles_sum; v[3] += (double)part[i].excit_gen_doubles;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 4; ++k) { v[k] = warp_sum_d(v[k]); if (lane == 0) sh[k][warp] = v[k]; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
        acc[threadIdx.x] = acc[threadIdx.x] + t;
    }
}

template <int W>
__global__ void __launch_bounds__(256)      // 80 registers, 3 blocks/SM: 64 (4 blocks) and 99 (2 blocks) are both ~18 % slower
k_ccmc_cluster(Sys s, Params p, CcmcArgs a, const uint64_t* __restrict__ states, const int64_t* __restrict__ pops,
               const double* __restrict__ dat, const long long* __restrict__ cum_enc, int64_t* __restrict__ spawn,
               unsigned long long* __restrict__ head, long long block_size, const int* __restrict__ proc_map,
               CcmcPartials* __restrict__ partials, int* __restrict__ err) {
    __shared__ double sd[2][8];
    __shared__ long long sl[2][8];
    __shared__ unsigned short sperm[256];
    __shared__ unsigned char swc[8][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // The attempts of a block are dealt to its threads grouped by cluster size (a stable counting sort on the size each
    // attempt's stream will draw first), so that the lanes of a warp walk select_cluster / collapse_cluster in step:
    // half of all attempts are the empty cluster, a quarter single excitors, ...  Which thread runs an attempt changes
    // neither its random stream nor its result.
    long long idx;
    {
        const long long idx0 = (long long)blockIdx.x * blockDim.x + tid;
        int cls = 7;
        if (idx0 < a.nattempts) {
            if (idx0 >= a.nattempts - a.nD0_select) {
                cls = 0;
            } else {
                PhiloxStream r0;
                r0.begin(p.seed, p.cycle, RNG_SPAWN, det_hash64<W>(p.f0) + (uint64_t)p.iproc * 0x9E3779B97F4A7C15ull,
                         (uint32_t)(idx0 + 1));
                const double rand = r0.next();
                double psize = 0.0;
                int n = -1;
                for (int i = 0; i <= a.max_cluster_size - a.min_cluster_size - 1; ++i) {
                    psize = psize + 1.0 / (double)(1ll << (i + 1));
                    if (rand < psize) { n = i + a.min_cluster_size; break; }
                }
                if (n == -1) n = a.max_cluster_size;
                cls = min(max(n, 0), 6);
            }
        }
        int rnk = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == k);
            if (cls == k) rnk = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) swc[warp][k] = (unsigned char)__popc(m);
        }
        __syncthreads();
        int base = 0;
        for (int k = 0; k < cls; ++k)
            for (int w = 0; w < 8; ++w) base += swc[w][k];
        for (int w = 0; w < warp; ++w) base += swc[w][cls];
        sperm[base + rnk] = (unsigned short)tid;
        __syncthreads();
        idx = (long long)blockIdx.x * blockDim.x + sperm[tid];
    }
    double pe = 0.0, d0 = 0.0;
    long long ndeath = 0, nas = 0;
    int64_t nspawn = 0, nkill = 0;
    double ps_hs = 0.0, ps_hd = 0.0;
    int ps_ns = 0, ps_nd = 0;
    uint64_t cf[W], child[W];
#pragma unroll
    for (int k = 0; k < W; ++k) { cf[k] = 0; child[k] = 0; }
    int dest_s = 0, dest_k = 0;
    if (idx < a.nattempts) {
        PhiloxStream rng;
        rng.begin(p.seed, p.cycle, RNG_SPAWN, det_hash64<W>(p.f0) + (uint64_t)p.iproc * 0x9E3779B97F4A7C15ull,
                  (uint32_t)(idx + 1));
        Cluster cl;
        const bool det_D0 = idx >= a.nattempts - a.nD0_select;    // deterministic selections of the reference (full_nc)
        if (det_D0) {
            // create_null_cluster(prob = nprocs * nD0_select) (src/ccmc.f90:803-812)
#pragma unroll
            for (int k = 0; k < W; ++k) cf[k] = p.f0[k];
            cl.nexcitors = 0; cl.excitation_level = 0; cl.sign = 1; cl.first_pos = 0;
            cl.amplitude = a.D0_normalisation;
            cl.pselect = (double)a.nprocs * (double)a.nD0_select;
        } else {
            ccmc_select_cluster<W>(rng, p, a, states, pops, cum_enc, cf, cl);
        }
        if (cl.excitation_level >= 0 && cl.excitation_level <= a.ex_level + 2) {
            uint8_t occ[HB_MAXNEL], su[64];
            decode_det<W>(cf, occ);
            if (s.kind == SYS_READ_IN && p.excit_gen != EXCIT_GEN_NO_RENORM && p.excit_gen != EXCIT_GEN_NO_RENORM_SPIN &&
                p.excit_gen != EXCIT_GEN_HEAT_BATH)
                build_symunocc_masks<W>(s, cf, su);
            // do_ccmc_accumulation (src/ccmc.f90:1007-1101)
            bool is_ref;
            const double hm0 = proj_energy_hmatel<W>(s, p, cf, occ, is_ref);
            const double wpop = cl.amplitude * cl.sign / cl.pselect;
            if (is_ref) d0 = wpop; else pe = hm0 * wpop;
            nas = 1;
            // spawner_ccmc (src/ccmc_death_spawning.f90:11-211)
            Gen g;
            gen_excit<W>(rng, s, p, cf, occ, su, g);
            const double hmatel = g.hmatel * cl.amplitude * 1.0 * cl.sign;
            const double pgen = g.pgen * cl.pselect * 1;
            if (p.ps_part && g.allowed) {   // src/ccmc_death_spawning.f90:150-157
                if (g.nexcit == 2) { ps_hd = (fabs(hmatel) * p.pattempt_double) / pgen; ps_nd = 1; }
                else { ps_hs = (fabs(hmatel) * p.pattempt_single) / pgen; ps_ns = 1; }
            }
            nspawn = attempt_to_spawn(rng, p, hmatel, pgen, (int64_t)1);
            if (nspawn != 0) {
                make_child<W>(cf, g, child);
                const int lvl = excit_level<W>(child, p.f0);
                if (ccmc_excitor_sign<W>(p.f0, child, lvl) < 0) nspawn = -nspawn;
                if (p.trunc_level >= 0 && lvl > p.trunc_level) nspawn = 0;   // create_spawned_particle_truncated
                else dest_s = (p.nprocs > 1) ? proc_map[owner_slot_shift<W>(child, s.nbasis, p.hash_seed, p.ccmc_shift, p.ccmc_freq,
                                                                            p.nprocs, p.nslots)] : 0;
            }
            // stochastic_ccmc_death + stochastic_death_attempt (src/ccmc_death_spawning.f90:213-441)
            if (!det_D0 && cl.excitation_level <= a.ex_level && (cl.nexcitors >= 2 || !a.full_nc)) {
                const double pe_old = p.proj_energy_old;
                double KiiAi;
                if (cl.nexcitors == 0) KiiAi = ((-pe_old) * 1.0 + (pe_old - p.shift) * 1.0) * cl.amplitude;
                else if (cl.nexcitors == 1) KiiAi = ((dat[cl.first_pos - 1] - pe_old) * 1.0 + (pe_old - p.shift) * 1.0) * cl.amplitude;
                else {
                    const double hii = (s.kind == SYS_UEG) ? slater_condon0_ueg(s, occ) : slater_condon0(s, occ);
                    KiiAi = ((hii - p.H00) - pe_old) * 1.0 * cl.amplitude;
                }
                KiiAi = 1.0 * (double)p.real_factor * KiiAi;
                KiiAi = KiiAi * p.tau / cl.pselect;
                double pdeath = fabs(KiiAi);
                if (pdeath < (double)p.spawn_cutoff) {
                    nkill = (pdeath > rng.next() * (double)p.spawn_cutoff) ? p.spawn_cutoff : 0;
                } else {
                    nkill = (int64_t)pdeath;
                    pdeath = pdeath - (double)nkill;
                    if (pdeath > rng.next()) nkill++;
                }
                ndeath = nkill;
                if (nkill != 0) {
                    if (KiiAi > 0) nkill = -nkill;
                    dest_k = (p.nprocs > 1) ? proc_map[owner_slot_shift<W>(cf, s.nbasis, p.hash_seed, p.ccmc_shift, p.ccmc_freq,
                                                                            p.nprocs, p.nslots)] : 0;
                }
            }
        }
    }
    __syncwarp();
    append_spawn_warp<W>(child, nspawn, dest_s, p.nprocs, spawn, head, block_size, err);
    append_spawn_warp<W>(cf, nkill, dest_k, p.nprocs, spawn, head, block_size, err);
    if (p.ps_part) ps_block_reduce(p.ps_part, ps_hs, ps_hd, ps_ns, ps_nd);
    const double r0 = warp_sum_d(pe), r1 = warp_sum_d(d0);
    const long long r2 = warp_sum_ll(ndeath), r3 = warp_sum_ll(nas);
    if (lane == 0) { sd[0][warp] = r0; sd[1][warp] = r1; sl[0][warp] = r2; sl[1][warp] = r3; }
    __syncthreads();
    if (tid == 0) {
        CcmcPartials out;
        out.pe = 0.0; out.d0 = 0.0; out.ndeath = 0; out.nattempts_spawn = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { out.pe += sd[0][w]; out.d0 += sd[1][w]; out.ndeath += sl[0][w]; out.nattempts_spawn += sl[1][w]; }
        partials[blockIdx.x] = out;
    }
}
// full_nc: every excitor is a non-composite cluster of its own - select_nc_cluster (src/ccmc_selection.f90:462-561),
// do_nc_ccmc_propagation (src/ccmc.f90:1275-1360) - and every excip (the reference included) dies in place through
// stochastic_ccmc_death_nc (src/ccmc_death_spawning.f90:443-547).  Thread per excitor; launched after k_ccmc_cluster,
// which reads the populations this kernel changes.
template <int W>
__global__ void __launch_bounds__(256)
k_ccmc_nc(Sys s, Params p, CcmcArgs a, const uint64_t* __restrict__ states, int64_t* __restrict__ pops,
          const double* __restrict__ dat, int64_t* __restrict__ spawn, unsigned long long* __restrict__ head,
          long long block_size, const int* __restrict__ proc_map, CcmcPartials* __restrict__ partials,
          long long* __restrict__ ndeath_nc_out, int* __restrict__ err) {
    __shared__ double sd[2][8];
    __shared__ long long sl[2][8];
    constexpr int E = W + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long i = (long long)blockIdx.x * blockDim.x + tid;
    double pe = 0.0, d0 = 0.0;
    long long ndeath_nc = 0, nas = 0;
    double ps_hs = 0.0, ps_hd = 0.0;
    int ps_ns = 0, ps_nd = 0;
    if (i < a.nstates) {
        uint64_t f[W];
        load_det<W>(states + i * W, f);
        const int64_t pop = pops[i];
        const uint64_t h = det_hash64<W>(f);
        const bool isD0 = (i + 1 == a.D0_pos);
        PhiloxStream rng;
        if (!isD0) {
            const double amp = (double)pop / (double)p.real_factor;
            const int level = excit_level<W>(f, p.f0);
            const int sign = ccmc_excitor_sign<W>(p.f0, f, level);
            uint8_t occ[HB_MAXNEL], su[64];
            decode_det<W>(f, occ);
            if (s.kind == SYS_READ_IN && p.excit_gen != EXCIT_GEN_NO_RENORM && p.excit_gen != EXCIT_GEN_NO_RENORM_SPIN &&
                p.excit_gen != EXCIT_GEN_HEAT_BATH)
                build_symunocc_masks<W>(s, f, su);
            bool is_ref;
            const double hm0 = proj_energy_hmatel<W>(s, p, f, occ, is_ref);
            pe = hm0 * (amp * sign / 1.0);
            rng.begin(p.seed, p.cycle, RNG_NATTEMPTS, h, 0);
            const int nsp = decide_nattempts(rng, fabs(amp) / 1.0);
            nas = nsp;
            const double unit = amp / fabs(amp);
            for (int ip = 0; ip < nsp; ++ip) {
                rng.begin(p.seed, p.cycle, RNG_SPAWN, h, (uint32_t)ip);
                Gen g;
                gen_excit<W>(rng, s, p, f, occ, su, g);
                const double hmatel = g.hmatel * unit * 1.0 * sign;
                const double pgen = g.pgen * 1.0 * 1;
                if (p.ps_part && g.allowed) {
                    if (g.nexcit == 2) { ps_hd = ps_hd + (fabs(hmatel) * p.pattempt_double) / pgen; ps_nd += 1; }
                    else { ps_hs = ps_hs + (fabs(hmatel) * p.pattempt_single) / pgen; ps_ns += 1; }
                }
                int64_t nspawn = attempt_to_spawn(rng, p, hmatel, pgen, (int64_t)1);
                if (nspawn != 0) {
                    uint64_t child[W];
                    make_child<W>(f, g, child);
                    const int lvl = excit_level<W>(child, p.f0);
                    if (ccmc_excitor_sign<W>(p.f0, child, lvl) < 0) nspawn = -nspawn;
                    if (!(p.trunc_level >= 0 && lvl > p.trunc_level)) {
                        const int dest = (p.nprocs > 1) ? proc_map[owner_slot_shift<W>(child, s.nbasis, p.hash_seed, p.ccmc_shift,
                                                                                         p.ccmc_freq, p.nprocs, p.nslots)] : 0;
                        const long long slot = (long long)atomicAdd(&head[dest], 1ull);
                        if (slot < block_size) {
                            int64_t* dst = spawn + ((long long)dest * block_size + slot) * E;
#pragma unroll
                            for (int k = 0; k < W; ++k) dst[k] = (int64_t)child[k];
                            dst[W] = nspawn;
                            dst[W + 1] = 0;
                        } else {
                            atomicOr(err, 1);
                        }
                    }
                }
            }
        }
        // stochastic_ccmc_death_nc
        {
            const double pe_old = p.proj_energy_old;
            double KiiAi;
            if (isD0) KiiAi = ((-pe_old) * 1.0 + (pe_old - p.shift) * 1.0) * (double)pop;
            else KiiAi = ((dat[i] - pe_old) * 1.0 + (pe_old - p.shift) * 1.0) * (double)pop;
            KiiAi = KiiAi * 1.0;
            double pdeath = p.tau * fabs(KiiAi);
            int64_t nkill = (int64_t)pdeath;
            pdeath = pdeath - (double)nkill;
            rng.begin(p.seed, p.cycle, RNG_DEATH, h, 0);
            if (pdeath > rng.next()) nkill = nkill + 1;
            if (nkill != 0) {
                if (KiiAi > 0) nkill = -nkill;
                pops[i] = pop + nkill;
                ndeath_nc = nkill < 0 ? -nkill : nkill;
            }
        }
    }
    if (p.ps_part) ps_block_reduce(p.ps_part, ps_hs, ps_hd, ps_ns, ps_nd);
    const double r0 = warp_sum_d(pe), r1 = warp_sum_d(d0);
    const long long r2 = warp_sum_ll(ndeath_nc), r3 = warp_sum_ll(nas);
    if (lane == 0) { sd[0][warp] = r0; sd[1][warp] = r1; sl[0][warp] = r2; sl[1][warp] = r3; }
    __syncthreads();
    if (tid == 0) {
        CcmcPartials out;
        out.pe = 0.0; out.d0 = 0.0; out.ndeath = 0; out.nattempts_spawn = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { out.pe += sd[0][w]; out.d0 += sd[1][w]; out.ndeath += sl[0][w]; out.nattempts_spawn += sl[1][w]; }
        partials[blockIdx.x] = out;
    }
    (void)ndeath_nc_out;
}
__global__ void k_ccmc_reduce(const CcmcPartials* __restrict__ partials, int n, CcmcPartials* out) {
    __shared__ double sd[2][32];
    __shared__ long long sl[2][32];
    double pe = 0.0, d0 = 0.0;
    long long nd = 0, na = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { pe += partials[i].pe; d0 += partials[i].d0; nd += partials[i].ndeath; na += partials[i].nattempts_spawn; }
    pe = warp_sum_d(pe); d0 = warp_sum_d(d0); nd = warp_sum_ll(nd); na = warp_sum_ll(na);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sd[0][warp] = pe; sd[1][warp] = d0; sl[0][warp] = nd; sl[1][warp] = na; }
    __syncthreads();
    if (threadIdx.x == 0) {
        CcmcPartials o; o.pe = 0; o.d0 = 0; o.ndeath = 0; o.nattempts_spawn = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { o.pe += sd[0][w]; o.d0 += sd[1][w]; o.ndeath += sl[0][w]; o.nattempts_spawn += sl[1][w]; }
        *out = o;
    }
}
// redistribute_particles (src/qmc_common.F90:505-595): excips whose owner under the current hash shift is another
// rank are moved to that rank's block of the spawn list and zeroed in the main list
template <int W>
__global__ void __launch_bounds__(256)
k_ccmc_redistribute(Sys s, Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ pops, long long n,
                    int64_t* __restrict__ spawn, unsigned long long* __restrict__ head, long long block_size,
                    const int* __restrict__ proc_map, int* __restrict__ err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t f[W];
#pragma unroll
    for (int k = 0; k < W; ++k) f[k] = 0;
    int64_t pop = 0;
    int dest = 0;
    if (i < n) {
        load_det<W>(states + i * W, f);
        dest = proc_map[owner_slot_shift<W>(f, s.nbasis, p.hash_seed, p.ccmc_shift, p.ccmc_freq, p.nprocs, p.nslots)];
        if (dest != p.iproc) {
            pop = pops[i];
            pops[i] = 0;
        }
    }
    append_spawn_warp<W>(f, pop, dest, p.nprocs, spawn, head, block_size, err);
}
// find_D0 (src/ccmc_utils.F90:21-67): position (1-based, 0 = absent) and population of f0 in the sorted main list
template <int W>
__global__ void k_find_det(Params p, const uint64_t* __restrict__ states, const int64_t* __restrict__ pops, long long n,
                           long long* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long pos = lower_bound_det<W>(states, n, p.f0);
    bool hit = false;
    if (pos < n) {
        uint64_t f[W];
        load_det<W>(states + pos * W, f);
        hit = det_eq<W>(f, p.f0);
    }
    out[0] = hit ? pos + 1 : 0;
    out[1] = hit ? pops[pos] : 0;
}
// inclusive prefix sums of |pop| (encoded) with the reference skipped (cumulative_population): exact integer scan
constexpr int SCAN64_ITEMS = 8;
__global__ void __launch_bounds__(256) k_cum_block(const int64_t* __restrict__ pops, long long n, long long skip,
                                                  long long* __restrict__ out, long long* __restrict__ block_sums) {
    __shared__ long long sw[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long base = ((long long)blockIdx.x * 256 + tid) * SCAN64_ITEMS;
    long long v[SCAN64_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN64_ITEMS; ++k) {
        const long long i = base + k;
        long long x = (i < n && i != skip) ? pops[i] : 0;
        x = x < 0 ? -x : x;
        sum += x;
        v[k] = sum;
    }
    long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sw[warp] = incl;
    __syncthreads();
    long long off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) off += sw[w]; tot += sw[w]; }
    const long long excl = off + incl - sum;
#pragma unroll
    for (int k = 0; k < SCAN64_ITEMS; ++k)
        if (base + k < n) out[base + k] = excl + v[k];
    if (tid == 0) block_sums[blockIdx.x] = tot;
}
__global__ void k_cum_sums(long long* data, int m) {   // in-place exclusive scan of the block sums, single block
    __shared__ long long swarp[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < m; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = (i < m) ? data[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        long long off = 0, tot = 0;
        for (int w = 0; w < nw; ++w) { const long long sv = swarp[w]; if (w < warp) off += sv; tot += sv; }
        const long long c = carry;
        if (i < m) data[i] = c + off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_cum_add(long long* __restrict__ out, long long n, const long long* __restrict__ offs) {
    const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * SCAN64_ITEMS;
    const long long off = offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN64_ITEMS; ++k)
        if (base + k < n) out[base + k] += off;
}

struct CycleStats {
    double pe, d0;              // this cycle
    long long ndeath, npart_after_death, nattempts_spawn;
    long long nkept, npart_new; // after merge
};

__global__ void k_reduce_partials(const SpawnPartials* __restrict__ partials, int n, CycleStats* st) {
    __shared__ double sd[2][32];
    __shared__ long long sl[3][32];
    double pe = 0.0, d0 = 0.0;
    long long nd = 0, np = 0, na = 0;
    // fixed assignment of partials to threads + fixed-order tree => run-to-run deterministic
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        pe += partials[i].pe; d0 += partials[i].d0;
        nd += partials[i].ndeath; np += partials[i].npart; na += partials[i].nattempts;
    }
    pe = warp_sum_d(pe); d0 = warp_sum_d(d0);
    nd = warp_sum_ll(nd); np = warp_sum_ll(np); na = warp_sum_ll(na);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sd[0][warp] = pe; sd[1][warp] = d0; sl[0][warp] = nd; sl[1][warp] = np; sl[2][warp] = na; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0; long long c = 0, d = 0, e = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sd[0][w]; b += sd[1][w]; c += sl[0][w]; d += sl[1][w]; e += sl[2][w]; }
        st->pe = a; st->d0 = b; st->ndeath = c; st->npart_after_death = d; st->nattempts_spawn = e;
    }
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort of spawn elements (replaces qsort, src/sort.f90:213-395) on the key words
// ------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;

template <int E>
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_hist(const int64_t* __restrict__ in, long long n, int word, int shift, unsigned* __restrict__ hist, int nblk,
             long long chunk) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const long long start = (long long)blockIdx.x * chunk;
    const long long end = min(n, start + chunk);
    for (long long i = start + threadIdx.x; i < end; i += SORT_THREADS) {
        unsigned d = (unsigned)(((uint64_t)in[i * E + word] >> shift) & 0xFFu);
        atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan of m unsigned values in place, single block
__global__ void k_scan_u32_single(unsigned* data, long long m) {
    __shared__ unsigned swarp[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long base = 0; base < m; base += blockDim.x) {
        long long i = base + threadIdx.x;
        unsigned v = (i < m) ? data[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        unsigned off = 0, tot = 0;
        for (int w = 0; w < nw; ++w) { unsigned sv = swarp[w]; if (w < warp) off += sv; tot += sv; }
        unsigned c = carry;
        if (i < m) data[i] = c + off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
}

template <int E>
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_scatter(const int64_t* __restrict__ in, int64_t* __restrict__ out, long long n, int word, int shift,
                const unsigned* __restrict__ hist, int nblk, long long chunk) {
    __shared__ unsigned sbase[256];
    __shared__ unsigned swc[SORT_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    sbase[tid] = hist[(size_t)tid * nblk + blockIdx.x];
    const long long start = (long long)blockIdx.x * chunk;
    const long long end = min(n, start + chunk);
    for (long long t0 = start; t0 < end; t0 += SORT_THREADS) {
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) swc[w][tid] = 0;
        __syncthreads();
        const long long i = t0 + tid;
        const bool valid = i < end;
        int64_t el[E];
        unsigned d = 256u + (unsigned)lane;  // invalid lanes never match a real digit nor each other
        if (valid) {
#pragma unroll
            for (int k = 0; k < E; ++k) el[k] = in[i * E + k];
            d = (unsigned)(((uint64_t)el[word] >> shift) & 0xFFu);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) swc[warp][d] = __popc(peers);
        __syncthreads();
        {   // digit `tid`: exclusive prefix over warps on top of the running base (stable order)
            unsigned run = sbase[tid];
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) {
                unsigned c = swc[w][tid];
                swc[w][tid] = run;
                run += c;
            }
            sbase[tid] = run;
        }
        __syncthreads();
        if (valid) {
            const long long pos = (long long)swc[warp][d] + rank;
            int64_t* dst = out + pos * E;
            if (E == 4) {
                reinterpret_cast<longlong2*>(dst)[0] = make_longlong2(el[0], el[1]);
                reinterpret_cast<longlong2*>(dst)[1] = make_longlong2(el[2], el[3]);
            } else {
#pragma unroll
                for (int k = 0; k < E; ++k) dst[k] = el[k];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// generic exclusive scan of int32 (two-level recursion): used for the insert flags and tile counts
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = TILE * SCAN_ITEMS;

__global__ void __launch_bounds__(TILE) k_scan_block(const int* __restrict__ in, int* __restrict__ out, long long n,
                                                     int* __restrict__ block_sums) {
    __shared__ int swarp[8];
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int tot;
    int excl = block_excl_scan(sum, swarp, &tot);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(TILE) k_scan_add(int* __restrict__ out, long long n, const int* __restrict__ block_offs) {
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    const int off = block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += off;
}
// single-block scan for small arrays; also writes the total to *total
__global__ void k_scan_small(const int* __restrict__ in, int* __restrict__ out, long long n, int* total) {
    __shared__ int swarp[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long base = 0; base < n; base += blockDim.x) {
        long long i = base + threadIdx.x;
        int v = (i < n) ? in[i] : 0;
        int incl = warp_incl_scan(v);
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < nw; ++w) { int sv = swarp[w]; if (w < warp) off += sv; tot += sv; }
        int c = carry;
        if (i < n) out[i] = c + off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// ------------------------------------------------------------------------------------------------
// Kernel: annihilate_spawn_t[_initiator] + annihilate_main_list[_initiator] + round_low_population_spawns
// (src/spawn_data.F90:859-1101, src/annihilation.f90:294-486, 600-675) on the sorted spawn list.
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
k_annihilate(Params p, int64_t* __restrict__ sp, long long n, const uint64_t* __restrict__ states,
             int64_t* __restrict__ pops, long long nstates, int* __restrict__ ins_flag, long long* __restrict__ ins_pos) {
    constexpr int E = W + 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ins_flag[i] = 0;
    uint64_t key[W];
#pragma unroll
    for (int k = 0; k < W; ++k) key[k] = (uint64_t)sp[i * E + k];
    if (i > 0) {
        bool same = true;
#pragma unroll
        for (int k = 0; k < W; ++k) same = same && ((uint64_t)sp[(i - 1) * E + k] == key[k]);
        if (same) return;  // not the head of its segment
    }
    long long pop = 0, initiator_pop = 0, events = 0;
    for (long long j = i; j < n; ++j) {
        if (j > i) {
            bool same = true;
#pragma unroll
            for (int k = 0; k < W; ++k) same = same && ((uint64_t)sp[j * E + k] == key[k]);
            if (!same) break;
        }
        const long long pj = sp[j * E + W];
        pop += pj;
        if (p.initiator) {
            if (!(sp[j * E + W + 1] & 1)) initiator_pop += pj;
            else events += (pj < 0) ? -1 : ((pj > 0) ? 1 : 0);
        }
    }
    int flag = 0;
    if (p.initiator) {
        const bool sgn_tot = pop >= 0, sgn_ini = initiator_pop >= 0;  // Fortran sign(1,0) = +1
        const bool keep = (initiator_pop != 0 && sgn_tot == sgn_ini) || ((events < 0 ? -events : events) > 1);
        flag = keep ? 0 : 1;
    }
    if (pop == 0) return;
    const long long pos = lower_bound_det<W>(states, nstates, key);
    bool hit = false;
    if (pos < nstates) {
        uint64_t f[W];
        load_det<W>(states + pos * W, f);
        hit = det_eq<W>(f, key);
    }
    if (hit) {
        const long long cur = pops[pos];
        if (!p.initiator) pops[pos] = cur + pop;
        else if (cur != 0) pops[pos] = cur + pop;
        else if (!flag) pops[pos] = pop;
        return;
    }
    if (p.initiator && flag) return;  // spawned by non-initiators onto an unoccupied determinant
    if (p.real_amplitudes) {
        PhiloxStream rng;
        rng.begin(p.seed, p.cycle, RNG_ROUND_SPAWN, det_hash64<W>(key), 0);
        pop = stochastic_round(rng, (int64_t)pop, p.real_factor);
        if (pop == 0) return;
    }
    sp[i * E + W] = pop;
    ins_flag[i] = 1;
    ins_pos[i] = pos;
}

// compaction of the surviving new determinants: ins[k] = [f, pop, pos]
template <int W>
__global__ void __launch_bounds__(256)
k_compact_inserts(const int64_t* __restrict__ sp, long long n, const int* __restrict__ ins_flag,
                  const int* __restrict__ ins_idx, const long long* __restrict__ ins_pos, int64_t* __restrict__ ins) {
    constexpr int E = W + 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !ins_flag[i]) return;
    const long long k = ins_idx[i];
#pragma unroll
    for (int w = 0; w < W; ++w) ins[k * E + w] = sp[i * E + w];
    ins[k * E + W] = sp[i * E + W];
    ins[k * E + W + 1] = ins_pos[i];
}

// insert_new_walker: dat(1) = sc0_ptr(f) - H00 (src/annihilation.f90:820-901)
template <int W>
__global__ void __launch_bounds__(256)
k_sc0(Sys s, double H00, const uint64_t* __restrict__ dets, long long stride_words, long long n, double* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t f[W];
#pragma unroll
    for (int w = 0; w < W; ++w) f[w] = dets[k * stride_words + w];
    uint8_t occ[HB_MAXNEL];
    decode_det<W>(f, occ);
    out[k] = ((s.kind == SYS_UEG) ? slater_condon0_ueg(s, occ) : slater_condon0(s, occ)) - H00;
}

// remove_unoccupied_dets, first half (src/annihilation.f90:537-598): stochastic rounding of main-list
// populations (real amplitudes) and per-tile survivor counts.
template <int W>
__global__ void __launch_bounds__(TILE)
k_round_count(Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ pops, long long nstates,
              int* __restrict__ tile_keep) {
    __shared__ int swarp[8];
    const long long i = (long long)blockIdx.x * TILE + threadIdx.x;
    int keep = 0;
    if (i < nstates) {
        int64_t pop = pops[i];
        if (p.real_amplitudes) {
            const int64_t ap = pop < 0 ? -pop : pop;
            if (pop != 0 && ap < p.real_factor) {
                uint64_t f[W];
                load_det<W>(states + i * W, f);
                PhiloxStream rng;
                rng.begin(p.seed, p.cycle, RNG_ROUND_MAIN, det_hash64<W>(f), 0);
                pop = stochastic_round(rng, pop, p.real_factor);
                pops[i] = pop;
            }
        }
        keep = pop != 0;
    }
    int tot;
    block_excl_scan(keep, swarp, &tot);
    if (threadIdx.x == 0) tile_keep[blockIdx.x] = tot;
}

// remove_unoccupied_dets (compaction) + insert_new_walkers (src/annihilation.f90:537-598, 677-818) as ONE
// out-of-place merge: tile of TILE old states + the new determinants whose insertion point falls in the tile.
template <int W>
__global__ void __launch_bounds__(TILE)
k_merge(const uint64_t* __restrict__ states, const int64_t* __restrict__ pops, const double* __restrict__ dat,
        long long nstates, const int* __restrict__ tile_off, const int64_t* __restrict__ ins,
        const double* __restrict__ ins_dat, long long nins, uint64_t* __restrict__ ostates, int64_t* __restrict__ opops,
        double* __restrict__ odat, long long* __restrict__ part_npart, int ntiles) {
    constexpr int E = W + 2;
    __shared__ int swarp[8];
    __shared__ int skept[TILE + 1];
    __shared__ long long sk[2];
    __shared__ long long sred[8];
    const int tid = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * TILE;
    const long long t1 = min(nstates, t0 + TILE);
    const bool last = (blockIdx.x == ntiles - 1);
    if (tid < 2) {
        // inserts with pos in [t0, t1) (last tile: also pos == nstates) are a contiguous range [k_lo, k_hi)
        const long long target = (tid == 0) ? t0 : t1;
        long long lo = 0, hi = nins;
        if (tid == 1 && last) lo = nins;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (ins[mid * E + W + 1] < target) lo = mid + 1; else hi = mid;
        }
        sk[tid] = lo;
    }
    const long long m = t0 + tid;
    int keep = 0;
    int64_t pop = 0;
    if (m < t1) { pop = pops[m]; keep = pop != 0; }
    int tot;
    const int kb = block_excl_scan(keep, swarp, &tot);
    skept[tid] = kb;
    if (tid == 0) skept[TILE] = tot;
    __syncthreads();
    const long long k_lo = sk[0], k_hi = sk[1];
    const long long ns = k_hi - k_lo;
    const long long out_base = (long long)tile_off[blockIdx.x] + k_lo;
    long long npart = 0;
    if (keep) {
        // number of new determinants in this tile inserted at or before old state m
        long long lo = 0, hi = ns;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (ins[(k_lo + mid) * E + W + 1] <= m) lo = mid + 1; else hi = mid;
        }
        const long long o = out_base + kb + lo;
        uint64_t f[W];
        load_det<W>(states + m * W, f);
        store_det<W>(ostates + o * W, f);
        opops[o] = pop;
        odat[o] = dat[m];
        npart += pop < 0 ? -pop : pop;
    }
    for (long long j = tid; j < ns; j += TILE) {
        const long long k = k_lo + j;
        const long long pos = ins[k * E + W + 1];
        const int loc = (int)(pos - t0);  // 0..TILE (TILE only for pos == nstates in the last tile)
        const long long o = out_base + skept[loc] + j;
        uint64_t f[W];
#pragma unroll
        for (int w = 0; w < W; ++w) f[w] = (uint64_t)ins[k * E + w];
        store_det<W>(ostates + o * W, f);
        const int64_t ip = ins[k * E + W];
        opops[o] = ip;
        odat[o] = ins_dat[k];
        npart += ip < 0 ? -ip : ip;
    }
    npart = warp_sum_ll(npart);
    if ((tid & 31) == 0) sred[tid >> 5] = npart;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < TILE / 32; ++w) t += sred[w];
        part_npart[blockIdx.x] = t;
    }
}

// sum |pop| over a list: per-block partials (fixed order) -> k_reduce_ll
__global__ void __launch_bounds__(TILE) k_abs_sum(const int64_t* __restrict__ pops, long long n, long long* __restrict__ part) {
    __shared__ long long sl[TILE / 32];
    long long v = 0;
    for (long long i = (long long)blockIdx.x * TILE + threadIdx.x; i < n; i += (long long)gridDim.x * TILE) {
        const long long p = pops[i];
        v += p < 0 ? -p : p;
    }
    v = warp_sum_ll(v);
    if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < TILE / 32; ++w) t += sl[w];
        part[blockIdx.x] = t;
    }
}

__global__ void k_reduce_ll(const long long* __restrict__ in, int n, long long* out) {
    __shared__ long long sl[32];
    long long v = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += in[i];
    v = warp_sum_ll(v);
    if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sl[w];
        *out = t;
    }
}

// ------------------------------------------------------------------------------------------------
// System-table kernels: J/K diagonal tables and the heat-bath builder
// (init_excit_mol_heat_bath, src/excit_gen_heat_bath_mol.F90:14-256; sums in the reference's order)
// ------------------------------------------------------------------------------------------------
__global__ void k_build_JK(Sys s, double* J, double* K) {
    const int nb = s.nbasis;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * nb) return;
    const int i = t / nb + 1, j = t % nb + 1;
    J[t] = two_body(s, i, j, i, j);
    K[t] = two_body(s, i, j, j, i);
}

// single-excitation row tables C(i,a,j) = <ij|aj>, X(i,a,j) = <ij|ja> (see hb_core.cuh Sys::sc1C)
__global__ void k_build_sc1_tables(Sys s, int NT, D2* CX) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)NT * NT * NT) return;
    const int tj = (int)(t % NT), ta = (int)((t / NT) % NT), ti = (int)(t / ((long long)NT * NT));
    const int i = s.uhf ? ti + 1 : 2 * ti + 1, a = s.uhf ? ta + 1 : 2 * ta + 1, j = s.uhf ? tj + 1 : 2 * tj + 1;
    D2 v;
    v.x = two_body(s, i, j, a, j);
    v.y = two_body(s, i, j, j, a);
    CX[t] = v;
}

// ijab_w(b,a,j,i) = |<ij||ab>| for allowed (spin, symmetry, distinct) index quadruples, else 0
__global__ void k_hb_ijab_w(Sys s, double* __restrict__ w) {
    const long long nb = s.nbasis;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * nb * nb * nb) return;
    const int b = (int)(t % nb) + 1, a = (int)((t / nb) % nb) + 1, j = (int)((t / (nb * nb)) % nb) + 1,
              i = (int)(t / (nb * nb * nb)) + 1;
    double val = 0.0;
    if (i != j && a != i && a != j) {
        const int it = min(i, j), jt = max(i, j);
        const int ij_sym = sym_conj(s, cross_product(s, s.bf_sym[it], s.bf_sym[jt]));
        const int isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
        const bool spin_ok = (s.bf_ms[it] == s.bf_ms[a] && s.bf_ms[jt] == s.bf_ms[b]) ||
                             (s.bf_ms[it] == s.bf_ms[b] && s.bf_ms[jt] == s.bf_ms[a]);
        if (spin_ok && s.bf_sym[b] == isymb && b != a && b != i && b != j) {
            const int at = min(a, b), bt = max(a, b);
            val = fabs(slater_condon2_excit(s, it, jt, at, bt, false));
        }
    }
    w[t] = val;
}
// ijab_tot(a,j,i) = sum_b ijab_w (sequential in b) ; ija_w(a,j,i) is the same sum
__global__ void k_hb_ijab_tot(int nb_, const double* __restrict__ w, double* __restrict__ tot, double* __restrict__ ija_w) {
    const long long nb = nb_;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * nb * nb) return;
    double sum = 0.0;
    const double* row = w + t * nb;
    for (int b = 0; b < nb; ++b) sum = sum + row[b];
    tot[t] = sum;
    ija_w[t] = sum;
}
// ija_tot(j,i) = sum_a ija_w (sequential) ; ij_w(j,i) = flat sequential sum over (a,b) of ijab_w
__global__ void k_hb_ij(int nb_, const double* __restrict__ ijab_w, const double* __restrict__ ija_w,
                        double* __restrict__ ija_tot, double* __restrict__ ij_w) {
    const long long nb = nb_;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * nb) return;
    double sum = 0.0;
    for (int a = 0; a < nb; ++a) sum = sum + ija_w[t * nb + a];
    ija_tot[t] = sum;
    double flat = 0.0;
    const double* base = ijab_w + t * nb * nb;
    for (long long ab = 0; ab < nb * nb; ++ab) {
        const double v = base[ab];
        if (v != 0.0) flat = flat + v;  // the reference only adds allowed terms; adding 0.0 would be identical
    }
    ij_w[t] = flat;
}
__global__ void k_hb_i(int nb, const double* __restrict__ ij_w, double* __restrict__ i_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    double sum = 0.0;
    for (int j = 0; j < nb; ++j)
        if (j != i) sum = sum + ij_w[(long long)i * nb + j];
    i_w[i] = sum;
}
// alias tables per row of length nb: rows = nb^2 (ija) or nb^3 (ijab); scratch: 2 ints per element
__global__ void k_hb_alias(int nb, long long nrows, const double* __restrict__ w, const double* __restrict__ tot,
                           double* __restrict__ U, int* __restrict__ K, int* __restrict__ scratch) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    if (!(fabs(tot[r]) > 0.0)) return;
    generate_alias_tables(nb, w + r * nb, tot[r], U + r * nb, K + r * nb, scratch + 2 * r * nb, scratch + 2 * r * nb + nb);
}

// ------------------------------------------------------------------------------------------------
// Probe kernels (parity tests through the C ABI)
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void k_gen_excit_batch(Sys s, Params p, const uint64_t* __restrict__ states, const int64_t* __restrict__ pops,
                                  const uint32_t* __restrict__ attempt, long long n, const int* __restrict__ proc_map,
                                  int* __restrict__ iout, double* __restrict__ dout, int64_t* __restrict__ nspawn) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t f[W];
#pragma unroll
    for (int k = 0; k < W; ++k) f[k] = states[t * W + k];
    uint8_t occ[HB_MAXNEL], su[64];
    decode_det<W>(f, occ);
    if (s.kind != SYS_UEG) build_symunocc(s, occ, su);
    PhiloxStream rng;
    rng.begin(p.seed, p.cycle, RNG_SPAWN, det_hash64<W>(f), attempt[t]);
    Gen g;
    gen_excit<W>(rng, s, p, f, occ, su, g);
    const int64_t ns = attempt_to_spawn(rng, p, g.hmatel, g.pgen, pops[t]);
    int own = -1;
    if (g.allowed) {
        uint64_t child[W];
#pragma unroll
        for (int k = 0; k < W; ++k) child[k] = f[k];
        child[(g.from1 - 1) >> 6] &= ~(1ull << ((g.from1 - 1) & 63));
        child[(g.to1 - 1) >> 6] |= (1ull << ((g.to1 - 1) & 63));
        if (g.nexcit == 2) {
            child[(g.from2 - 1) >> 6] &= ~(1ull << ((g.from2 - 1) & 63));
            child[(g.to2 - 1) >> 6] |= (1ull << ((g.to2 - 1) & 63));
        }
        own = proc_map[owner_slot(child, s.nbasis, p.hash_seed, p.nprocs, p.nslots)];
    }
    int* io = iout + t * 8;
    io[0] = g.nexcit; io[1] = g.from1; io[2] = g.from2; io[3] = g.to1; io[4] = g.to2; io[5] = g.perm; io[6] = g.allowed;
    io[7] = own;
    dout[t * 2] = g.pgen; dout[t * 2 + 1] = g.hmatel;
    nspawn[t] = ns;
}

// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
struct hb200_engine {
    hb200_config cfg;
    int W = 1, E = 3;
    cudaStream_t stream = nullptr;
    Sys sys;
    Params par;
    bool have_sys = false, have_hb = false, have_ref = false, have_ppn = false, have_pp = false;
    // owned device buffers for system tables
    std::vector<void*> owned;
    int* d_proc_map = nullptr;
    // main list (double buffered)
    // buffers 0/1: current list and the merge output (swapped every cycle); buffer 2 (allocated on first use): staging
    // area of the asynchronous upload, rotated in by hb200_upload_psips_commit
    uint64_t* d_states[3] = {nullptr, nullptr, nullptr};
    int64_t* d_pops[3] = {nullptr, nullptr, nullptr};
    double* d_dat[3] = {nullptr, nullptr, nullptr};
    int cur = 0, alt = 1, stg = 2;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done = nullptr;
    long long stg_n = -1;
    long long nstates = 0;
    long long nparticles_enc = 0;  // sum |pop| (encoded) of the current list
    // spawn store
    int64_t* d_spawn[2] = {nullptr, nullptr};
    int sp_cur = 0;      // buffer holding the current stage's list
    long long sp_n = 0;  // number of elements in it (contiguous from 0) after comm
    bool sp_blocked = true;  // true: still partitioned in per-destination blocks (before comm)
    long long block_size = 0;
    unsigned long long* d_head = nullptr;
    std::vector<unsigned long long> h_head;
    int* d_err = nullptr;
    // scratch
    SpawnPartials* d_partials = nullptr;
    long long max_tiles = 0;
    CycleStats* d_stats = nullptr;
    unsigned* d_hist = nullptr;
    long long hist_cap = 0;
    int* d_ins_flag = nullptr;
    int* d_ins_idx = nullptr;
    long long* d_ins_pos = nullptr;
    double* d_ins_dat = nullptr;
    int* d_tile_keep = nullptr;
    int* d_tile_off = nullptr;
    int* d_scan_l1 = nullptr;
    int* d_scan_l1o = nullptr;
    int* d_total = nullptr;   // [4] small ints
    long long* d_part_ll = nullptr;
    long long* d_ll = nullptr;  // [4]
    bool ccmc_full_nc = false;                     // ccmc_in%full_nc
    int ccmc_hash_shift = 0, ccmc_move_freq = 5;   // spawn%hash_shift (+1 per cycle), spawn%move_freq
    // CCMC scratch
    long long* d_cum = nullptr;        // [walker_length] inclusive prefix sums of |pop| (reference skipped)
    long long* d_cum_blk = nullptr;
    CcmcPartials* d_cc_part = nullptr;
    PsPartials* d_ps_part = nullptr;   // pattempt_update: per-block sums of one launch
    double* d_ps_acc = nullptr;        // [4] running totals since the last hb200_get_ps_stats(reset)
    size_t ps_part_cap = 0;
    CcmcPartials* d_cc_tot = nullptr;
    // NCCL
    ncclComm_t comm = nullptr;
    long long* d_counts = nullptr;  // [nprocs*nprocs]
    // timing / counters
    cudaEvent_t ev[6];
    cudaEvent_t evk[2];           // brackets the k_spawn_death launch alone (roofline timing)
    float spawn_kernel_ms = 0.f;  // accumulated over the cycles of the last hb200_iterate
    double ms[8] = {0};
    long long launches = 0, spawn_launches = 0;
};


// All host<->device copies are issued on the engine's own (non-blocking) stream and then synchronised: a plain
// cudaMemcpy runs on the legacy stream, which is NOT ordered with kernels on a cudaStreamNonBlocking stream, and a
// pageable H2D copy may return before its DMA has landed.
static cudaError_t copy_sync(hb200_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
template <class T>
static int dalloc(hb200_engine* e, T** p, size_t n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    *p = (T*)q;
    e->owned.push_back(q);
    return 0;
}
template <class T>
static int dupload(hb200_engine* e, const T** dst, const T* src, size_t n) {
    T* q = nullptr;
    if (dalloc(e, &q, n)) return 1;
    if (n) CK(copy_sync(e, q, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = q;
    return 0;
}

static cudaError_t copy_sync(hb200_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    cudaError_t r = cudaMemcpyAsync(dst, src, bytes, kind, e->stream);
    if (r != cudaSuccess) return r;
    return cudaStreamSynchronize(e->stream);
}

static bool uses_heat_bath_tables(const hb200_engine* e) {
    const int eg = e->cfg.excit_gen;
    return eg == HB200_EXCIT_GEN_HEAT_BATH || eg == HB200_EXCIT_GEN_HEAT_BATH_UNIFORM || eg == HB200_EXCIT_GEN_HEAT_BATH_SINGLE ||
           eg == HB200_EXCIT_GEN_POWER_PITZER_OCC_IJ || eg == HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ;
}
static size_t spawn_smem_bytes(const hb200_engine* e) {
    const int eg = e->cfg.excit_gen;
    const int nsu = (eg == HB200_EXCIT_GEN_POWER_PITZER_ORDERN) ? e->sys.nel :
                    (e->sys.kind == SYS_READ_IN && eg != HB200_EXCIT_GEN_POWER_PITZER && eg != HB200_EXCIT_GEN_NO_RENORM && eg != HB200_EXCIT_GEN_NO_RENORM_SPIN &&
                     eg != HB200_EXCIT_GEN_HEAT_BATH &&
                     eg != HB200_EXCIT_GEN_HEAT_BATH_SINGLE)
                        ? 2 * e->sys.nsym_tot : 0;
    const bool hb = eg == HB200_EXCIT_GEN_HEAT_BATH;
    return SpawnSmem(e->W, e->sys.nel, nsu, e->sys.nbasis, hb, uses_heat_bath_tables(e), !hb && e->par.ps_part != nullptr,
                     e->par.qn != 0).total;
}

extern "C" {

const char* hb200_last_error(void) { return g_err.c_str(); }

hb200_engine* hb200_create(const hb200_config* cfg) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_err = "hb200_create: no CUDA device (the engine has no CPU fallback)";
        return nullptr;
    }
    if (cfg->nel > HB_MAXNEL || cfg->nbasis > 64 * HB_MAXW || cfg->nbasis > 255) {
        g_err = "hb200_create: nel/nbasis beyond compiled limits (HB_MAXNEL, HB_MAXW)";
        return nullptr;
    }
    hb200_engine* e = new hb200_engine();
    e->cfg = *cfg;
    e->W = (cfg->nbasis + 63) / 64;
    e->E = e->W + 2;
    auto fail = [&](const char* what) -> hb200_engine* {
        if (g_err.empty()) g_err = what;
        hb200_destroy(e);
        return nullptr;
    };
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail("cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    for (int i = 0; i < 6; ++i) cudaEventCreate(&e->ev[i]);
    for (int i = 0; i < 2; ++i) cudaEventCreate(&e->evk[i]);
    memset(&e->sys, 0, sizeof(Sys));
    memset(&e->par, 0, sizeof(Params));
    Params& p = e->par;
    p.excit_gen = cfg->excit_gen;
    p.pattempt_single = cfg->pattempt_single;
    p.pattempt_double = cfg->pattempt_double;
    p.real_amplitudes = cfg->real_amplitudes;
    p.real_factor = cfg->real_amplitudes ? (1ll << 31) : 1;  // particle_t_utils.f90 (POP_SIZE=64)
    {
        double c = cfg->real_amplitudes ? cfg->spawn_cutoff : 0.0;
        p.spawn_cutoff = (int64_t)ceil(c * (double)p.real_factor);  // src/spawn_data.F90:215
    }
    p.initiator = cfg->initiator_approx;
    p.initiator_pop = cfg->initiator_pop;
    p.trunc_level = cfg->trunc_level;
    p.seed = cfg->rng_seed;
    p.hash_seed = (uint32_t)cfg->hash_seed;
    p.nprocs = std::max(1, cfg->nprocs);
    p.iproc = cfg->iproc;
    p.nslots = std::max(1, cfg->nslots);
    const long long cap = cfg->walker_length;
    long long scap = cfg->spawned_walker_length;
    if (scap % p.nprocs != 0) scap = ((scap + p.nprocs - 1) / p.nprocs) * p.nprocs;  // src/qmc.F90:1461-1468
    e->cfg.spawned_walker_length = scap;
    e->block_size = scap / p.nprocs;
    const int W = e->W, E = e->E;
    for (int b = 0; b < 2; ++b) {
        if (dalloc(e, &e->d_states[b], (size_t)cap * W)) return fail("alloc states");
        if (dalloc(e, &e->d_pops[b], (size_t)cap)) return fail("alloc pops");
        if (dalloc(e, &e->d_dat[b], (size_t)cap)) return fail("alloc dat");
        if (dalloc(e, &e->d_spawn[b], (size_t)scap * E)) return fail("alloc spawn");
    }
    e->max_tiles = (cap + TILE - 1) / TILE + 1;
    if (dalloc(e, &e->d_head, (size_t)p.nprocs)) return fail("alloc");
    e->h_head.assign(p.nprocs, 0);
    if (dalloc(e, &e->d_err, 4)) return fail("alloc");
    if (dalloc(e, &e->d_partials, (size_t)e->max_tiles)) return fail("alloc");
    if (dalloc(e, &e->d_stats, 1)) return fail("alloc");
    e->hist_cap = 256ll * ((scap + 2047) / 2048 + 1);
    if (dalloc(e, &e->d_hist, (size_t)e->hist_cap)) return fail("alloc");
    if (dalloc(e, &e->d_ins_flag, (size_t)scap)) return fail("alloc");
    if (dalloc(e, &e->d_ins_idx, (size_t)scap)) return fail("alloc");
    if (dalloc(e, &e->d_ins_pos, (size_t)scap)) return fail("alloc");
    if (dalloc(e, &e->d_ins_dat, (size_t)scap)) return fail("alloc");
    if (dalloc(e, &e->d_tile_keep, (size_t)e->max_tiles)) return fail("alloc");
    if (dalloc(e, &e->d_tile_off, (size_t)e->max_tiles)) return fail("alloc");
    const long long l1 = std::max(scap, e->max_tiles) / SCAN_BLOCK + 2;
    if (dalloc(e, &e->d_scan_l1, (size_t)l1)) return fail("alloc");
    if (dalloc(e, &e->d_scan_l1o, (size_t)l1)) return fail("alloc");
    if (dalloc(e, &e->d_total, 4)) return fail("alloc");
    if (dalloc(e, &e->d_part_ll, (size_t)e->max_tiles)) return fail("alloc");
    if (dalloc(e, &e->d_ll, 4)) return fail("alloc");
    if (dalloc(e, &e->d_counts, (size_t)p.nprocs * p.nprocs)) return fail("alloc");
    {
        std::vector<int> map((size_t)p.nprocs * p.nslots);
        for (size_t i = 0; i < map.size(); ++i) map[i] = (int)(i % p.nprocs);  // src/load_balancing.F90:170
        if (dalloc(e, &e->d_proc_map, map.size())) return fail("alloc");
        if (copy_sync(e, e->d_proc_map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
            return fail("memcpy proc_map");
    }
    if (cudaMemsetAsync(e->d_err, 0, 4 * sizeof(int), e->stream) != cudaSuccess) return fail("memset");
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return fail("sync");
    return e;
}

void hb200_destroy(hb200_engine* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (void* q : e->owned) cudaFree(q);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->copy_done) cudaEventDestroy(e->copy_done);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int hb200_set_system_read_in(hb200_engine* e, const hb200_system_read_in* in) {
    CK(cudaSetDevice(e->cfg.device));
    if (in->nbasis != e->cfg.nbasis || in->nel != e->cfg.nel) FAIL("set_system: nbasis/nel differ from hb200_create");
    Sys& s = e->sys;
    s.kind = SYS_READ_IN;
    s.nbasis = in->nbasis; s.nel = in->nel; s.W = e->W;
    s.nsym_tot = in->nsym_tot; s.sym0 = in->sym0; s.sym_max = in->sym_max; s.pg_mask = in->pg_mask;
    s.Lz_mask = in->Lz_mask; s.Lz_offset = in->Lz_offset; s.gamma_sym = in->gamma_sym; s.uhf = in->uhf;
    s.nvirt = in->nvirt; s.nvirt_alpha = in->nvirt_alpha; s.nvirt_beta = in->nvirt_beta; s.max_nbss = in->max_nbss;
    s.Ecore = in->Ecore;
    if (2 * s.nsym_tot > 64) FAIL("set_system: too many irreps for the symunocc scratch");
    const int nb = s.nbasis;
    std::vector<uint8_t> sym(nb + 1, 0);
    std::vector<int8_t> ms(nb + 1, 0);
    std::vector<uint16_t> sp(nb + 1, 0);
    for (int i = 1; i <= nb; ++i) { sym[i] = (uint8_t)in->bf_sym[i]; ms[i] = (int8_t)in->bf_ms[i]; sp[i] = (uint16_t)in->bf_spatial[i]; }
    if (dupload(e, &s.bf_sym, sym.data(), sym.size())) return 1;
    if (dupload(e, &s.bf_ms, ms.data(), ms.size())) return 1;
    if (dupload(e, &s.bf_spatial, sp.data(), sp.size())) return 1;
    if (dupload(e, &s.nbss, in->nbasis_sym_spin, (size_t)2 * s.nsym_tot)) return 1;
    if (dupload(e, &s.ssbf, in->sym_spin_basis_fns, (size_t)s.max_nbss * 2 * s.nsym_tot)) return 1;
    {
        std::vector<uint64_t> mask((size_t)2 * s.nsym_tot * e->W, 0);
        for (int i = 1; i <= nb; ++i) {
            const int c = (in->bf_ms[i] > 0 ? 1 : 0) + 2 * in->bf_sym[i];
            mask[(size_t)c * e->W + ((i - 1) >> 6)] |= 1ull << ((i - 1) & 63);
        }
        if (dupload(e, &s.su_mask, mask.data(), mask.size())) return 1;
    }
    if (dupload(e, &s.h1, in->one_body, (size_t)nb * nb)) return 1;
    for (int c = 0; c < (s.uhf ? 4 : 1); ++c)
        if (dupload(e, &s.v2[c], in->two_body[c], (size_t)in->nintgrls)) return 1;
    double *J = nullptr, *K = nullptr;
    if (dalloc(e, &J, (size_t)nb * nb) || dalloc(e, &K, (size_t)nb * nb)) return 1;
    k_build_JK<<<(nb * nb + 255) / 256, 256, 0, e->stream>>>(s, J, K);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    s.Jd = J; s.Kd = K;
    {
        const int NT = s.uhf ? nb : nb / 2;
        const long long n3 = (long long)NT * NT * NT;
        D2* CX = nullptr;
        if (dalloc(e, &CX, (size_t)n3)) return 1;
        k_build_sc1_tables<<<(unsigned)((n3 + 255) / 256), 256, 0, e->stream>>>(s, NT, CX);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        s.sc1CX = CX; s.NT = NT;
    }
    e->have_sys = true;
    return 0;
}

int hb200_set_system_ueg(hb200_engine* e, const hb200_system_ueg* in) {
    CK(cudaSetDevice(e->cfg.device));
    if (in->nbasis != e->cfg.nbasis || in->nel != e->cfg.nel) FAIL("set_system_ueg: nbasis/nel differ from hb200_create");
    if (uses_heat_bath_tables(e)) FAIL("set_system_ueg: heat_bath is a molecular generator");
    Sys& s = e->sys;
    s.kind = SYS_UEG;
    s.nbasis = in->nbasis; s.nel = in->nel; s.W = e->W;
    s.nsym_tot = 1; s.sym0 = 0; s.sym_max = 0; s.pg_mask = 0; s.Lz_mask = 0; s.Lz_offset = 0; s.gamma_sym = 0; s.uhf = 0;
    const int nb = s.nbasis;
    std::vector<uint8_t> sym(nb + 1, 0);
    std::vector<int8_t> ms(nb + 1, 0);
    std::vector<uint16_t> sp(nb + 1, 0);
    std::vector<K4> kv(nb + 1);
    for (int i = 1; i <= nb; ++i) {
        ms[i] = (int8_t)((i & 1) ? 1 : -1);
        sp[i] = (uint16_t)((i + 1) / 2);
        kv[i].x = in->kvec[3 * i]; kv[i].y = in->kvec[3 * i + 1]; kv[i].z = in->kvec[3 * i + 2]; kv[i].w = 0;
    }
    kv[0].x = kv[0].y = kv[0].z = kv[0].w = 0;
    if (dupload(e, &s.bf_sym, sym.data(), sym.size())) return 1;
    if (dupload(e, &s.bf_ms, ms.data(), ms.size())) return 1;
    if (dupload(e, &s.bf_spatial, sp.data(), sp.size())) return 1;
    if (dupload(e, &s.ueg_k, kv.data(), kv.size())) return 1;
    if (dupload(e, &s.sp_eigv, in->sp_eigv, (size_t)nb + 1)) return 1;
    if (dupload(e, &s.ueg_lookup, in->lookup, (size_t)in->n_lookup + 1)) return 1;
    const size_t tD = 2 * (size_t)in->tern_kmax + 1;
    if (dupload(e, &s.ueg_tern, in->ternary_conserve, (size_t)(e->W + 1) * tD * tD * tD)) return 1;
    s.ueg_piL = 3.1415926535897931 * in->box_length;   // pi*cell_param, as evaluated first in coulomb_int_ueg_3d
    s.ueg_kmax = in->kmax; s.ueg_offset = in->offset;
    for (int d = 0; d < 3; ++d) s.ueg_oi[d] = in->offset_inds[d];
    s.ueg_tK = in->tern_kmax; s.ueg_tD = (int)tD;
    e->have_sys = true;
    return 0;
}

int hb200_build_heat_bath(hb200_engine* e) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("build_heat_bath: system not set");
    Sys& s = e->sys;
    const long long nb = s.nbasis;
    const long long n2 = nb * nb, n3 = n2 * nb, n4 = n3 * nb;
    double *i_w, *ij_w, *ija_w, *ija_U, *ija_tot, *ijab_w, *ijab_U, *ijab_tot;
    int *ija_K, *ijab_K, *scratch;
    if (dalloc(e, &i_w, nb) || dalloc(e, &ij_w, n2) || dalloc(e, &ija_w, n3) || dalloc(e, &ija_U, n3) ||
        dalloc(e, &ija_K, n3) || dalloc(e, &ija_tot, n2) || dalloc(e, &ijab_w, n4) || dalloc(e, &ijab_U, n4) ||
        dalloc(e, &ijab_K, n4) || dalloc(e, &ijab_tot, n3))
        return 1;
    void* sc = nullptr;
    CK(cudaMalloc(&sc, (size_t)2 * n4 * sizeof(int)));
    scratch = (int*)sc;
    cudaStream_t st = e->stream;
    CK(cudaMemsetAsync(ija_U, 0, n3 * sizeof(double), st));
    CK(cudaMemsetAsync(ija_K, 0, n3 * sizeof(int), st));
    CK(cudaMemsetAsync(ijab_U, 0, n4 * sizeof(double), st));
    CK(cudaMemsetAsync(ijab_K, 0, n4 * sizeof(int), st));
    k_hb_ijab_w<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(s, ijab_w);
    k_hb_ijab_tot<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>((int)nb, ijab_w, ijab_tot, ija_w);
    k_hb_ij<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>((int)nb, ijab_w, ija_w, ija_tot, ij_w);
    k_hb_i<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>((int)nb, ij_w, i_w);
    k_hb_alias<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>((int)nb, n2, ija_w, ija_tot, ija_U, ija_K, scratch);
    k_hb_alias<<<(unsigned)((n3 + 127) / 128), 128, 0, st>>>((int)nb, n3, ijab_w, ijab_tot, ijab_U, ijab_K, scratch);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(sc));
    s.hb_i_w = i_w; s.hb_ij_w = ij_w; s.hb_ija_w = ija_w; s.hb_ija_U = ija_U; s.hb_ija_K = ija_K; s.hb_ija_tot = ija_tot;
    s.hb_ijab_w = ijab_w; s.hb_ijab_U = ijab_U; s.hb_ijab_K = ijab_K; s.hb_ijab_tot = ijab_tot;
    e->have_hb = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// power_pitzer_orderN tables (init_excit_mol_power_pitzer_orderN, src/excit_gen_power_pitzer_mol.F90:215-572), built on
// the device: every weight is one thread's sequential sum in the reference's order (bit-identical to the CPU tables),
// then one thread per column applies check_min_weight_ratio (:140-213) and generate_alias_tables.
// ------------------------------------------------------------------------------------------------
struct PpnBuild { double* w[6]; double* U[6]; int* K[6]; double* tot[6]; const int* occ; double min_weight; };

// single_excitation_weight_mol (src/hamiltonian_molecular.f90:444-523); occ0 ascending
__device__ double ppn_single_excitation_weight(const Sys& s, const int* occ0, int i, int a) {
    const int nel = s.nel, nb = s.nbasis;
    int n_jb = 0;
    double weight = 0.0;
    for (int j = 0; j < nel; ++j) {
        const int oj = occ0[j];
        const double t1 = two_body_real(s, i, oj, oj, a) - two_body_real(s, i, oj, a, oj);
        int op = 0;
        for (int pos = 1; pos <= nb; ++pos) {            // the virtual orbitals of the reference, ascending
            if (op < nel && occ0[op] == pos) { op++; continue; }
            n_jb++;
            weight = weight + fabs(t1 + two_body_real(s, i, pos, a, pos) - two_body_real(s, i, pos, pos, a));
        }
    }
    return weight / (double)n_jb;
}
// init_double_weights_ab (src/excit_gen_utils.f90:68-140): accumulates onto weight
__device__ void ppn_double_weights_ab(const Sys& s, int i, int j, double& weight) {
    const int it = min(i, j), jt = max(i, j);
    const int ij_sym = sym_conj(s, cross_product(s, s.bf_sym[it], s.bf_sym[jt]));
    for (int a = 1; a <= s.nbasis; ++a) {
        if (a == it || a == jt) continue;
        const int isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
        for (int b = 1; b <= s.nbasis; ++b) {
            const bool spin_ok = (ms_of(it) == ms_of(a) && ms_of(jt) == ms_of(b)) || (ms_of(it) == ms_of(b) && ms_of(jt) == ms_of(a));
            if (spin_ok && s.bf_sym[b] == isymb && a != b && b != it && b != jt)
                weight = weight + fabs(slater_condon2_excit(s, it, jt, min(a, b), max(a, b), false));
        }
    }
}
__global__ void k_ppn_weights(Sys s, PpnBuild t) {
    const int nel = s.nel, nb = s.nbasis, mv = s.max_nbss, nsym = s.nsym_tot, nall = nb / 2;
    const long long nA = nel, nB = nel, nC = (long long)nb * mv, nD = (long long)nb * nel, nE = (long long)nb * nall,
                    nF = (long long)nb * nsym * mv;
    long long job = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int* occ = t.occ;
    const double depsilon = 1.e-12;
    if (job < nA) {                       // i in a single excitation
        const int oi = occ[job];
        const int isyma = cross_product(s, s.bf_sym[oi], s.gamma_sym);
        double w = 0.0;
        for (int a = 1; a <= nb; ++a)
            if (a != oi && s.bf_sym[a] == isyma && ms_of(a) == ms_of(oi)) w = w + ppn_single_excitation_weight(s, occ, oi, a);
        if (w < depsilon) w = 10.0 * depsilon;
        t.w[PPN_IS][job] = w;
        return;
    }
    job -= nA;
    if (job < nB) {                       // i in a double excitation: one running sum over all j, a, b
        double w = 0.0;
        for (int j = 0; j < nel; ++j)
            if (j != job) ppn_double_weights_ab(s, occ[job], occ[j], w);
        if (w < depsilon) w = 10.0 * depsilon;
        t.w[PPN_ID][job] = w;
        return;
    }
    job -= nB;
    if (job < nC) {                       // a given i, single excitation
        const int i = (int)(job / mv) + 1, a = (int)(job % mv) + 1;
        const int imsa = ims_of(i), isyma = cross_product(s, s.bf_sym[i], s.gamma_sym);
        double w = 0.0;
        if (a <= nbss(s, imsa, isyma)) {
            const int oa = ssbf(s, a, imsa, isyma);
            if (oa != i) {
                w = ppn_single_excitation_weight(s, occ, i, oa);
                if (w < depsilon) w = 10.0 * depsilon;
            }
        }
        t.w[PPN_IAS][(size_t)mv * i + a - 1] = w;
        return;
    }
    job -= nC;
    if (job < nD) {                       // j given i, double excitation
        const int i = (int)(job / nel) + 1, j = (int)(job % nel);
        double w = 0.0;
        if (occ[j] != i) ppn_double_weights_ab(s, i, occ[j], w);
        if (w < depsilon) w = 10.0 * depsilon;
        t.w[PPN_IJD][(size_t)nel * i + j] = w;
        return;
    }
    job -= nD;
    if (job < nE) {                       // a given i: sqrt|<ia|ai>| over the orbitals of the spin of i
        const int i = (int)(job / nall) + 1, k = (int)(job % nall) + 1;
        const int a = (ms_of(i) < 0) ? 2 * k : 2 * k - 1;
        t.w[PPN_IAD][(size_t)nall * i + k - 1] = (a != i) ? pp_weight(s, false, i, a) : 0.0;
        return;
    }
    job -= nE;
    if (job < nF) {                       // b given j, per symmetry class
        const int i = (int)(job / ((long long)nsym * mv)) + 1;
        const int bsym = (int)((job / mv) % nsym), k = (int)(job % mv) + 1;
        const int ims = ims_of(i);
        double w = 0.0;
        if (k <= nbss(s, ims, bsym)) {
            const int a = ssbf(s, k, ims, bsym);
            if (a != i) w = pp_weight(s, false, i, a);
        }
        t.w[PPN_JBD][(size_t)mv * (bsym + (size_t)nsym * i) + k - 1] = w;
    }
}
__device__ void ppn_check_min_weight_ratio(double* weights, double& weights_tot, int n, double min_ratio) {
    double min_weight_tmp = 0.0;
    int nonzero = 0;
    if (!(weights_tot > 0.0 && min_ratio > 0.0)) return;
    for (int i = 0; i < n; ++i) if (weights[i] > 0.0) nonzero++;
    double min_weight = (min_ratio / (float)nonzero) * weights_tot;
    while (fabs(min_weight_tmp - min_weight) > 1.e-12) {
        min_weight_tmp = min_weight;
        double keep = 0.0;
        int cnt = 0;
        for (int k = 0; k < n; ++k) {
            if (weights[k] > 0.0 && weights[k] < min_weight) cnt++;
            else keep = keep + weights[k];
        }
        if (cnt == nonzero) break;
        min_weight = (min_ratio / (float)nonzero) * (keep / (1 - (min_ratio * (float)cnt / (float)nonzero)));
    }
    weights_tot = 0.0;
    for (int j = 0; j < n; ++j) {
        if (weights[j] > 0.0 && weights[j] < min_weight) weights[j] = min_weight;
        weights_tot = weights_tot + weights[j];
    }
}
__global__ void k_ppn_alias(Sys s, PpnBuild t) {
    const int nel = s.nel, nb = s.nbasis, mv = s.max_nbss, nsym = s.nsym_tot, nall = nb / 2;
    long long job = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int which, n;
    size_t col, stride;
    bool min_ratio = true;
    if (job == 0) { which = PPN_IS; col = 0; stride = nel; n = nel; }
    else if (job == 1) { which = PPN_ID; col = 0; stride = nel; n = nel; }
    else if ((job -= 2) < nb) {
        const int i = (int)job + 1;
        which = PPN_IAS; col = i; stride = mv; n = nbss(s, ims_of(i), cross_product(s, s.bf_sym[i], s.gamma_sym));
    } else if ((job -= nb) < nb) { which = PPN_IJD; col = (size_t)job + 1; stride = nel; n = nel; }
    else if ((job -= nb) < nb) { which = PPN_IAD; col = (size_t)job + 1; stride = nall; n = nall; min_ratio = false; }
    else if ((job -= nb) < (long long)nb * nsym) {
        const int i = (int)(job / nsym) + 1, bsym = (int)(job % nsym);
        which = PPN_JBD; col = (size_t)bsym + (size_t)nsym * i; stride = mv; n = nbss(s, ims_of(i), bsym); min_ratio = false;
    } else return;
    double* w = t.w[which] + stride * col;
    double tot = 0.0;
    for (int k = 0; k < n; ++k) tot = tot + w[k];
    if (min_ratio) ppn_check_min_weight_ratio(w, tot, n, t.min_weight);
    t.tot[which][col] = tot;
    if (n > 0) {
        int under[HB_MAXLIST], over[HB_MAXLIST];
        generate_alias_tables(n, w, tot, t.U[which] + stride * col, t.K[which] + stride * col, under, over);
    }
}

int hb200_build_power_pitzer_orderN(hb200_engine* e, double min_weight) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys || e->sys.kind != SYS_READ_IN) FAIL("build_power_pitzer_orderN: needs a read_in system");
    if (!e->have_ref) FAIL("build_power_pitzer_orderN: reference not set (call hb200_set_reference first)");
    Sys& s = e->sys;
    const int nel = s.nel, nb = s.nbasis, mv = s.max_nbss, nsym = s.nsym_tot, nall = nb / 2;
    if (nall > HB_MAXLIST || mv > HB_MAXLIST || nel > HB_MAXLIST) FAIL("build_power_pitzer_orderN: basis too large");
    const size_t len[6] = {(size_t)nel, (size_t)mv * (nb + 1), (size_t)nel, (size_t)nel * (nb + 1), (size_t)nall * (nb + 1),
                           (size_t)mv * nsym * (nb + 1)};
    const size_t ncol[6] = {1, (size_t)nb + 1, 1, (size_t)nb + 1, (size_t)nb + 1, (size_t)nsym * (nb + 1)};
    PpnBuild t;
    cudaStream_t st = e->stream;
    for (int k = 0; k < 6; ++k) {
        if (dalloc(e, &t.w[k], len[k]) || dalloc(e, &t.U[k], len[k]) || dalloc(e, &t.K[k], len[k]) || dalloc(e, &t.tot[k], ncol[k]))
            return 1;
        CK(cudaMemsetAsync(t.w[k], 0, len[k] * sizeof(double), st));
        CK(cudaMemsetAsync(t.U[k], 0, len[k] * sizeof(double), st));
        CK(cudaMemsetAsync(t.K[k], 0, len[k] * sizeof(int), st));
        CK(cudaMemsetAsync(t.tot[k], 0, ncol[k] * sizeof(double), st));
    }
    // the reference's occupied orbitals, ascending (pp%occ_list)
    std::vector<int> occ0;
    for (int o = 1; o <= nb; ++o)
        if ((e->par.f0[(o - 1) >> 6] >> ((o - 1) & 63)) & 1ull) occ0.push_back(o);
    if ((int)occ0.size() != nel) FAIL("build_power_pitzer_orderN: reference does not have nel electrons");
    int* d_occ = nullptr;
    if (dalloc(e, &d_occ, (size_t)nel)) return 1;
    CK(cudaMemcpyAsync(d_occ, occ0.data(), sizeof(int) * nel, cudaMemcpyHostToDevice, st));
    t.occ = d_occ;
    t.min_weight = min_weight;
    const long long njobs = 2LL * nel + (long long)nb * mv + (long long)nb * nel + (long long)nb * nall + (long long)nb * nsym * mv;
    k_ppn_weights<<<(unsigned)((njobs + 63) / 64), 64, 0, st>>>(s, t);
    const long long ncols = 2 + 3LL * nb + (long long)nb * nsym;
    k_ppn_alias<<<(unsigned)((ncols + 63) / 64), 64, 0, st>>>(s, t);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    for (int k = 0; k < 6; ++k) { s.ppn[k].w = t.w[k]; s.ppn[k].U = t.U[k]; s.ppn[k].K = t.K[k]; s.ppn[k].tot = t.tot[k]; }
    s.ppn_occ = d_occ;
    e->have_ppn = true;
    return 0;
}

// power_pitzer tables (init_excit_mol_power_pitzer_occ_ref, src/excit_gen_power_pitzer_mol.F90:19-138): weights over the
// reference's virtual orbitals (pp_ia_d) and over each symmetry class (pp_jb_d) for every reference electron
struct PpBuild { double* w[2]; double* U[2]; int* K[2]; double* tot[2]; const int* occ; const int* virt[2]; int nvirt[2];
                 int sia; double min_weight; };
__global__ void k_pp_weights(Sys s, PpBuild t) {
    const int nel = s.nel, mv = s.max_nbss, nsym = s.nsym_tot;
    long long job = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nA = (long long)nel * t.sia, nB = (long long)nel * nsym * mv;
    if (job < nA) {
        const int i = (int)(job / t.sia), k = (int)(job % t.sia);
        const int oj = t.occ[i], sp = (ms_of(oj) < 0) ? 0 : 1;
        t.w[0][job] = (k < t.nvirt[sp]) ? pp_weight(s, false, oj, t.virt[sp][k]) : 0.0;
        return;
    }
    job -= nA;
    if (job < nB) {
        const int i = (int)(job / ((long long)nsym * mv)), bsym = (int)((job / mv) % nsym), k = (int)(job % mv) + 1;
        const int oj = t.occ[i], ims = ims_of(oj);
        t.w[1][job] = (k <= nbss(s, ims, bsym)) ? pp_weight(s, false, oj, ssbf(s, k, ims, bsym)) : 0.0;
    }
}
__global__ void k_pp_alias(Sys s, PpBuild t) {
    const int nel = s.nel, mv = s.max_nbss, nsym = s.nsym_tot;
    long long job = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int which, n;
    size_t col, stride;
    if (job < nel) {
        which = 0; col = (size_t)job; stride = t.sia; n = t.nvirt[(ms_of(t.occ[job]) < 0) ? 0 : 1];
    } else if ((job -= nel) < (long long)nel * nsym) {
        const int i = (int)(job / nsym), bsym = (int)(job % nsym);
        which = 1; col = (size_t)job; stride = mv; n = nbss(s, ims_of(t.occ[i]), bsym);
    } else return;
    if (n <= 0) return;
    double* w = t.w[which] + stride * col;
    double tot = 0.0;
    for (int k = 0; k < n; ++k) tot = tot + w[k];
    ppn_check_min_weight_ratio(w, tot, n, t.min_weight);
    t.tot[which][col] = tot;
    int under[HB_MAXLIST], over[HB_MAXLIST];
    generate_alias_tables(n, w, tot, t.U[which] + stride * col, t.K[which] + stride * col, under, over);
}
// init_excit_ueg_power_pitzer (src/excit_gen_ueg.f90:362-408): one thread per orbital i builds its column
__global__ void k_ueg_pp_build(Sys s, double* w, double* U, int* K, double* tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > s.nbasis) return;
    const int maxv = s.nbasis / 2;
    double* wc = w + (size_t)maxv * i;
    double t = 0.0;
    for (int j = 1; j <= maxv; ++j) {
        const int a = j * 2 - (i & 1);
        const double weight = (a != i) ? fabs(ueg_coulomb(s, i, a)) : 0.0;
        wc[j - 1] = weight;
        t = t + weight;
    }
    tot[i] = t;
    int under[HB_MAXLIST], over[HB_MAXLIST];
    generate_alias_tables(maxv, wc, t, U + (size_t)maxv * i, K + (size_t)maxv * i, under, over);
}
int hb200_build_power_pitzer(hb200_engine* e, double min_weight) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("build_power_pitzer: system not set");
    if (e->sys.kind == SYS_UEG) {
        Sys& s = e->sys;
        const int nb = s.nbasis, maxv = nb / 2;
        if (maxv > HB_MAXLIST) FAIL("build_power_pitzer: basis too large");
        double *w, *U, *tot;
        int* K;
        const size_t len = (size_t)maxv * (nb + 1);
        if (dalloc(e, &w, len) || dalloc(e, &U, len) || dalloc(e, &K, len) || dalloc(e, &tot, (size_t)nb + 1)) return 1;
        CK(cudaMemsetAsync(w, 0, len * sizeof(double), e->stream));
        CK(cudaMemsetAsync(U, 0, len * sizeof(double), e->stream));
        CK(cudaMemsetAsync(K, 0, len * sizeof(int), e->stream));
        CK(cudaMemsetAsync(tot, 0, ((size_t)nb + 1) * sizeof(double), e->stream));
        k_ueg_pp_build<<<(nb + 63) / 64, 64, 0, e->stream>>>(s, w, U, K, tot);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        s.pp_ia.w = w; s.pp_ia.U = U; s.pp_ia.K = K; s.pp_ia.tot = tot;
        e->have_pp = true;
        return 0;
    }
    if (!e->have_ref) FAIL("build_power_pitzer: reference not set (call hb200_set_reference first)");
    Sys& s = e->sys;
    const int nel = s.nel, nb = s.nbasis, mv = s.max_nbss, nsym = s.nsym_tot;
    std::vector<int> occ0, virt[2];
    for (int o = 1; o <= nb; ++o) {
        if ((e->par.f0[(o - 1) >> 6] >> ((o - 1) & 63)) & 1ull) occ0.push_back(o);
        else virt[(o & 1) ? 1 : 0].push_back(o);          // odd orbitals are alpha (ms = +1)
    }
    if ((int)occ0.size() != nel) FAIL("build_power_pitzer: reference does not have nel electrons");
    const int sia = std::max<int>(1, (int)std::max(virt[0].size(), virt[1].size()));
    if (sia > HB_MAXLIST || mv > HB_MAXLIST) FAIL("build_power_pitzer: basis too large");
    PpBuild t;
    cudaStream_t st = e->stream;
    const size_t len[2] = {(size_t)nel * sia, (size_t)nel * nsym * mv}, ncol[2] = {(size_t)nel, (size_t)nel * nsym};
    for (int k = 0; k < 2; ++k) {
        if (dalloc(e, &t.w[k], len[k]) || dalloc(e, &t.U[k], len[k]) || dalloc(e, &t.K[k], len[k]) || dalloc(e, &t.tot[k], ncol[k]))
            return 1;
        CK(cudaMemsetAsync(t.U[k], 0, len[k] * sizeof(double), st));
        CK(cudaMemsetAsync(t.K[k], 0, len[k] * sizeof(int), st));
        CK(cudaMemsetAsync(t.tot[k], 0, ncol[k] * sizeof(double), st));
    }
    int *d_occ = nullptr, *d_virt[2] = {nullptr, nullptr};
    if (dalloc(e, &d_occ, (size_t)nel) || dalloc(e, &d_virt[0], virt[0].size()) || dalloc(e, &d_virt[1], virt[1].size())) return 1;
    CK(cudaMemcpyAsync(d_occ, occ0.data(), sizeof(int) * nel, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 2; ++k)
        if (!virt[k].empty()) CK(cudaMemcpyAsync(d_virt[k], virt[k].data(), sizeof(int) * virt[k].size(), cudaMemcpyHostToDevice, st));
    t.occ = d_occ; t.virt[0] = d_virt[0]; t.virt[1] = d_virt[1];
    t.nvirt[0] = (int)virt[0].size(); t.nvirt[1] = (int)virt[1].size();
    t.sia = sia; t.min_weight = min_weight;
    const long long njobs = (long long)(len[0] + len[1]);
    k_pp_weights<<<(unsigned)((njobs + 127) / 128), 128, 0, st>>>(s, t);
    const long long ncols = (long long)(ncol[0] + ncol[1]);
    k_pp_alias<<<(unsigned)((ncols + 63) / 64), 64, 0, st>>>(s, t);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    s.pp_ia.w = t.w[0]; s.pp_ia.U = t.U[0]; s.pp_ia.K = t.K[0]; s.pp_ia.tot = t.tot[0];
    s.pp_jb.w = t.w[1]; s.pp_jb.U = t.U[1]; s.pp_jb.K = t.K[1]; s.pp_jb.tot = t.tot[1];
    s.pp_virt[0] = d_virt[0]; s.pp_virt[1] = d_virt[1]; s.pp_nvirt[0] = t.nvirt[0]; s.pp_nvirt[1] = t.nvirt[1]; s.pp_sia = sia;
    s.ppn_occ = d_occ;
    e->have_pp = true;
    return 0;
}

int hb200_download_heat_bath(hb200_engine* e, int which, void* out, int64_t n) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_hb) FAIL("heat-bath tables not built");
    const Sys& s = e->sys;
    const void* src[10] = {s.hb_i_w, s.hb_ij_w, s.hb_ija_w, s.hb_ija_U, s.hb_ija_tot, s.hb_ijab_w, s.hb_ijab_U,
                           s.hb_ijab_tot, s.hb_ija_K, s.hb_ijab_K};
    if (which < 0 || which > 9) FAIL("download_heat_bath: bad table id");
    const size_t esz = which >= 8 ? sizeof(int) : sizeof(double);
    CK(copy_sync(e, out, src[which], (size_t)n * esz, cudaMemcpyDeviceToHost));
    return 0;
}

int hb200_set_reference(hb200_engine* e, const uint64_t* f0, double H00) {
    for (int k = 0; k < HB_MAXW; ++k) e->par.f0[k] = (k < e->W) ? f0[k] : 0;
    e->par.H00 = H00;
    e->have_ref = true;
    return 0;
}

int hb200_set_proc_map(hb200_engine* e, const int32_t* map, int32_t n) {
    CK(cudaSetDevice(e->cfg.device));
    if (n != e->par.nprocs * e->par.nslots) FAIL("set_proc_map: wrong length");
    CK(copy_sync(e, e->d_proc_map, map, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

int hb200_upload_psips(hb200_engine* e, const uint64_t* states, const int64_t* pops, const double* dat, int64_t n) {
    CK(cudaSetDevice(e->cfg.device));
    if (n > e->cfg.walker_length) FAIL("upload_psips: more states than walker_length");
    const int c = e->cur;
    long long s = 0;
    CK(cudaMemsetAsync(e->d_err, 0, 4 * sizeof(int), e->stream));  // a new list starts a new calculation
    if (n) {
        cudaStream_t st = e->stream;
        CK(cudaMemcpyAsync(e->d_states[c], states, (size_t)n * e->W * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(e->d_pops[c], pops, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(e->d_dat[c], dat, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        const int nb = (int)std::min<long long>(1184, (n + TILE - 1) / TILE);
        k_abs_sum<<<nb, TILE, 0, st>>>(e->d_pops[c], n, e->d_part_ll);
        k_reduce_ll<<<1, 1024, 0, st>>>(e->d_part_ll, nb, e->d_ll);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&s, e->d_ll, sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    e->nstates = n;
    e->nparticles_enc = s;
    return 0;
}

// Asynchronous variant for hosts that keep particle_t on the CPU: the copy of the NEXT list runs on a second stream
// into a third buffer while hb200_iterate works on the current one; hb200_upload_psips_commit makes it current.
int hb200_upload_psips_begin(hb200_engine* e, const uint64_t* states, const int64_t* pops, const double* dat, int64_t n) {
    CK(cudaSetDevice(e->cfg.device));
    if (n > e->cfg.walker_length) FAIL("upload_psips_begin: more states than walker_length");
    if (!e->copy_stream) {
        const size_t cap = (size_t)e->cfg.walker_length;
        if (dalloc(e, &e->d_states[2], cap * e->W)) return 1;
        if (dalloc(e, &e->d_pops[2], cap)) return 1;
        if (dalloc(e, &e->d_dat[2], cap)) return 1;
        CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&e->copy_done, cudaEventDisableTiming));
    }
    const int g = e->stg;
    if (n) {
        CK(cudaMemcpyAsync(e->d_states[g], states, (size_t)n * e->W * 8, cudaMemcpyHostToDevice, e->copy_stream));
        CK(cudaMemcpyAsync(e->d_pops[g], pops, (size_t)n * 8, cudaMemcpyHostToDevice, e->copy_stream));
        CK(cudaMemcpyAsync(e->d_dat[g], dat, (size_t)n * 8, cudaMemcpyHostToDevice, e->copy_stream));
    }
    CK(cudaEventRecord(e->copy_done, e->copy_stream));
    e->stg_n = n;
    return 0;
}
int hb200_upload_psips_commit(hb200_engine* e) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->stg_n < 0) FAIL("upload_psips_commit: no upload in flight");
    cudaStream_t st = e->stream;
    CK(cudaStreamWaitEvent(st, e->copy_done, 0));
    const long long n = e->stg_n;
    const int g = e->stg;
    e->stg = e->cur; e->cur = g; e->stg_n = -1;
    long long s = 0;
    CK(cudaMemsetAsync(e->d_err, 0, 4 * sizeof(int), st));
    if (n) {
        const int nb = (int)std::min<long long>(1184, (n + TILE - 1) / TILE);
        k_abs_sum<<<nb, TILE, 0, st>>>(e->d_pops[g], n, e->d_part_ll);
        k_reduce_ll<<<1, 1024, 0, st>>>(e->d_part_ll, nb, e->d_ll);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&s, e->d_ll, sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    e->nstates = n;
    e->nparticles_enc = s;
    return 0;
}

int hb200_download_psips(hb200_engine* e, uint64_t* states, int64_t* pops, double* dat, int64_t capacity, int64_t* nstates) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    const long long n = e->nstates;
    *nstates = n;
    if (n > capacity) FAIL("download_psips: capacity too small");
    const int c = e->cur;
    if (n) {
        cudaStream_t st = e->stream;
        CK(cudaMemcpyAsync(states, e->d_states[c], (size_t)n * e->W * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(pops, e->d_pops[c], (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(dat, e->d_dat[c], (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

int64_t hb200_nstates(hb200_engine* e) { return e->nstates; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// stage drivers
// ------------------------------------------------------------------------------------------------
#define DISPATCH_W(e, ...)                                    \
    switch ((e)->W) {                                         \
        case 1: { constexpr int WW = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int WW = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int WW = 3; __VA_ARGS__; } break; \
        default: { constexpr int WW = 4; __VA_ARGS__; } break; \
    }

static int stage_spawn_death(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, CycleStats* hst) {
    Params& p = e->par;
    p.tau = in->tau; p.shift = in->shift; p.proj_energy_old = in->proj_energy_old; p.cycle = cycle;
    cudaStream_t st = e->stream;
    CK(cudaMemsetAsync(e->d_head, 0, sizeof(unsigned long long) * p.nprocs, st));
    const long long n = e->nstates;
    const int ntiles = (int)((n + TILE - 1) / TILE);
    if (ntiles > 0) {
        const size_t smem = spawn_smem_bytes(e);
        const int c = e->cur;
        CK(cudaEventRecord(e->evk[0], st));
#define LAUNCH_SPAWN(GG)                                                                                          \
    DISPATCH_W(e, {                                                                                               \
        static bool attr_set = false;                                                                             \
        if (!attr_set) {                                                                                          \
            CK(cudaFuncSetAttribute(k_spawn_death<WW, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
            attr_set = true;                                                                                      \
        }                                                                                                         \
        k_spawn_death<WW, GG><<<ntiles, TILE, smem, st>>>(e->sys, p, e->d_states[c], e->d_pops[c], e->d_dat[c], n, \
                                                          e->d_spawn[0], e->d_head, e->block_size, e->d_proc_map, \
                                                          e->d_partials, e->d_err);                               \
    })
        if (e->sys.kind == SYS_UEG) {
            if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER) { LAUNCH_SPAWN(GEN_UEG_PP); }
            else { LAUNCH_SPAWN(GEN_UEG); }
        }
        else switch (e->cfg.excit_gen) {
            case HB200_EXCIT_GEN_NO_RENORM: LAUNCH_SPAWN(EXCIT_GEN_NO_RENORM); break;
            case HB200_EXCIT_GEN_RENORM: LAUNCH_SPAWN(EXCIT_GEN_RENORM); break;
            case HB200_EXCIT_GEN_RENORM_SPIN: LAUNCH_SPAWN(EXCIT_GEN_RENORM_SPIN); break;
            case HB200_EXCIT_GEN_POWER_PITZER_ORDERN: LAUNCH_SPAWN(EXCIT_GEN_POWER_PITZER_ORDERN); break;
            case HB200_EXCIT_GEN_POWER_PITZER: LAUNCH_SPAWN(EXCIT_GEN_POWER_PITZER); break;
            case HB200_EXCIT_GEN_NO_RENORM_SPIN: LAUNCH_SPAWN(EXCIT_GEN_NO_RENORM_SPIN); break;
            case HB200_EXCIT_GEN_HEAT_BATH: LAUNCH_SPAWN(EXCIT_GEN_HEAT_BATH); break;
            case HB200_EXCIT_GEN_HEAT_BATH_UNIFORM: LAUNCH_SPAWN(EXCIT_GEN_HEAT_BATH_UNIFORM); break;
            case HB200_EXCIT_GEN_HEAT_BATH_SINGLE: LAUNCH_SPAWN(EXCIT_GEN_HEAT_BATH_SINGLE); break;
            case HB200_EXCIT_GEN_POWER_PITZER_OCC:
            case HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC: LAUNCH_SPAWN(EXCIT_GEN_POWER_PITZER_OCC); break;
            case HB200_EXCIT_GEN_POWER_PITZER_OCC_IJ:
            case HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ: LAUNCH_SPAWN(EXCIT_GEN_POWER_PITZER_OCC_IJ); break;
            default: FAIL("spawn_death: excitation generator not implemented");
        }
#undef LAUNCH_SPAWN
        CK(cudaGetLastError());
        CK(cudaEventRecord(e->evk[1], st));
        e->launches++; e->spawn_launches++;
    }
    k_reduce_partials<<<1, 1024, 0, st>>>(e->d_partials, ntiles, e->d_stats);
    CK(cudaGetLastError());
    e->launches++;
    if (p.ps_part && ntiles > 0) {
        k_reduce_ps<<<1, 1024, 0, st>>>(e->d_ps_part, ntiles, e->d_ps_acc);
        CK(cudaGetLastError());
        e->launches++;
    }
    CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * p.nprocs, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hst, e->d_stats, sizeof(CycleStats), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ntiles > 0) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e->evk[0], e->evk[1]);
        e->spawn_kernel_ms += t;
    }
    for (int d = 0; d < p.nprocs; ++d)
        if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;  // overflow: dropped
    e->sp_cur = 0;
    e->sp_blocked = true;
    e->sp_n = (p.nprocs == 1) ? (long long)e->h_head[0] : 0;
    if (p.nprocs == 1) e->sp_blocked = false;
    return 0;
}

static int stage_comm(hb200_engine* e) {
    Params& p = e->par;
    if (p.nprocs == 1) { e->sp_blocked = false; e->sp_n = (long long)e->h_head[0]; e->sp_cur = 0; return 0; }
    if (!e->comm) FAIL("comm_spawn: nprocs > 1 but hb200_comm_init was not called");
    cudaStream_t st = e->stream;
    const int np = p.nprocs, me = p.iproc, E = e->E;
    // MPI_Alltoall of the counts (src/spawn_data.F90:693) as an all-gather of each rank's row
    std::vector<long long> row(np);
    for (int d = 0; d < np; ++d) row[d] = (long long)e->h_head[d];
    CK(cudaMemcpyAsync(e->d_counts + (size_t)me * np, row.data(), sizeof(long long) * np, cudaMemcpyHostToDevice, st));
    NCK(g_nccl.AllGather(e->d_counts + (size_t)me * np, e->d_counts, np, ncclInt64, e->comm, st));
    std::vector<long long> counts((size_t)np * np);
    CK(cudaMemcpyAsync(counts.data(), e->d_counts, sizeof(long long) * np * np, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // MPI_Alltoallv (src/spawn_data.F90:721): receive blocks ordered by source rank
    long long off = 0;
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < np; ++r) {
        const long long nsend = counts[(size_t)me * np + r], nrecv = counts[(size_t)r * np + me];
        if (r == me) {
            if (nsend) CK(cudaMemcpyAsync(e->d_spawn[1] + off * E, e->d_spawn[0] + (long long)r * e->block_size * E,
                                          (size_t)nsend * E * 8, cudaMemcpyDeviceToDevice, st));
        } else {
            if (nsend) NCK(g_nccl.Send(e->d_spawn[0] + (long long)r * e->block_size * E, (size_t)nsend * E, ncclInt64, r, e->comm, st));
            if (nrecv) NCK(g_nccl.Recv(e->d_spawn[1] + off * E, (size_t)nrecv * E, ncclInt64, r, e->comm, st));
        }
        off += nrecv;
    }
    NCK(g_nccl.GroupEnd());
    if (off > e->cfg.spawned_walker_length) FAIL("comm_spawn: received more than spawned_walker_length");
    e->sp_cur = 1;
    e->sp_n = off;
    e->sp_blocked = false;
    return 0;
}

static int device_scan(hb200_engine* e, const int* d_in, int* d_out, long long n, int slot);

static int stage_sort(hb200_engine* e) {
    const long long n = e->sp_n;
    if (n <= 1) return 0;
    cudaStream_t st = e->stream;
    const int E = e->E;
    const long long chunk = std::max<long long>(2048, ((n + 1183) / 1184 + SORT_THREADS - 1) / SORT_THREADS * SORT_THREADS);
    const int nblk = (int)((n + chunk - 1) / chunk);
    if (256ll * nblk > e->hist_cap) FAIL("sort: histogram scratch too small");
    const int npass = (e->cfg.nbasis + 7) / 8;
    for (int ps = 0; ps < npass; ++ps) {
        const int word = (8 * ps) / 64, shift = (8 * ps) % 64;
        const int64_t* src = e->d_spawn[e->sp_cur];
        int64_t* dst = e->d_spawn[e->sp_cur ^ 1];
        switch (E) {
            case 3: k_radix_hist<3><<<nblk, SORT_THREADS, 0, st>>>(src, n, word, shift, e->d_hist, nblk, chunk); break;
            case 4: k_radix_hist<4><<<nblk, SORT_THREADS, 0, st>>>(src, n, word, shift, e->d_hist, nblk, chunk); break;
            case 5: k_radix_hist<5><<<nblk, SORT_THREADS, 0, st>>>(src, n, word, shift, e->d_hist, nblk, chunk); break;
            default: k_radix_hist<6><<<nblk, SORT_THREADS, 0, st>>>(src, n, word, shift, e->d_hist, nblk, chunk); break;
        }
        if (256ll * nblk <= 8192) {
            k_scan_u32_single<<<1, 1024, 0, st>>>(e->d_hist, 256ll * nblk);
        } else {
            // multi-block exclusive scan (counts < 2^31, so the int scan is bit-identical); in place
            if (device_scan(e, (const int*)e->d_hist, (int*)e->d_hist, 256ll * nblk, 2)) return 1;
            e->launches += 2;
        }
        switch (E) {
            case 3: k_radix_scatter<3><<<nblk, SORT_THREADS, 0, st>>>(src, dst, n, word, shift, e->d_hist, nblk, chunk); break;
            case 4: k_radix_scatter<4><<<nblk, SORT_THREADS, 0, st>>>(src, dst, n, word, shift, e->d_hist, nblk, chunk); break;
            case 5: k_radix_scatter<5><<<nblk, SORT_THREADS, 0, st>>>(src, dst, n, word, shift, e->d_hist, nblk, chunk); break;
            default: k_radix_scatter<6><<<nblk, SORT_THREADS, 0, st>>>(src, dst, n, word, shift, e->d_hist, nblk, chunk); break;
        }
        CK(cudaGetLastError());
        e->launches += 3;
        e->sp_cur ^= 1;
    }
    return 0;
}

// exclusive scan of n ints (d_in -> d_out), total to d_total[slot]
static int device_scan(hb200_engine* e, const int* d_in, int* d_out, long long n, int slot) {
    cudaStream_t st = e->stream;
    if (n <= 0) { CK(cudaMemsetAsync(e->d_total + slot, 0, sizeof(int), st)); return 0; }
    if (n <= 4 * SCAN_BLOCK) {
        k_scan_small<<<1, 1024, 0, st>>>(d_in, d_out, n, e->d_total + slot);
        e->launches++;
    } else {
        const long long nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
        k_scan_block<<<(unsigned)nb, TILE, 0, st>>>(d_in, d_out, n, e->d_scan_l1);
        k_scan_small<<<1, 1024, 0, st>>>(e->d_scan_l1, e->d_scan_l1o, nb, e->d_total + slot);
        k_scan_add<<<(unsigned)nb, TILE, 0, st>>>(d_out, n, e->d_scan_l1o);
        e->launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}

static int stage_annihilate_main(hb200_engine* e, uint32_t cycle, CycleStats* hst) {
    Params& p = e->par;
    p.cycle = cycle;
    cudaStream_t st = e->stream;
    const long long n = e->sp_n, ns = e->nstates;
    const int c = e->cur, o = e->alt;
    int64_t* sp = e->d_spawn[e->sp_cur];
    int64_t* ins = e->d_spawn[e->sp_cur ^ 1];
    int h_tot[2] = {0, 0};
    if (n > 0) {
        DISPATCH_W(e, k_annihilate<WW><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, sp, n, e->d_states[c], e->d_pops[c], ns,
                                                                                   e->d_ins_flag, e->d_ins_pos));
        CK(cudaGetLastError());
        e->launches++;
        if (device_scan(e, e->d_ins_flag, e->d_ins_idx, n, 0)) return 1;
        DISPATCH_W(e, k_compact_inserts<WW><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sp, n, e->d_ins_flag, e->d_ins_idx,
                                                                                        e->d_ins_pos, ins));
        CK(cudaGetLastError());
        e->launches++;
    } else {
        CK(cudaMemsetAsync(e->d_total, 0, sizeof(int), st));
    }
    const int ntiles = std::max<int>(1, (int)((ns + TILE - 1) / TILE));
    DISPATCH_W(e, k_round_count<WW><<<ntiles, TILE, 0, st>>>(p, e->d_states[c], e->d_pops[c], ns, e->d_tile_keep));
    CK(cudaGetLastError());
    e->launches++;
    if (device_scan(e, e->d_tile_keep, e->d_tile_off, ntiles, 1)) return 1;
    CK(cudaMemcpyAsync(h_tot, e->d_total, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    long long nins = h_tot[0];
    const long long nkept = h_tot[1];
    // insert_new_walkers capacity check (src/annihilation.f90:750-771)
    if (nkept + nins > e->cfg.walker_length) {
        int one = 1;
        CK(cudaMemcpyAsync(e->d_err + 1, &one, sizeof(int), cudaMemcpyHostToDevice, st));
        nins = 0;
    }
    if (nins > 0) {
        DISPATCH_W(e, k_sc0<WW><<<(unsigned)((nins + 255) / 256), 256, 0, st>>>(e->sys, p.H00, (const uint64_t*)ins, e->E, nins,
                                                                               e->d_ins_dat));
        CK(cudaGetLastError());
        e->launches++;
    }
    DISPATCH_W(e, k_merge<WW><<<ntiles, TILE, 0, st>>>(e->d_states[c], e->d_pops[c], e->d_dat[c], ns, e->d_tile_off, ins,
                                                        e->d_ins_dat, nins, e->d_states[o], e->d_pops[o], e->d_dat[o],
                                                        e->d_part_ll, ntiles));
    CK(cudaGetLastError());
    k_reduce_ll<<<1, 1024, 0, st>>>(e->d_part_ll, ntiles, e->d_ll);
    CK(cudaGetLastError());
    e->launches += 2;
    long long npart = 0;
    CK(cudaMemcpyAsync(&npart, e->d_ll, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    e->cur = o;
    e->alt = c;
    e->nstates = nkept + nins;
    e->nparticles_enc = npart;
    hst->nkept = nkept;
    hst->npart_new = npart;
    e->sp_n = 0;
    return 0;
}

static void fill_out(hb200_engine* e, hb200_iter_out* out, const CycleStats& st, long long nattempts) {
    out->nparticles = (double)e->nparticles_enc / (double)e->par.real_factor;
    out->nstates = e->nstates;
    out->ndeath = st.ndeath;
    out->nattempts = nattempts;
    int herr[2] = {0, 0};
    copy_sync(e, herr, e->d_err, 2 * sizeof(int), cudaMemcpyDeviceToHost);
    out->spawn_error = herr[0];
    out->psip_error = herr[1];
}

extern "C" {

int hb200_spawn_death(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, hb200_iter_out* out) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("spawn_death: system not set");
    if (uses_heat_bath_tables(e) && !e->have_hb) FAIL("spawn_death: heat-bath tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("spawn_death: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("spawn_death: power_pitzer tables not built");
    CycleStats st;
    memset(&st, 0, sizeof(st));
    const long long nattempts = llround(2.0 * ((double)e->nparticles_enc / (double)e->par.real_factor));
    if (stage_spawn_death(e, in, cycle, &st)) return 1;
    e->nparticles_enc = st.npart_after_death;
    if (out) {
        memset(out, 0, sizeof(*out));
        out->proj_energy = st.pe; out->D0_population = st.d0;
        long long ev = 0;
        for (int d = 0; d < e->par.nprocs; ++d) ev += (long long)e->h_head[d];
        out->nspawn_events = ev;
        out->nattempts_spawn = st.nattempts_spawn;
        fill_out(e, out, st, nattempts);
    }
    return 0;
}

// One CCMC cycle up to (not including) annihilation: get_D0_info, init_mc_cycle, cumulative_population,
// set_cluster_selections and the iattempt loop of do_ccmc (src/ccmc.f90:625-857).
int hb200_ccmc_spawn(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, int32_t ex_level, hb200_ccmc_out* out) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("ccmc_spawn: system not set");
    if (uses_heat_bath_tables(e) && !e->have_hb) FAIL("ccmc_spawn: heat-bath tables not built");
    if (e->par.qn) FAIL("ccmc_spawn: the quasi-Newton propagator is only implemented for FCIQMC");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("ccmc_spawn: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("ccmc_spawn: power_pitzer tables not built");
    if (e->par.nprocs > 1 && !e->comm) FAIL("ccmc_spawn: nprocs > 1 but hb200_comm_init was not called");
    if (e->cfg.initiator_approx) FAIL("ccmc_spawn: the initiator approximation is not implemented for CCMC");
    Params& p = e->par;
    p.tau = in->tau; p.shift = in->shift; p.proj_energy_old = in->proj_energy_old; p.cycle = cycle;
    cudaStream_t st = e->stream;
    const long long n = e->nstates;
    const int c = e->cur;
    memset(out, 0, sizeof(*out));
    if (!e->d_cum) {
        const long long cap = e->cfg.walker_length;
        if (dalloc(e, &e->d_cum, (size_t)cap)) return 1;
        if (dalloc(e, &e->d_cum_blk, (size_t)(cap / (256 * SCAN64_ITEMS) + 2))) return 1;
        if (dalloc(e, &e->d_cc_part, 2 * (size_t)(e->cfg.spawned_walker_length / 256 + cap / 256 + 4))) return 1;
        if (dalloc(e, &e->d_cc_tot, 2)) return 1;
    }
    CK(cudaMemsetAsync(e->d_head, 0, sizeof(unsigned long long) * p.nprocs, st));
    // get_D0_info (src/ccmc_utils.F90:69-130): owner of the reference under the current hash shift, its position and
    // population there, broadcast to every rank (MPI_Bcast -> ncclBroadcast)
    p.ccmc_shift = e->ccmc_hash_shift; p.ccmc_freq = e->ccmc_move_freq;
    int D0_proc = 0;
    if (p.nprocs > 1) {
        std::vector<int> map((size_t)p.nprocs * p.nslots);
        CK(cudaMemcpyAsync(map.data(), e->d_proc_map, map.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int slot = 0;
        DISPATCH_W(e, slot = owner_slot_shift<WW>(p.f0, e->sys.nbasis, p.hash_seed, p.ccmc_shift, p.ccmc_freq, p.nprocs, p.nslots));
        D0_proc = map[slot];
    }
    long long d0info[2] = {0, 0};
    const bool have_D0 = (p.iproc == D0_proc);
    if (have_D0) {
        if (n > 0) {
            DISPATCH_W(e, k_find_det<WW><<<1, 32, 0, st>>>(p, e->d_states[c], e->d_pops[c], n, e->d_ll));
            CK(cudaGetLastError());
            e->launches++;
        } else {
            CK(cudaMemsetAsync(e->d_ll, 0, 2 * sizeof(long long), st));
        }
    }
    if (p.nprocs > 1) NCK(g_nccl.Broadcast(e->d_ll, e->d_ll, 2, ncclInt64, D0_proc, e->comm, st));
    CK(cudaMemcpyAsync(d0info, e->d_ll, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (d0info[0] == 0) FAIL("ccmc_spawn: find_D0: cannot find the reference in the excip list");
    e->ccmc_hash_shift += 1;                    // src/ccmc.f90:625
    p.ccmc_shift = e->ccmc_hash_shift;
    CcmcArgs a;
    a.nstates = n; a.D0_pos = have_D0 ? d0info[0] : -1;
    a.D0_normalisation = (double)d0info[1] / (double)p.real_factor;
    a.ex_level = ex_level; a.nprocs = p.nprocs;
    a.max_cluster_size = (int)std::min<long long>(std::min(e->sys.nel, ex_level + 2), n - (have_D0 ? 1 : 0));
    // init_mc_cycle (src/qmc_common.F90:950-1017, ccmc branch) with min_attempts = nint(|D0_normalisation|)
    long long nattempts = (long long)((double)e->nparticles_enc / (double)p.real_factor);
    nattempts = std::max<long long>(nattempts, llround(fabs(a.D0_normalisation)));
    // cumulative_population (src/ccmc_utils.F90:427-563)
    long long tot_enc = 0;
    if (n > 0) {
        const int nb = (int)((n + 256 * SCAN64_ITEMS - 1) / (256 * SCAN64_ITEMS));
        k_cum_block<<<nb, 256, 0, st>>>(e->d_pops[c], n, have_D0 ? a.D0_pos - 1 : -1, e->d_cum, e->d_cum_blk);
        k_cum_sums<<<1, 1024, 0, st>>>(e->d_cum_blk, nb);
        k_cum_add<<<nb, 256, 0, st>>>(e->d_cum, n, e->d_cum_blk);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&tot_enc, e->d_cum + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        e->launches += 3;
    }
    a.tot_abs_real_pop = (double)tot_enc / (double)p.real_factor;
    // set_cluster_selections (src/ccmc_selection.f90:874-948)
    a.full_nc = e->ccmc_full_nc ? 1 : 0;
    if (e->ccmc_full_nc) {
        a.min_cluster_size = 2;
        a.nD0_select = llround(fabs(a.D0_normalisation));
        a.nstochastic = (long long)ceil(a.tot_abs_real_pop);
        nattempts = llround(a.tot_abs_real_pop) + a.nD0_select + a.nstochastic;   // estimators%nattempts
    } else {
        a.min_cluster_size = 0;
        a.nD0_select = 0;
        a.nstochastic = nattempts;
    }
    a.nattempts = a.nstochastic + a.nD0_select;
    CcmcPartials tot, tot_nc;
    memset(&tot, 0, sizeof(tot));
    memset(&tot_nc, 0, sizeof(tot_nc));
    const size_t part_cap = (size_t)(e->cfg.spawned_walker_length / 256 + e->cfg.walker_length / 256 + 4);
    if (a.nattempts > 0) {
        const long long nblk = (a.nattempts + 255) / 256;
        if ((size_t)nblk > part_cap)
            FAIL("ccmc_spawn: more cluster selections than the partial-sum scratch holds");
        DISPATCH_W(e, k_ccmc_cluster<WW><<<(unsigned)nblk, 256, 0, st>>>(e->sys, p, a, e->d_states[c], e->d_pops[c], e->d_dat[c],
                                                                          e->d_cum, e->d_spawn[0], e->d_head, e->block_size,
                                                                          e->d_proc_map, e->d_cc_part, e->d_err));
        CK(cudaGetLastError());
        k_ccmc_reduce<<<1, 1024, 0, st>>>(e->d_cc_part, (int)nblk, e->d_cc_tot);
        CK(cudaGetLastError());
        e->launches += 2;
        if (p.ps_part) {
            k_reduce_ps<<<1, 1024, 0, st>>>(e->d_ps_part, (int)nblk, e->d_ps_acc);
            CK(cudaGetLastError());
            e->launches++;
        }
        CK(cudaMemcpyAsync(&tot, e->d_cc_tot, sizeof(tot), cudaMemcpyDeviceToHost, st));
    }
    if (e->ccmc_full_nc && n > 0) {
        // non-composite clusters + in-place death; after k_ccmc_cluster, which reads the populations changed here
        const long long nblk = (n + 255) / 256;
        Params pn = p;
        if (pn.ps_part) pn.ps_part += part_cap;
        DISPATCH_W(e, k_ccmc_nc<WW><<<(unsigned)nblk, 256, 0, st>>>(e->sys, pn, a, e->d_states[c], e->d_pops[c], e->d_dat[c],
                                                                     e->d_spawn[0], e->d_head, e->block_size, e->d_proc_map,
                                                                     e->d_cc_part + part_cap, nullptr, e->d_err));
        CK(cudaGetLastError());
        k_ccmc_reduce<<<1, 1024, 0, st>>>(e->d_cc_part + part_cap, (int)nblk, e->d_cc_tot + 1);
        CK(cudaGetLastError());
        e->launches += 2;
        if (pn.ps_part) {
            k_reduce_ps<<<1, 1024, 0, st>>>(pn.ps_part, (int)nblk, e->d_ps_acc);
            CK(cudaGetLastError());
            e->launches++;
        }
        CK(cudaMemcpyAsync(&tot_nc, e->d_cc_tot + 1, sizeof(tot_nc), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * p.nprocs, cudaMemcpyDeviceToHost, st));
    int herr[2] = {0, 0};
    CK(cudaMemcpyAsync(herr, e->d_err, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    long long nspawn_events = 0;                 // calc_events_spawn_t: counted before redistribute_particles
    for (int d = 0; d < p.nprocs; ++d) {
        if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;
        nspawn_events += (long long)e->h_head[d];
    }
    if (p.nprocs > 1 && n > 0) {
        DISPATCH_W(e, k_ccmc_redistribute<WW><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->sys, p, e->d_states[c], e->d_pops[c], n,
                                                                                          e->d_spawn[0], e->d_head, e->block_size,
                                                                                          e->d_proc_map, e->d_err));
        CK(cudaGetLastError());
        e->launches++;
        CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * p.nprocs, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(herr, e->d_err, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int d = 0; d < p.nprocs; ++d)
            if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;
    }
    e->sp_cur = 0;
    e->sp_blocked = p.nprocs > 1;
    e->sp_n = (p.nprocs == 1) ? (long long)e->h_head[0] : 0;
    out->proj_energy = tot.pe + tot_nc.pe; out->D0_population = tot.d0 + tot_nc.d0;
    out->D0_normalisation = a.D0_normalisation;
    out->nattempts = nattempts; out->nattempts_spawn = tot.nattempts_spawn + tot_nc.nattempts_spawn;
    out->ndeath = tot.ndeath; out->ndeath_nc = tot_nc.ndeath;
    out->nspawn_events = nspawn_events; out->tot_abs_real_pop = a.tot_abs_real_pop;
    out->spawn_error = herr[0]; out->psip_error = herr[1];
    return 0;
}

// find_parallel_spin_prob_mol (src/qmc_common.F90:262-377): sum |<ij|H|ab>| over all orbital quadruples, split by
// parallel / anti-parallel ij.  One block per i; fixed-order reductions.
__global__ void __launch_bounds__(256) k_parallel_spin_prob(Sys s, double* __restrict__ part) {
    __shared__ double sh[2][8];
    const int nb = s.nbasis;
    const int i = blockIdx.x + 1;
    double par = 0.0, ortho = 0.0;
    for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) {
        const int j = t / nb + 1, a = t % nb + 1;
        if (i == j || a == i || a == j) continue;
        const int it = min(i, j), jt = max(i, j);
        const int ij_sym = sym_conj(s, cross_product(s, s.bf_sym[it], s.bf_sym[jt]));
        const int isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
        for (int b = 1; b <= nb; ++b) {
            const bool spin_ok = (ms_of(it) == ms_of(a) && ms_of(jt) == ms_of(b)) || (ms_of(it) == ms_of(b) && ms_of(jt) == ms_of(a));
            if (!(spin_ok && s.bf_sym[b] == isymb && b != a && b != i && b != j)) continue;
            const double h = fabs(slater_condon2_excit(s, it, jt, min(a, b), max(a, b), false));
            if (ms_of(it) == ms_of(jt)) par = par + h; else ortho = ortho + h;
        }
    }
    par = warp_sum_d(par); ortho = warp_sum_d(ortho);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = par; sh[1][warp] = ortho; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
        part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    }
}
// qmc_in%pattempt_parallel (src/qmc.F90:974-988) for excit_gen = renorm_spin / no_renorm_spin: a negative value asks
// for find_parallel_spin_prob_mol.
int hb200_set_pattempt_parallel(hb200_engine* e, double pattempt_parallel) {
    if (pattempt_parallel < 0.0) {
        if (e->sys.kind != SYS_READ_IN || e->sys.nbasis <= 0) FAIL("set_pattempt_parallel: needs a read_in system");
        const int nb = e->sys.nbasis;
        double* d_part = nullptr;
        CK(cudaMalloc(&d_part, sizeof(double) * 2 * nb));
        k_parallel_spin_prob<<<nb, 256, 0, e->stream>>>(e->sys, d_part);
        std::vector<double> h(2 * (size_t)nb);
        CK(cudaMemcpyAsync(h.data(), d_part, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaFree(d_part));
        double par = 0.0, ortho = 0.0;
        for (int i = 0; i < nb; ++i) { par += h[2 * i]; ortho += h[2 * i + 1]; }
        pattempt_parallel = par / (par + ortho);
    }
    e->par.pattempt_parallel = pattempt_parallel;
    return 0;
}
double hb200_get_pattempt_parallel(hb200_engine* e) { return e->par.pattempt_parallel; }

// qmc = { quasi_newton = true } (propagator_t, src/qmc_data.f90:866-884; init_sp_fock / init_quasi_newton,
// src/qmc.F90:1064-1160): sp_fock[0..nbasis] (entry 0 unused; null switches the propagator off), the reference's
// fock_sum and the threshold / value / population-control scalars.  FCIQMC only.
int hb200_set_quasi_newton(hb200_engine* e, const double* sp_fock, double ref_fock_sum, double threshold, double value,
                           double pop_control) {
    CK(cudaSetDevice(e->cfg.device));
    Params& p = e->par;
    if (!sp_fock) { p.qn = 0; p.sp_fock = nullptr; return 0; }
    if (!e->have_sys) FAIL("set_quasi_newton: system not set");
    double* d = nullptr;
    if (dalloc(e, &d, (size_t)e->sys.nbasis + 1)) return 1;
    CK(copy_sync(e, d, sp_fock, ((size_t)e->sys.nbasis + 1) * sizeof(double), cudaMemcpyHostToDevice));
    p.sp_fock = d; p.ref_fock_sum = ref_fock_sum; p.qn_threshold = threshold; p.qn_value = value; p.qn_pop_control = pop_control;
    p.qn = 1;
    return 0;
}

// qmc_in%pattempt_update (src/qmc.F90:1049-1060, src/spawning.F90:2139-2372): the engine holds pattempt_single /
// pattempt_double and, while `accumulate` is set, sums |H_ij| pattempt / pgen and the counts of the allowed single and
// double excitations it generates; the host reads the sums once per report loop, allreduces them and sets the new
// probabilities (update_pattempt_single).
int hb200_set_pattempt(hb200_engine* e, double pattempt_single, double pattempt_double, int32_t accumulate) {
    if (accumulate) {
        // src/check_input.F90:192-197
        if (e->sys.kind != SYS_READ_IN) FAIL("pattempt_update only used in read_in systems.");
        if (e->cfg.excit_gen == HB200_EXCIT_GEN_HEAT_BATH) FAIL("pattempt_update is not used with heat bath excitation generator.");
        if (!e->d_ps_part) {
            const size_t cc = 2 * (size_t)(e->cfg.spawned_walker_length / 256 + e->cfg.walker_length / 256 + 4);
            e->ps_part_cap = std::max<size_t>((size_t)e->max_tiles, cc);
            if (dalloc(e, &e->d_ps_part, e->ps_part_cap)) return 1;
            if (dalloc(e, &e->d_ps_acc, 4)) return 1;
            CK(cudaMemsetAsync(e->d_ps_acc, 0, 4 * sizeof(double), e->stream));
        }
    }
    e->par.pattempt_single = pattempt_single;
    e->par.pattempt_double = pattempt_double;
    e->cfg.pattempt_single = pattempt_single;
    e->cfg.pattempt_double = pattempt_double;
    e->par.ps_part = accumulate ? e->d_ps_part : nullptr;
    return 0;
}
// out[0..3] = h_pgen_singles_sum, excit_gen_singles, h_pgen_doubles_sum, excit_gen_doubles accumulated on this rank since
// the last reset (p_single_double_coll_t rep_accum, src/excit_gens.f90:13-27)
int hb200_get_ps_stats(hb200_engine* e, double* out, int32_t reset) {
    if (!e->d_ps_acc) { for (int k = 0; k < 4; ++k) out[k] = 0.0; return 0; }
    CK(cudaMemcpyAsync(out, e->d_ps_acc, 4 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (reset) CK(cudaMemsetAsync(e->d_ps_acc, 0, 4 * sizeof(double), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int hb200_ccmc_set_full_nc(hb200_engine* e, int32_t full_nc) {
    e->ccmc_full_nc = full_nc != 0;
    return 0;
}
int hb200_ccmc_set_hash_shift(hb200_engine* e, int32_t hash_shift, int32_t move_freq) {
    e->ccmc_hash_shift = hash_shift;
    e->ccmc_move_freq = move_freq;
    return 0;
}

// ncycles full CCMC cycles (the icycle loop of do_ccmc, src/ccmc.f90:603-896): cluster selection / spawning / death,
// then direct_annihilation and end_mc_cycle.
int hb200_ccmc_iterate(hb200_engine* e, int32_t ncycles, const hb200_iter_in* in, int32_t ex_level, hb200_iter_out* out) {
    CK(cudaSetDevice(e->cfg.device));
    memset(out, 0, sizeof(*out));
    CycleStats cs;
    memset(&cs, 0, sizeof(cs));
    hb200_ccmc_out co;
    long long nattempts = 0;
    for (int c = 0; c < ncycles; ++c) {
        const uint32_t cycle = in->first_cycle + (uint32_t)c;
        out->walker_iterations += (double)e->nparticles_enc / (double)e->par.real_factor;
        if (hb200_ccmc_spawn(e, in, cycle, ex_level, &co)) return 1;
        nattempts = co.nattempts;
        out->proj_energy += co.proj_energy;
        out->D0_population += co.D0_population;
        out->nattempts_spawn += co.nattempts_spawn;
        out->nspawn_events = co.nspawn_events;
        out->ndeath = co.ndeath;
        if (stage_comm(e)) return 1;
        if (stage_sort(e)) return 1;
        if (stage_annihilate_main(e, cycle, &cs)) return 1;
        // end_mc_cycle(nspawn_events, ndeath_nc, real_factor, nattempts_spawn, rspawn)
        if (co.nattempts_spawn > 0)
            out->rspawn += ((double)co.nspawn_events + (double)co.ndeath_nc / (double)e->par.real_factor) / (double)co.nattempts_spawn;
    }
    fill_out(e, out, cs, nattempts);
    out->ndeath = co.ndeath;
    return 0;
}

int hb200_comm_spawn(hb200_engine* e) {
    CK(cudaSetDevice(e->cfg.device));
    return stage_comm(e);
}

int hb200_annihilate_spawn(hb200_engine* e) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->sp_blocked) FAIL("annihilate_spawn: call hb200_comm_spawn first");
    if (stage_sort(e)) return 1;
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int hb200_annihilate_main(hb200_engine* e, uint32_t cycle, hb200_iter_out* out) {
    CK(cudaSetDevice(e->cfg.device));
    CycleStats st;
    memset(&st, 0, sizeof(st));
    if (stage_annihilate_main(e, cycle, &st)) return 1;
    if (out) fill_out(e, out, st, 0);
    return 0;
}

int hb200_download_spawn(hb200_engine* e, int64_t* sdata, int64_t capacity, int64_t* n) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    const int E = e->E;
    if (e->sp_blocked) {
        // still partitioned by destination: concatenate the blocks
        long long tot = 0;
        for (int d = 0; d < e->par.nprocs; ++d) tot += (long long)e->h_head[d];
        *n = tot;
        if (tot > capacity) FAIL("download_spawn: capacity too small");
        long long off = 0;
        for (int d = 0; d < e->par.nprocs; ++d) {
            const long long c = (long long)e->h_head[d];
            if (c) CK(copy_sync(e, sdata + off * E, e->d_spawn[0] + (long long)d * e->block_size * E, (size_t)c * E * 8,
                                 cudaMemcpyDeviceToHost));
            off += c;
        }
        return 0;
    }
    *n = e->sp_n;
    if (e->sp_n > capacity) FAIL("download_spawn: capacity too small");
    if (e->sp_n) CK(copy_sync(e, sdata, e->d_spawn[e->sp_cur], (size_t)e->sp_n * E * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int hb200_upload_spawn(hb200_engine* e, const int64_t* sdata, int64_t n) {
    CK(cudaSetDevice(e->cfg.device));
    if (n > e->cfg.spawned_walker_length) FAIL("upload_spawn: too many elements");
    if (n) CK(copy_sync(e, e->d_spawn[0], sdata, (size_t)n * e->E * 8, cudaMemcpyHostToDevice));
    e->sp_cur = 0; e->sp_n = n; e->sp_blocked = false;
    return 0;
}

int hb200_iterate(hb200_engine* e, int32_t ncycles, const hb200_iter_in* in, hb200_iter_out* out) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("iterate: system not set");
    if (uses_heat_bath_tables(e) && !e->have_hb) FAIL("iterate: heat-bath tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("iterate: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("iterate: power_pitzer tables not built");
    memset(out, 0, sizeof(*out));
    cudaStream_t st = e->stream;
    float acc[4] = {0, 0, 0, 0};
    e->spawn_kernel_ms = 0.f;
    CycleStats cs;
    memset(&cs, 0, sizeof(cs));
    long long nattempts = 0;
    CK(cudaEventRecord(e->ev[5], st));
    for (int c = 0; c < ncycles; ++c) {
        const uint32_t cycle = in->first_cycle + (uint32_t)c;
        // init_mc_cycle (src/qmc_common.F90:950-1017)
        nattempts = llround(2.0 * ((double)e->nparticles_enc / (double)e->par.real_factor));
        out->walker_iterations += (double)e->nparticles_enc / (double)e->par.real_factor;
        CK(cudaEventRecord(e->ev[0], st));
        if (stage_spawn_death(e, in, cycle, &cs)) return 1;
        CK(cudaEventRecord(e->ev[1], st));
        out->proj_energy += cs.pe;
        out->D0_population += cs.d0;
        out->nattempts_spawn += cs.nattempts_spawn;
        long long ev = 0;
        for (int d = 0; d < e->par.nprocs; ++d) ev += (long long)e->h_head[d];
        out->nspawn_events = ev;
        if (stage_comm(e)) return 1;
        CK(cudaEventRecord(e->ev[2], st));
        if (stage_sort(e)) return 1;
        CK(cudaEventRecord(e->ev[3], st));
        if (stage_annihilate_main(e, cycle, &cs)) return 1;
        CK(cudaEventRecord(e->ev[4], st));
        CK(cudaEventSynchronize(e->ev[4]));
        for (int k = 0; k < 4; ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, e->ev[k], e->ev[k + 1]);
            acc[k] += t;
        }
        // end_mc_cycle / spawning_rate (src/qmc_common.F90:1240-1304)
        const double ndeath_real = (double)cs.ndeath / (double)e->par.real_factor;
        if (nattempts > 0) out->rspawn += ((double)ev + ndeath_real) / (double)nattempts;
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, e->ev[5], e->ev[4]);
    for (int k = 0; k < 4; ++k) e->ms[k] = acc[k];
    e->ms[4] = tot;
    e->ms[5] = e->spawn_kernel_ms;
    fill_out(e, out, cs, nattempts);
    return 0;
}

int hb200_sc0_batch(hb200_engine* e, const uint64_t* states, int64_t n, double* out) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("sc0_batch: system not set");
    if (n == 0) return 0;
    uint64_t* d_f = nullptr;
    double* d_o = nullptr;
    CK(cudaMalloc((void**)&d_f, (size_t)n * e->W * 8));
    CK(cudaMalloc((void**)&d_o, (size_t)n * 8));
    CK(copy_sync(e, d_f, states, (size_t)n * e->W * 8, cudaMemcpyHostToDevice));
    DISPATCH_W(e, k_sc0<WW><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->sys, 0.0, d_f, e->W, n, d_o));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(copy_sync(e, out, d_o, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_f); cudaFree(d_o);
    return 0;
}

int hb200_gen_excit_batch(hb200_engine* e, const uint64_t* states, const int64_t* pops, const uint32_t* attempt,
                          int64_t n, uint32_t cycle, double tau, int32_t* iout, double* dout, int64_t* nspawn) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("gen_excit_batch: system not set");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("gen_excit_batch: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("gen_excit_batch: power_pitzer tables not built");
    if (n == 0) return 0;
    Params p = e->par;
    p.cycle = cycle; p.tau = tau;
    uint64_t* d_f; int64_t* d_p; uint32_t* d_a; int* d_io; double* d_do; int64_t* d_ns;
    CK(cudaMalloc((void**)&d_f, (size_t)n * e->W * 8));
    CK(cudaMalloc((void**)&d_p, (size_t)n * 8));
    CK(cudaMalloc((void**)&d_a, (size_t)n * 4));
    CK(cudaMalloc((void**)&d_io, (size_t)n * 8 * 4));
    CK(cudaMalloc((void**)&d_do, (size_t)n * 2 * 8));
    CK(cudaMalloc((void**)&d_ns, (size_t)n * 8));
    CK(copy_sync(e, d_f, states, (size_t)n * e->W * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, d_p, pops, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, d_a, attempt, (size_t)n * 4, cudaMemcpyHostToDevice));
    DISPATCH_W(e, k_gen_excit_batch<WW><<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(e->sys, p, d_f, d_p, d_a, n,
                                                                                           e->d_proc_map, d_io, d_do, d_ns));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(copy_sync(e, iout, d_io, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, dout, d_do, (size_t)n * 2 * 8, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, nspawn, d_ns, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_f); cudaFree(d_p); cudaFree(d_a); cudaFree(d_io); cudaFree(d_do); cudaFree(d_ns);
    return 0;
}

int hb200_get_unique_id(uint8_t id[128]) {
    ncclUniqueId uid;
    if (!g_nccl.load(g_err)) return 1;
    NCK(g_nccl.GetUniqueId(&uid));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id, &uid, 128);
    return 0;
}

int hb200_comm_init(hb200_engine* e, const uint8_t id[128]) {
    CK(cudaSetDevice(e->cfg.device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    if (!g_nccl.load(g_err)) return 1;
    NCK(g_nccl.CommInitRank(&e->comm, e->par.nprocs, uid, e->par.iproc));
    return 0;
}

int hb200_last_timing(hb200_engine* e, double ms[8], int64_t cnt[4]) {
    for (int k = 0; k < 8; ++k) ms[k] = e->ms[k];
    cnt[0] = e->spawn_launches; cnt[1] = e->launches; cnt[2] = 0; cnt[3] = 0;
    return 0;
}

}  // extern "C"
