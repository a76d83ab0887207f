// hande_b200: declarations shared by the translation units of libhande_b200.so (engine state, error macros, small
// device helpers).  The library is split into several .cu files so that they compile in parallel (hande_b200/build.py):
//   hb_engine.cu     C ABI, stage drivers, sort / annihilation / merge kernels, table builders
//   hb_spawn_tu.cu   k_spawn_death instantiations of one (W, generator group) per object file
//   hb_ccmc_tu.cu    CCMC kernels and the excitation-generator probe kernel of one W per object file
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>   // types only: the library is bound at run time (see NcclApi in hb_engine.cu)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/hande_b200.h"
#include "hb_core.cuh"

using namespace hb;

extern thread_local std::string g_err;
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + std::to_string(__LINE__); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)
#define FAIL(msg)        \
    do {                 \
        g_err = (msg);   \
        return 1;        \
    } while (0)

constexpr int ANN_SHORT = 32;   // k_annihilate: runs of equal keys up to this length are summed by the thread of their first element
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

// The element count of the spawn list is read on the device (*pn, clamped to cap): hb200_iterate never brings it to the
// host, so grids are sized from a host-side upper bound and the chunk of a block follows from the real count.
__device__ __forceinline__ long long dev_count(const unsigned long long* __restrict__ pn, long long cap) {
    const unsigned long long v = *pn;
    return v > (unsigned long long)cap ? cap : (long long)v;
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

struct SpawnPartials {  // one per block; reduced in fixed order by k_reduce_partials
    double pe, d0;
    long long ndeath, npart, nattempts;
};
struct CcmcPartials { double pe, d0; long long ndeath, nattempts_spawn; };
struct CycleStats {
    double pe, d0;              // this cycle
    long long ndeath, npart_after_death, nattempts_spawn;
    long long nkept, npart_new; // after merge
};

// layout of the pinned host block the results of a cycle are copied to
struct HostOut {
    CycleStats st;
    long long npart_new;
    int tot[4];            // [0] new determinants, [1] surviving states
    int err[4];
    unsigned long long spn;
    unsigned long long head[1];   // [nprocs]
};


// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
// semi_stoch_t (src/semi_stoch.F90:40-108), separate annihilation; see hb_semistoch.cuh
struct SemiStoch {
    bool on = false;
    int tot = 0, nloc = 0, maxsz = 0, displ = 0;   // determ%tot_size, sizes(iproc), max(sizes), displs(iproc)
    std::vector<int> sizes;
    uint64_t* d_all = nullptr;      // [tot][W] determ%dets: rank by rank, each rank's part sorted
    uint64_t* d_local = nullptr;    // = d_all + displ * W
    uint64_t* d_sorted = nullptr;   // [tot][W] sorted copy (check_if_determ)
    int* d_pad = nullptr;           // [tot] position of row i in the all-gathered vector (rank * maxsz + local index)
    long long* d_idx = nullptr;     // [nloc] determ%indices
    uint32_t* d_bits = nullptr;     // one bit per main-list state: determ%flags == 0
    double* d_full = nullptr;       // [nprocs * maxsz] determ%full_vector (padded per rank)
    double* d_vec = nullptr;        // [nloc] determ%vector after the projection
    double* d_rho = nullptr;        // [nloc] determ%rho_minus_qn_weight
    long long* d_colptr = nullptr;  // [nloc + 1] determ%hamil by column
    int* d_row = nullptr;
    double* d_val = nullptr;
    long long nnz = 0;
    int* d_miss = nullptr;
    std::vector<void*> bufs;        // everything above (freed when the space is replaced)
};

struct hb200_engine {
    hb200_config cfg;
    SemiStoch ss;
    int W = 1, E = 3;            // words per determinant / per spawn element on the DEVICE (1..4, or 32 = wide layout)
    int We = 1;                  // words per determinant in the host's layout, ceil(nbasis/64)
    const struct ListOps* ops = nullptr;
    int64_t* d_items[2] = {nullptr, nullptr};   // wide layout: compressed sort keys
    int key_bits = 0, key_words = 0;
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
    int64_t* sp_ptr[2] = {nullptr, nullptr};   // the two buffers the sort ping-pongs between (d_spawn, or a receive
                                               // buffer of the peer-to-peer exchange and d_spawn[0])
    int sp_cur = 0;      // buffer (index into sp_ptr) holding the current stage's list
    long long sp_n = 0;  // number of elements in it (contiguous from 0) after comm - staged calls only; hb200_iterate
                         // keeps the count on the device:
    unsigned long long* d_spn = nullptr;       // [1] element count written by the host for the staged calls
    const unsigned long long* sp_pn = nullptr; // where the sort / annihilation kernels read the element count
    long long sp_cap = 0;                      // ... and the bound they clamp it to
    long long last_spn = 1 << 20;              // spawn-list length of the previous cycle (sizes the sort's grid)
    unsigned long long* d_tile_state = nullptr; // fused merge: per-tile {flag, survivor count} words of the look-back
    unsigned* d_ticket = nullptr;
    long long* d_long_q = nullptr;             // k_annihilate: heads of the long runs of equal keys
    unsigned* d_long_n = nullptr;
    // pinned host block the cycle's results are copied into (one synchronisation per cycle)
    unsigned char* h_out = nullptr;
    // peer-to-peer exchange of the spawn blocks (hb200_p2p_export / hb200_p2p_import): receive buffers of this rank,
    // double buffered by cycle parity, exported through CUDA IPC; peer pointers of every rank
    unsigned char* d_p2p_block = nullptr;      // [heads (256 B) | recv 0 | recv 1], one allocation = one IPC handle
    int64_t* d_recv[2] = {nullptr, nullptr};
    unsigned long long* d_recv_head = nullptr; // [2]
    std::vector<void*> peer_base;              // mapped base of every rank's block (own: d_p2p_block)
    int64_t** d_peer_recv = nullptr;           // device array [2][nprocs]
    unsigned long long** d_peer_head = nullptr;// device array [nprocs] (head of parity k at +k)
    unsigned long long* d_snap = nullptr;      // [(chunks + 1)][nprocs] d_head after each spawn chunk
    long long* d_push = nullptr;               // [3][nprocs] lo, count, remote offset of the chunk being pushed
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_comm = nullptr;
    bool p2p = false;
    hb200_barrier_fn host_barrier = nullptr;   // ends an exchange instead of the NCCL collective (hb200_set_host_barrier)
    void* host_barrier_arg = nullptr;
    int xparity = 0;                           // receive buffer of the next exchange
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
    int npartials = 0;            // SpawnPartials written by the last spawn launch
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
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evk[2] = {nullptr, nullptr};           // brackets the k_spawn_death launch alone (roofline timing)
    float spawn_kernel_ms = 0.f;  // accumulated over the cycles of the last hb200_iterate
    double ms[8] = {0};
    long long launches = 0, spawn_launches = 0;
};

// GEN: compile-time generator of a k_spawn_death instantiation: EXCIT_GEN_* for read_in systems, GEN_UEG* for the UEG
enum { GEN_UEG = 100, GEN_UEG_PP = 101 };
// generator groups = object files of hb_spawn_tu.cu
enum { SPAWN_GROUP_HEAT_BATH = 0, SPAWN_GROUP_HB_UNIFORM = 1, SPAWN_GROUP_PP_OCC = 2, SPAWN_GROUP_UNIFORM = 3,
       SPAWN_GROUP_TABLES = 4, SPAWN_NGROUPS = 5 };
// launchers defined in hb_spawn_tu.cu (one per W and group) and hb_ccmc_tu.cu (one per W)
struct SpawnLaunch {
    int gen;            // GEN of the instantiation
    int ntiles;         // tiles of this launch
    size_t smem;
    long long n;        // states of the whole list
    int tile0;          // first tile of this launch (the spawning step of a multi-rank cycle is launched in chunks)
};
typedef int (*hb_spawn_fn)(hb200_engine* e, const Params& p, const SpawnLaunch& L);
#define HB_DECL_SPAWN(W, G) int hb_spawn_w##W##_g##G(hb200_engine* e, const Params& p, const SpawnLaunch& L);
#define HB_DECL_SPAWN_W(W) HB_DECL_SPAWN(W, 0) HB_DECL_SPAWN(W, 1) HB_DECL_SPAWN(W, 2) HB_DECL_SPAWN(W, 3) HB_DECL_SPAWN(W, 4)
HB_DECL_SPAWN_W(1) HB_DECL_SPAWN_W(2) HB_DECL_SPAWN_W(3) HB_DECL_SPAWN_W(4) HB_DECL_SPAWN(32, 4)
size_t hb_spawn_smem_bytes(const hb200_engine* e);
bool hb_uses_heat_bath_tables(const hb200_engine* e);

// list kernels of one bit-string width (hb_list.cuh / hb_list_tu.cu)
struct ListOps {
    int (*annihilate)(hb200_engine* e, const Params& p, int64_t* sp, long long bound);
    int (*compact)(hb200_engine* e, const int64_t* sp, long long bound, int64_t* ins);
    int (*round_count)(hb200_engine* e, const Params& p, int ntiles);
    int (*sc0)(hb200_engine* e, double H00, const uint64_t* dets, long long stride_words, long long n, double* out, const int* pn);
    int (*merge)(hb200_engine* e, const Params& p, const int64_t* ins, int ntiles, bool fused);
    int (*slot_pop)(hb200_engine* e, unsigned long long* d);
    int (*compress)(hb200_engine* e, const int64_t* sp, long long bound, int bits, int kw, int64_t* items);
    int (*gather)(hb200_engine* e, const int64_t* sp, int kw, const int64_t* items, int64_t* out);
    int (*owner_slot_shift)(const hb200_engine* e, const Params& p);
    // semi-stochastic projection (hb_semistoch.cuh)
    int (*ss_locate)(hb200_engine* e, int buf, const int* ntot2);
    int (*ss_hamil)(hb200_engine* e, int pass);
    int (*ss_project)(hb200_engine* e, const Params& p);
};
const ListOps* hb_list_ops_w1();
const ListOps* hb_list_ops_w2();
const ListOps* hb_list_ops_w3();
const ListOps* hb_list_ops_w4();
const ListOps* hb_list_ops_w32();

struct CcmcLaunch {
    CcmcArgs a;
    long long nblk;
    CcmcPartials* partials;
    const long long* cum;
};
#define HB_DECL_CCMC(W)                                                                                              \
    int hb_ccmc_cluster_w##W(hb200_engine* e, const Params& p, const CcmcLaunch& L);                                 \
    int hb_ccmc_nc_w##W(hb200_engine* e, const Params& p, const CcmcLaunch& L);                                      \
    int hb_ccmc_redistribute_w##W(hb200_engine* e, const Params& p);                                                 \
    int hb_ccmc_find_det_w##W(hb200_engine* e, const Params& p);                                                     \
    int hb_gen_excit_batch_w##W(hb200_engine* e, const Params& p, const uint64_t* d_f, const int64_t* d_p,           \
                                const uint32_t* d_a, long long n, int* d_io, double* d_do, int64_t* d_ns,             \
                                const double* d_rn, int nrn, int* d_nused);
HB_DECL_CCMC(1) HB_DECL_CCMC(2) HB_DECL_CCMC(3) HB_DECL_CCMC(4) HB_DECL_CCMC(32)
