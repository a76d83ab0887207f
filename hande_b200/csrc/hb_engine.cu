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
//   k_spawn_death     (hb_spawn.cuh) fused: decode, initiator flag, projected energy, decide_nattempts, spawning
//                     attempts (load-balanced over a 256-state tile), stochastic death; warp-aggregated append
//   k_ccmc_*          (hb_ccmc.cuh) CCMC cluster selection / spawning / death
//   k_radix_*         LSD radix sort of the spawn list on the bit-string key
//   k_annihilate      segmented sum of equal keys + initiator flag algebra + binary search into the main list
//   k_round_count     stochastic rounding of main-list populations + per-tile survivor counts
//   k_merge           single pass merge of survivors and new determinants into the other main-list buffer
//   k_sc0             <D|H|D> - H00 for new determinants
#include <dlfcn.h>
#include "hb_common.cuh"

thread_local std::string g_err;
// NCCL is resolved lazily with dlopen/dlsym instead of a link-time dependency: a host process that also uses
// torch (bench.py, the multi-GPU tests) must end up with ONE libnccl.so.2, and torch's bundled copy (2.28) is newer
// than the system one (2.27).  Order: a copy already loaded in the process, $HB200_NCCL_LIB, the system library.
struct NcclApi {
    void* h = nullptr;
    bool loaded = false;   // set only when every symbol resolved
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
        if (loaded) return true;
        if (h) { dlclose(h); h = nullptr; }
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) { const char* p = getenv("HB200_NCCL_LIB"); if (p && *p) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define HB_SYM(field, name) field = (decltype(field))dlsym(h, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; dlclose(h); h = nullptr; return false; }
        HB_SYM(GetUniqueId, "ncclGetUniqueId") HB_SYM(CommInitRank, "ncclCommInitRank") HB_SYM(CommDestroy, "ncclCommDestroy")
        HB_SYM(AllGather, "ncclAllGather") HB_SYM(Broadcast, "ncclBroadcast") HB_SYM(Send, "ncclSend") HB_SYM(Recv, "ncclRecv")
        HB_SYM(GroupStart, "ncclGroupStart") HB_SYM(GroupEnd, "ncclGroupEnd") HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
        loaded = true;
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

__device__ __forceinline__ long long sort_chunk(long long n, int nblk) {
    const long long c = (n + nblk - 1) / nblk;
    return (c + SORT_THREADS - 1) / SORT_THREADS * SORT_THREADS;
}

template <int E>
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_hist(const int64_t* __restrict__ in, const unsigned long long* __restrict__ pn, long long cap, int word, int shift,
             unsigned* __restrict__ hist, int nblk) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const long long n = dev_count(pn, cap);
    const long long chunk = sort_chunk(n, nblk);
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
k_radix_scatter(const int64_t* __restrict__ in, int64_t* __restrict__ out, const unsigned long long* __restrict__ pn,
                long long cap, int word, int shift, const unsigned* __restrict__ hist, int nblk) {
    __shared__ unsigned sbase[256];
    __shared__ unsigned swc[SORT_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    sbase[tid] = hist[(size_t)tid * nblk + blockIdx.x];
    const long long n = dev_count(pn, cap);
    const long long chunk = sort_chunk(n, nblk);
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

// pn != nullptr: the element count is min(*pn, n) (read on the device), n the bound the grid was sized from
__global__ void __launch_bounds__(TILE) k_scan_block(const int* __restrict__ in, int* __restrict__ out, long long n,
                                                     int* __restrict__ block_sums, const unsigned long long* __restrict__ pn) {
    __shared__ int swarp[8];
    if (pn) n = dev_count(pn, n);
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
__global__ void __launch_bounds__(TILE) k_scan_add(int* __restrict__ out, long long n, const int* __restrict__ block_offs,
                                                   const unsigned long long* __restrict__ pn) {
    if (pn) n = dev_count(pn, n);
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    const int off = block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += off;
}
// single-block scan for small arrays; also writes the total to *total
__global__ void k_scan_small(const int* __restrict__ in, int* __restrict__ out, long long n, int* total,
                             const unsigned long long* __restrict__ pn) {
    __shared__ int swarp[32];
    __shared__ int carry;
    if (pn) n = dev_count(pn, n);
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
// create_weighted_excitation_list_mol's weights (src/hamiltonian_molecular.f90:348-390) for every orbital pair
__global__ void k_build_ppw(Sys s, double* w) {
    const int nb = s.nbasis;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * nb * nb) return;
    const int cs = t / (nb * nb), r = t % (nb * nb), i = r / nb + 1, a = r % nb + 1;
    w[t] = sqrt(fabs(cs ? two_body(s, i, a, i, a) : two_body(s, i, a, a, i)));
}
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


// All host<->device copies are issued on the engine's own (non-blocking) stream and then synchronised: a plain
// cudaMemcpy runs on the legacy stream, which is NOT ordered with kernels on a cudaStreamNonBlocking stream, and a
// pageable H2D copy may return before its DMA has landed.
static cudaError_t copy_sync(hb200_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
static int ss_relocate(hb200_engine* e, int buf, const int* ntot2, bool check);
static int stage_determ(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, const double* full_host);
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

// determinants host <-> device: the wide layout pads every determinant to 32 words
static cudaError_t copy_states_h2d(hb200_engine* e, uint64_t* dst, const uint64_t* src, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (e->W == e->We) return cudaMemcpyAsync(dst, src, (size_t)n * e->W * 8, cudaMemcpyHostToDevice, st);
    cudaError_t r = cudaMemsetAsync(dst, 0, (size_t)n * e->W * 8, st);
    if (r != cudaSuccess) return r;
    return cudaMemcpy2DAsync(dst, (size_t)e->W * 8, src, (size_t)e->We * 8, (size_t)e->We * 8, (size_t)n, cudaMemcpyHostToDevice, st);
}
static cudaError_t copy_states_d2h(hb200_engine* e, uint64_t* dst, const uint64_t* src, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (e->W == e->We) return cudaMemcpyAsync(dst, src, (size_t)n * e->W * 8, cudaMemcpyDeviceToHost, st);
    return cudaMemcpy2DAsync(dst, (size_t)e->We * 8, src, (size_t)e->W * 8, (size_t)e->We * 8, (size_t)n, cudaMemcpyDeviceToHost, st);
}

// spawn elements [f(1:W), population, flag] host <-> device (wide layout: f padded to 32 words)
static cudaError_t copy_spawn(hb200_engine* e, int64_t* dst, const int64_t* src, long long n, bool to_device, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (e->W == e->We) return cudaMemcpyAsync(dst, src, (size_t)n * e->E * 8, kind, st);
    const size_t hs = (size_t)(e->We + 2) * 8, ds = (size_t)e->E * 8;
    cudaError_t r;
    if (to_device) {
        r = cudaMemsetAsync(dst, 0, (size_t)n * ds, st);
        if (r != cudaSuccess) return r;
        r = cudaMemcpy2DAsync(dst, ds, src, hs, (size_t)e->We * 8, (size_t)n, kind, st);
        if (r != cudaSuccess) return r;
        return cudaMemcpy2DAsync(dst + e->W, ds, src + e->We, hs, 16, (size_t)n, kind, st);
    }
    r = cudaMemcpy2DAsync(dst, hs, src, ds, (size_t)e->We * 8, (size_t)n, kind, st);
    if (r != cudaSuccess) return r;
    return cudaMemcpy2DAsync(dst + e->We, hs, src + e->W, ds, 16, (size_t)n, kind, st);
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
bool hb_uses_heat_bath_tables(const hb200_engine* e) { return uses_heat_bath_tables(e); }

extern "C" {

const char* hb200_last_error(void) { return g_err.c_str(); }

hb200_engine* hb200_create(const hb200_config* cfg) {
    g_err.clear();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_err = "hb200_create: no CUDA device (the engine has no CPU fallback)";
        return nullptr;
    }
    if (cfg->nel > HB_MAXNEL || cfg->nbasis > 64 * HB_MAXW) {
        g_err = "hb200_create: nel/nbasis beyond compiled limits (HB_MAXNEL, HB_MAXW)";
        return nullptr;
    }
    hb200_engine* e = new hb200_engine();
    e->cfg = *cfg;
    // host layout: We = ceil(nbasis/64) words per determinant (particle_t%states).  Device layout: the same for
    // We <= 4 (nbasis <= 254: byte occupied lists); wider bit strings - the plane-wave bases of the UEG - use the wide
    // layout, 32 words per determinant (zero padded) with 16-bit occupied lists and a compressed-key sort.
    e->We = (cfg->nbasis + 63) / 64;
    e->W = (e->We <= 4 && cfg->nbasis <= 254) ? e->We : 32;
    e->E = e->W + 2;
    switch (e->W) {
        case 1: e->ops = hb_list_ops_w1(); break;
        case 2: e->ops = hb_list_ops_w2(); break;
        case 3: e->ops = hb_list_ops_w3(); break;
        case 4: e->ops = hb_list_ops_w4(); break;
        default: e->ops = hb_list_ops_w32(); break;
    }
    if (e->W > 4) {
        int b = 1;
        while ((1 << b) < cfg->nbasis + 1) ++b;
        e->key_bits = b;
        e->key_words = (cfg->nel * b + 63) / 64;
        if (e->key_words > 5) {
            g_err = "hb200_create: wide layout: nel * bits(nbasis) exceeds the 320-bit compressed sort key";
            delete e;
            return nullptr;
        }
    }
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
    p.we = e->We;
    p.cheby_weight = 1.0;
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
    if (dalloc(e, &e->d_spn, 2)) return fail("alloc");
    if (e->W > 4)
        for (int b = 0; b < 2; ++b)
            if (dalloc(e, &e->d_items[b], (size_t)scap * (e->key_words + 1))) return fail("alloc sort keys");
    if (dalloc(e, &e->d_long_q, (size_t)(scap / ANN_SHORT + 2))) return fail("alloc");
    if (dalloc(e, &e->d_long_n, 2)) return fail("alloc");
    if (dalloc(e, &e->d_tile_state, (size_t)e->max_tiles)) return fail("alloc");
    if (dalloc(e, &e->d_ticket, 2)) return fail("alloc");
    if (cudaMallocHost((void**)&e->h_out, sizeof(HostOut) + sizeof(unsigned long long) * p.nprocs) != cudaSuccess) return fail("alloc pinned");
    e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
    e->sp_pn = e->d_spn;
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
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->comm_stream) cudaStreamSynchronize(e->comm_stream);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (size_t r = 0; r < e->peer_base.size(); ++r)
        if (e->peer_base[r] && e->peer_base[r] != (void*)e->d_p2p_block) cudaIpcCloseMemHandle(e->peer_base[r]);
    if (e->h_out) cudaFreeHost(e->h_out);
    for (int i = 0; i < 8; ++i) if (e->ev_chunk[i]) cudaEventDestroy(e->ev_chunk[i]);
    if (e->ev_comm) cudaEventDestroy(e->ev_comm);
    if (e->comm_stream) cudaStreamDestroy(e->comm_stream);
    for (void* q : e->ss.bufs) cudaFree(q);
    for (void* q : e->owned) cudaFree(q);
    for (int i = 0; i < 6; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (int i = 0; i < 2; ++i) if (e->evk[i]) cudaEventDestroy(e->evk[i]);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->copy_done) cudaEventDestroy(e->copy_done);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int hb200_set_system_read_in(hb200_engine* e, const hb200_system_read_in* in) {
    CK(cudaSetDevice(e->cfg.device));
    if (in->nbasis != e->cfg.nbasis || in->nel != e->cfg.nel) FAIL("set_system: nbasis/nel differ from hb200_create");
    if (e->W > 4) FAIL("set_system: read_in systems are limited to 254 spin-orbitals (the wide layout is built for the UEG generators)");
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
    {
        const int eg = e->cfg.excit_gen;
        if (eg == HB200_EXCIT_GEN_POWER_PITZER_OCC || eg == HB200_EXCIT_GEN_POWER_PITZER_OCC_IJ ||
            eg == HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC || eg == HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ) {
            double* w = nullptr;
            if (dalloc(e, &w, (size_t)2 * nb * nb)) return 1;
            k_build_ppw<<<(2 * nb * nb + 255) / 256, 256, 0, e->stream>>>(s, w);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(e->stream));
            s.ppw = w;
        }
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
    if (e->W == e->We) {
        if (dupload(e, &s.ueg_tern, in->ternary_conserve, (size_t)(e->W + 1) * tD * tD * tD)) return 1;
    } else {     // wide layout: (0:We) words per entry -> (0:32), zero padded
        std::vector<uint64_t> t((size_t)(e->W + 1) * tD * tD * tD, 0ull);
        for (size_t k = 0; k < tD * tD * tD; ++k)
            for (int w = 0; w <= e->We; ++w) t[k * (e->W + 1) + w] = in->ternary_conserve[k * (e->We + 1) + w];
        if (dupload(e, &s.ueg_tern, t.data(), t.size())) return 1;
    }
    s.ueg_piL = 3.1415926535897931 * in->box_length;   // pi*cell_param, as evaluated first in coulomb_int_ueg_3d
    s.ueg_kmax = in->kmax; s.ueg_offset = in->offset;
    for (int d = 0; d < 3; ++d) s.ueg_oi[d] = in->offset_inds[d];
    s.ueg_tK = in->tern_kmax; s.ueg_tD = (int)tD;
    e->have_sys = true;
    return 0;
}

// {aliasU, weight, aliasK} of every row entry packed into one 32-byte record (hb_core.cuh HbRec)
__global__ void k_hb_pack(long long n, int nb, const double* __restrict__ U, const double* __restrict__ w, const int* __restrict__ K,
                          const double* __restrict__ tot, HbRec* __restrict__ rec) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    HbRec r;
    const double rt = tot[t / nb];
    r.U = U[t]; r.w = w[t]; r.K = K[t]; r.pad0 = 0;
    r.p = (rt != 0.0) ? w[t] / rt : 0.0;
    rec[t] = r;
}
// single-excitation rows for branch-free occupied-list sums (hb_core.cuh Sys::sc1T)
__global__ void k_build_sc1T(Sys s, int A, D2* T) {
    const long long nb = s.nbasis;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * A * (nb + 1)) return;
    const int j = (int)(t % (nb + 1)), ta = (int)((t / (nb + 1)) % A), i = (int)(t / ((nb + 1) * A)) + 1;
    const int a = s.uhf ? ta + 1 : 2 * ta + 2 - (i & 1);       // RHF: the orbital of spatial index ta with the spin of i
    D2 v; v.x = 0.0; v.y = 0.0;
    if (j != 0 && j != i) {
        v.x = two_body(s, i, j, a, j);
        if (((j ^ i) & 1) == 0) v.y = two_body(s, i, j, j, a);
    }
    T[t] = v;
}

struct HbArrays { double *i_w, *ij_w, *ija_w, *ija_U, *ija_tot, *ijab_w, *ijab_U, *ijab_tot; int *ija_K, *ijab_K; };

static int hb_alloc_arrays(hb200_engine* e, HbArrays& t) {
    const long long nb = e->sys.nbasis, n2 = nb * nb, n3 = n2 * nb, n4 = n3 * nb;
    if (dalloc(e, &t.i_w, nb) || dalloc(e, &t.ij_w, n2) || dalloc(e, &t.ija_w, n3) || dalloc(e, &t.ija_U, n3) ||
        dalloc(e, &t.ija_K, n3) || dalloc(e, &t.ija_tot, n2) || dalloc(e, &t.ijab_w, n4) || dalloc(e, &t.ijab_U, n4) ||
        dalloc(e, &t.ijab_K, n4) || dalloc(e, &t.ijab_tot, n3))
        return 1;
    return 0;
}
// the derived layouts of excit_gen = heat_bath (packed records, branch-free single-excitation rows) and the hand-over
// of the tables to the kernels
static int hb_finish_tables(hb200_engine* e, const HbArrays& t) {
    Sys& s = e->sys;
    cudaStream_t st = e->stream;
    const long long nb = s.nbasis, n2 = nb * nb, n3 = n2 * nb, n4 = n3 * nb;
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_HEAT_BATH_SINGLE) {
        // heat_bath_single evaluates nel x nvirt single-excitation matrix elements per single attempt: the branch-free rows
        const int A = s.uhf ? (int)nb : (int)nb / 2;
        const long long nT = nb * A * (nb + 1);
        D2* T = nullptr;
        if (dalloc(e, &T, (size_t)nT)) return 1;
        k_build_sc1T<<<(unsigned)((nT + 255) / 256), 256, 0, st>>>(s, A, T);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        s.sc1T = T; s.sc1A = A;
    }
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_HEAT_BATH) {
        // The packed hb_ija rows (nb^3 records) and the single-excitation rows sc1T share one allocation.  (Pinning it in
        // L2 with a persisting access-policy window was measured and made the spawning kernels 10-17 % SLOWER on B200 -
        // the carve-out costs the other gathers more than it saves - so no window is set; see DESIGN.md.)
        const int A = s.uhf ? (int)nb : (int)nb / 2;
        const long long nT = nb * A * (nb + 1);
        const size_t bytes_rec = (size_t)n3 * sizeof(HbRec), bytes_T = ((size_t)nT * sizeof(D2) + 255) & ~(size_t)255;
        unsigned char* arena = nullptr;
        if (dalloc(e, &arena, bytes_rec + bytes_T)) return 1;
        HbRec* ija_rec = reinterpret_cast<HbRec*>(arena);
        D2* T = reinterpret_cast<D2*>(arena + bytes_rec);
        HbRec* ijab_rec;
        if (dalloc(e, &ijab_rec, n4)) return 1;
        k_hb_pack<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(n3, (int)nb, t.ija_U, t.ija_w, t.ija_K, t.ija_tot, ija_rec);
        k_hb_pack<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(n4, (int)nb, t.ijab_U, t.ijab_w, t.ijab_K, t.ijab_tot, ijab_rec);
        k_build_sc1T<<<(unsigned)((nT + 255) / 256), 256, 0, st>>>(s, A, T);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        s.hb_ija_rec = ija_rec; s.hb_ijab_rec = ijab_rec; s.sc1T = T; s.sc1A = A;
    }
    s.hb_i_w = t.i_w; s.hb_ij_w = t.ij_w; s.hb_ija_w = t.ija_w; s.hb_ija_U = t.ija_U; s.hb_ija_K = t.ija_K; s.hb_ija_tot = t.ija_tot;
    s.hb_ijab_w = t.ijab_w; s.hb_ijab_U = t.ijab_U; s.hb_ijab_K = t.ijab_K; s.hb_ijab_tot = t.ijab_tot;
    e->have_hb = true;
    return 0;
}

int hb200_build_heat_bath(hb200_engine* e) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("build_heat_bath: system not set");
    Sys& s = e->sys;
    const long long nb = s.nbasis;
    const long long n2 = nb * nb, n3 = n2 * nb, n4 = n3 * nb;
    HbArrays t;
    if (hb_alloc_arrays(e, t)) return 1;
    void* sc = nullptr;
    CK(cudaMalloc(&sc, (size_t)2 * n4 * sizeof(int)));
    int* scratch = (int*)sc;
    cudaStream_t st = e->stream;
    CK(cudaMemsetAsync(t.ija_U, 0, n3 * sizeof(double), st));
    CK(cudaMemsetAsync(t.ija_K, 0, n3 * sizeof(int), st));
    CK(cudaMemsetAsync(t.ijab_U, 0, n4 * sizeof(double), st));
    CK(cudaMemsetAsync(t.ijab_K, 0, n4 * sizeof(int), st));
    k_hb_ijab_w<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(s, t.ijab_w);
    k_hb_ijab_tot<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>((int)nb, t.ijab_w, t.ijab_tot, t.ija_w);
    k_hb_ij<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>((int)nb, t.ijab_w, t.ija_w, t.ija_tot, t.ij_w);
    k_hb_i<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>((int)nb, t.ij_w, t.i_w);
    k_hb_alias<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>((int)nb, n2, t.ija_w, t.ija_tot, t.ija_U, t.ija_K, scratch);
    k_hb_alias<<<(unsigned)((n3 + 127) / 128), 128, 0, st>>>((int)nb, n3, t.ijab_w, t.ijab_tot, t.ijab_U, t.ijab_K, scratch);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(sc));
    return hb_finish_tables(e, t);
}

// excit_gen_heat_bath_t as the HOST built it (init_excit_mol_heat_bath, src/excit_gen_heat_bath_mol.F90:14-256): the
// tables are uploaded as they are, nothing is recomputed.  Arrays in the reference's (column-major) order:
// i_weights(nb), ij_weights(nb,nb), hb_ija%{weights, aliasU, aliasK}(nb,nb,nb), hb_ija%weights_tot(nb,nb),
// hb_ijab%{weights, aliasU, aliasK}(nb,nb,nb,nb), hb_ijab%weights_tot(nb,nb,nb); aliasK 1-based as in the reference.
int hb200_set_excit_tables(hb200_engine* e, const hb200_heat_bath_tables* in) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("set_excit_tables: system not set");
    if (!uses_heat_bath_tables(e)) FAIL("set_excit_tables: the configured excitation generator does not use the heat-bath tables");
    const long long nb = e->sys.nbasis, n2 = nb * nb, n3 = n2 * nb, n4 = n3 * nb;
    HbArrays t;
    if (hb_alloc_arrays(e, t)) return 1;
    CK(copy_sync(e, t.i_w, in->i_weights, (size_t)nb * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ij_w, in->ij_weights, (size_t)n2 * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ija_w, in->ija_weights, (size_t)n3 * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ija_U, in->ija_aliasU, (size_t)n3 * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ija_K, in->ija_aliasK, (size_t)n3 * 4, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ija_tot, in->ija_weights_tot, (size_t)n2 * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ijab_w, in->ijab_weights, (size_t)n4 * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ijab_U, in->ijab_aliasU, (size_t)n4 * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ijab_K, in->ijab_aliasK, (size_t)n4 * 4, cudaMemcpyHostToDevice));
    CK(copy_sync(e, t.ijab_tot, in->ijab_weights_tot, (size_t)n3 * 8, cudaMemcpyHostToDevice));
    return hb_finish_tables(e, t);
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
    CK(cudaSetDevice(e->cfg.device));
    for (int k = 0; k < HB_MAXW; ++k) e->par.f0[k] = (k < e->We) ? f0[k] : 0;
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
        CK(copy_states_h2d(e, e->d_states[c], states, n, st));
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
    if (e->ss.on && ss_relocate(e, e->cur, nullptr, true)) return 1;   // a new list must still hold the deterministic states
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
        CK(copy_states_h2d(e, e->d_states[g], states, n, e->copy_stream));
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
    if (e->ss.on && ss_relocate(e, e->cur, nullptr, true)) return 1;   // as hb200_upload_psips: the space follows the new list
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
        CK(copy_states_d2h(e, states, e->d_states[c], n, st));
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
// launchers of hb_ccmc_tu.cu (one object file per W)
#define DISPATCH_CCMC(e, rc, fn, ...)               \
    switch ((e)->W) {                               \
        case 1: rc = fn##1(__VA_ARGS__); break;     \
        case 2: rc = fn##2(__VA_ARGS__); break;     \
        case 3: rc = fn##3(__VA_ARGS__); break;     \
        case 4: rc = fn##4(__VA_ARGS__); break;     \
        default: rc = fn##32(__VA_ARGS__); break;   \
    }

// ---- peer-to-peer exchange kernels (hb200_iterate with nprocs > 1 after hb200_p2p_import) ---------------------------
// Reserve room for this chunk's elements in every destination rank's receive buffer: one system-scope atomic per
// destination on the REMOTE head counter (NVLink atomics), replacing the count exchange of MPI_Alltoall
// (src/spawn_data.F90:693).  snap_lo / snap_hi: d_head before / after the spawn chunk.
__global__ void k_push_reserve(const unsigned long long* __restrict__ snap_lo, const unsigned long long* __restrict__ snap_hi,
                               unsigned long long* const* __restrict__ peer_head, int parity, int np, long long block_size,
                               long long cap, long long* __restrict__ push, int* __restrict__ err) {
    const int d = threadIdx.x;
    if (d >= np) return;
    const long long lo = min((long long)snap_lo[d], block_size), hi = min((long long)snap_hi[d], block_size);
    long long cnt = hi - lo, off = 0;
    if (cnt > 0) {
        off = (long long)atomicAdd_system(peer_head[d] + parity, (unsigned long long)cnt);
        if (off + cnt > cap) {          // the destination's receive buffer is full: spawn%error
            atomicOr(err, 1);
            cnt = max(0ll, cap - off);
        }
    }
    push[d] = lo; push[np + d] = cnt; push[2 * np + d] = off;
}
// Copy the chunk's elements of every per-destination block straight into the destination rank's receive buffer
// (peer stores over NVLink; blockIdx.y = destination).  Replaces MPI_Alltoallv (src/spawn_data.F90:721).
__global__ void __launch_bounds__(256) k_push_copy(const int64_t* __restrict__ blocks, long long block_size, int E, int np,
                                                   const long long* __restrict__ push, int64_t* const* __restrict__ peer_recv,
                                                   int parity) {
    const int d = blockIdx.y;
    const long long lo = push[d], cnt = push[np + d], off = push[2 * np + d];
    if (cnt <= 0) return;
    const int64_t* src = blocks + ((long long)d * block_size + lo) * E;
    int64_t* dst = peer_recv[parity * np + d] + off * E;
    const long long nw = cnt * E;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if ((E & 1) == 0) {       // 16-byte elements of an even-E list are 16-byte aligned on both sides
        const longlong2* s2 = reinterpret_cast<const longlong2*>(src);
        longlong2* d2 = reinterpret_cast<longlong2*>(dst);
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nw / 2; k += stride) d2[k] = s2[k];
    } else {
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nw; k += stride) dst[k] = src[k];
    }
    __threadfence_system();
}
// insert_new_walkers capacity check (src/annihilation.f90:750-771) on the device: tot[0] = new determinants,
// tot[1] = surviving states
__global__ void k_cap_check(int* __restrict__ tot, long long walker_length, int* __restrict__ err) {
    if ((long long)tot[0] + (long long)tot[1] > walker_length) { err[1] = 1; tot[0] = 0; }
}

static int set_count_host(hb200_engine* e, long long n) {     // staged calls: the host knows the element count
    e->sp_n = n;
    unsigned long long v = (unsigned long long)n;
    CK(cudaMemcpyAsync(e->d_spn, &v, sizeof(v), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));       // v is a stack variable
    e->sp_pn = e->d_spn;
    e->sp_cap = std::max<long long>(n, 0);
    return 0;
}

static int spawn_dispatch(hb200_engine* e, int tile0, int ntiles, long long n) {
    Params& p = e->par;
    // (generator group, GEN) of this calculation -> the object file holding its k_spawn_death instantiation
    int group, gen;
    if (e->sys.kind == SYS_UEG) {
        group = SPAWN_GROUP_TABLES;
        gen = (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER) ? (int)GEN_UEG_PP : (int)GEN_UEG;
    } else switch (e->cfg.excit_gen) {
        case HB200_EXCIT_GEN_NO_RENORM: case HB200_EXCIT_GEN_RENORM: case HB200_EXCIT_GEN_RENORM_SPIN:
        case HB200_EXCIT_GEN_NO_RENORM_SPIN: group = SPAWN_GROUP_UNIFORM; gen = e->cfg.excit_gen; break;
        case HB200_EXCIT_GEN_POWER_PITZER_ORDERN: case HB200_EXCIT_GEN_POWER_PITZER:
            group = SPAWN_GROUP_TABLES; gen = e->cfg.excit_gen; break;
        case HB200_EXCIT_GEN_HEAT_BATH: group = SPAWN_GROUP_HEAT_BATH; gen = EXCIT_GEN_HEAT_BATH; break;
        case HB200_EXCIT_GEN_HEAT_BATH_UNIFORM: case HB200_EXCIT_GEN_HEAT_BATH_SINGLE:
            group = SPAWN_GROUP_HB_UNIFORM; gen = e->cfg.excit_gen; break;
        case HB200_EXCIT_GEN_POWER_PITZER_OCC:
        case HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC: group = SPAWN_GROUP_PP_OCC; gen = EXCIT_GEN_POWER_PITZER_OCC; break;
        case HB200_EXCIT_GEN_POWER_PITZER_OCC_IJ:
        case HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ: group = SPAWN_GROUP_PP_OCC; gen = EXCIT_GEN_POWER_PITZER_OCC_IJ; break;
        default: FAIL("spawn_death: excitation generator not implemented");
    }
#define HB_ROW(W) {hb_spawn_w##W##_g0, hb_spawn_w##W##_g1, hb_spawn_w##W##_g2, hb_spawn_w##W##_g3, hb_spawn_w##W##_g4}
    // wide layout (W = 32): only the generators that need no nbasis^3 / nbasis^4 tables are built (the UEG's)
    static const hb_spawn_fn table[5][SPAWN_NGROUPS] = {HB_ROW(1), HB_ROW(2), HB_ROW(3), HB_ROW(4),
                                                        {nullptr, nullptr, nullptr, nullptr, hb_spawn_w32_g4}};
#undef HB_ROW
    SpawnLaunch L;
    L.gen = gen; L.ntiles = ntiles; L.smem = 0; L.n = n; L.tile0 = tile0;
    const hb_spawn_fn fn = table[e->W <= 4 ? e->W - 1 : 4][group];
    if (!fn) FAIL("spawn_death: this excitation generator is not built for the wide layout (nbasis > 254)");
    if (fn(e, p, L)) return 1;
    CK(cudaGetLastError());
    e->launches++; e->spawn_launches++;
    return 0;
}

// The spawning step.  overlap = false: one launch (staged calls, single rank).  overlap = true (peer-to-peer exchange
// set up): the tile range is launched in chunks; as soon as a chunk has finished, its part of every per-destination
// block is pushed into the destination rank's receive buffer on the communication stream while the next chunk spawns.
static int stage_spawn_launch(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, bool overlap) {
    Params& p = e->par;
    p.tau = in->tau; p.shift = in->shift; p.proj_energy_old = in->proj_energy_old; p.cycle = cycle;
    cudaStream_t st = e->stream;
    const int np = p.nprocs;
    CK(cudaMemsetAsync(e->d_head, 0, sizeof(unsigned long long) * np, st));
    const long long n = e->nstates;
    const int ntiles = (int)((n + TILE - 1) / TILE);
    CK(cudaEventRecord(e->evk[0], st));
    if (!overlap) {
        if (ntiles > 0 && spawn_dispatch(e, 0, ntiles, n)) return 1;
    } else {
        const int par = e->xparity;
        static const int min_tiles = getenv("HB200_P2P_MIN_TILES") ? atoi(getenv("HB200_P2P_MIN_TILES")) : 4096;
        const int nchunk = (ntiles >= min_tiles) ? 4 : 1;
        CK(cudaMemsetAsync(e->d_snap, 0, sizeof(unsigned long long) * np, st));
        for (int c = 0; c < nchunk; ++c) {
            const int t0 = (int)((long long)ntiles * c / nchunk), t1 = (int)((long long)ntiles * (c + 1) / nchunk);
            if (t1 > t0 && spawn_dispatch(e, t0, t1 - t0, n)) return 1;
            CK(cudaMemcpyAsync(e->d_snap + (size_t)(c + 1) * np, e->d_head, sizeof(unsigned long long) * np,
                               cudaMemcpyDeviceToDevice, st));
            CK(cudaEventRecord(e->ev_chunk[c], st));
            CK(cudaStreamWaitEvent(e->comm_stream, e->ev_chunk[c], 0));
            k_push_reserve<<<1, std::max(32, np), 0, e->comm_stream>>>(e->d_snap + (size_t)c * np, e->d_snap + (size_t)(c + 1) * np,
                                                                        e->d_peer_head, par, np, e->block_size,
                                                                        e->cfg.spawned_walker_length, e->d_push, e->d_err);
            k_push_copy<<<dim3(32, np), 256, 0, e->comm_stream>>>(e->d_spawn[0], e->block_size, e->E, np, e->d_push,
                                                                   e->d_peer_recv, par);
            CK(cudaGetLastError());
            e->launches += 2;
        }
        if (e->host_barrier) {
            // the host's own barrier (MPI_Barrier in a Fortran host without NCCL): this rank's pushes have landed when
            // its communication stream has drained, everybody's after the barrier
            CK(cudaStreamSynchronize(e->comm_stream));
            e->host_barrier(e->host_barrier_arg);
        } else {
            // every rank's pushes have landed once this collective completes (it also hands every rank the count matrix)
            CK(cudaMemcpyAsync(e->d_counts + (size_t)p.iproc * np, e->d_head, sizeof(long long) * np, cudaMemcpyDeviceToDevice,
                               e->comm_stream));
            NCK(g_nccl.AllGather(e->d_counts + (size_t)p.iproc * np, e->d_counts, np, ncclInt64, e->comm, e->comm_stream));
        }
        CK(cudaEventRecord(e->ev_comm, e->comm_stream));
    }
    CK(cudaEventRecord(e->evk[1], st));
    k_reduce_partials<<<1, 1024, 0, st>>>(e->d_partials, ntiles > 0 ? e->npartials : 0, e->d_stats);
    CK(cudaGetLastError());
    e->launches++;
    if (p.ps_part && ntiles > 0) {
        k_reduce_ps<<<1, 1024, 0, st>>>(e->d_ps_part, ntiles, e->d_ps_acc);
        CK(cudaGetLastError());
        e->launches++;
    }
    return 0;
}

// staged call: spawning step + the host reads the counts
static int stage_spawn_death(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, CycleStats* hst) {
    Params& p = e->par;
    cudaStream_t st = e->stream;
    if (stage_spawn_launch(e, in, cycle, false)) return 1;
    CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * p.nprocs, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hst, e->d_stats, sizeof(CycleStats), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
        float t = 0.f;
        cudaEventElapsedTime(&t, e->evk[0], e->evk[1]);
        e->spawn_kernel_ms += t;
    }
    for (int d = 0; d < p.nprocs; ++d)
        if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;  // overflow: dropped
    e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
    e->sp_cur = 0;
    e->sp_blocked = true;
    e->sp_n = 0;
    if (p.nprocs == 1) {
        e->sp_blocked = false;
        if (set_count_host(e, (long long)e->h_head[0])) return 1;
    }
    return 0;
}

static int stage_comm(hb200_engine* e) {
    Params& p = e->par;
    if (p.nprocs == 1) {
        e->sp_blocked = false; e->sp_cur = 0;
        e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
        return set_count_host(e, (long long)e->h_head[0]);
    }
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
    // MPI_Alltoallv (src/spawn_data.F90:721): receive blocks ordered by source rank.  Every rank evaluates the same
    // capacity checks on the same count matrix BEFORE any send/recv is posted, so a failure is collective.
    long long off = 0;
    for (int r = 0; r < np; ++r) off += counts[(size_t)r * np + me];
    bool too_many = false;
    for (int d = 0; d < np; ++d) {
        long long t = 0;
        for (int r = 0; r < np; ++r) t += counts[(size_t)r * np + d];
        if (t > e->cfg.spawned_walker_length) too_many = true;
    }
    if (too_many) FAIL("comm_spawn: a rank would receive more than spawned_walker_length");
    off = 0;
    int rc = 0;
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < np && !rc; ++r) {
        const long long nsend = counts[(size_t)me * np + r], nrecv = counts[(size_t)r * np + me];
        if (r == me) {
            if (nsend && cudaMemcpyAsync(e->d_spawn[1] + off * E, e->d_spawn[0] + (long long)r * e->block_size * E,
                                         (size_t)nsend * E * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = 1;
        } else {
            if (nsend && g_nccl.Send(e->d_spawn[0] + (long long)r * e->block_size * E, (size_t)nsend * E, ncclInt64, r, e->comm, st) != ncclSuccess) rc = 1;
            if (nrecv && g_nccl.Recv(e->d_spawn[1] + off * E, (size_t)nrecv * E, ncclInt64, r, e->comm, st) != ncclSuccess) rc = 1;
        }
        off += nrecv;
    }
    NCK(g_nccl.GroupEnd());       // always reached: the group is never left open
    if (rc) FAIL("comm_spawn: posting the send/recv of the spawn blocks failed");
    e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
    e->sp_cur = 1;
    e->sp_blocked = false;
    return set_count_host(e, off);
}

static int device_scan(hb200_engine* e, const int* d_in, int* d_out, long long n, int slot, const unsigned long long* pn);

// LSD radix sort of the current spawn list.  bound: host-side upper bound of the element count (the grids are sized from
// it); the count itself is read by the kernels from e->sp_pn.
static int stage_sort(hb200_engine* e, long long bound) {
    if (bound <= 1) return 0;
    cudaStream_t st = e->stream;
    const int nblk = (int)std::max<long long>(1, std::min<long long>(1184, (bound + 2047) / 2048));
    if (256ll * nblk > e->hist_cap) FAIL("sort: histogram scratch too small");
    const unsigned long long* pn = e->sp_pn;
    const long long cap = e->sp_cap;
    const bool wide = e->W > 4;
    // wide layout: sort compressed keys (hb_list.cuh: the occupied orbitals packed key_bits apiece, nel * key_bits bits
    // instead of nbasis) with the element index as payload, then move the elements once
    const int E = wide ? e->key_words + 1 : e->E;
    const int npass = wide ? (e->cfg.nel * e->key_bits + 7) / 8 : (e->cfg.nbasis + 7) / 8;
    int icur = 0;
    if (wide) {
        if (e->ops->compress(e, e->sp_ptr[e->sp_cur], bound, e->key_bits, e->key_words, e->d_items[0])) return 1;
        e->launches++;
    }
    for (int ps = 0; ps < npass; ++ps) {
        const int word = (8 * ps) / 64, shift = (8 * ps) % 64;
        const int64_t* src = wide ? e->d_items[icur] : e->sp_ptr[e->sp_cur];
        int64_t* dst = wide ? e->d_items[icur ^ 1] : e->sp_ptr[e->sp_cur ^ 1];
        switch (E) {
            case 3: k_radix_hist<3><<<nblk, SORT_THREADS, 0, st>>>(src, pn, cap, word, shift, e->d_hist, nblk); break;
            case 4: k_radix_hist<4><<<nblk, SORT_THREADS, 0, st>>>(src, pn, cap, word, shift, e->d_hist, nblk); break;
            case 5: k_radix_hist<5><<<nblk, SORT_THREADS, 0, st>>>(src, pn, cap, word, shift, e->d_hist, nblk); break;
            default: k_radix_hist<6><<<nblk, SORT_THREADS, 0, st>>>(src, pn, cap, word, shift, e->d_hist, nblk); break;
        }
        if (256ll * nblk <= 8192) {
            k_scan_u32_single<<<1, 1024, 0, st>>>(e->d_hist, 256ll * nblk);
        } else {
            // multi-block exclusive scan (counts < 2^31, so the int scan is bit-identical); in place
            if (device_scan(e, (const int*)e->d_hist, (int*)e->d_hist, 256ll * nblk, 2, nullptr)) return 1;
            e->launches += 2;
        }
        switch (E) {
            case 3: k_radix_scatter<3><<<nblk, SORT_THREADS, 0, st>>>(src, dst, pn, cap, word, shift, e->d_hist, nblk); break;
            case 4: k_radix_scatter<4><<<nblk, SORT_THREADS, 0, st>>>(src, dst, pn, cap, word, shift, e->d_hist, nblk); break;
            case 5: k_radix_scatter<5><<<nblk, SORT_THREADS, 0, st>>>(src, dst, pn, cap, word, shift, e->d_hist, nblk); break;
            default: k_radix_scatter<6><<<nblk, SORT_THREADS, 0, st>>>(src, dst, pn, cap, word, shift, e->d_hist, nblk); break;
        }
        CK(cudaGetLastError());
        e->launches += 3;
        if (wide) icur ^= 1; else e->sp_cur ^= 1;
    }
    if (wide) {
        if (e->ops->gather(e, e->sp_ptr[e->sp_cur], e->key_words, e->d_items[icur], e->sp_ptr[e->sp_cur ^ 1])) return 1;
        e->launches++;
        e->sp_cur ^= 1;
    }
    return 0;
}

// exclusive scan of n ints (d_in -> d_out), total to d_total[slot]; pn != nullptr: the count is min(*pn, n) on the device
static int device_scan(hb200_engine* e, const int* d_in, int* d_out, long long n, int slot, const unsigned long long* pn) {
    cudaStream_t st = e->stream;
    if (n <= 0) { CK(cudaMemsetAsync(e->d_total + slot, 0, sizeof(int), st)); return 0; }
    if (n <= 4 * SCAN_BLOCK) {
        k_scan_small<<<1, 1024, 0, st>>>(d_in, d_out, n, e->d_total + slot, pn);
        e->launches++;
    } else {
        const long long nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
        k_scan_block<<<(unsigned)nb, TILE, 0, st>>>(d_in, d_out, n, e->d_scan_l1, pn);
        k_scan_small<<<1, 1024, 0, st>>>(e->d_scan_l1, e->d_scan_l1o, nb, e->d_total + slot, nullptr);
        k_scan_add<<<(unsigned)nb, TILE, 0, st>>>(d_out, n, e->d_scan_l1o, pn);
        e->launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}

// annihilate_main_list + remove_unoccupied_dets + insert_new_walkers, all counts kept on the device: nothing is read
// back here.  bound: host-side upper bound of the spawn-list length.
static int stage_annihilate_launch(hb200_engine* e, uint32_t cycle, long long bound) {
    Params& p = e->par;
    p.cycle = cycle;
    cudaStream_t st = e->stream;
    const long long ns = e->nstates;
    int64_t* sp = e->sp_ptr[e->sp_cur];
    int64_t* ins = e->sp_ptr[e->sp_cur ^ 1];
    const unsigned long long* pn = e->sp_pn;
    if (bound > 0) {
        CK(cudaMemsetAsync(e->d_long_n, 0, sizeof(unsigned), st));
        if (e->ops->annihilate(e, p, sp, bound)) return 1;
        e->launches += 2;
        if (device_scan(e, e->d_ins_flag, e->d_ins_idx, bound, 0, pn)) return 1;
        if (e->ops->compact(e, sp, bound, ins)) return 1;
        e->launches++;
    } else {
        CK(cudaMemsetAsync(e->d_total, 0, sizeof(int), st));
    }
    const int ntiles = std::max<int>(1, (int)((ns + TILE - 1) / TILE));
    // Wide layout, when the list cannot overflow (survivors <= states, new determinants <= spawn-list bound): the
    // rounding, the survivor counts and the merge are ONE pass over the 272-byte states (k_merge<W, true>, decoupled
    // look-back over the tiles).  With 32-byte states the look-back chain over 4e5 tiles costs more than the second pass
    // it saves (measured at 1e8 walkers: 6.8 ms against 2.9 ms), so the counts come first there - as they must whenever
    // insert_new_walkers' capacity check (src/annihilation.f90:750-771) may have to drop the new determinants.
    const bool fused = (e->W > 4 || getenv("HB200_FUSED_MERGE")) && ns + std::min<long long>(bound, e->sp_cap) <= e->cfg.walker_length;
    if (!fused) {
        if (e->ops->round_count(e, p, ntiles)) return 1;
        e->launches++;
        if (device_scan(e, e->d_tile_keep, e->d_tile_off, ntiles, 1, nullptr)) return 1;
        k_cap_check<<<1, 1, 0, st>>>(e->d_total, e->cfg.walker_length, e->d_err);
    }
    if (bound > 0) {
        if (e->ops->sc0(e, p.H00, (const uint64_t*)ins, e->E, bound, e->d_ins_dat, e->d_total)) return 1;
        e->launches++;
    }
    if (e->ops->merge(e, p, ins, ntiles, fused)) return 1;
    k_reduce_ll<<<1, 1024, 0, st>>>(e->d_part_ll, ntiles, e->d_ll);
    CK(cudaGetLastError());
    e->launches += 3;
    return 0;
}

// host bookkeeping once the counts of the merge are known
static void finish_merge(hb200_engine* e, CycleStats* hst, long long nins, long long nkept, long long npart) {
    const int c = e->cur, o = e->alt;
    e->cur = o;
    e->alt = c;
    e->nstates = nkept + nins;
    e->nparticles_enc = npart;
    hst->nkept = nkept;
    hst->npart_new = npart;
    e->sp_n = 0;
}

// staged call: annihilation + merge, then the host reads the counts
static int stage_annihilate_main(hb200_engine* e, uint32_t cycle, CycleStats* hst) {
    cudaStream_t st = e->stream;
    if (stage_annihilate_launch(e, cycle, e->sp_n)) return 1;
    int h_tot[2] = {0, 0};
    long long npart = 0;
    CK(cudaMemcpyAsync(h_tot, e->d_total, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&npart, e->d_ll, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    finish_merge(e, hst, h_tot[0], h_tot[1], npart);
    if (e->ss.on && ss_relocate(e, e->cur, nullptr, true)) return 1;   // determ%indices / determ%flags follow the new list
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
    if (e->par.cheby_weight != 1.0) FAIL("ccmc_spawn: the wall-Chebyshev propagator is only implemented for FCIQMC");
    if (e->ss.on) FAIL("ccmc_spawn: the semi-stochastic projection is an FCIQMC feature (clear the deterministic space)");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("ccmc_spawn: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("ccmc_spawn: power_pitzer tables not built");
    if (e->par.nprocs > 1 && !e->comm) FAIL("ccmc_spawn: nprocs > 1 but hb200_comm_init was not called");
    if (e->cfg.initiator_approx) FAIL("ccmc_spawn: the initiator approximation is not implemented for CCMC");
    if (std::min(e->sys.nel, ex_level + 2) > HB_MAX_CLUSTER)
        FAIL("ccmc_spawn: clusters of more than HB_MAX_CLUSTER (8) excitors are not supported (ex_level + 2 <= 8)");
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
        slot = e->ops->owner_slot_shift(e, p);
        D0_proc = map[slot];
    }
    long long d0info[2] = {0, 0};
    const bool have_D0 = (p.iproc == D0_proc);
    if (have_D0) {
        if (n > 0) {
            { int rc = 0; DISPATCH_CCMC(e, rc, hb_ccmc_find_det_w, e, p); if (rc) return 1; }
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
        {
            CcmcLaunch L; L.a = a; L.nblk = nblk; L.partials = e->d_cc_part; L.cum = e->d_cum;
            int rc = 0; DISPATCH_CCMC(e, rc, hb_ccmc_cluster_w, e, p, L); if (rc) return 1;
        }
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
        {
            CcmcLaunch L; L.a = a; L.nblk = nblk; L.partials = e->d_cc_part + part_cap; L.cum = e->d_cum;
            int rc = 0; DISPATCH_CCMC(e, rc, hb_ccmc_nc_w, e, pn, L); if (rc) return 1;
        }
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
        { int rc = 0; DISPATCH_CCMC(e, rc, hb_ccmc_redistribute_w, e, p); if (rc) return 1; }
        e->launches++;
        CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * p.nprocs, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(herr, e->d_err, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int d = 0; d < p.nprocs; ++d)
            if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;
    }
    e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
    e->sp_cur = 0;
    e->sp_blocked = p.nprocs > 1;
    if (set_count_host(e, (p.nprocs == 1) ? (long long)e->h_head[0] : 0)) return 1;
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
    CK(cudaSetDevice(e->cfg.device));
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

// Wall-Chebyshev propagator (src/propagators.f90:11-208): the host owns the spectral range, the zeroes and the weights
// (init_chebyshev / update_chebyshev) and loops over the `order` sub-cycles of an MC cycle (src/fciqmc.f90:298-299); the
// engine applies the weight of the current sub-cycle in attempt_to_spawn and stochastic_death.  1.0 = linear projector.
int hb200_set_propagator_weight(hb200_engine* e, double weight) {
    if (e->cfg.excit_gen < 0) FAIL("set_propagator_weight: engine not configured");
    e->par.cheby_weight = weight;
    return 0;
}

// qmc_in%pattempt_update (src/qmc.F90:1049-1060, src/spawning.F90:2139-2372): the engine holds pattempt_single /
// pattempt_double and, while `accumulate` is set, sums |H_ij| pattempt / pgen and the counts of the allowed single and
// double excitations it generates; the host reads the sums once per report loop, allreduces them and sets the new
// probabilities (update_pattempt_single).
int hb200_set_pattempt(hb200_engine* e, double pattempt_single, double pattempt_double, int32_t accumulate) {
    CK(cudaSetDevice(e->cfg.device));
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
    CK(cudaSetDevice(e->cfg.device));
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
        if (stage_sort(e, e->sp_n)) return 1;
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

// initialise_slot_pop (src/load_balancing.F90:624-654): this rank's population in each of the nprocs * nslots
// load-balancing slots; the host sums the ranks (MPI_AllReduce) and runs the policy (do_load_balancing).
int hb200_slot_populations(hb200_engine* e, double* slot_pop, int32_t n) {
    CK(cudaSetDevice(e->cfg.device));
    const int ns = e->par.nprocs * e->par.nslots;
    if (n != ns) FAIL("slot_populations: wrong length");
    unsigned long long* d = nullptr;
    CK(cudaMalloc((void**)&d, sizeof(unsigned long long) * ns));
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * ns, e->stream));
    const long long m = e->nstates;
    if (m > 0) {
        if (e->ops->slot_pop(e, d)) return 1;
    }
    std::vector<unsigned long long> h(ns);
    CK(copy_sync(e, h.data(), d, sizeof(unsigned long long) * ns, cudaMemcpyDeviceToHost));
    cudaFree(d);
    for (int k = 0; k < ns; ++k) slot_pop[k] = (double)h[k] / (double)e->par.real_factor;
    return 0;
}

// redistribute_particles (src/qmc_common.F90:505-595) after the host changed proc_map (hb200_set_proc_map): every
// determinant whose owner is now another rank goes, with its whole population and flag 0, into that rank's block of
// the spawn list and is zeroed here.  Leaves the engine where hb200_spawn_death leaves it: the host continues with
// hb200_comm_spawn / hb200_annihilate_spawn / hb200_annihilate_main (direct_annihilation in
// redistribute_load_balancing_dets, src/qmc_common.F90:1332-1390).  nsent: population that left (real units).
int hb200_redistribute_particles(hb200_engine* e, double* nsent) {
    CK(cudaSetDevice(e->cfg.device));
    // redistribute_load_balancing_dets annihilates the moved determinants without the deterministic flags and
    // redistribute_semi_stoch_t (src/qmc_common.F90:597-650) then rebuilds the space under the new proc_map: the host
    // switches the projection off (hb200_set_determ_space with all sizes zero), redistributes, and sets the space again
    // from determ%dets (hande_b200/fciqmc.py does)
    if (e->ss.on) FAIL("redistribute_particles: a deterministic space is set (clear it first, set it again afterwards)");
    Params p = e->par;
    p.ccmc_shift = 0; p.ccmc_freq = 0;
    cudaStream_t st = e->stream;
    CK(cudaMemsetAsync(e->d_head, 0, sizeof(unsigned long long) * p.nprocs, st));
    const long long before = e->nparticles_enc;
    long long after = before;
    if (e->nstates > 0 && p.nprocs > 1) {
        { int rc = 0; DISPATCH_CCMC(e, rc, hb_ccmc_redistribute_w, e, p); if (rc) return 1; }
        const int c = e->cur;
        const int nb = (int)std::min<long long>(1184, (e->nstates + TILE - 1) / TILE);
        k_abs_sum<<<nb, TILE, 0, st>>>(e->d_pops[c], e->nstates, e->d_part_ll);
        k_reduce_ll<<<1, 1024, 0, st>>>(e->d_part_ll, nb, e->d_ll);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&after, e->d_ll, sizeof(long long), cudaMemcpyDeviceToHost, st));
        e->launches += 3;
    }
    CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * p.nprocs, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int d = 0; d < p.nprocs; ++d)
        if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;
    e->nparticles_enc = after;
    if (nsent) *nsent = (double)(before - after) / (double)p.real_factor;
    e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
    e->sp_cur = 0;
    e->sp_blocked = p.nprocs > 1;
    e->sp_n = 0;
    if (p.nprocs == 1 && set_count_host(e, 0)) return 1;
    return 0;
}

int hb200_annihilate_spawn(hb200_engine* e) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->sp_blocked) FAIL("annihilate_spawn: call hb200_comm_spawn first");
    if (stage_sort(e, e->sp_n)) return 1;
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
    const int E = e->E, Eh = e->We + 2;      // element length on the device / in the host's layout
    if (e->sp_blocked) {
        // still partitioned by destination: concatenate the blocks
        long long tot = 0;
        for (int d = 0; d < e->par.nprocs; ++d) tot += (long long)e->h_head[d];
        *n = tot;
        if (tot > capacity) FAIL("download_spawn: capacity too small");
        long long off = 0;
        for (int d = 0; d < e->par.nprocs; ++d) {
            const long long c = (long long)e->h_head[d];
            if (c) CK(copy_spawn(e, sdata + off * Eh, e->d_spawn[0] + (long long)d * e->block_size * E, c, false, e->stream));
            off += c;
        }
        CK(cudaStreamSynchronize(e->stream));
        return 0;
    }
    *n = e->sp_n;
    if (e->sp_n > capacity) FAIL("download_spawn: capacity too small");
    if (e->sp_n) CK(copy_spawn(e, sdata, e->sp_ptr[e->sp_cur], e->sp_n, false, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int hb200_spawn_counts(hb200_engine* e, int64_t* counts, int32_t nprocs) {
    if (nprocs != e->par.nprocs) FAIL("spawn_counts: nprocs differs from hb200_create");
    if (!e->sp_blocked && e->par.nprocs > 1) FAIL("spawn_counts: the spawn list is no longer partitioned by destination");
    for (int d = 0; d < nprocs; ++d) counts[d] = (int64_t)e->h_head[d];
    return 0;
}

int hb200_upload_spawn(hb200_engine* e, const int64_t* sdata, int64_t n) {
    CK(cudaSetDevice(e->cfg.device));
    if (n > e->cfg.spawned_walker_length) FAIL("upload_spawn: too many elements");
    if (n) CK(copy_spawn(e, e->d_spawn[0], sdata, n, true, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
    e->sp_cur = 0; e->sp_blocked = false;
    return set_count_host(e, n);
}

// ncycles MC cycles with ONE host synchronisation per cycle: every count a later stage needs (spawn-list length, new
// determinants, surviving states) stays on the device, grids are sized from host-side upper bounds, and the results
// of the cycle (estimators, counts, error flags) arrive in one pinned block after the merge.  With the peer-to-peer
// exchange set up (hb200_p2p_import) the spawn blocks travel to their owner ranks while the spawning step is still
// running (stage_spawn_launch); otherwise comm_spawn_t's NCCL send/recv path (stage_comm) is used.
int hb200_iterate(hb200_engine* e, int32_t ncycles, const hb200_iter_in* in, hb200_iter_out* out) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("iterate: system not set");
    if (uses_heat_bath_tables(e) && !e->have_hb) FAIL("iterate: heat-bath tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("iterate: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("iterate: power_pitzer tables not built");
    const int np = e->par.nprocs;
    const bool overlap = np > 1 && e->p2p && (e->comm || e->host_barrier);
    if (np > 1 && !e->comm && !overlap) FAIL("iterate: nprocs > 1 but hb200_comm_init was not called");
    memset(out, 0, sizeof(*out));
    cudaStream_t st = e->stream;
    float acc[4] = {0, 0, 0, 0};
    e->spawn_kernel_ms = 0.f;
    CycleStats cs;
    memset(&cs, 0, sizeof(cs));
    long long nattempts = 0;
    HostOut* ho = reinterpret_cast<HostOut*>(e->h_out);
    CK(cudaEventRecord(e->ev[5], st));
    for (int c = 0; c < ncycles; ++c) {
        const uint32_t cycle = in->first_cycle + (uint32_t)c;
        // init_mc_cycle (src/qmc_common.F90:950-1017)
        const double npart_real = (double)e->nparticles_enc / (double)e->par.real_factor;
        nattempts = llround(2.0 * npart_real);
        out->walker_iterations += npart_real;
        CK(cudaEventRecord(e->ev[0], st));
        if (stage_spawn_launch(e, in, cycle, overlap)) return 1;
        CK(cudaEventRecord(e->ev[1], st));
        long long bound;
        if (np == 1) {
            // one spawn-list element per successful attempt; attempts per state <= |population| + 1
            bound = std::min<long long>(e->block_size, (long long)ceil(npart_real) + e->nstates + 1);
            e->sp_ptr[0] = e->d_spawn[0]; e->sp_ptr[1] = e->d_spawn[1];
            e->sp_cur = 0; e->sp_blocked = false;
            e->sp_pn = e->d_head; e->sp_cap = e->block_size;
        } else if (overlap) {
            const int par = e->xparity;
            CK(cudaStreamWaitEvent(st, e->ev_comm, 0));
            bound = e->cfg.spawned_walker_length;
            e->sp_ptr[0] = e->d_recv[par]; e->sp_ptr[1] = e->d_spawn[0];
            e->sp_cur = 0; e->sp_blocked = false;
            e->sp_pn = e->d_recv_head + par; e->sp_cap = e->cfg.spawned_walker_length;
        } else {
            CK(cudaMemcpyAsync(e->h_head.data(), e->d_head, sizeof(unsigned long long) * np, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (int d = 0; d < np; ++d)
                if ((long long)e->h_head[d] > e->block_size) e->h_head[d] = (unsigned long long)e->block_size;
            if (stage_comm(e)) return 1;
            bound = e->sp_n;
        }
        // determ_projection + deterministic_annihilation (src/fciqmc.f90:394, src/annihilation.f90:64-65); its all-gather
        // is ordered after the exchange's collective, and its time is reported with the exchange (comm_ms)
        if (e->ss.on && stage_determ(e, in, cycle, nullptr)) return 1;
        CK(cudaEventRecord(e->ev[2], st));
        // the sort kernels split the list by its real length whatever their grid is, so the grid follows the length of
        // the previous cycle's list (a hint) rather than the loose upper bound
        if (stage_sort(e, std::min(bound, std::max<long long>(4096, 2 * e->last_spn + 1024)))) return 1;
        CK(cudaEventRecord(e->ev[3], st));
        if (stage_annihilate_launch(e, cycle, bound)) return 1;
        if (e->ss.on && ss_relocate(e, e->alt, e->d_total, false)) return 1;   // on the merged list, counts still on the device
        CK(cudaEventRecord(e->ev[4], st));
        // the cycle's results: one pinned block, one synchronisation
        CK(cudaMemcpyAsync(&ho->st, e->d_stats, sizeof(CycleStats), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ho->npart_new, e->d_ll, sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ho->tot, e->d_total, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ho->err, e->d_err, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ho->head, e->d_head, sizeof(unsigned long long) * np, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ho->spn, e->sp_pn, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (overlap) {
            // this receive buffer is next written two exchanges from now; peers cannot get there before this rank has
            // joined the next exchange's collective, which is stream-ordered after this reset
            CK(cudaMemsetAsync(e->d_recv_head + e->xparity, 0, sizeof(unsigned long long), st));
            e->xparity ^= 1;
        }
        CK(cudaStreamSynchronize(st));
        for (int k = 0; k < 4; ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, e->ev[k], e->ev[k + 1]);
            acc[k] += t;
        }
        {
            float t = 0.f;
            cudaEventElapsedTime(&t, e->evk[0], e->evk[1]);
            e->spawn_kernel_ms += t;
        }
        cs = ho->st;
        long long ev = 0;
        for (int d = 0; d < np; ++d) {
            e->h_head[d] = std::min<unsigned long long>(ho->head[d], (unsigned long long)e->block_size);
            ev += (long long)e->h_head[d];
        }
        out->proj_energy += cs.pe;
        out->D0_population += cs.d0;
        out->nattempts_spawn += cs.nattempts_spawn;
        out->nspawn_events = ev;
        finish_merge(e, &cs, ho->tot[0], ho->tot[1], ho->npart_new);
        e->last_spn = (long long)std::min<unsigned long long>(ho->spn, (unsigned long long)e->sp_cap);
        e->sp_pn = e->d_spn; e->sp_cap = 0;
        // end_mc_cycle / spawning_rate (src/qmc_common.F90:1240-1304)
        const double ndeath_real = (double)cs.ndeath / (double)e->par.real_factor;
        if (nattempts > 0) out->rspawn += ((double)ev + ndeath_real) / (double)nattempts;
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, e->ev[5], e->ev[4]);
    for (int k = 0; k < 4; ++k) e->ms[k] = acc[k];
    e->ms[4] = tot;
    e->ms[5] = e->spawn_kernel_ms;
    out->nparticles = (double)e->nparticles_enc / (double)e->par.real_factor;
    out->nstates = e->nstates;
    out->ndeath = cs.ndeath;
    out->nattempts = nattempts;
    out->spawn_error = ho->err[0];
    out->psip_error = ho->err[1];
    return 0;
}

// ---- peer-to-peer exchange set-up -------------------------------------------------------------------------------
// hb200_p2p_export: allocate this rank's receive block and return its CUDA IPC handle (64 bytes); the host gathers the
// handles of all ranks (MPI_Allgather / torch.distributed) and passes them to hb200_p2p_import, which maps every
// peer's block.  One process per GPU, all on one node (NVLink / NVSwitch or PCIe peer access).
int hb200_p2p_export(hb200_engine* e, uint8_t handle[64]) {
    CK(cudaSetDevice(e->cfg.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    const size_t bufb = (size_t)e->cfg.spawned_walker_length * e->E * 8;
    const size_t bufa = (bufb + 255) & ~(size_t)255;
    if (!e->d_p2p_block) {
        void* q = nullptr;
        CK(cudaMalloc(&q, 256 + 2 * bufa));
        e->owned.push_back(q);
        e->d_p2p_block = (unsigned char*)q;
        CK(cudaMemset(q, 0, 256));
        e->d_recv_head = (unsigned long long*)q;
        e->d_recv[0] = (int64_t*)(e->d_p2p_block + 256);
        e->d_recv[1] = (int64_t*)(e->d_p2p_block + 256 + bufa);
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->d_p2p_block));
    memcpy(handle, &h, 64);
    return 0;
}

// Switch the peer-to-peer exchange off (or back on after a successful hb200_p2p_import): every rank must use the same
// exchange, so a host whose import failed on one rank (no peer access between two of the GPUs) disables it everywhere.
int hb200_p2p_enable(hb200_engine* e, int32_t on) {
    if (on && !e->d_peer_recv) FAIL("p2p_enable: hb200_p2p_import has not succeeded on this rank");
    e->p2p = on != 0;
    return 0;
}

int hb200_set_host_barrier(hb200_engine* e, hb200_barrier_fn fn, void* arg) {
    e->host_barrier = fn;
    e->host_barrier_arg = arg;
    return 0;
}

int hb200_p2p_import(hb200_engine* e, const uint8_t* handles, int32_t nprocs) {
    CK(cudaSetDevice(e->cfg.device));
    if (nprocs != e->par.nprocs) FAIL("p2p_import: nprocs differs from hb200_create");
    if (!e->d_p2p_block) FAIL("p2p_import: call hb200_p2p_export first");
    const int np = nprocs, me = e->par.iproc;
    const size_t bufb = (size_t)e->cfg.spawned_walker_length * e->E * 8;
    const size_t bufa = (bufb + 255) & ~(size_t)255;
    e->peer_base.assign(np, nullptr);
    for (int r = 0; r < np; ++r) {
        if (r == me) { e->peer_base[r] = e->d_p2p_block; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* q = nullptr;
        CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
        e->peer_base[r] = q;
    }
    std::vector<int64_t*> pr((size_t)2 * np);
    std::vector<unsigned long long*> ph(np);
    for (int r = 0; r < np; ++r) {
        unsigned char* b = (unsigned char*)e->peer_base[r];
        ph[r] = (unsigned long long*)b;
        pr[r] = (int64_t*)(b + 256);
        pr[np + r] = (int64_t*)(b + 256 + bufa);
    }
    if (dalloc(e, &e->d_peer_recv, (size_t)2 * np)) return 1;
    if (dalloc(e, &e->d_peer_head, (size_t)np)) return 1;
    if (dalloc(e, &e->d_snap, (size_t)9 * np)) return 1;
    if (dalloc(e, &e->d_push, (size_t)3 * np)) return 1;
    CK(copy_sync(e, e->d_peer_recv, pr.data(), sizeof(int64_t*) * 2 * np, cudaMemcpyHostToDevice));
    CK(copy_sync(e, e->d_peer_head, ph.data(), sizeof(unsigned long long*) * np, cudaMemcpyHostToDevice));
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&e->comm_stream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&e->ev_chunk[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming));
    e->xparity = 0;
    e->p2p = true;
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
    CK(copy_states_h2d(e, d_f, states, n, e->stream));
    if (e->ops->sc0(e, 0.0, d_f, e->W, n, d_o, nullptr)) return 1;
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
    CK(copy_states_h2d(e, d_f, states, n, e->stream));
    CK(copy_sync(e, d_p, pops, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, d_a, attempt, (size_t)n * 4, cudaMemcpyHostToDevice));
    { int rc = 0; DISPATCH_CCMC(e, rc, hb_gen_excit_batch_w, e, p, d_f, d_p, d_a, n, d_io, d_do, d_ns, nullptr, 0, nullptr); if (rc) return 1; }
    CK(cudaStreamSynchronize(e->stream));
    CK(copy_sync(e, iout, d_io, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, dout, d_do, (size_t)n * 2 * 8, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, nspawn, d_ns, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_f); cudaFree(d_p); cudaFree(d_a); cudaFree(d_io); cudaFree(d_do); cudaFree(d_ns);
    return 0;
}

// Level-1 parity hook (SURVEY 8b "inject_rng"): the same as hb200_gen_excit_batch with INJECTED uniform random numbers -
// attempt k consumes rn[k][0], rn[k][1], ... in order (excitation generator, then attempt_to_spawn) instead of the Philox
// stream; nused[k] = how many it drew.  A host compares its own generator fed the same numbers.
int hb200_gen_excit_batch_rn(hb200_engine* e, const uint64_t* states, const int64_t* pops, const double* rn, int32_t nrn,
                             int64_t n, double tau, int32_t* iout, double* dout, int64_t* nspawn, int32_t* nused) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys) FAIL("gen_excit_batch_rn: system not set");
    if (uses_heat_bath_tables(e) && !e->have_hb) FAIL("gen_excit_batch_rn: heat-bath tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER_ORDERN && !e->have_ppn) FAIL("gen_excit_batch_rn: power_pitzer_orderN tables not built");
    if (e->cfg.excit_gen == HB200_EXCIT_GEN_POWER_PITZER && !e->have_pp) FAIL("gen_excit_batch_rn: power_pitzer tables not built");
    if (n == 0) return 0;
    Params p = e->par;
    p.tau = tau;
    uint64_t* d_f; int64_t* d_p; double* d_rn; int* d_io; double* d_do; int64_t* d_ns; int* d_nu;
    CK(cudaMalloc((void**)&d_f, (size_t)n * e->W * 8));
    CK(cudaMalloc((void**)&d_p, (size_t)n * 8));
    CK(cudaMalloc((void**)&d_rn, (size_t)n * nrn * 8));
    CK(cudaMalloc((void**)&d_io, (size_t)n * 8 * 4));
    CK(cudaMalloc((void**)&d_do, (size_t)n * 2 * 8));
    CK(cudaMalloc((void**)&d_ns, (size_t)n * 8));
    CK(cudaMalloc((void**)&d_nu, (size_t)n * 4));
    CK(copy_states_h2d(e, d_f, states, n, e->stream));
    CK(copy_sync(e, d_p, pops, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(copy_sync(e, d_rn, rn, (size_t)n * nrn * 8, cudaMemcpyHostToDevice));
    { int rc = 0; DISPATCH_CCMC(e, rc, hb_gen_excit_batch_w, e, p, d_f, d_p, nullptr, n, d_io, d_do, d_ns, d_rn, nrn, d_nu); if (rc) return 1; }
    CK(cudaStreamSynchronize(e->stream));
    CK(copy_sync(e, iout, d_io, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, dout, d_do, (size_t)n * 2 * 8, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, nspawn, d_ns, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(copy_sync(e, nused, d_nu, (size_t)n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_f); cudaFree(d_p); cudaFree(d_rn); cudaFree(d_io); cudaFree(d_do); cudaFree(d_ns); cudaFree(d_nu);
    return 0;
}

}  // extern "C"

// ---- semi-stochastic projection (src/semi_stoch.F90; separate annihilation, the reference's default) --------------------
__global__ void __launch_bounds__(256)
k_ss_gather_pops(const int64_t* __restrict__ pops, const long long* __restrict__ idx, int nloc, double real_factor,
                 double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nloc) out[j] = (double)pops[idx[j]] / real_factor;   // set_determ_info (src/semi_stoch.F90:826-857)
}
static void ss_free(hb200_engine* e) {
    for (void* q : e->ss.bufs) cudaFree(q);
    e->ss = SemiStoch();
    e->par.ss_bits = nullptr; e->par.ss_sorted = nullptr; e->par.ss_tot = 0;
}
template <class T>
static int ss_alloc(hb200_engine* e, T** p, size_t n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    *p = (T*)q;
    e->ss.bufs.push_back(q);
    return 0;
}
// positions of this rank's deterministic states in the current list + the flag bits; fails if one is missing
static int ss_relocate(hb200_engine* e, int buf, const int* ntot2, bool check) {
    if (e->ops->ss_locate(e, buf, ntot2)) return 1;
    e->launches++;
    if (check) {
        int miss = 0;
        CK(copy_sync(e, &miss, e->ss.d_miss, sizeof(int), cudaMemcpyDeviceToHost));
        if (miss) FAIL("semi-stochastic: deterministic states of this rank are missing from the main list");
    }
    return 0;
}
// gather this rank's deterministic amplitudes into its slot of the all-gather buffer (set_determ_info)
static int ss_gather(hb200_engine* e) {
    SemiStoch& S = e->ss;
    if (S.nloc > 0) {
        k_ss_gather_pops<<<(S.nloc + 255) / 256, 256, 0, e->stream>>>(e->d_pops[e->cur], S.d_idx, S.nloc, (double)e->par.real_factor,
                                                                      S.d_full + (size_t)e->par.iproc * S.maxsz);
        CK(cudaGetLastError());
        e->launches++;
    }
    return 0;
}
// determ_projection + deterministic_annihilation of one cycle; full == nullptr: the vector is all-gathered with NCCL
static int stage_determ(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, const double* full_host) {
    SemiStoch& S = e->ss;
    Params& p = e->par;
    p.tau = in->tau; p.shift = in->shift; p.proj_energy_old = in->proj_energy_old; p.cycle = cycle;
    const int np = p.nprocs;
    if (full_host) {
        // the host's own mpi_allgatherv (src/semi_stoch.F90:1061-1063): sizes(r) amplitudes of every rank, rank by rank
        size_t k = 0;
        for (int r = 0; r < np; ++r) {
            if (S.sizes[r])
                CK(cudaMemcpyAsync(S.d_full + (size_t)r * S.maxsz, full_host + k, sizeof(double) * S.sizes[r], cudaMemcpyHostToDevice,
                                   e->stream));
            k += (size_t)S.sizes[r];
        }
    } else {
        if (ss_gather(e)) return 1;
        if (np > 1) {
            if (!e->comm) FAIL("semi-stochastic: nprocs > 1 needs hb200_comm_init (or the staged hb200_determ_project call)");
            NCK(g_nccl.AllGather(S.d_full + (size_t)p.iproc * S.maxsz, S.d_full, (size_t)S.maxsz, ncclDouble, e->comm, e->stream));
        }
    }
    if (e->ops->ss_project(e, p)) return 1;
    e->launches++;
    return 0;
}

extern "C" {

// init_semi_stoch_t (src/semi_stoch.F90:134-377) with the space chosen by the host (create_high_pop_space,
// create_ci_determ_space and read_determ_from_file are host logic): dets = determ%dets, every deterministic
// determinant rank by rank in the host's layout (We words each), every rank's part in ascending list order;
// sizes = determ%sizes.  Builds the hash-free membership table, this rank's slice of the deterministic Hamiltonian (on the
// device, create_determ_hamil) and adds the rank's deterministic states that are not in its main list with zero
// population (add_determ_dets_to_psip_list).  tot = 0 switches the projection off.
int hb200_set_determ_space(hb200_engine* e, const uint64_t* dets, const int32_t* sizes) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->have_sys || !e->have_ref) FAIL("set_determ_space: system / reference not set");
    CK(cudaStreamSynchronize(e->stream));
    ss_free(e);
    const int np = e->par.nprocs, ip = e->par.iproc, We = e->We, W = e->W;
    long long tot = 0;
    int maxsz = 0, displ = 0;
    for (int r = 0; r < np; ++r) {
        if (sizes[r] < 0) FAIL("set_determ_space: negative size");
        if (r < ip) displ += sizes[r];
        tot += sizes[r];
        maxsz = std::max(maxsz, (int)sizes[r]);
    }
    if (tot == 0) return 0;
    if (tot > (1ll << 30)) FAIL("set_determ_space: deterministic space too large");
    SemiStoch& S = e->ss;
    S.tot = (int)tot; S.nloc = sizes[ip]; S.maxsz = maxsz; S.displ = displ;
    S.sizes.assign(sizes, sizes + np);
    auto less = [We](const uint64_t* a, const uint64_t* b) {
        for (int k = We - 1; k >= 0; --k) {
            if (a[k] < b[k]) return true;
            if (a[k] > b[k]) return false;
        }
        return false;
    };
    const uint64_t* mine = dets + (size_t)displ * We;
    for (int j = 1; j < S.nloc; ++j)
        if (!less(mine + (size_t)(j - 1) * We, mine + (size_t)j * We)) FAIL("set_determ_space: this rank's determinants are not in ascending order");
    // --- add_determ_dets_to_psip_list: merge the missing states into the list with zero population
    {
        const long long n = e->nstates;
        std::vector<uint64_t> st((size_t)(n + S.nloc) * We), st2;
        std::vector<int64_t> po((size_t)(n + S.nloc)), po2;
        std::vector<double> da((size_t)(n + S.nloc)), da2;
        int64_t got = 0;
        if (hb200_download_psips(e, st.data(), po.data(), da.data(), n + S.nloc, &got)) return 1;
        std::vector<int> missing;
        {
            long long i = 0;
            for (int j = 0; j < S.nloc; ++j) {
                const uint64_t* f = mine + (size_t)j * We;
                while (i < n && less(st.data() + (size_t)i * We, f)) ++i;
                if (i >= n || less(f, st.data() + (size_t)i * We)) missing.push_back(j);
            }
        }
        if (!missing.empty()) {
            const long long m = (long long)missing.size();
            if (n + m > e->cfg.walker_length) FAIL("set_determ_space: the deterministic states do not fit the main list");
            std::vector<uint64_t> mf((size_t)m * We);
            std::vector<double> md((size_t)m);
            for (long long k = 0; k < m; ++k) memcpy(mf.data() + (size_t)k * We, mine + (size_t)missing[(size_t)k] * We, (size_t)We * 8);
            if (hb200_sc0_batch(e, mf.data(), m, md.data())) return 1;
            st2.resize((size_t)(n + m) * We); po2.resize((size_t)(n + m)); da2.resize((size_t)(n + m));
            long long i = 0, o = 0;
            for (long long k = 0; k <= m; ++k) {
                const uint64_t* f = (k < m) ? mf.data() + (size_t)k * We : nullptr;
                while (i < n && (!f || less(st.data() + (size_t)i * We, f))) {
                    memcpy(st2.data() + (size_t)o * We, st.data() + (size_t)i * We, (size_t)We * 8);
                    po2[(size_t)o] = po[(size_t)i]; da2[(size_t)o] = da[(size_t)i];
                    ++i; ++o;
                }
                if (f) {
                    memcpy(st2.data() + (size_t)o * We, f, (size_t)We * 8);
                    po2[(size_t)o] = 0; da2[(size_t)o] = md[(size_t)k] - e->par.H00;   // insert_new_walker: sc0 - H00
                    ++o;
                }
            }
            int herr[4] = {0, 0, 0, 0};
            CK(copy_sync(e, herr, e->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost));
            if (hb200_upload_psips(e, st2.data(), po2.data(), da2.data(), n + m)) return 1;
            CK(copy_sync(e, e->d_err, herr, 4 * sizeof(int), cudaMemcpyHostToDevice));   // upload resets the error flags
        }
    }
    // --- device tables
    if (ss_alloc(e, &S.d_all, (size_t)tot * W) || ss_alloc(e, &S.d_sorted, (size_t)tot * W) || ss_alloc(e, &S.d_pad, (size_t)tot) ||
        ss_alloc(e, &S.d_idx, (size_t)S.nloc) || ss_alloc(e, &S.d_bits, ((size_t)e->cfg.walker_length + 31) / 32 + 1) ||
        ss_alloc(e, &S.d_full, (size_t)np * maxsz) || ss_alloc(e, &S.d_vec, (size_t)S.nloc) || ss_alloc(e, &S.d_rho, (size_t)S.nloc) ||
        ss_alloc(e, &S.d_colptr, (size_t)S.nloc + 1) || ss_alloc(e, &S.d_miss, 1))
        return 1;
    S.d_local = S.d_all + (size_t)displ * W;
    CK(cudaMemsetAsync(S.d_full, 0, sizeof(double) * (size_t)np * maxsz, e->stream));
    CK(copy_states_h2d(e, S.d_all, dets, tot, e->stream));
    {
        std::vector<const uint64_t*> ptr((size_t)tot);
        for (long long i = 0; i < tot; ++i) ptr[(size_t)i] = dets + (size_t)i * We;
        std::sort(ptr.begin(), ptr.end(), less);
        std::vector<uint64_t> sorted((size_t)tot * We);
        for (long long i = 0; i < tot; ++i) memcpy(sorted.data() + (size_t)i * We, ptr[(size_t)i], (size_t)We * 8);
        for (long long i = 1; i < tot; ++i)
            if (!less(sorted.data() + (size_t)(i - 1) * We, sorted.data() + (size_t)i * We)) FAIL("set_determ_space: repeated determinant");
        CK(copy_states_h2d(e, S.d_sorted, sorted.data(), tot, e->stream));
        std::vector<int> pad((size_t)tot);
        long long i = 0;
        for (int r = 0; r < np; ++r)
            for (int k = 0; k < sizes[r]; ++k, ++i) pad[(size_t)i] = r * maxsz + k;
        CK(cudaMemcpyAsync(S.d_pad, pad.data(), sizeof(int) * (size_t)tot, cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    // --- create_determ_hamil: count, scan, fill
    if (e->ops->ss_hamil(e, 0)) return 1;
    {
        std::vector<long long> cp((size_t)S.nloc + 1, 0);
        if (S.nloc) CK(copy_sync(e, cp.data(), S.d_colptr, sizeof(long long) * (size_t)S.nloc, cudaMemcpyDeviceToHost));
        long long run = 0;
        for (int j = 0; j <= S.nloc; ++j) { const long long c = (j < S.nloc) ? cp[(size_t)j] : 0; cp[(size_t)j] = run; run += c; }
        S.nnz = run;
        CK(copy_sync(e, S.d_colptr, cp.data(), sizeof(long long) * ((size_t)S.nloc + 1), cudaMemcpyHostToDevice));
        if (ss_alloc(e, &S.d_row, (size_t)S.nnz) || ss_alloc(e, &S.d_val, (size_t)S.nnz)) return 1;
    }
    if (e->ops->ss_hamil(e, 1)) return 1;
    CK(cudaStreamSynchronize(e->stream));
    // --- determ%indices / determ%flags on the current list
    if (ss_relocate(e, e->cur, nullptr, true)) return 1;
    S.on = true;
    e->par.ss_bits = S.d_bits; e->par.ss_sorted = S.d_sorted; e->par.ss_tot = S.tot;
    return 0;
}

// determ%hamil of this rank by column (create_determ_hamil): nnz is returned; with non-null arrays col_ptr[sizes(iproc)+1],
// row[nnz] (row index 0..tot-1 in determ%dets order) and val[nnz] are filled.
int64_t hb200_determ_hamil(hb200_engine* e, int64_t* col_ptr, int32_t* row, double* val) {
    SemiStoch& S = e->ss;
    if (!S.on) return 0;
    if (col_ptr) {
        cudaSetDevice(e->cfg.device);
        if (copy_sync(e, col_ptr, S.d_colptr, sizeof(long long) * ((size_t)S.nloc + 1), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (S.nnz) {
            std::vector<int> pr((size_t)S.nnz);
            if (copy_sync(e, pr.data(), S.d_row, sizeof(int) * (size_t)S.nnz, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
            if (copy_sync(e, val, S.d_val, sizeof(double) * (size_t)S.nnz, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
            std::vector<int> displs(S.sizes.size(), 0);
            for (size_t r = 1; r < S.sizes.size(); ++r) displs[r] = displs[r - 1] + S.sizes[r - 1];
            for (long long z = 0; z < S.nnz; ++z) row[z] = displs[(size_t)(pr[(size_t)z] / S.maxsz)] + pr[(size_t)z] % S.maxsz;
        }
    }
    return S.nnz;
}

// determ%vector of this rank: which = 0 the amplitudes of its deterministic states now (what set_determ_info collects
// during the spawning loop; the host all-gathers these for hb200_determ_project), which = 1 the result of the last
// projection, -tau (H - S) v restricted to the rank.
int hb200_determ_vector(hb200_engine* e, int32_t which, double* vec) {
    CK(cudaSetDevice(e->cfg.device));
    SemiStoch& S = e->ss;
    if (!S.on) FAIL("determ_vector: no deterministic space");
    if (S.nloc == 0) return 0;
    if (which == 0) {
        if (ss_gather(e)) return 1;
        CK(copy_sync(e, vec, S.d_full + (size_t)e->par.iproc * S.maxsz, sizeof(double) * (size_t)S.nloc, cudaMemcpyDeviceToHost));
    } else {
        CK(copy_sync(e, vec, S.d_vec, sizeof(double) * (size_t)S.nloc, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// Staged call between hb200_spawn_death and hb200_annihilate_main: determ_projection + deterministic_annihilation
// (src/semi_stoch.F90:1009-1094, src/annihilation.f90:488-535).  full_vector = determ%full_vector gathered by the host
// (tot doubles, rank by rank), or NULL to gather on the device (one rank, or NCCL).
int hb200_determ_project(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, const double* full_vector) {
    CK(cudaSetDevice(e->cfg.device));
    if (!e->ss.on) FAIL("determ_project: no deterministic space");
    if (stage_determ(e, in, cycle, full_vector)) return 1;
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

}  // extern "C"

extern "C" {

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

