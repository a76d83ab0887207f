// hande_b200: kernels that work on the walker lists with the bit-string width W as a template parameter - annihilation,
// rounding, merge, <D|H|D> of new determinants, slot populations, and the key compression of the wide layout - and their
// launchers.  Included by hb_list_tu.cu, which is compiled once per W (1..4 with byte occupied lists, 32 with 16-bit ones).
#pragma once
#include "hb_common.cuh"

// ------------------------------------------------------------------------------------------------
// Kernel: annihilate_spawn_t[_initiator] + annihilate_main_list[_initiator] + round_low_population_spawns
// (src/spawn_data.F90:859-1101, src/annihilation.f90:294-486, 600-675) on the sorted spawn list.
// ------------------------------------------------------------------------------------------------

// What happens to one distinct spawned determinant once its events are summed: annihilation against the main list or
// a new entry of it.  pop / initiator_pop / events: totals over the run that starts at element i.
template <int W>
__device__ __forceinline__ void annihilate_apply(const Params& p, int64_t* __restrict__ sp, long long i, const uint64_t* key,
                                                 long long pop, long long initiator_pop, long long events,
                                                 const uint64_t* __restrict__ states, int64_t* __restrict__ pops,
                                                 long long nstates, int* __restrict__ ins_flag, long long* __restrict__ ins_pos) {
    constexpr int E = W + 2;
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
        rng.begin(p.seed, p.cycle, RNG_ROUND_SPAWN, det_hash64<W>(key, HB_NW(p)), 0);
        pop = stochastic_round(rng, (int64_t)pop, p.real_factor);
        if (pop == 0) return;
    }
    sp[i * E + W] = pop;
    ins_flag[i] = 1;
    ins_pos[i] = pos;
}

template <int W>
__global__ void __launch_bounds__(256)
k_annihilate(Params p, int64_t* __restrict__ sp, const unsigned long long* __restrict__ pn, long long cap,
             const uint64_t* __restrict__ states, int64_t* __restrict__ pops, long long nstates, int* __restrict__ ins_flag,
             long long* __restrict__ ins_pos, long long* __restrict__ long_q, unsigned* __restrict__ long_n) {
    constexpr int E = W + 2;
    const long long n = dev_count(pn, cap);
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
    long long j = i;
    for (; j < n && j < i + ANN_SHORT; ++j) {
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
    if (j == i + ANN_SHORT && j < n) {
        bool same = true;
#pragma unroll
        for (int k = 0; k < W; ++k) same = same && ((uint64_t)sp[j * E + k] == key[k]);
        if (same) {     // a long run (e.g. the reference determinant near convergence): one warp sums it, k_annihilate_long
            long_q[atomicAdd(long_n, 1u)] = i;
            return;
        }
    }
    annihilate_apply<W>(p, sp, i, key, pop, initiator_pop, events, states, pops, nstates, ins_flag, ins_pos);
}

// The long runs queued by k_annihilate: one warp per run (grid-stride over the queue); the end of the run is found by a
// galloping + binary search on the sorted keys, the sums are integer (order-independent).
template <int W>
__global__ void __launch_bounds__(256)
k_annihilate_long(Params p, int64_t* __restrict__ sp, const unsigned long long* __restrict__ pn, long long cap,
                  const uint64_t* __restrict__ states, int64_t* __restrict__ pops, long long nstates,
                  int* __restrict__ ins_flag, long long* __restrict__ ins_pos, const long long* __restrict__ long_q,
                  const unsigned* __restrict__ long_n) {
    constexpr int E = W + 2;
    const long long n = dev_count(pn, cap);
    const unsigned nq = *long_n;
    const int lane = threadIdx.x & 31;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned q = gw; q < nq; q += nw) {
        const long long i = long_q[q];
        uint64_t key[W];
#pragma unroll
        for (int k = 0; k < W; ++k) key[k] = (uint64_t)sp[i * E + k];
        auto same_at = [&](long long j) {
            bool same = true;
#pragma unroll
            for (int k = 0; k < W; ++k) same = same && ((uint64_t)sp[j * E + k] == key[k]);
            return same;
        };
        long long lo = i + ANN_SHORT, step = ANN_SHORT;      // element lo belongs to the run
        long long hi = lo + step;
        while (hi < n && same_at(hi)) { lo = hi; step <<= 1; hi = lo + step; }
        if (hi > n) hi = n;                                   // first element not in the run lies in (lo, hi]
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (same_at(mid)) lo = mid; else hi = mid;
        }
        long long pop = 0, initiator_pop = 0, events = 0;
        for (long long j = i + lane; j < hi; j += 32) {
            const long long pj = sp[j * E + W];
            pop += pj;
            if (p.initiator) {
                if (!(sp[j * E + W + 1] & 1)) initiator_pop += pj;
                else events += (pj < 0) ? -1 : ((pj > 0) ? 1 : 0);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pop += __shfl_xor_sync(0xffffffffu, pop, o);
            initiator_pop += __shfl_xor_sync(0xffffffffu, initiator_pop, o);
            events += __shfl_xor_sync(0xffffffffu, events, o);
        }
        if (lane == 0) annihilate_apply<W>(p, sp, i, key, pop, initiator_pop, events, states, pops, nstates, ins_flag, ins_pos);
        __syncwarp();
    }
}

// compaction of the surviving new determinants: ins[k] = [f, pop, pos]
template <int W>
__global__ void __launch_bounds__(256)
k_compact_inserts(const int64_t* __restrict__ sp, const unsigned long long* __restrict__ pn, long long cap,
                  const int* __restrict__ ins_flag, const int* __restrict__ ins_idx, const long long* __restrict__ ins_pos,
                  int64_t* __restrict__ ins) {
    constexpr int E = W + 2;
    const long long n = dev_count(pn, cap);
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
k_sc0(Sys s, double H00, const uint64_t* __restrict__ dets, long long stride_words, long long n, double* __restrict__ out,
      const int* __restrict__ pn) {
    if (pn) n = min(n, (long long)*pn);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t f[W];
#pragma unroll
    for (int w = 0; w < W; ++w) f[w] = dets[k * stride_words + w];
    occ_t occ[HB_MAXNEL];
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
        // deterministic states are neither rounded nor removed (src/annihilation.f90:572-586)
        const bool determ_det = p.ss_bits != nullptr && ((p.ss_bits[i >> 5] >> (i & 31)) & 1u);
        if (p.real_amplitudes && !determ_det) {
            const int64_t ap = pop < 0 ? -pop : pop;
            if (pop != 0 && ap < p.real_factor) {
                uint64_t f[W];
                load_det<W>(states + i * W, f);
                PhiloxStream rng;
                rng.begin(p.seed, p.cycle, RNG_ROUND_MAIN, det_hash64<W>(f, HB_NW(p)), 0);
                pop = stochastic_round(rng, pop, p.real_factor);
                pops[i] = pop;
            }
        }
        keep = pop != 0 || determ_det;
    }
    int tot;
    block_excl_scan(keep, swarp, &tot);
    if (threadIdx.x == 0) tile_keep[blockIdx.x] = tot;
}

// remove_unoccupied_dets (compaction) + insert_new_walkers (src/annihilation.f90:537-598, 677-818) as ONE
// out-of-place merge: tile of TILE old states + the new determinants whose insertion point falls in the tile.
// FUSED: the stochastic rounding of remove_unoccupied_dets and the survivor counts (k_round_count + the scan of the
// tile counts) happen here as well - one pass over the old list instead of two.  Tiles take a ticket (so that every
// predecessor has started), publish their survivor count in a packed {flag, value} word and look back over their
// predecessors for the exclusive prefix (decoupled look-back, as in single-pass scans); the last tile writes the
// total.  The spin is bounded: a protocol failure sets the error flag instead of hanging the device.
constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_MASK = (1ull << 62) - 1ull;
template <int W, bool FUSED>
__global__ void __launch_bounds__(TILE)
k_merge(Params p, const uint64_t* __restrict__ states, const int64_t* __restrict__ pops, const double* __restrict__ dat,
        long long nstates, const int* __restrict__ tile_off, const int64_t* __restrict__ ins,
        const double* __restrict__ ins_dat, const int* __restrict__ pnins, uint64_t* __restrict__ ostates,
        int64_t* __restrict__ opops, double* __restrict__ odat, long long* __restrict__ part_npart, int ntiles,
        unsigned long long* __restrict__ tile_state, unsigned* __restrict__ ticket, int* __restrict__ total_kept,
        int* __restrict__ err) {
    constexpr int E = W + 2;
    const long long nins = *pnins;
    __shared__ int swarp[8];
    __shared__ int skept[TILE + 1];
    __shared__ long long sk[2];
    __shared__ long long sred[8];
    __shared__ unsigned s_tile;
    __shared__ long long s_excl;
    const int tid = threadIdx.x;
    unsigned tile = blockIdx.x;
    if (FUSED) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        tile = s_tile;
    }
    const long long t0 = (long long)tile * TILE;
    const long long t1 = min(nstates, t0 + TILE);
    const bool last = ((int)tile == ntiles - 1);
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
    if (m < t1) {
        pop = pops[m];
        const bool determ_det = p.ss_bits != nullptr && ((p.ss_bits[m >> 5] >> (m & 31)) & 1u);
        if (FUSED && p.real_amplitudes && !determ_det) {
            // remove_unoccupied_dets: stochastic rounding of |population| < 1 (src/annihilation.f90:537-598)
            const int64_t ap = pop < 0 ? -pop : pop;
            if (pop != 0 && ap < p.real_factor) {
                uint64_t f[W];
                load_det<W>(states + m * W, f);
                PhiloxStream rng;
                rng.begin(p.seed, p.cycle, RNG_ROUND_MAIN, det_hash64<W>(f, HB_NW(p)), 0);
                pop = stochastic_round(rng, pop, p.real_factor);
            }
        }
        keep = pop != 0 || determ_det;
    }
    int tot;
    const int kb = block_excl_scan(keep, swarp, &tot);
    skept[tid] = kb;
    if (tid == 0) skept[TILE] = tot;
    if (FUSED && tid == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            atomicExch(&tile_state[0], LB_PREFIX | (unsigned long long)tot);
        } else {
            atomicExch(&tile_state[tile], LB_AGG | (unsigned long long)tot);
            long long j = (long long)tile - 1;
            unsigned spins = 0;
            for (;;) {
                const unsigned long long v = atomicAdd(&tile_state[j], 0ull);
                if ((v >> 62) == 0) {
                    if (++spins > (1u << 28)) { atomicOr(err + 1, 2); break; }      // never in a correct run
                    continue;
                }
                excl += v & LB_MASK;
                if (v & LB_PREFIX) break;
                --j;
            }
            atomicExch(&tile_state[tile], LB_PREFIX | (excl + (unsigned long long)tot));
        }
        s_excl = (long long)excl;
        if (last) *total_kept = (int)(excl + (unsigned long long)tot);
    }
    __syncthreads();
    const long long k_lo = sk[0], k_hi = sk[1];
    const long long ns = k_hi - k_lo;
    const long long out_base = (FUSED ? s_excl : (long long)tile_off[tile]) + k_lo;
    long long npart = 0;
    long long o_out = 0;
    if (keep) {
        // number of new determinants in this tile inserted at or before old state m
        long long lo = 0, hi = ns;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (ins[(k_lo + mid) * E + W + 1] <= m) lo = mid + 1; else hi = mid;
        }
        const long long o = out_base + kb + lo;
        o_out = o;
        if (W <= 4) {
            uint64_t f[W];
            load_det<W>(states + m * W, f);
            store_det<W>(ostates + o * W, f);
        }
        opops[o] = pop;
        odat[o] = dat[m];
        npart += pop < 0 ? -pop : pop;
    }
    if (W > 4) {
        // wide layout: the warp moves the kept determinants one after the other, lane k word k (coalesced 256-byte rows
        // instead of 32 strided 8-byte accesses per lane)
        const int lane = tid & 31;
        unsigned mm = __ballot_sync(0xffffffffu, keep != 0);
        while (mm) {
            const int src = __ffs(mm) - 1;
            mm &= mm - 1;
            const long long ms = __shfl_sync(0xffffffffu, m, src), os = __shfl_sync(0xffffffffu, o_out, src);
            for (int k = lane; k < W; k += 32) ostates[os * W + k] = __ldcs(states + ms * W + k);
        }
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
        part_npart[tile] = t;
    }
}

// initialise_slot_pop (src/load_balancing.F90:624-654): encoded |population| per load-balancing slot
// (slot = modulo(hash(f), nprocs * nslots)); integer atomics, so the sums are exact and order-independent
template <int W>
__global__ void __launch_bounds__(256)
k_slot_pop(const uint64_t* __restrict__ states, const int64_t* __restrict__ pops, long long n, int nbasis, uint32_t seed,
           int nprocs, int nslots, unsigned long long* __restrict__ slot_pop) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t f[W];
    load_det<W>(states + i * W, f);
    const long long pp = pops[i];
    if (pp != 0) atomicAdd(&slot_pop[owner_slot(f, nbasis, seed, nprocs, nslots)], (unsigned long long)(pp < 0 ? -pp : pp));
}


// ------------------------------------------------------------------------------------------------
// Wide layout (W = 32): the spawn list is sorted through compressed keys.  A determinant of nel electrons is the list of
// its occupied orbitals; bit_str_cmp order (unsigned, last word most significant) is the lexicographic order of that
// list read from the largest orbital down, i.e. the order of the integer whose field m (bits [m b, (m+1) b), b =
// bits per orbital index) holds the m-th smallest orbital.  nel * b bits (154 for 14 electrons in 2000 plane waves)
// are radix-sorted instead of nbasis bits; the elements themselves move once.
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
k_wide_compress(const int64_t* __restrict__ sp, const unsigned long long* __restrict__ pn, long long cap, int bits, int kw,
                int64_t* __restrict__ items) {
    constexpr int E = W + 2;
    const long long n = dev_count(pn, cap);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key[5] = {0, 0, 0, 0, 0};
    int off = 0;
    for (int w = 0; w < W; ++w) {
        uint64_t x = (uint64_t)sp[i * E + w];
        while (x) {
            const uint64_t v = (uint64_t)(w * 64 + ctz64(x) + 1);
            x &= x - 1;
            const int word = off >> 6, sh = off & 63;
            key[word] |= v << sh;
            if (sh + bits > 64) key[word + 1] |= v >> (64 - sh);
            off += bits;
        }
    }
    for (int k = 0; k < kw; ++k) items[i * (kw + 1) + k] = (int64_t)key[k];
    items[i * (kw + 1) + kw] = i;
}
template <int W>
__global__ void __launch_bounds__(256)
k_wide_gather(const int64_t* __restrict__ sp, const unsigned long long* __restrict__ pn, long long cap, int kw,
              const int64_t* __restrict__ items, int64_t* __restrict__ out) {
    constexpr int E = W + 2;
    const long long n = dev_count(pn, cap);
    const long long total = n * E, stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const long long j = t / E;
        const int w = (int)(t - j * E);
        out[t] = sp[items[j * (kw + 1) + kw] * E + w];
    }
}

// ------------------------------------------------------------------------------------------------
// launchers (one set per W, reached through ListOps)
// ------------------------------------------------------------------------------------------------
template <int W>
static int list_annihilate(hb200_engine* e, const Params& p, int64_t* sp, long long bound) {
    const int c = e->cur;
    const unsigned nb = (unsigned)((bound + 255) / 256);
    k_annihilate<W><<<nb, 256, 0, e->stream>>>(p, sp, e->sp_pn, e->sp_cap, e->d_states[c], e->d_pops[c], e->nstates, e->d_ins_flag,
                                                e->d_ins_pos, e->d_long_q, e->d_long_n);
    k_annihilate_long<W><<<std::min(nb, 592u), 256, 0, e->stream>>>(p, sp, e->sp_pn, e->sp_cap, e->d_states[c], e->d_pops[c],
                                                                    e->nstates, e->d_ins_flag, e->d_ins_pos, e->d_long_q,
                                                                    e->d_long_n);
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int list_compact(hb200_engine* e, const int64_t* sp, long long bound, int64_t* ins) {
    k_compact_inserts<W><<<(unsigned)((bound + 255) / 256), 256, 0, e->stream>>>(sp, e->sp_pn, e->sp_cap, e->d_ins_flag, e->d_ins_idx,
                                                                                  e->d_ins_pos, ins);
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int list_round_count(hb200_engine* e, const Params& p, int ntiles) {
    const int c = e->cur;
    k_round_count<W><<<ntiles, TILE, 0, e->stream>>>(p, e->d_states[c], e->d_pops[c], e->nstates, e->d_tile_keep);
    CK(cudaGetLastError());
    return 0;
}
// dets: n determinants, stride_words apart; pn != nullptr: the count is min(n, *pn) on the device
template <int W>
static int list_sc0(hb200_engine* e, double H00, const uint64_t* dets, long long stride_words, long long n, double* out,
                    const int* pn) {
    k_sc0<W><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->sys, H00, dets, stride_words, n, out, pn);
    CK(cudaGetLastError());
    return 0;
}
// fused = true: rounding + survivor counts + merge in one pass (k_round_count and the tile scan are not launched)
template <int W>
static int list_merge(hb200_engine* e, const Params& p, const int64_t* ins, int ntiles, bool fused) {
    const int c = e->cur, o = e->alt;
    if (fused) {
        CK(cudaMemsetAsync(e->d_tile_state, 0, sizeof(unsigned long long) * (size_t)ntiles, e->stream));
        CK(cudaMemsetAsync(e->d_ticket, 0, sizeof(unsigned), e->stream));
        k_merge<W, true><<<ntiles, TILE, 0, e->stream>>>(p, e->d_states[c], e->d_pops[c], e->d_dat[c], e->nstates, e->d_tile_off, ins,
                                                         e->d_ins_dat, e->d_total, e->d_states[o], e->d_pops[o], e->d_dat[o],
                                                         e->d_part_ll, ntiles, e->d_tile_state, e->d_ticket, e->d_total + 1, e->d_err);
    } else {
        k_merge<W, false><<<ntiles, TILE, 0, e->stream>>>(p, e->d_states[c], e->d_pops[c], e->d_dat[c], e->nstates, e->d_tile_off, ins,
                                                          e->d_ins_dat, e->d_total, e->d_states[o], e->d_pops[o], e->d_dat[o],
                                                          e->d_part_ll, ntiles, nullptr, nullptr, nullptr, e->d_err);
    }
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int list_slot_pop(hb200_engine* e, unsigned long long* d) {
    const long long m = e->nstates;
    k_slot_pop<W><<<(unsigned)((m + 255) / 256), 256, 0, e->stream>>>(e->d_states[e->cur], e->d_pops[e->cur], m, e->sys.nbasis,
                                                                       e->par.hash_seed, e->par.nprocs, e->par.nslots, d);
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int list_compress(hb200_engine* e, const int64_t* sp, long long bound, int bits, int kw, int64_t* items) {
    k_wide_compress<W><<<(unsigned)((bound + 255) / 256), 256, 0, e->stream>>>(sp, e->sp_pn, e->sp_cap, bits, kw, items);
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int list_gather(hb200_engine* e, const int64_t* sp, int kw, const int64_t* items, int64_t* out) {
    k_wide_gather<W><<<1184, 256, 0, e->stream>>>(sp, e->sp_pn, e->sp_cap, kw, items, out);
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int list_owner_slot_shift(const hb200_engine* e, const Params& p) {
    return owner_slot_shift<W>(p.f0, e->sys.nbasis, p.hash_seed, p.ccmc_shift, p.ccmc_freq, p.nprocs, p.nslots);
}
#include "hb_semistoch.cuh"

template <int W>
static const ListOps* list_ops() {
    static const ListOps ops = {list_annihilate<W>, list_compact<W>, list_round_count<W>, list_sc0<W>, list_merge<W>,
                                list_slot_pop<W>, list_compress<W>, list_gather<W>, list_owner_slot_shift<W>,
                                ss_locate<W>, ss_hamil<W>, ss_project<W>};
    return &ops;
}
