// hande_b200: the spawning step of the original heat-bath generator (excit_gen = heat_bath, the bench headline) as a
// WAVEFRONT of four kernels that hand compact records to each other through HBM.
//
// Measured on B200 (profiles/r2_*): the per-attempt work is ~3,000 dependent scalar operations with 10-30 cycle
// latencies; a warp alone issues 0.07-0.08 instructions per cycle, and the issue rate of an SM sub-partition grows
// linearly with the number of resident warps up to at least 8.  One fused kernel (hb_spawn_hb.cuh) needs 120 registers
// and 12 KB of shared memory per warp for its slot state and queues, i.e. 4 warps per sub-partition, and its 150 KB of
// code thrashes the instruction cache unless the warps are kept in step.  Split by phase, every kernel is small, needs no
// queue in shared memory (the queue IS the record array: compaction over the whole grid, every lane of every warp busy)
// and runs at 6-8 warps per sub-partition; the records cost ~90 bytes of HBM traffic per attempt, 1-2 ms at 1e8 attempts
// against the ~25 ms they save.
//   K1 k_wf_select  per state: decode, set_parent_flag, update_proj_energy_mol, decide_nattempts, stochastic_death;
//                   per attempt: i, j (alias selections over the occupied orbitals), a (precomputed alias row) -> RecA
//   K2 k_wf_coin    slater_condon1(i->a), single/double coin, b (one packed 32-byte record of the 2 GB hb_ijab row)
//                   -> RecD (double excitations) / RecS (single excitations)
//   K3 k_wf_double  |slater_condon1| of the other orderings, four-ordering pgen, H_ij, attempt_to_spawn, append
//   K4 k_wf_single  nel-term pgen sum, attempt_to_spawn, append
// Reference: gen_excit_mol_heat_bath (src/excit_gen_heat_bath_mol.F90:258-548), do_fciqmc_spawning_attempt
// (src/fciqmc.f90:635-769).  Every number is computed by the same operations in the same order as in the oracle.
#pragma once
#include "hb_common.cuh"
#include "hb_spawn_hb.cuh"     // leaf routines shared with the fused kernel (heavy determinants still go through it)

namespace hbwf {
using hbw::u01;
using hbw::OneDraw;
using hbw::HeavyItem;
using hbw::HeavyQueue;

constexpr int K1_WARPS = 8;          // warps per block of k_wf_select
constexpr int HEAVY = 1024;          // attempts of one state above which it is deferred to the fused heavy kernel

enum { RA_NEED_IA = 1, RA_NEG = 2, RA_FLAG = 4 };   // RecA/RecD/RecS flags: i->a allowed; parent population negative; parent not an initiator

// Records: a fixed header followed by the determinant's bit string and its occupied list (four orbitals to a word, zero
// padded), so that the consumer needs neither a dependent gather from the main list, nor a decode, nor shared memory;
// the stride is rounded to 16 bytes.
enum { RD_KB = 0x10, RD_KC = 0x20, RD_KD = 0x40, RD_WKNOWN = 0x80, RS_PERM = 0x10 };
struct alignas(16) RecA {            // an attempt that survived the choice of i, j, a
    uint32_t state, att;             // index in the main list; attempt number (keys the random stream)
    uint8_t i, j, a, flags;
    uint32_t pad;
    double i_tot, ij_tot, r3;        // sums of S_i and ij_weights(:, i) over the occupied orbitals; the stream's draw 3
};
static_assert(sizeof(RecA) == 48, "RecA layout");
struct alignas(16) RecD {            // a double excitation i, j -> a, b before its generation probability is known
    uint32_t state;
    uint8_t i, j, a, b;
    uint8_t flags, pad[7];           // RA_NEG, RA_FLAG, RD_K*: orderings i->b, j->a, j->b allowed as singles, RD_WKNOWN
    double i_tot, ij_tot, psingle, wab, pab, rs;   // rs: the uniform attempt_to_spawn draws
};
static_assert(sizeof(RecD) == 64, "RecD layout");
struct alignas(16) RecS {            // a single excitation i -> a
    uint32_t state;
    uint8_t i, a, flags, pad0;       // RA_NEG, RA_FLAG, RS_PERM
    uint32_t pad1[2];
    double i_tot, ij_tot, h_ia, rs;
};
static_assert(sizeof(RecS) == 48, "RecS layout");
// header | f(W words) | occupied list
__host__ __device__ inline int rec_stride(int header, int nel, int W) { return (header + 8 * W + 4 * ((nel + 3) >> 2) + 15) & ~15; }

struct Counters { unsigned nA, nD, nS, pad; };

// Records are written once and read once: streamed (evict-first) so that they do not push the L2-resident tables out.
template <class T>
__device__ __forceinline__ T load_rec(const unsigned char* src) {
    static_assert(sizeof(T) % 16 == 0, "record headers are multiples of 16 bytes");
    T r;
    int4* d = reinterpret_cast<int4*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); ++k) d[k] = __ldcs(reinterpret_cast<const int4*>(src) + k);
    return r;
}
template <class T>
__device__ __forceinline__ void store_rec(unsigned char* dst, const T& r) {
    const int4* d = reinterpret_cast<const int4*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); ++k) __stcs(reinterpret_cast<int4*>(dst) + k, d[k]);
}
template <int W>
__device__ __forceinline__ void load_tail(const unsigned char* rec, int header, int noccw, uint64_t* f, uint32_t* occw) {
    const uint2* fs = reinterpret_cast<const uint2*>(rec + header);
#pragma unroll
    for (int k = 0; k < W; ++k) { const uint2 v = __ldcs(fs + k); f[k] = ((uint64_t)v.y << 32) | v.x; }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rec + header + 8 * W);
    for (int w = 0; w < noccw; ++w) occw[w] = __ldcs(src + w);
}
template <int W>
__device__ __forceinline__ void store_tail(unsigned char* rec, int header, int noccw, const uint64_t* f, const uint32_t* occw) {
    uint2* fd = reinterpret_cast<uint2*>(rec + header);
#pragma unroll
    for (int k = 0; k < W; ++k) __stcs(fd + k, make_uint2((uint32_t)f[k], (uint32_t)(f[k] >> 32)));
    uint32_t* dst = reinterpret_cast<uint32_t*>(rec + header + 8 * W);
    for (int w = 0; w < noccw; ++w) __stcs(dst + w, occw[w]);
}
// byte offsets inside a warp's private shared-memory region of k_wf_select
struct SelSmem {
    int stage, sf, shash, sscan, socc, sbits, total, noccw;
    __host__ __device__ SelSmem(int W, int nel) {
        int o = 0;
        stage = o;  o += nel * 32 * 8;               // [q][lane] weights of the list being selected from
        sf = o;     o += 32 * W * 8;
        shash = o;  o += 32 * 8;
        sscan = o;  o += 33 * 4; o = (o + 7) & ~7;
        noccw = (nel + 3) >> 2;
        socc = o;   o += 32 * noccw * 4;             // occupied lists, four orbitals to a word
        sbits = o;  o += 32;
        total = (o + 15) & ~15;
    }
};

template <int W, class Mask>
__global__ void __launch_bounds__(K1_WARPS * 32, 4)
k_wf_select(Sys s, Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ pops, const double* __restrict__ dat,
            long long state_lo, long long state_hi, unsigned char* __restrict__ recA, Counters* __restrict__ cnt, unsigned capA,
            SpawnPartials* __restrict__ partials, int* __restrict__ err, HeavyQueue hq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int nel = s.nel;
    const int64_t nb = s.nbasis;
    const SelSmem L(W, nel);
    const int strideA = rec_stride(sizeof(RecA), nel, W);
    double* siw = reinterpret_cast<double*>(smem_raw);
    const int siw_bytes = (s.nbasis * 8 + 15) & ~15;
    for (int k = tid; k < s.nbasis; k += K1_WARPS * 32) siw[k] = s.hb_i_w[k];
    __syncthreads();
    const double* siw1 = siw - 1;
    unsigned char* wb = smem_raw + siw_bytes + warp * L.total;
    double* wq = reinterpret_cast<double*>(wb + L.stage) + lane;
    uint64_t* sf = reinterpret_cast<uint64_t*>(wb + L.sf);
    uint64_t* shash = reinterpret_cast<uint64_t*>(wb + L.shash);
    int* sscan = reinterpret_cast<int*>(wb + L.sscan);
    uint32_t* socc = reinterpret_cast<uint32_t*>(wb + L.socc);
    uint8_t* sbits = wb + L.sbits;
    double pe = 0.0, d0 = 0.0;
    long long ndeath = 0, npart = 0, nattempts = 0;
    const long long ntile = (state_hi - state_lo + 31) / 32;
    const long long gwarp = (long long)blockIdx.x * K1_WARPS + warp, nwarps = (long long)gridDim.x * K1_WARPS;
    for (long long t = gwarp; t < ntile; t += nwarps) {
        // ---- the idet loop's per-determinant part (src/fciqmc.f90:315-371)
        const long long idx = state_lo + t * 32 + lane;
        int natt = 0;
        if (idx < state_hi) {
            uint64_t f[W];
            load_det<W>(states + idx * W, f);
            const int64_t pop = __ldcs(pops + idx);
            const double Kii = __ldcs(dat + idx);
#pragma unroll
            for (int k = 0; k < W; ++k) sf[lane * W + k] = f[k];
            uint8_t* occ = reinterpret_cast<uint8_t*>(socc + lane * L.noccw);
            socc[lane * L.noccw + L.noccw - 1] = 0u;
            decode_det<W>(f, occ);
            const uint64_t h = det_hash64<W>(f);
            shash[lane] = h;
            const double real_pop = (double)pop / (double)p.real_factor;
            // set_parent_flag (src/ifciqmc.f90:13-57)
            sbits[lane] = (uint8_t)((pop < 0 ? RA_NEG : 0) | ((fabs(real_pop) > p.initiator_pop) ? 0 : RA_FLAG));
            // update_proj_energy_mol (src/energy_evaluation.F90:906-986)
            bool is_ref;
            const double hm = proj_energy_hmatel<W>(s, p, f, occ, is_ref);
            if (is_ref) d0 += real_pop; else pe += hm * real_pop;
            // decide_nattempts (src/qmc_common.F90:379-406)
            natt = (int)real_pop;
            if (natt < 0) natt = -natt;
            const double pextra = fabs(real_pop) - natt;
            if (fabs(pextra) > 1.e-12) {
                const uint4 r = hbw::philox_block(((uint32_t)RNG_NATTEMPTS << 24) | 0u, 0u, (uint32_t)h, (uint32_t)(h >> 32), p.seed, p.cycle);
                if (pextra > u01(r.x, r.y)) natt++;
            }
            // stochastic_death (src/death.f90:11-130)
            const uint4 r = hbw::philox_block(((uint32_t)RNG_DEATH << 24) | 0u, 0u, (uint32_t)h, (uint32_t)(h >> 32), p.seed, p.cycle);
            OneDraw rng{u01(r.x, r.y)};
            int64_t kill_abs;
            const double death_weight = p.qn ? qn_weighting(p, qn_fock_sum(s, p, occ)) : 1.0;
            const int64_t newpop = stochastic_death(rng, p, Kii, pop, kill_abs, death_weight);
            pops[idx] = newpop;
            ndeath += kill_abs;
            npart += newpop < 0 ? -newpop : newpop;
            nattempts += natt;
            if (natt > HEAVY) {     // a determinant with a huge population: its attempts are spread over the whole grid later
                const unsigned k = atomicAdd(hq.count, 1u);
                if (k < hq.cap) { HeavyItem hi; hi.state = idx; hi.pop = pop; hi.natt = natt; hi.pad = 0; hq.items[k] = hi; natt = 0; }
            }
        }
        const int incl = warp_incl_scan(natt);
        const int T = __shfl_sync(0xffffffffu, incl, 31);
        sscan[lane] = incl - natt;
        if (lane == 0) sscan[32] = T;
        __syncwarp();
        const int nst = (int)min(32LL, state_hi - state_lo - t * 32);
        // ---- the spawning attempts of the tile, 32 at a time: i, j, a
        for (int base = 0; base < T; base += 32) {
            const int a_idx = base + lane;
            bool allowed = false, need_ia = false;
            int lo = 0, i = 0, j = 0, a = 0;
            uint32_t att = 0;
            double i_tot = 0.0, ij_tot = 0.0, r3 = 0.0;
            if (a_idx < T) {
                int hi = nst;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (sscan[mid] <= a_idx) lo = mid; else hi = mid;
                }
                att = (uint32_t)(a_idx - sscan[lo]);
                const uint32_t* occw = socc + lo * L.noccw;
                const uint8_t* occ = reinterpret_cast<const uint8_t*>(occw);
                const uint64_t hsh = shash[lo];
                const uint4 ra = hbw::philox_block(((uint32_t)RNG_SPAWN << 24) | 0u, att, (uint32_t)hsh, (uint32_t)(hsh >> 32), p.seed, p.cycle);
                const uint4 rb = hbw::philox_block(((uint32_t)RNG_SPAWN << 24) | 1u, att, (uint32_t)hsh, (uint32_t)(hsh >> 32), p.seed, p.cycle);
                r3 = u01(rb.z, rb.w);
                // i from S_i over the occupied orbitals (select_ij_heat_bath, src/excit_gen_utils.f90:9-66)
                i_tot = hbw::gather_occ<true>(siw1, occw, nel, wq);
                double x = u01(ra.x, ra.y) * nel;
                int k = (int)x;
                x = x - k;
                i = occ[alias_select_fast<Mask>(nel, wq, 32, nel / i_tot, k, x) - 1];
                // j from ij_weights(:, i) over the occupied orbitals
                ij_tot = hbw::gather_occ<true>(s.hb_ij_w + nb * (i - 1) - 1, occw, nel, wq);
                if (ij_tot > 0.0) {
                    x = u01(ra.z, ra.w) * nel;
                    k = (int)x;
                    x = x - k;
                    j = occ[alias_select_fast<Mask>(nel, wq, 32, nel / ij_tot, k, x) - 1];
                    allowed = fabs(s.hb_ija_tot[HB_I2(j, i)]) > 0.0;
                }
                if (allowed) {
                    // a from the precomputed alias row hb_ija(:, j, i)
                    x = u01(rb.x, rb.y) * (int)nb;
                    const int K = (int)floor(x);
                    x = x - K;
                    const HbRec* rec = s.hb_ija_rec + HB_I3(1, j, i) + K;
                    a = (x < rec->U) ? K + 1 : rec->K;
                    if (hbw::smem_det_test(sf + lo * W, a)) allowed = false;
                    else need_ia = hb_single_allowed(s, i, a);
                }
            }
            // ---- the surviving attempts go to the record array: one pointer bump per warp
            const unsigned m = __ballot_sync(0xffffffffu, allowed);
            if (m != 0) {
                unsigned slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&cnt->nA, (unsigned)__popc(m));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (allowed) {
                    const unsigned k = slot0 + __popc(m & lt);
                    if (k < capA) {
                        RecA r;
                        r.state = (uint32_t)(state_lo + t * 32 + lo); r.att = att;
                        r.i = (uint8_t)i; r.j = (uint8_t)j; r.a = (uint8_t)a;
                        r.flags = (uint8_t)((need_ia ? RA_NEED_IA : 0) | sbits[lo]);
                        r.pad = 0; r.i_tot = i_tot; r.ij_tot = ij_tot; r.r3 = r3;
                        unsigned char* dst = recA + (size_t)k * strideA;
                        const int4* d4 = reinterpret_cast<const int4*>(&r);
#pragma unroll
                        for (int w = 0; w < 3; ++w) __stcs(reinterpret_cast<int4*>(dst) + w, d4[w]);
                        store_tail<W>(dst, sizeof(RecA), L.noccw, sf + lo * W, socc + lo * L.noccw);
                    } else {
                        atomicOr(err, 4);      // record array too small: an engine bug, not a run-time condition
                    }
                }
            }
        }
        __syncwarp();
    }
    // deterministic block reduction of the estimators
    __shared__ double sred[2][K1_WARPS];
    __shared__ long long lred[3][K1_WARPS];
    const double r0 = warp_sum_d(pe), r1 = warp_sum_d(d0);
    const long long r2 = warp_sum_ll(ndeath), r3 = warp_sum_ll(npart), r4 = warp_sum_ll(nattempts);
    if (lane == 0) { sred[0][warp] = r0; sred[1][warp] = r1; lred[0][warp] = r2; lred[1][warp] = r3; lred[2][warp] = r4; }
    __syncthreads();
    if (tid == 0) {
        SpawnPartials out;
        out.pe = 0.0; out.d0 = 0.0; out.ndeath = 0; out.npart = 0; out.nattempts = 0;
        for (int w = 0; w < K1_WARPS; ++w) {
            out.pe += sred[0][w]; out.d0 += sred[1][w];
            out.ndeath += lred[0][w]; out.npart += lred[1][w]; out.nattempts += lred[2][w];
        }
        partials[blockIdx.x] = out;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// thread-per-record kernels
// ------------------------------------------------------------------------------------------------------------------
constexpr int MAXOCCW = HB_MAXNEL / 4;

// warp-aggregated append of one record of `bytes` (multiple of 16) per participating lane; returns the slot or ~0u
__device__ __forceinline__ unsigned claim_slot(bool want, unsigned* counter) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return ~0u;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned slot0 = 0;
    if (lane == leader) slot0 = atomicAdd(counter, (unsigned)__popc(m));
    slot0 = __shfl_sync(0xffffffffu, slot0, leader);
    return want ? slot0 + __popc(m & ((1u << lane) - 1u)) : ~0u;
}

// K2: slater_condon1(i -> a) where that single excitation is allowed, the single/double coin, b
template <int W>
__global__ void __launch_bounds__(256, 4)
k_wf_coin(Sys s, Params p, const uint64_t* __restrict__ states, const unsigned char* __restrict__ recA, Counters* __restrict__ cnt,
          unsigned char* __restrict__ recD, unsigned char* __restrict__ recS, unsigned cap, int* __restrict__ err) {
    const int nel = s.nel, noccw = (nel + 3) >> 2;
    const int64_t nb = s.nbasis;
    const int strideA = rec_stride(sizeof(RecA), nel, W), strideD = rec_stride(sizeof(RecD), nel, W), strideS = rec_stride(sizeof(RecS), nel, W);
    const unsigned nA = min(cnt->nA, cap);
    const unsigned nround = (nA + 31u) & ~31u;        // whole warps stay in the loop (ballots)
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < nround; r += gridDim.x * blockDim.x) {
        bool is_dbl = false, is_sgl = false;
        RecA ra;
        RecD rd;
        RecS rs_;
        uint32_t occw[MAXOCCW];
        uint64_t f[W];
        if (r < nA) {
            const unsigned char* src = recA + (size_t)r * strideA;
            ra = load_rec<RecA>(src);
            load_tail<W>(src, sizeof(RecA), noccw, f, occw);
            const int i = ra.i, j = ra.j, a = ra.a;
            const uint64_t hsh = det_hash64<W>(f);
            const uint4 rc = hbw::philox_block(((uint32_t)RNG_SPAWN << 24) | 2u, ra.att, (uint32_t)hsh, (uint32_t)(hsh >> 32), p.seed, p.cycle);
            const double r3 = ra.r3, r4 = u01(rc.x, rc.y), r5 = u01(rc.z, rc.w);
            bool dbl = true, perm_ia = false;
            double psingle = 0.0, rb = r3, rsp = r4, h_ia = 0.0;
            const bool need_ia = (ra.flags & RA_NEED_IA) != 0;
            // the draw that selects b is known before the coin is: fetch its 32-byte record from the 2 GB hb_ijab rows now
            // (streamed, no reuse: {aliasU, weight}, {aliasK, -, weight / weights_tot}), under the slater_condon1 sum
            double xb = (need_ia ? r4 : r3) * (int)nb;
            const int K = (int)floor(xb);
            xb = xb - K;
            const HbRec* rec = s.hb_ijab_rec + HB_I4(1, a, j, i) + K;
            const double2 uw = __ldcs(reinterpret_cast<const double2*>(rec));
            const int4 kp = __ldcs(reinterpret_cast<const int4*>(rec) + 1);
            const double wt = s.hb_ija_rec[HB_I3(a, j, i)].w;          // = hb_ijab%weights_tot(a,j,i)
            if (need_ia) {
                perm_ia = excit_perm1<W>(f, i, a);
                const double h = hbw::sc1_lean(s, occw, noccw, i, a);
                h_ia = perm_ia ? -h : h;
                const double hmod = fabs(h_ia);
                if (hmod < wt) psingle = hmod / (wt + hmod); else psingle = 0.5;
                dbl = !(r3 < psingle);
                rb = r4; rsp = dbl ? r5 : r4;
            }
            (void)rb;
            if (dbl) {
                const int b = (xb < uw.x) ? K + 1 : kp.x;
                if (!det_test(f, b)) {
                    is_dbl = true;
                    rd.state = ra.state; rd.i = ra.i; rd.j = ra.j; rd.a = ra.a; rd.b = (uint8_t)b;
                    uint8_t fl = ra.flags & (RA_NEG | RA_FLAG);
                    if (b == K + 1) { fl |= RD_WKNOWN; rd.wab = uw.y; rd.pab = __hiloint2double(kp.w, kp.z); }
                    else { rd.wab = 0.0; rd.pab = 0.0; }
                    if (hb_single_allowed(s, i, b)) fl |= RD_KB;
                    if (hb_single_allowed(s, j, a)) fl |= RD_KC;
                    if (hb_single_allowed(s, j, b)) fl |= RD_KD;
                    rd.flags = fl;
                    for (int k = 0; k < 7; ++k) rd.pad[k] = 0;
                    rd.i_tot = ra.i_tot; rd.ij_tot = ra.ij_tot; rd.psingle = psingle; rd.rs = rsp;
                }
            } else {
                is_sgl = true;
                rs_.state = ra.state; rs_.i = ra.i; rs_.a = ra.a;
                rs_.flags = (uint8_t)((ra.flags & (RA_NEG | RA_FLAG)) | (perm_ia ? RS_PERM : 0));
                rs_.pad0 = 0; rs_.pad1[0] = 0; rs_.pad1[1] = 0;
                rs_.i_tot = ra.i_tot; rs_.ij_tot = ra.ij_tot; rs_.h_ia = h_ia; rs_.rs = rsp;
            }
        }
        const unsigned kd = claim_slot(is_dbl, &cnt->nD);
        if (is_dbl) {
            if (kd < cap) {
                unsigned char* dst = recD + (size_t)kd * strideD;
                store_rec(dst, rd);
                store_tail<W>(dst, sizeof(RecD), noccw, f, occw);
            } else atomicOr(err, 4);
        }
        const unsigned ks = claim_slot(is_sgl, &cnt->nS);
        if (is_sgl) {
            if (ks < cap) {
                unsigned char* dst = recS + (size_t)ks * strideS;
                store_rec(dst, rs_);
                store_tail<W>(dst, sizeof(RecS), noccw, f, occw);
            } else atomicOr(err, 4);
        }
    }
}

// attempt_to_spawn, create_spawned_particle[_initiator][_truncated], assign_particle_processor and the warp-aggregated
// add_[flagged_]spawned_particle (src/spawning.F90:711-766, 770-838, 907-1319) for one generated excitation per lane;
// every lane of the warp must call it
template <int W>
__device__ __forceinline__ void spawn_and_append(const Sys& s, const Params& p, bool live, const uint64_t* f, const uint32_t* occw,
                                                 const Gen& g, double rs, uint8_t flags, int64_t* __restrict__ spawn,
                                                 unsigned long long* __restrict__ head, long long block_size,
                                                 const int* __restrict__ proc_map, int* __restrict__ err) {
    constexpr int E = W + 2;
    const int lane = threadIdx.x & 31;
    int64_t nspawn = 0;
    uint64_t child[W];
    int dest = 0, pflag = 0;
    if (live) {
        double hmq = g.hmatel;
        if (p.qn) {      // spawn_standard (src/spawning.F90:101-103): cdet%fock_sum from the occupied list
            double fs = 0.0;
            const uint8_t* occ = reinterpret_cast<const uint8_t*>(occw);
            for (int k = 0; k < s.nel; ++k) fs = fs + p.sp_fock[occ[k]];
            hmq = hmq * qn_spawned_weighting(p, fs - p.ref_fock_sum, g);
        }
        OneDraw rng{rs};
        nspawn = attempt_to_spawn(rng, p, hmq, g.pgen, (flags & RA_NEG) ? (int64_t)-1 : (int64_t)1);
        if (nspawn != 0) {
            make_child<W>(f, g, child);
            if (p.trunc_level >= 0 && excit_level<W>(child, p.f0) > p.trunc_level) {
                nspawn = 0;
            } else {
                dest = (p.nprocs > 1) ? proc_map[owner_slot(child, s.nbasis, p.hash_seed, p.nprocs, p.nslots)] : 0;
                pflag = p.initiator ? ((flags & RA_FLAG) ? 1 : 0) : 0;
            }
        }
    }
    const unsigned has = __ballot_sync(0xffffffffu, nspawn != 0);
    if (nspawn != 0) {
        const unsigned peers = (p.nprocs > 1) ? __match_any_sync(has, dest) : has;
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        unsigned long long slot0 = 0;
        if (lane == leader) slot0 = atomicAdd(&head[dest], (unsigned long long)__popc(peers));
        slot0 = __shfl_sync(peers, slot0, leader);
        const long long sl = (long long)slot0 + rank;
        if (sl < block_size) {
            int64_t* dst = spawn + ((long long)dest * block_size + sl) * E;
            if (W == 2) {
                reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(child[0], child[1]);
                reinterpret_cast<longlong2*>(dst)[1] = make_longlong2((long long)nspawn, (long long)pflag);
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w) dst[w] = (int64_t)child[w];
                dst[W] = nspawn;
                dst[W + 1] = pflag;
            }
        } else {
            atomicOr(err, 1);  // spawn%error: no space left in the spawning array
        }
    }
}

// K3: double excitations - |slater_condon1| of the orderings i->b, j->a, j->b, the four-ordering generation probability
// (src/excit_gen_heat_bath_mol.F90:408-487), H_ij, spawn
template <int W>
__global__ void __launch_bounds__(256, 3)
k_wf_double(Sys s, Params p, const uint64_t* __restrict__ states, const unsigned char* __restrict__ recD, const Counters* __restrict__ cnt,
            unsigned cap, int64_t* __restrict__ spawn, unsigned long long* __restrict__ head, long long block_size,
            const int* __restrict__ proc_map, int* __restrict__ err) {
    const int nel = s.nel, noccw = (nel + 3) >> 2;
    const int64_t nb = s.nbasis;
    const int strideD = rec_stride(sizeof(RecD), nel, W);
    const unsigned nD = min(cnt->nD, cap);
    const unsigned nround = (nD + 31u) & ~31u;
    const double* siw1 = s.hb_i_w - 1;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < nround; r += gridDim.x * blockDim.x) {
        const bool live = r < nD;
        RecD rd;
        uint32_t occw[MAXOCCW];
        uint64_t f[W];
        Gen g;
        g.allowed = true; g.nexcit = 2; g.pgen = 1.0; g.hmatel = 0.0; g.from1 = g.from2 = g.to1 = g.to2 = 0; g.perm = false;
        if (live) {
            const unsigned char* src = recD + (size_t)r * strideD;
            rd = load_rec<RecD>(src);
            load_tail<W>(src, sizeof(RecD), noccw, f, occw);
            const int i = rd.i, j = rd.j, a = rd.a, b = rd.b;
            // hb_ija%weights(a,j,i) = hb_ijab%weights_tot(a,j,i) = ...(a,i,j), and the same divided by
            // hb_ija%weights_tot(j,i) = ...(i,j)
            const HbRec* ra_ = s.hb_ija_rec + HB_I3(a, j, i);
            const HbRec* rb_ = s.hb_ija_rec + HB_I3(b, j, i);
            const double Ta = ra_->w, ra = ra_->p, Tb = rb_->w, rb = rb_->p;
            double wab, wa;                                             // hb_ijab%weights(b,a,j,i) and it / weights_tot(a,j,i)
            if (rd.flags & RD_WKNOWN) { wab = rd.wab; wa = rd.pab; }
            else {
                const HbRec* rw = s.hb_ijab_rec + HB_I4(b, a, j, i);
                wab = __ldcs(&rw->w); wa = __ldcs(&rw->p);
            }
            const double wij = s.hb_ij_w[HB_I2(j, i)];                  // = ij_weights(i,j)
            // one pass over the occupied list for the three slater_condon1 sums (orderings i->b, j->a, j->b; each in
            // occ_list order) and for ji_weights_occ_tot: 16 independent loads per step instead of four passes
            const bool kb = (rd.flags & RD_KB) != 0, kc = (rd.flags & RD_KC) != 0, kd = (rd.flags & RD_KD) != 0;
            double h0 = kb ? one_body(s, i, b) : 0.0, h1 = kc ? one_body(s, j, a) : 0.0, h2 = kd ? one_body(s, j, b) : 0.0;
            double ji_tot = 0.0;
            {
                const unsigned tA = s.uhf ? (unsigned)(a - 1) : ((unsigned)(a - 1) >> 1);
                const unsigned tB = s.uhf ? (unsigned)(b - 1) : ((unsigned)(b - 1) >> 1);
                const unsigned rl = (unsigned)(s.nbasis + 1);
                const D2* __restrict__ r0 = s.sc1T + ((size_t)((unsigned)(i - 1) * (unsigned)s.sc1A + tB)) * rl;
                const D2* __restrict__ r1 = s.sc1T + ((size_t)((unsigned)(j - 1) * (unsigned)s.sc1A + tA)) * rl;
                const D2* __restrict__ r2 = s.sc1T + ((size_t)((unsigned)(j - 1) * (unsigned)s.sc1A + tB)) * rl;
                const double* __restrict__ col1 = s.hb_ij_w + nb * (j - 1) - 1;
                const D2 zero = {0.0, 0.0};
#pragma unroll 1
                for (int w = 0; w < noccw; ++w) {
                    const uint32_t o4 = occw[w];
                    D2 v0[4], v1[4], v2[4];
                    double cv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t o = (o4 >> (8 * k)) & 0xffu;
                        v0[k] = kb ? r0[o] : zero;
                        v1[k] = kc ? r1[o] : zero;
                        v2[k] = kd ? r2[o] : zero;
                        cv[k] = o ? col1[o] : 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        h0 = h0 + v0[k].x; h0 = h0 - v0[k].y;
                        h1 = h1 + v1[k].x; h1 = h1 - v1[k].y;
                        h2 = h2 + v2[k].x; h2 = h2 - v2[k].y;
                        ji_tot = ji_tot + cv[k];
                    }
                }
            }
            double ps[3] = {0.0, 0.0, 0.0};
            if (kb) { const double hm = fabs(h0); if (hm < Tb) ps[0] = hm / (Tb + hm); else ps[0] = 0.5; }
            if (kc) { const double hm = fabs(h1); if (hm < Ta) ps[1] = hm / (Ta + hm); else ps[1] = 0.5; }
            if (kd) { const double hm = fabs(h2); if (hm < Tb) ps[2] = hm / (Tb + hm); else ps[2] = 0.5; }
            const double pi_ = siw1[i] / rd.i_tot;
            const double pj_ = siw1[j] / rd.i_tot;
            const double pij = wij / rd.ij_tot;
            const double pji = wij / ji_tot;
            const double wb = wab / Tb;
            const double pgen_ija = ((pi_) * (pij)) * ra * (1.0 - rd.psingle) * wa;
            const double pgen_ijb = ((pi_) * (pij)) * rb * (1.0 - ps[0]) * wb;
            const double pgen_jia = ((pj_) * (pji)) * ra * (1.0 - ps[1]) * wa;
            const double pgen_jib = ((pj_) * (pji)) * rb * (1.0 - ps[2]) * wb;
            g.pgen = pgen_ija + pgen_ijb + pgen_jia + pgen_jib;
            g.from1 = (i < j) ? i : j; g.from2 = (i < j) ? j : i;
            g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
            g.perm = excit_perm2<W>(f, g.from1, g.from2, g.to1, g.to2);
            g.hmatel = slater_condon2_excit(s, g.from1, g.from2, g.to1, g.to2, g.perm);
        }
        spawn_and_append<W>(s, p, live, f, occw, g, live ? rd.rs : 0.0, live ? rd.flags : 0, spawn, head, block_size, proc_map, err);
    }
}

// K4: single excitations - the generation probability sums over the spectator orbital (src/excit_gen_heat_bath_mol.F90:
// 489-536), spawn
template <int W>
__global__ void __launch_bounds__(256, 4)
k_wf_single(Sys s, Params p, const uint64_t* __restrict__ states, const unsigned char* __restrict__ recS, const Counters* __restrict__ cnt,
            unsigned cap, int64_t* __restrict__ spawn, unsigned long long* __restrict__ head, long long block_size,
            const int* __restrict__ proc_map, int* __restrict__ err) {
    const int nel = s.nel, noccw = (nel + 3) >> 2;
    const int64_t nb = s.nbasis;
    const int strideS = rec_stride(sizeof(RecS), nel, W);
    const unsigned nS = min(cnt->nS, cap);
    const unsigned nround = (nS + 31u) & ~31u;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < nround; r += gridDim.x * blockDim.x) {
        const bool live = r < nS;
        RecS rs_;
        uint32_t occw[MAXOCCW];
        uint64_t f[W];
        Gen g;
        g.allowed = true; g.nexcit = 1; g.pgen = 1.0; g.hmatel = 0.0; g.from1 = g.from2 = g.to1 = g.to2 = 0; g.perm = false;
        if (live) {
            const unsigned char* src = recS + (size_t)r * strideS;
            rs_ = load_rec<RecS>(src);
            load_tail<W>(src, sizeof(RecS), noccw, f, occw);
            const int i = rs_.i, a = rs_.a;
            const double hmod = fabs(rs_.h_ia);
            const uint8_t* occ = reinterpret_cast<const uint8_t*>(occw);
            const double* ijcol1 = s.hb_ij_w + nb * (i - 1) - 1;
            const double* ijat1 = s.hb_ija_tot + nb * (i - 1) - 1;
            double psum = 0.0;
            (void)occ; (void)ijat1;
            const HbRec* __restrict__ recs = s.hb_ija_rec + HB_I3(a, 1, i) - nb;     // entry (a, oq, i) at recs + nb * oq
#pragma unroll 1
            for (int w = 0; w < noccw; ++w) {       // four spectator orbitals per step: their loads go out together
                const uint32_t o4 = occw[w];
                double Tq[4], pq[4], wq4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t oq = (o4 >> (8 * k)) & 0xffu;
                    const bool on = oq != 0u && oq != (uint32_t)i && oq != (uint32_t)a;
                    const HbRec* rec = recs + nb * (int64_t)oq;
                    Tq[k] = on ? rec->w : 1.0;      // hb_ija%weights(a,oq,i) = hb_ijab%weights_tot(a,oq,i)
                    pq[k] = on ? rec->p : 0.0;      // hb_ija%weights(a,oq,i) / hb_ija%weights_tot(oq,i)
                    wq4[k] = on ? ijcol1[oq] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t oq = (o4 >> (8 * k)) & 0xffu;
                    const bool on = oq != 0u && oq != (uint32_t)i && oq != (uint32_t)a;
                    if (on) {                       // the reference skips the other terms
                        double psq;
                        if (hmod < Tq[k]) psq = hmod / (Tq[k] + hmod); else psq = 0.5;
                        psum = psum + (psq * (wq4[k] / rs_.ij_tot) * pq[k]);
                    }
                }
            }
            g.from1 = i; g.to1 = a;
            g.perm = (rs_.flags & RS_PERM) != 0;
            g.hmatel = rs_.h_ia;
            g.pgen = psum * (s.hb_i_w[i - 1] / rs_.i_tot);
        }
        spawn_and_append<W>(s, p, live, f, occw, g, live ? rs_.rs : 0.0, live ? rs_.flags : 0, spawn, head, block_size, proc_map, err);
    }
}

}  // namespace hbwf
