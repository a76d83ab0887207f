// hande_b200: the fused spawn + death + estimators kernel (k_spawn_death) - included by hb_spawn_tu.cu, which
// instantiates it for one (W, generator group).
#pragma once
#include "hb_common.cuh"

// ------------------------------------------------------------------------------------------------
// Kernel: fused spawn + death + estimators  (src/fciqmc.f90:315-371, 635-769)
// ------------------------------------------------------------------------------------------------
constexpr int SINGLES_CHUNK = 64;  // single excitations whose pgen terms are staged in shared memory at a time

// Shared-memory carve-up of k_spawn_death (same arithmetic on host and device).  For the heat-bath generator the
// phase-A staging area (column i of hb_ij_w at the occupied orbitals, [q][thread]) shares its storage with the buffers
// that are only live in the later phases.
struct SpawnSmem {
    size_t sf, shash, ssign, sscan, swarp, sred, siw, sw, sh1, shm, ssp, spsum, sterm, sok, sq, scnt, sflag, slo, sperm, ssq,
        ssi, socc, ssu, sps, sdf, hs_wi, hs_un, hs_q, hs_lo, total;
    // heat_bath: the original heat-bath generator (phase buffers); hb_stage: any generator that selects i, j from the
    // heat-bath weights (needs the hb_i_w copy and the per-thread staging area of nel doubles)
    // ps: pattempt_update statistics are accumulated (per thread: two doubles and two counters)
    // qn: quasi-Newton propagator (fock_sum of each state of the tile)
    // hbs: heat_bath_single (block-cooperative exact single-excitation weights: per batch slot the row sums, the
    // unoccupied list, the queue of single-excitation attempts and their tile states)
    __host__ __device__ SpawnSmem(int W, int nel, int nsu, int nb, bool heat_bath, bool hb_stage, bool ps, bool qn, bool hbs = false) {
        size_t o = 0;
        sf = o;     o += (size_t)TILE * (W > 4 ? W + 1 : W) * 8;   // wide rows are padded by one word (bank conflicts)
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
        o = (o + 3) & ~(size_t)3;
        socc = o;   o += (size_t)TILE * nel * sizeof(occ_t);
        ssu = o;    o += (size_t)TILE * nsu;
        o = (o + 7) & ~(size_t)7;
        sps = o;    o += ps ? (size_t)TILE * 24 : 0;
        sdf = o;    o += qn ? (size_t)TILE * 8 : 0;
        hs_wi = o;  o += hbs ? (size_t)8 * 64 * 8 : 0;
        hs_un = o;  o += hbs ? (size_t)8 * 256 : 0;
        hs_q = o;   o += hbs ? TILE : 0;
        hs_lo = o;  o += hbs ? TILE : 0;
        total = (o + 15) & ~(size_t)15;
    }
};

// GEN: compile-time generator of this instantiation (one kernel per generator keeps the code - and the instruction
// cache footprint - to what the run actually executes): EXCIT_GEN_* for read_in systems, GEN_UEG for the UEG.
template <int W, int GEN>
// wide layout (W = 32): the 64 KB of staged determinants per block allow two blocks per SM anyway - take the registers
__global__ void __launch_bounds__(TILE, (W > 4) ? 2 : ((GEN == EXCIT_GEN_HEAT_BATH || GEN == EXCIT_GEN_HEAT_BATH_UNIFORM) ? 3 : 4) * (256 / TILE))
k_spawn_death(Sys s, Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ pops,
              const double* __restrict__ dat, long long nstates, int64_t* __restrict__ spawn,
              unsigned long long* __restrict__ head, long long block_size, const int* __restrict__ proc_map,
              SpawnPartials* __restrict__ partials, int* __restrict__ err, int tile0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nel = s.nel;
    constexpr int SW = (W > 4) ? W + 1 : W;     // words between the staged determinants of a tile
    constexpr bool heat_bath = (GEN == EXCIT_GEN_HEAT_BATH);
    constexpr bool hb_stage = heat_bath || (GEN == EXCIT_GEN_HEAT_BATH_UNIFORM) || (GEN == EXCIT_GEN_HEAT_BATH_SINGLE) ||
                              (GEN == EXCIT_GEN_POWER_PITZER_OCC_IJ);
    const int nsu = (GEN == EXCIT_GEN_POWER_PITZER_ORDERN) ? nel : (GEN == EXCIT_GEN_RENORM || GEN == EXCIT_GEN_RENORM_SPIN || GEN == EXCIT_GEN_HEAT_BATH_UNIFORM || GEN == EXCIT_GEN_POWER_PITZER_OCC ||
                     GEN == EXCIT_GEN_POWER_PITZER_OCC_IJ) ? 2 * s.nsym_tot : 0;
    const bool ps_on = !heat_bath && p.ps_part != nullptr;
    constexpr bool hbs = (GEN == EXCIT_GEN_HEAT_BATH_SINGLE);
    const SpawnSmem L(W, nel, nsu, s.nbasis, heat_bath, hb_stage, ps_on, p.qn != 0, hbs);
    double* hs_wi = reinterpret_cast<double*>(smem_raw + L.hs_wi);
    uint8_t* hs_un = smem_raw + L.hs_un;
    uint8_t* hs_q = smem_raw + L.hs_q;
    uint8_t* hs_lo = smem_raw + L.hs_lo;
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
    occ_t* socc = reinterpret_cast<occ_t*>(smem_raw + L.socc);
    uint8_t* ssu = smem_raw + L.ssu;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const long long idx = ((long long)blockIdx.x + tile0) * TILE + tid;
    const int E = W + 2;

    double pe = 0.0, d0 = 0.0;
    long long ndeath = 0, npart = 0;
    int natt = 0;
    if (hb_stage)
        for (int k = tid; k < s.nbasis; k += TILE) siw[k] = s.hb_i_w[k];
    if (W > 4) {
        // wide layout: a thread's 256-byte determinant is 8 sectors - the tile is staged with coalesced loads instead
        const long long base = ((long long)blockIdx.x + tile0) * TILE;
        const int nst_tile = (int)min((long long)TILE, nstates - base);
        for (int t = tid; t < nst_tile * W; t += TILE) {
            const int l = t / W, k = t - l * W;
            sf[l * SW + k] = __ldcs(states + base * W + t);
        }
        __syncthreads();
    }
    if (idx < nstates) {
        uint64_t f[W];
        if (W > 4) {
#pragma unroll
            for (int k = 0; k < W; ++k) f[k] = sf[tid * SW + k];
        } else {
            load_det<W>(states + idx * W, f);
        }
        const int64_t pop = __ldcs(pops + idx);
        const double Kii = __ldcs(dat + idx);
#pragma unroll
        for (int k = 0; k < W; ++k) sf[tid * SW + k] = f[k];
        ssign[tid] = pop < 0;
        occ_t* occ = socc + tid * nel;
        decode_det<W>(f, occ);
        if (GEN == EXCIT_GEN_POWER_PITZER_ORDERN) find_diff_ref_cdet<W>(s, p.f0, f, ssu + tid * nsu);
        else if (nsu) build_symunocc_masks<W>(s, f, ssu + tid * nsu);
        const uint64_t h = det_hash64<W>(f, HB_NW(p));
        if (!hb_stage) shash[tid] = h;
        const double real_pop = (double)pop / (double)p.real_factor;
        // set_determ_info (src/semi_stoch.F90:826-857): bit 1 of the staged flag marks a deterministic parent
        const bool determ_parent = p.ss_bits != nullptr && ((p.ss_bits[idx >> 5] >> (idx & 31)) & 1u);
        // set_parent_flag (src/ifciqmc.f90:13-57): deterministic states are always initiators
        sflag[tid] = ((fabs(real_pop) > p.initiator_pop || determ_parent) ? 0 : 1) | (determ_parent ? 2 : 0);
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
        // no death step for deterministic states (src/fciqmc.f90:368): the projection carries their diagonal
        int64_t newpop = pop;
        kill_abs = 0;
        if (!determ_parent) newpop = stochastic_death(rng, p, Kii, pop, kill_abs, death_weight);
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
        for (int k = 0; k < W; ++k) f[k] = sf[lo * SW + k];
        PhiloxStream rng;
        rng.begin(p.seed, p.cycle, RNG_SPAWN, hb_stage ? det_hash64<W>(f, HB_NW(p)) : shash[lo], (uint32_t)att);
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
                for (int k = 0; k < W; ++k) ff[k] = sf[l * SW + k];
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
                for (int kk = 0; kk < W; ++kk) ff[kk] = sf[l * SW + kk];
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
            hb_phase_f<W>(s, f, socc + lo * nel, st, hmk, psum, siw, g);
        } else if (hbs) {
            // ---- heat_bath_single: the double excitations are generated per thread; the single excitations (probability
            // pattempt_single) need |<D|H|D_i^a>| for ALL nel x nvirt (i, a) pairs of their determinant
            // (find_ia_single_weights, src/excit_gen_utils.f90:162-218; the reference caches them per determinant).
            // They are queued and the whole block evaluates the weights of a batch of them, one weight per thread, into
            // the staging area; row sums and the two alias selections follow in the reference's order.
            if (tid == 0) scnt[0] = 0;
            __syncthreads();
            bool single = false;
            if (active) {
                single = rng.next() < p.pattempt_single;
                if (single) {
                    const int q = atomicAdd(&scnt[0], 1);
                    hs_q[q] = (uint8_t)tid;
                    hs_lo[tid] = (uint8_t)lo;
                } else {
                    gen_double_heat_bath_uniform<W>(rng, s, p, f, socc + lo * nel, siw, sw + tid, TILE, g);
                }
            }
            __syncthreads();       // the staging area is free again; the queue is complete
            const int nq = scnt[0];
            const int nvirt = s.nbasis - nel, npair = nel * nvirt;
            const int B = max(1, min(8, (TILE * nel) / npair));
            for (int q0 = 0; q0 < nq; q0 += B) {
                const int nbat = min(B, nq - q0);
                for (int t = tid; t < nbat * nvirt; t += TILE) {
                    const int b = t / nvirt, k = t - b * nvirt, l = hs_lo[hs_q[q0 + b]];
                    uint64_t ff[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) ff[w] = sf[l * SW + w];
                    hs_un[b * 256 + k] = (uint8_t)nth_unocc<W>(ff, k + 1);
                }
                __syncthreads();
                for (int t = tid; t < nbat * npair; t += TILE) {
                    const int b = t / npair, r = t - b * npair, q = r / nvirt, k = r - q * nvirt, l = hs_lo[hs_q[q0 + b]];
                    sw[t] = hb_single_weight(s, socc + l * nel, socc[l * nel + q], hs_un[b * 256 + k]);
                }
                __syncthreads();
                for (int t = tid; t < nbat * nel; t += TILE) {
                    const int b = t / nel, q = t - b * nel;
                    const double* row = sw + b * npair + q * nvirt;
                    double wsum = 0.0;
                    for (int k = 0; k < nvirt; ++k) wsum = wsum + row[k];
                    hs_wi[b * 64 + q] = wsum;
                }
                __syncthreads();
                if (active && single)
                    for (int b = 0; b < nbat; ++b)
                        if (hs_q[q0 + b] == tid)
                            gen_single_heat_bath_select<W>(rng, s, p, f, socc + lo * nel, hs_wi + b * 64, sw + b * npair,
                                                           hs_un + b * 256, g);
                __syncthreads();
            }
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
            hmq = hmq * p.cheby_weight;                                              // (src/spawning.F90:117-118)
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
                } else if ((sflag[lo] & 2) && ss_check_if_determ<W>(p.ss_sorted, p.ss_tot, child)) {
                    nspawn = 0;   // deterministic -> deterministic: the projection's job (src/fciqmc.f90:726-736)
                } else {
                    // assign_particle_processor (src/spawning.F90:770-838)
                    dest = (p.nprocs > 1) ? proc_map[owner_slot(child, s.nbasis, p.hash_seed, p.nprocs, p.nslots)] : 0;
                    pflag = p.initiator ? (sflag[lo] & 1) : 0;
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
            p.ps_part[tile0 + blockIdx.x] = out;
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
        partials[tile0 + blockIdx.x] = out;
    }
}
