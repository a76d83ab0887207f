// hande_b200: CCMC kernels (cluster selection / spawning / death), redistribute_particles, find_D0 and the
// excitation-generator probe kernel - included by hb_ccmc_tu.cu, which instantiates them for one W.
#pragma once
#include "hb_common.cuh"

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
                r0.begin(p.seed, p.cycle, RNG_SPAWN, det_hash64<W>(p.f0, HB_NW(p)) + (uint64_t)p.iproc * 0x9E3779B97F4A7C15ull,
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
        rng.begin(p.seed, p.cycle, RNG_SPAWN, det_hash64<W>(p.f0, HB_NW(p)) + (uint64_t)p.iproc * 0x9E3779B97F4A7C15ull,
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
            occ_t occ[HB_MAXNEL]; uint8_t su[64];
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
            // quasi-Newton (src/ccmc_death_spawning.f90:141-144,295): cdet%fock_sum = sum_fock_values_occ_list - ref%fock_sum
            const double dfock = p.qn ? qn_fock_sum(s, p, occ) : 0.0;
            const double invd_s = (p.qn && g.allowed) ? qn_spawned_weighting(p, dfock, g) : 1.0;
            const double hmatel = g.hmatel * cl.amplitude * invd_s * cl.sign;
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
                const double invd = p.qn ? qn_weighting(p, dfock) : 1.0, pc = p.qn ? p.qn_pop_control : 1.0;
                double KiiAi;
                if (cl.nexcitors == 0) KiiAi = ((-pe_old) * invd + (pe_old - p.shift) * pc) * cl.amplitude;
                else if (cl.nexcitors == 1) KiiAi = ((dat[cl.first_pos - 1] - pe_old) * invd + (pe_old - p.shift) * pc) * cl.amplitude;
                else {
                    const double hii = (s.kind == SYS_UEG) ? slater_condon0_ueg(s, occ) : slater_condon0(s, occ);
                    KiiAi = ((hii - p.H00) - pe_old) * invd * cl.amplitude;
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
        const uint64_t h = det_hash64<W>(f, HB_NW(p));
        const bool isD0 = (i + 1 == a.D0_pos);
        PhiloxStream rng;
        double dfock = 0.0;       // quasi-Newton: sum_fock_values_bit_string(f) - ref%fock_sum (0 for the reference)
        if (!isD0) {
            const double amp = (double)pop / (double)p.real_factor;
            const int level = excit_level<W>(f, p.f0);
            const int sign = ccmc_excitor_sign<W>(p.f0, f, level);
            occ_t occ[HB_MAXNEL]; uint8_t su[64];
            decode_det<W>(f, occ);
            if (s.kind == SYS_READ_IN && p.excit_gen != EXCIT_GEN_NO_RENORM && p.excit_gen != EXCIT_GEN_NO_RENORM_SPIN &&
                p.excit_gen != EXCIT_GEN_HEAT_BATH)
                build_symunocc_masks<W>(s, f, su);
            bool is_ref;
            const double hm0 = proj_energy_hmatel<W>(s, p, f, occ, is_ref);
            if (p.qn) dfock = qn_fock_sum(s, p, occ);
            pe = hm0 * (amp * sign / 1.0);
            rng.begin(p.seed, p.cycle, RNG_NATTEMPTS, h, 0);
            const int nsp = decide_nattempts(rng, fabs(amp) / 1.0);
            nas = nsp;
            const double unit = amp / fabs(amp);
            for (int ip = 0; ip < nsp; ++ip) {
                rng.begin(p.seed, p.cycle, RNG_SPAWN, h, (uint32_t)ip);
                Gen g;
                gen_excit<W>(rng, s, p, f, occ, su, g);
                const double invd_s = (p.qn && g.allowed) ? qn_spawned_weighting(p, dfock, g) : 1.0;
                const double hmatel = g.hmatel * unit * invd_s * sign;
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
            const double invd = p.qn ? qn_weighting(p, dfock) : 1.0, pc = p.qn ? p.qn_pop_control : 1.0;
            double KiiAi;
            if (isD0) KiiAi = ((-pe_old) * invd + (pe_old - p.shift) * pc) * (double)pop;
            else KiiAi = ((dat[i] - pe_old) * invd + (pe_old - p.shift) * pc) * (double)pop;
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
// ------------------------------------------------------------------------------------------------
// Probe kernels (parity tests through the C ABI)
// ------------------------------------------------------------------------------------------------
template <int W>
// rn != nullptr: INJECTED random numbers - attempt t draws rn[t*nrn], rn[t*nrn+1], ... in order (the excitation generator
// first, attempt_to_spawn after it) instead of the Philox stream, and nused[t] receives how many it drew: the level-1
// parity hook (a host feeds the numbers its own generator consumed and compares choice, pgen, H_ij and nspawn).
__global__ void k_gen_excit_batch(Sys s, Params p, const uint64_t* __restrict__ states, const int64_t* __restrict__ pops,
                                  const uint32_t* __restrict__ attempt, long long n, const int* __restrict__ proc_map,
                                  int* __restrict__ iout, double* __restrict__ dout, int64_t* __restrict__ nspawn,
                                  const double* __restrict__ rn, int nrn, int* __restrict__ nused) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (rn) {
        uint64_t f[W];
#pragma unroll
        for (int k = 0; k < W; ++k) f[k] = states[t * W + k];
        occ_t occ[HB_MAXNEL]; uint8_t su[64];
        decode_det<W>(f, occ);
        if (s.kind != SYS_UEG) build_symunocc(s, occ, su);
        ListStream rng{rn + t * nrn, nrn, 0};
        Gen g;
        gen_excit<W>(rng, s, p, f, occ, su, g);
        const int64_t ns = attempt_to_spawn(rng, p, g.hmatel, g.pgen, pops[t]);
        int* io = iout + t * 8;
        io[0] = g.nexcit; io[1] = g.from1; io[2] = g.from2; io[3] = g.to1; io[4] = g.to2; io[5] = g.perm; io[6] = g.allowed;
        io[7] = -1;
        dout[t * 2] = g.pgen; dout[t * 2 + 1] = g.hmatel;
        nspawn[t] = ns;
        nused[t] = rng.k;
        return;
    }
    uint64_t f[W];
#pragma unroll
    for (int k = 0; k < W; ++k) f[k] = states[t * W + k];
    occ_t occ[HB_MAXNEL]; uint8_t su[64];
    decode_det<W>(f, occ);
    if (s.kind != SYS_UEG) build_symunocc(s, occ, su);
    PhiloxStream rng;
    rng.begin(p.seed, p.cycle, RNG_SPAWN, det_hash64<W>(f, HB_NW(p)), attempt[t]);
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
