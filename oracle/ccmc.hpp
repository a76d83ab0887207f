// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the reference's CCMC propagation (stochastic cluster selection, no full_nc / linked / even
// selection / multi-reference / quasi-Newton; real orbitals):
//   driver            src/ccmc.f90:603-896 (do_ccmc icycle body), :897-905 (estimators), :1007-1273,:1362-1455
//   selection         src/ccmc_selection.f90:93-392 (select_cluster), :394-460 (create_null_cluster),
//                     :874-948 (set_cluster_selections)
//   cluster algebra   src/ccmc_utils.F90:69-130 (get_D0_info), :132-295 (collapse_cluster),
//                     :296-412 (convert_excitor_to_determinant), :427-590 (cumulative_population)
//   spawning / death  src/ccmc_death_spawning.f90:11-211 (spawner_ccmc), :213-361 (stochastic_ccmc_death),
//                     :363-441 (stochastic_death_attempt), :944-1014 (create_spawned_particle_ccmc)
//   cycle bookkeeping src/qmc_common.F90:950-1017 (init_mc_cycle, ccmc branch), :1240-1304 (end_mc_cycle),
//                     :799-925 (initial_cc_projected_energy: D0 only at the start of a calculation)
//   utilities         src/search.f90:378-481 (binary_search_real_p), src/sort.f90:827-852 (insert_sort_real_p)
// Random numbers are consumed in the reference's order: per cluster 1 (size) + nexcitors (excips), then the
// excitation generator's draws, 1 for attempt_to_spawn and 1 for the death attempt (if a death is attempted).
#pragma once
#include <complex>
#include "fciqmc.hpp"

namespace oracle {

struct Cluster {
    int nexcitors = 0;
    int excitation_level = 0;          // INT_MAX == huge(0): not allowed
    double pselect = 1.0;
    double amplitude = 0.0;
    int cluster_to_det_sign = 1;
    int64_t first_pos = 0;             // position (1-based) of the first excitor: cdet%data => psip_list%dat(:,pos)
};

struct CcmcStats {   // what the device engine is compared against, per cycle
    int64_t nattempts = 0, nattempts_spawn = 0, nspawn_events = 0, ndeath = 0, ndeath_nc = 0;
    double proj_energy = 0.0, D0_population = 0.0, D0_normalisation = 0.0;
};

struct OracleCcmc : Oracle {
    int move_freq = 5;                 // ccmc_in%move_freq (default)
    bool full_nc = false;              // ccmc_in%full_nc (full_non_composite)
    int hash_shift = 0;                // spawn%hash_shift: +1 per cycle (src/ccmc.f90:540,625)
    int64_t nattempts_last = 0;
    std::vector<double> cumulative_abs_real_pops;
    CcmcStats last;

    static void insert_sort(double* list, int n) {
        for (int i = 2; i <= n; ++i) {
            int j = i - 1;
            double tmp = list[i - 1];
            while (j >= 1) {
                if (list[j - 1] <= tmp) break;
                list[j] = list[j - 1];
                j--;
            }
            list[j] = tmp;
        }
    }
    // binary_search_real_p (src/search.f90:378-481); list is 1-based through L(i)
    static void binary_search_real(const std::vector<double>& list, double item, int istart, int iend, bool& hit, int& pos) {
        auto L = [&](int i) { return list[i - 1]; };
        if (istart > iend) { pos = istart; hit = false; return; }
        int lo = istart, hi = iend;
        hit = false;
        while (hi != lo) {
            pos = (hi + lo) / 2;
            double compare = item - L(pos);
            if (std::fabs(compare) < depsilon) { hit = true; break; }
            else if (compare > 0.0) lo = pos + 1;
            else hi = pos;
        }
        if (hi == lo) {
            double compare = item - L(hi);
            if (std::fabs(compare) < depsilon) { hit = true; pos = hi; }
            else if (compare > 0.0) pos = hi + 1;
            else pos = hi;
        }
    }

    // collapse_cluster + collapse_excitor_onto_cluster (src/ccmc_utils.F90:132-295)
    void collapse_cluster(const Det& excitor, double excitor_population, Det& cluster_excitor, double& cluster_population,
                          bool& allowed) const {
        const int W = sys.W;
        uint64_t ee[MAXW], ce[MAXW], ea[MAXW], ca[MAXW], ec[MAXW], cc[MAXW];
        bool clash = false;
        for (int k = 0; k < W; ++k) {
            ee[k] = f0.w[k] ^ excitor.w[k];
            ce[k] = f0.w[k] ^ cluster_excitor.w[k];
            ea[k] = ee[k] & f0.w[k];
            ca[k] = ce[k] & f0.w[k];
            ec[k] = ee[k] & excitor.w[k];
            cc[k] = ce[k] & cluster_excitor.w[k];
            if ((ec[k] & cc[k]) != 0 || (ea[k] & ca[k]) != 0) clash = true;
        }
        if (clash) {
            allowed = false;
            cluster_population = cluster_population * excitor_population;
            return;
        }
        allowed = true;
        cluster_population = cluster_population * excitor_population;
        for (int ib = 0; ib < W; ++ib)
            for (int bit = 0; bit < 64; ++bit) {
                if (!((ee[ib] >> bit) & 1ull)) continue;
                const int orb = ib * 64 + bit + 1;
                uint64_t mask[MAXW], perm[MAXW];
                sys.excit_mask(orb, mask);
                if ((f0.w[ib] >> bit) & 1ull) {
                    cluster_excitor.w[ib] &= ~(1ull << bit);
                    for (int k = 0; k < W; ++k) perm[k] = (mask[k] & ca[k]) | cc[k];
                } else {
                    cluster_excitor.w[ib] |= (1ull << bit);
                    for (int k = 0; k < W; ++k) perm[k] = (~mask[k]) & cc[k];
                    perm[ib] &= ~(1ull << bit);
                }
                int n = 0;
                for (int k = 0; k < W; ++k) n += __builtin_popcountll(perm[k]);
                if (n % 2 == 1) cluster_population = -cluster_population;
            }
    }
    // convert_excitor_to_determinant (src/ccmc_utils.F90:296-412)
    int convert_excitor_to_determinant(const Det& excitor, int excitor_level) const {
        int nann = excitor_level, ncre = excitor_level, sign = 1;
        for (int ib = 0; ib < sys.W; ++ib) {
            const uint64_t ex = f0.w[ib] ^ excitor.w[ib];
            for (int bit = 0; bit < 64; ++bit) {
                if ((f0.w[ib] >> bit) & 1ull) {
                    if ((ex >> bit) & 1ull) nann--;
                    else if ((nann + ncre) % 2 == 1) sign = -sign;
                } else if ((ex >> bit) & 1ull) {
                    ncre--;
                }
            }
        }
        return sign;
    }

    // create_null_cluster (src/ccmc_selection.f90:394-460)
    void create_null_cluster(double prob, double D0_normalisation, DetInfo& cdet, Cluster& cl) const {
        cl.pselect = prob;
        cl.nexcitors = 0;
        cdet.initiator_flag = (std::fabs(D0_normalisation) <= in.initiator_pop) ? 3 : 0;
        cl.excitation_level = 0;
        cl.amplitude = D0_normalisation;
        cl.cluster_to_det_sign = 1;
        cl.first_pos = 0;
        decode_for(sys, EG(), f0, cdet);
    }

    // select_cluster (src/ccmc_selection.f90:93-392), linked_ccmc = false, discard_threshold = huge
    void select_cluster(Rng& rng, const RankState& r, int ex_level, int64_t nattempts, double normalisation,
                        double tot_excip_pop, int min_size, int max_size, DetInfo& cdet, Cluster& cl) const {
        cl.pselect = (double)(nattempts * in.nprocs);
        double rand = rng.next();
        double psize = 0.0;
        cl.nexcitors = -1;
        for (int i = 0; i <= max_size - min_size - 1; ++i) {
            psize = psize + 1.0 / (double)(1ll << (i + 1));
            if (rand < psize) {
                cl.nexcitors = i + min_size;
                cl.pselect = cl.pselect / (double)(1ll << (i + 1));
                break;
            }
        }
        if (cl.nexcitors == -1) {
            cl.nexcitors = max_size;
            cl.pselect = cl.pselect * (1.0 - psize);
        }
        cdet.initiator_flag = 0;
        bool allowed = min_size <= max_size;
        if (cl.nexcitors == 0) {
            create_null_cluster(cl.pselect, normalisation, cdet, cl);
            return;
        }
        double pop[16];
        for (int i = 0; i < cl.nexcitors; ++i) pop[i] = rng.next() * tot_excip_pop;
        insert_sort(pop, cl.nexcitors);
        int prev_pos = 1;
        double cluster_population = 0.0;
        Det cf;
        for (int i = 1; i <= cl.nexcitors; ++i) {
            bool hit;
            int pos;
            binary_search_real(cumulative_abs_real_pops, pop[i - 1], prev_pos, (int)r.nstates, hit, pos);
            for (;;) {
                if (pos == 1) break;
                if (std::fabs(cumulative_abs_real_pops[pos - 1] - cumulative_abs_real_pops[pos - 2]) > depsilon) break;
                pos = pos - 1;
            }
            const double excitor_pop = (double)r.pops[pos - 1] / (double)pop_real_factor;
            if (i == 1) {
                cf = r.states[pos - 1];
                cl.first_pos = pos;
                cluster_population = excitor_pop;
                cl.pselect = cl.pselect / in.nprocs;
            } else {
                collapse_cluster(r.states[pos - 1], excitor_pop, cf, cluster_population, allowed);
                if (!allowed) break;
                if (pos != prev_pos) cl.pselect = cl.pselect / in.nprocs;
            }
            if (std::fabs(excitor_pop) <= in.initiator_pop) cdet.initiator_flag = 3;
            cl.pselect = (cl.pselect * std::fabs(excitor_pop)) / tot_excip_pop;
            prev_pos = pos;
        }
        if (allowed) {
            cl.excitation_level = sys.excitation_level(f0, cf);
            allowed = cl.excitation_level <= ex_level + 2;
        }
        if (allowed) {
            static const double fact[13] = {1, 1, 2, 6, 24, 120, 720, 5040, 40320, 362880, 3628800, 39916800, 479001600};
            cl.pselect = cl.pselect * fact[cl.nexcitors];
            cl.cluster_to_det_sign = convert_excitor_to_determinant(cf, cl.excitation_level);
            decode_for(sys, EG(), cf, cdet);
            double norm_pow = 1.0;   // normalisation**(nexcitors-1), integer power by repeated multiplication
            for (int k = 0; k < cl.nexcitors - 1; ++k) norm_pow = norm_pow * normalisation;
            cl.amplitude = cluster_population / norm_pow;
        } else {
            cl.excitation_level = INT32_MAX;
        }
    }

    // cumulative_population (src/ccmc_utils.F90:427-563), calc_dist = false, real populations
    double cumulative_population(const RankState& r, int D0_proc, int D0_pos) {
        const int64_t n = r.nstates;
        cumulative_abs_real_pops.assign((size_t)std::max<int64_t>(n, 1), 0.0);
        auto contrib = [&](int64_t i) { return std::fabs((double)r.pops[i - 1]) / (double)pop_real_factor; };
        auto& c = cumulative_abs_real_pops;
        if (n == 0) return 0.0;
        c[0] = contrib(1);
        if (D0_proc == r.iproc) {
            for (int64_t i = 2; i <= D0_pos - 1; ++i) c[i - 1] = c[i - 2] + contrib(i);
            if (D0_pos == 1) c[0] = 0.0;
            if (D0_pos > 1) c[D0_pos - 1] = c[D0_pos - 2];
            for (int64_t i = D0_pos + 1; i <= n; ++i) c[i - 1] = c[i - 2] + contrib(i);
        } else {
            for (int64_t i = 2; i <= n; ++i) c[i - 1] = c[i - 2] + contrib(i);
        }
        return c[n - 1];
    }

    // assign_particle_processor with spawn%hash_shift / spawn%move_freq (src/spawning.F90:770-838)
    int owner_ccmc(const Det& f) const {
        return assign_particle_processor(f, sys.nbasis, hash_seed, hash_shift, move_freq, in.nprocs, proc_map.data(), in.nslots);
    }
    // redistribute_particles (src/qmc_common.F90:505-595)
    void redistribute_particles(RankState& r) {
        const int64_t block_size = in.spawned_walker_length / in.nprocs;
        double nsent = 0.0;
        for (int64_t i = 0; i < r.nstates; ++i) {
            const int pproc = owner_ccmc(r.states[i]);
            if (pproc != r.iproc) {
                if ((int64_t)r.send[pproc].size() + 1 > block_size) { r.spawn_error = true; }
                else { SpawnElem e; e.f = r.states[i]; e.pop = r.pops[i]; e.flag = 0; r.send[pproc].push_back(e); }
                nsent = nsent + std::fabs((double)r.pops[i]);
                r.pops[i] = 0;
            }
        }
        nsent = nsent / (double)pop_real_factor;
        r.nparticles = r.nparticles - nsent;
    }

    // add a particle to the spawn list of rank r through create_spawned_particle_truncated / create_spawned_particle
    // (src/spawning.F90:1074-1319) - same rule as the FCIQMC path of this oracle
    void add_spawn(RankState& r, const Det& fnew, int64_t nspawned) {
        if (in.ex_level >= 0 && sys.excitation_level(f0, fnew) > ref_ex_level) return;
        const int64_t block_size = in.spawned_walker_length / in.nprocs;
        int dest = owner_ccmc(fnew);
        if ((int64_t)r.send[dest].size() + 1 > block_size) { r.spawn_error = true; return; }
        SpawnElem e;
        e.f = fnew; e.pop = nspawned; e.flag = 0;
        r.send[dest].push_back(e);
    }

    // one MC cycle on one rank up to (not including) annihilation: src/ccmc.f90:625-857
    // get_D0_info (src/ccmc_utils.F90:69-130) with the current hash_shift; the population is broadcast from D0_proc
    void get_D0_info(int& D0_proc, int& D0_pos, double& D0_normalisation) const {
        D0_proc = owner_ccmc(f0);
        const RankState& r = ranks[D0_proc];
        bool hit;
        int pos;
        binary_search(r.states, f0, 1, (int)r.nstates, sys.W, hit, pos);
        if (!hit) throw std::runtime_error("find_D0: Cannot find reference!");
        D0_pos = pos;
        D0_normalisation = (double)r.pops[D0_pos - 1] / (double)pop_real_factor;
    }
    void ccmc_cycle_rank(RankState& r, uint32_t cycle_id, int D0_proc, int D0_pos_in, double D0_normalisation) {
        Rng& rng = *r.rng;
        rng.set_cycle(cycle_id);
        const int nD0_proc = (r.iproc == D0_proc) ? 1 : 0;
        const int D0_pos = nD0_proc ? D0_pos_in : -1;
        const int max_cluster_size = (int)std::min<int64_t>(std::min(sys.nel, ref_ex_level + 2), r.nstates - nD0_proc);
        // init_mc_cycle (src/qmc_common.F90:950-1017), ccmc branch
        for (auto& b : r.send) b.clear();
        r.ndeath = 0;
        int64_t nattempts = (int64_t)r.nparticles;
        nattempts = std::max<int64_t>(nattempts, (int64_t)std::llround(std::fabs(D0_normalisation)));
        r.nattempts = nattempts;
        const double tot_abs_real_pop = cumulative_population(r, D0_proc, D0_pos);
        // set_cluster_selections (src/ccmc_selection.f90:874-948)
        int min_cluster_size = 0;
        int64_t nD0_select = 0, nstochastic_clusters = nattempts, nsingle_excitors = 0;
        if (full_nc) {
            min_cluster_size = 2;
            nD0_select = (int64_t)std::llround(std::fabs(D0_normalisation));
            nstochastic_clusters = (int64_t)std::ceil(tot_abs_real_pop);
            nsingle_excitors = r.nstates;
            nattempts = (int64_t)std::llround(tot_abs_real_pop) + nD0_select + nstochastic_clusters;
            r.nattempts = nattempts;
        }
        int64_t nattempts_spawn = 0;
        int64_t ndeath_nc = 0;
        double proj_energy_cycle = 0.0, D0_population_cycle = 0.0;
        DetInfo cdet;
        Cluster cl;
        PsColl ps_stats;   // zero_ps_stats (src/ccmc_data.F90:184-207): per cycle
        // do_ccmc_accumulation (src/ccmc.f90:1007-1101)
        auto accumulate = [&]() {
            double d0 = 0.0, pe = 0.0;
            update_proj_energy(cdet, cl.amplitude * cl.cluster_to_det_sign / cl.pselect, d0, pe);
            D0_population_cycle = D0_population_cycle + d0;
            proj_energy_cycle = proj_energy_cycle + pe;
        };
        // perform_ccmc_spawning_attempt -> spawner_ccmc (src/ccmc_death_spawning.f90:11-211)
        auto spawn_attempt = [&](int nspawnings_total) {
            GenResult g = gen_excit_sys(rng, sys, EG(), cdet);
            double hmatel = g.hmatel;
            // calc_qn_spawned_weighting(propagator, cdet%fock_sum, connection) for an allowed excitation, else 1
            // (src/ccmc_death_spawning.f90:141-144); cdet%fock_sum = sum_fock_values_occ_list - ref%fock_sum (src/ccmc.f90:747-749)
            const double invdiagel = g.allowed ? qn_spawned_weighting(qn ? fock_sum_of(cdet.occ) : 0.0, g.conn) : 1.0;
            hmatel = hmatel * cl.amplitude * invdiagel * cl.cluster_to_det_sign;
            const double pgen = g.pgen * cl.pselect * nspawnings_total;
            if (g.allowed && vary_psingles) {   // update_p_single_double_data (src/spawning.F90:2139-2215)
                if (g.conn.nexcit == 1) {
                    ps_stats.h_pgen_singles_sum = ps_stats.h_pgen_singles_sum + ((std::fabs(hmatel) * eg.pattempt_single) / pgen);
                    ps_stats.excit_gen_singles = ps_stats.excit_gen_singles + 1.0;
                } else if (g.conn.nexcit == 2) {
                    ps_stats.h_pgen_doubles_sum = ps_stats.h_pgen_doubles_sum + ((std::fabs(hmatel) * eg.pattempt_double) / pgen);
                    ps_stats.excit_gen_doubles = ps_stats.excit_gen_doubles + 1.0;
                }
            }
            int64_t nspawn = attempt_to_spawn(rng, hmatel, pgen, 1);
            if (nspawn != 0) {
                Det fexcit = sys.create_excited_det(cdet.f, g.conn);
                const int excitor_level = sys.excitation_level(f0, fexcit);
                const int excitor_sign = convert_excitor_to_determinant(fexcit, excitor_level);
                if (excitor_sign < 0) nspawn = -nspawn;
                add_spawn(r, fexcit, nspawn);   // create_spawned_particle_ccmc
            }
        };
        // non-composite clusters, one per excitor (full_nc): select_nc_cluster (src/ccmc_selection.f90:462-561) +
        // do_nc_ccmc_propagation (src/ccmc.f90:1275-1360)
        for (int64_t iexcitor = 1; iexcitor <= nsingle_excitors; ++iexcitor) {
            if (iexcitor == D0_pos) continue;
            const Det& fx = r.states[iexcitor - 1];
            cl.pselect = 1.0;
            cl.nexcitors = 1;
            cl.first_pos = iexcitor;
            const double excitor_pop = (double)r.pops[iexcitor - 1] / (double)pop_real_factor;
            cl.excitation_level = sys.excitation_level(f0, fx);
            cl.amplitude = excitor_pop;
            cl.cluster_to_det_sign = convert_excitor_to_determinant(fx, cl.excitation_level);
            decode_for(sys, EG(), fx, cdet);
            accumulate();
            rng.begin(RNG_NATTEMPTS, fx, sys.W, 0);
            const int nspawnings_cluster = decide_nattempts(rng, std::fabs(cl.amplitude) / cl.pselect);
            nattempts_spawn += nspawnings_cluster;
            cl.amplitude = cl.amplitude / std::fabs(cl.amplitude);
            for (int ip = 0; ip < nspawnings_cluster; ++ip) {
                rng.begin(RNG_SPAWN, fx, sys.W, (uint32_t)ip);
                spawn_attempt(1);
            }
        }
        for (int64_t iattempt = 1; iattempt <= nstochastic_clusters; ++iattempt) {
            rng.begin(RNG_SPAWN, f0, sys.W, (uint32_t)iattempt);
            rng.mix((uint64_t)r.iproc * 0x9E3779B97F4A7C15ull);
            select_cluster(rng, r, ref_ex_level, nstochastic_clusters, D0_normalisation, tot_abs_real_pop, min_cluster_size,
                           max_cluster_size, cdet, cl);
            if (!(cl.excitation_level <= ref_ex_level + 2)) continue;
            accumulate();
            // do_stochastic_ccmc_propagation (src/ccmc.f90:1103-1191), cluster_multispawn_threshold = huge => 1 attempt
            const int nspawnings_cluster = 1;
            nattempts_spawn += nspawnings_cluster;
            const bool attempt_death = cl.excitation_level <= ref_ex_level;
            spawn_attempt(nspawnings_cluster);
            if (attempt_death && (cl.nexcitors >= 2 || !full_nc)) {
                // stochastic_ccmc_death (src/ccmc_death_spawning.f90:213-361), not linked
                const double pe_old = est.proj_energy_old;
                double KiiAi;
                // invdiagel = calc_qn_weighting(propagator, cdet%fock_sum) (src/ccmc_death_spawning.f90:295)
                const double invd = qn ? qn_weighting(fock_sum_of(cdet.occ)) : 1.0;
                if (cl.nexcitors == 0) KiiAi = ((-pe_old) * invd + (pe_old - shift) * qn_pop_control) * cl.amplitude;
                else if (cl.nexcitors == 1) KiiAi = ((r.dat[cl.first_pos - 1] - pe_old) * invd + (pe_old - shift) * qn_pop_control) * cl.amplitude;
                else KiiAi = ((diag_hmatel(sys, cdet.f) - H00) - pe_old) * invd * cl.amplitude;
                KiiAi = 1.0 * (double)pop_real_factor * KiiAi;
                KiiAi = KiiAi * tau / cl.pselect;
                // stochastic_death_attempt (:363-441)
                double pdeath = std::fabs(KiiAi);
                int64_t nkill;
                if (pdeath < (double)spawn_cutoff) {
                    nkill = (pdeath > rng.next() * (double)spawn_cutoff) ? spawn_cutoff : 0;
                } else {
                    nkill = (int64_t)pdeath;
                    pdeath = pdeath - (double)nkill;
                    if (pdeath > rng.next()) nkill++;
                }
                if (nkill != 0) {
                    if (KiiAi > 0) nkill = -nkill;
                    add_spawn(r, cdet.f, nkill);
                }
                r.ndeath += (nkill < 0 ? -nkill : nkill);
            }
        }
        // deterministic selections of the reference (full_nc; src/ccmc.f90:803-821)
        for (int64_t k = 1; k <= nD0_select; ++k) {
            if (k == 1) create_null_cluster((double)in.nprocs * (double)nD0_select, D0_normalisation, cdet, cl);
            accumulate();
            nattempts_spawn += 1;
            rng.begin(RNG_SPAWN, f0, sys.W, (uint32_t)(nstochastic_clusters + k));
            rng.mix((uint64_t)r.iproc * 0x9E3779B97F4A7C15ull);
            spawn_attempt(1);
        }
        // stochastic_ccmc_death_nc in place on every excip incl. the reference (full_nc; src/ccmc.f90:823-841,
        // src/ccmc_death_spawning.f90:443-547)
        if (full_nc) {
            double nparticles_change = 0.0;
            for (int64_t i = 1; i <= r.nstates; ++i) {
                const bool isD0 = (i == D0_pos);
                const double pe_old = est.proj_energy_old;
                int64_t& population = r.pops[i - 1];
                double KiiAi;
                // dfock = sum_fock_values_bit_string(states(:,i)) - ref%fock_sum (src/ccmc.f90:839-842);
                // invdiagel = calc_qn_weighting(propagator, dfock) (src/ccmc_death_spawning.f90:504)
                double invd = 1.0;
                if (qn) {
                    DetInfo dd;
                    decode_det_occ(sys, r.states[i - 1], dd);
                    invd = qn_weighting(fock_sum_of(dd.occ));
                }
                if (isD0) KiiAi = ((-pe_old) * invd + (pe_old - shift) * qn_pop_control) * (double)population;
                else KiiAi = ((r.dat[i - 1] - pe_old) * invd + (pe_old - shift) * qn_pop_control) * (double)population;
                const int64_t old_pop = population;
                KiiAi = KiiAi * 1.0;
                double pdeath = tau * std::fabs(KiiAi);
                int64_t nkill = (int64_t)pdeath;
                pdeath = pdeath - (double)nkill;
                rng.begin(RNG_DEATH, r.states[i - 1], sys.W, 0);
                if (pdeath > rng.next()) nkill = nkill + 1;
                if (nkill != 0) {
                    if (KiiAi > 0) nkill = -nkill;
                    population = population + nkill;
                    nparticles_change = nparticles_change +
                        (double)((population < 0 ? -population : population) - (old_pop < 0 ? -old_pop : old_pop)) / (double)pop_real_factor;
                    ndeath_nc += (nkill < 0 ? -nkill : nkill);
                }
            }
            r.nparticles = r.nparticles + nparticles_change;
        }
        r.D0_population = r.D0_population + D0_population_cycle;
        r.proj_energy = r.proj_energy + proj_energy_cycle;
        int ev = 0;
        for (auto& b : r.send) ev += (int)b.size();
        r.nspawn_events = ev;
        last.nattempts = nattempts; last.nattempts_spawn = nattempts_spawn; last.nspawn_events = ev; last.ndeath = r.ndeath;
        if (vary_psingles) {   // ps_stats_reduction_update (src/ccmc_data.F90:209-242)
            if ((int)ps_rep_accum.size() != in.nprocs) ps_rep_accum.assign(in.nprocs, PsColl());
            PsColl& a = ps_rep_accum[r.iproc];
            a.h_pgen_singles_sum = a.h_pgen_singles_sum + ps_stats.h_pgen_singles_sum;
            a.excit_gen_singles = a.excit_gen_singles + ps_stats.excit_gen_singles;
            a.h_pgen_doubles_sum = a.h_pgen_doubles_sum + ps_stats.h_pgen_doubles_sum;
            a.excit_gen_doubles = a.excit_gen_doubles + ps_stats.excit_gen_doubles;
        }
        last.proj_energy = proj_energy_cycle; last.D0_population = D0_population_cycle; last.D0_normalisation = D0_normalisation;
        // end_mc_cycle(nspawn_events, ndeath_nc, real_factor, nattempts_spawn, rspawn): spawning_rate
        r.rspawn = r.rspawn + ((nattempts_spawn > 0)
                                   ? ((double)ev + (double)ndeath_nc / (double)pop_real_factor) / (double)nattempts_spawn : 0.0);
        last.ndeath_nc = ndeath_nc;
    }
    static double invdiag() { return 1.0; }

    // spawn/death stage of one cycle on every rank (src/ccmc.f90:625-857) incl. redistribute_particles (:868-869)
    void ccmc_spawn_stage(uint32_t cycle_id) {
        int D0_proc, D0_pos;
        double D0_normalisation;
        get_D0_info(D0_proc, D0_pos, D0_normalisation);
        hash_shift = hash_shift + 1;
        int64_t na = 0;
        for (auto& r : ranks) { ccmc_cycle_rank(r, cycle_id, D0_proc, D0_pos, D0_normalisation); na += r.nattempts; }
        nattempts_last = na;
        if (in.nprocs > 1)
            for (auto& r : ranks) redistribute_particles(r);
    }
    void ccmc_cycle(uint32_t cycle_id) {
        ccmc_spawn_stage(cycle_id);
        comm_spawn();
        for (auto& r : ranks) annihilate_rank(r);
    }

    void ccmc_report_loop(int ireport) {
        est.proj_energy_old = (std::fabs(est.D0_population) < std::numeric_limits<double>::min())
                                  ? 0.0 : est.proj_energy / est.D0_population;
        est.D0_population_old = est.D0_population;
        for (auto& r : ranks) { r.rspawn = 0.0; r.proj_energy = 0.0; r.D0_population = 0.0; }
        for (int icycle = 1; icycle <= in.ncycles; ++icycle) {
            int iter = mc_cycles_done + (ireport - 1) * in.ncycles + icycle;
            ccmc_cycle((uint32_t)iter);
        }
        double pe = 0.0, d0 = 0.0, rsp = 0.0, ntot = 0.0;
        int64_t nst = 0, nev = 0;
        bool err = false;
        for (auto& r : ranks) {
            pe += r.proj_energy; d0 += r.D0_population; rsp += r.rspawn; ntot += r.nparticles;
            nst += r.nstates; nev += r.nspawn_events;
            err = err || r.spawn_error || r.psip_error;
        }
        est.proj_energy = pe / (in.ncycles * 1);
        est.D0_population = d0 / (in.ncycles * 1);
        rspawn_report = rsp / (in.ncycles * 1 * in.nprocs);
        est.tot_nstates = nst; est.tot_nspawn_events = nev;
        if (vary_shift) update_shift(ntot_particles_old, ntot, in.ncycles);
        est.D0_population_old = est.D0_population;
        ntot_particles_old = ntot;
        if (!vary_shift && ntot > in.target_particles) {
            vary_shift = true;
            if (in.vary_shift_from_proje) shift = est.proj_energy / est.D0_population;
            else shift = in.vary_shift_from;
        }
        error = err;
        end_report_loop_pattempt();
        ReportRow row;
        row.iter = mc_cycles_done + ireport * in.ncycles;
        row.shift = shift; row.proj_energy = est.proj_energy; row.D0_population = est.D0_population;
        row.nparticles = ntot_particles_old; row.nstates = est.tot_nstates; row.nspawn_events = est.tot_nspawn_events;
        row.rspawn = rspawn_report;
        rows.push_back(row);
        nattempts_rows.push_back(nattempts_last);
    }
    std::vector<int64_t> nattempts_rows;


    void run_ccmc() {
        // initial_cc_projected_energy (src/qmc_common.F90:799-925) for a population on the reference only:
        // proj_energy = 0, D0_population = D0_normalisation (its separate random stream does not touch the main one)
        est.proj_energy = 0.0;
        double d0 = 0.0, ntot = 0.0;
        int64_t nst = 0;
        for (auto& r : ranks) {
            for (int64_t i = 0; i < r.nstates; ++i)
                if (r.states[i] == f0) d0 += (double)r.pops[i] / (double)pop_real_factor;
            ntot += r.nparticles; nst += r.nstates;
        }
        est.D0_population = d0; est.tot_nstates = nst;
        ntot_particles_old = ntot;
        ReportRow row0;
        row0.iter = mc_cycles_done; row0.shift = shift; row0.proj_energy = est.proj_energy;
        row0.D0_population = est.D0_population; row0.nparticles = ntot_particles_old;
        row0.nstates = est.tot_nstates; row0.nspawn_events = 0; row0.rspawn = 0.0;
        rows.push_back(row0);
        nattempts_rows.assign(1, (int64_t)ntot);
        for (int ireport = 1; ireport <= in.nreport; ++ireport) {
            ccmc_report_loop(ireport);
            if (error) break;
        }
        mc_cycles_done += in.ncycles * in.nreport;
    }
};

}  // namespace oracle
