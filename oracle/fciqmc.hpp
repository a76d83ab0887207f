// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the reference's FCIQMC / iFCIQMC propagation loop:
//   do_fciqmc main loop               src/fciqmc.f90:276-407, 635-769
//   spawn_standard / attempt_to_spawn src/spawning.F90:34-124, 711-766
//   create_spawned_particle*          src/spawning.F90:907-1319
//   stochastic_round*                 src/stoch_utils.f90:15-136
//   stochastic_death                  src/death.f90:11-130
//   set_parent_flag                   src/ifciqmc.f90:13-57
//   direct_annihilation and friends   src/annihilation.f90:9-901
//   spawn_t store / sort / merge      src/spawn_data.F90:156-248, 359-465, 624-739, 859-1101
//   qsort                             src/sort.f90:213-395
//   binary_search                     src/search.f90:47-170
//   estimators / shift                src/energy_evaluation.F90:126-201, 320-711, 906-986, 1433-1456
//   cycle bookkeeping                 src/qmc_common.F90:379-406, 697-797, 929-1017, 1240-1304
//   initial distribution              src/qmc.F90:1507-1646
// Several MPI ranks are emulated in-process (one RankState each, own RNG seed+iproc) so the
// np2/np4 golden trajectories can be followed too.
#pragma once
#include <memory>
#include <functional>
#include <cmath>
#include "system.hpp"
#include "rng.hpp"
#include "excit_gen.hpp"
#include "ueg.hpp"

namespace oracle {

struct QmcIn {
    double tau = 0.001;
    int seed = 7;
    double D0_population = 10.0;
    int ncycles = 20;
    int nreport = 10;
    double target_particles = 1.e7;
    double initial_shift = 0.0;
    double shift_damping = 0.05;
    double vary_shift_from = 0.0;
    bool vary_shift_from_proje = false;
    bool initiator_approx = false;
    double initiator_pop = 3.0;
    bool real_amplitudes = false;
    double spawn_cutoff = 0.01;
    int excit_gen = EXCIT_GEN_RENORM;
    double pattempt_single = -1.0, pattempt_double = -1.0;
    double pattempt_parallel = -1.0;
    // quasi-Newton propagator (src/qmc_data.f90:227-241)
    bool quasi_newton = false;
    double quasi_newton_threshold = -1.0, quasi_newton_value = -1.0, quasi_newton_pop_control = -1.0;
    int64_t walker_length = 1 << 20;          // elements per rank
    int64_t spawned_walker_length = 1 << 18;  // elements per rank
    int ex_level = -1;                        // truncation level (reference%ex_level); -1 => none
    int nprocs = 1;
    int nslots = 1;
    int rng_kind = 0;  // 0 dSFMT (reference stream), 1 Philox (engine stream)
    // Literal int32 truncation of the first event in annihilate_spawn_t_initiator
    // (src/spawn_data.F90:1019, sign(1,int(spawn_parts))).  true = as the reference (gfortran wraps),
    // false = mathematically intended symmetric rule (what the GPU engine implements).
    bool literal_event_int32 = true;
    std::vector<int> ref_det;  // explicit reference determinant (reference = { det = {...} }); empty => Aufbau
    int pop_real_bits = 31;    // real amplitudes are stored times 2^31 (POP_SIZE=64); real_amplitude_force_32 => 2^11
    // semi_stoch_in_t (src/qmc_data.f90:305-338)
    int ss_space = 0;          // 0 empty_determ_space, 1 high_pop_determ_space ("high"), 2 ci_determ_space ("ci")
    int ss_target_size = 0;    // size: number of most populated determinants wanted (space = "high")
    int ss_start_iter = 1, ss_shift_iter = -1;
    int ss_ci_ex_level = -1;   // ci_space = { ex_level = ... }
    bool ss_separate_annihilation = true;   // projection_mode (src/qmc_data.f90:325)
};

struct SpawnElem {
    Det f;
    int64_t pop = 0;
    int64_t flag = 0;
};

struct Estimators {
    double proj_energy = 0.0, D0_population = 0.0;
    double proj_energy_old = 0.0, D0_population_old = 0.0;
    int64_t tot_nstates = 0, tot_nspawn_events = 0;
};

struct RankState {
    int iproc = 0;
    std::vector<Det> states;
    std::vector<int64_t> pops;
    std::vector<double> dat;
    int64_t nstates = 0;
    double nparticles = 0.0;
    std::vector<std::vector<SpawnElem>> send;  // per destination rank (block of sdata)
    std::vector<SpawnElem> recv;               // after comm: concatenated by source rank
    std::unique_ptr<Rng> rng;
    bool spawn_error = false, psip_error = false;
    double rspawn = 0.0;
    double proj_energy = 0.0, D0_population = 0.0;  // per-rank accumulators within a report loop
    int nspawn_events = 0;
    int64_t ndeath = 0;
    int64_t nattempts = 0;
    // semi_stoch_t, the per-process part (src/semi_stoch.F90:40-108)
    std::vector<uint8_t> dflag;          // determ%flags: 0 = deterministic state, 1 = not (one per state of the main list)
    std::vector<double> dvector;         // determ%vector
    std::vector<int64_t> dindices;       // determ%indices (0-based positions in the main list)
    std::vector<int> drow_ptr, dcol_ind; // determ%hamil (csrp_t): rows = every deterministic state, columns = this rank's
    std::vector<double> dmat;
    std::vector<double> drho_minus_qn;   // determ%rho_minus_qn_weight
};

struct ReportRow {
    int iter;
    double shift, proj_energy, D0_population, nparticles;
    int64_t nstates, nspawn_events;
    double rspawn;
};

// qsort_int_64_list (src/sort.f90:213-395) for spawn elements keyed on the bit string, ascending.
inline void qsort_spawn(std::vector<SpawnElem>& list, int head, int W) {
    auto gt = [&](const SpawnElem& a, const SpawnElem& b) { return det_less(b.f, a.f, W); };
    auto ge = [&](const SpawnElem& a, const SpawnElem& b) { return !det_less(a.f, b.f, W); };
    const int switch_threshold = 7;
    std::vector<std::pair<int, int>> stack;
    int lo = 1, hi = head;
    auto L = [&](int k) -> SpawnElem& { return list[k - 1]; };
    for (;;) {
        if (hi - lo < switch_threshold) {
            for (int j = lo + 1; j <= hi; ++j) {
                SpawnElem tmp = L(j);
                int i;
                for (i = j - 1; i >= 1; --i) {
                    if (ge(tmp, L(i))) break;
                    L(i + 1) = L(i);
                }
                L(i + 1) = tmp;
            }
            if (stack.empty()) break;
            lo = stack.back().first; hi = stack.back().second;
            stack.pop_back();
        } else {
            int pivot = (lo + hi) / 2;
            std::swap(L(pivot), L(lo + 1));
            if (gt(L(lo), L(hi))) std::swap(L(lo), L(hi));
            if (gt(L(lo + 1), L(hi))) std::swap(L(lo + 1), L(hi));
            if (gt(L(lo), L(lo + 1))) std::swap(L(lo), L(lo + 1));
            int i = lo + 1, j = hi;
            SpawnElem tmp = L(lo + 1);
            for (;;) {
                i++;
                while (gt(tmp, L(i))) i++;
                j--;
                while (gt(L(j), tmp)) j--;
                if (j < i) break;
                std::swap(L(i), L(j));
            }
            L(lo + 1) = L(j);
            L(j) = tmp;
            if (hi - i + 1 >= j - lo) {
                stack.push_back({i, hi});
                hi = j - 1;
            } else {
                stack.push_back({lo, j - 1});
                lo = i;
            }
        }
    }
}

// binary_search_i0_list (src/search.f90:47-170), ascending order; 1-based positions.
inline void binary_search(const std::vector<Det>& list, const Det& item, int istart, int iend, int W, bool& hit,
                          int& pos) {
    auto cmp = [&](const Det& b1, const Det& b2) -> int {  // bit_str_cmp(b1,b2): 1 if b1<b2, -1 if b1>b2
        if (det_less(b1, b2, W)) return 1;
        if (det_less(b2, b1, W)) return -1;
        return 0;
    };
    if (istart > iend) { pos = istart; hit = false; return; }
    int lo = istart, hi = iend;
    hit = false;
    pos = istart;
    while (hi != lo) {
        pos = (hi + lo) / 2;
        int c = cmp(list[pos - 1], item);
        if (c == 0) { hit = true; break; }
        else if (c == 1) lo = pos + 1;
        else hi = pos;
    }
    if (hi == lo) {
        int c = cmp(list[hi - 1], item);
        if (c == 0) { hit = true; pos = hi; }
        else if (c == 1) pos = hi + 1;
        else pos = hi;
    }
}

struct Oracle {
    System sys;
    QmcIn in;
    ExcitGenData eg;
    const ExcitGenData* eg_shared = nullptr;  // replicas of orc_cpu_baseline share the (2 GB) heat-bath tables
    const ExcitGenData& EG() const { return eg_shared ? *eg_shared : eg; }
    // qmc_in%pattempt_update (src/qmc.F90:1049-1060): p_single_double_t of src/excit_gens.f90:13-41
    struct PsColl { double h_pgen_singles_sum = 0.0, excit_gen_singles = 0.0, h_pgen_doubles_sum = 0.0, excit_gen_doubles = 0.0; };
    bool vary_psingles = false;
    std::vector<PsColl> ps_rep_accum;  // per rank, zeroed after each update
    PsColl ps_total;
    double ps_counter = 1.0;
    std::vector<double> pattempt_log;  // pattempt_single after each change ("# pattempt_single changed to be:")
    // propagator_t (src/qmc_data.f90:866-884): quasi-Newton weights
    bool qn = false;
    std::vector<double> sp_fock;      // 1-based
    double ref_fock_sum = 0.0, qn_threshold = 0.0, qn_value = 0.0, qn_pop_control = 1.0;
    // init_sp_fock + init_quasi_newton (src/qmc.F90:1064-1160), calc_fock_values_3d_ueg (src/hamiltonian_ueg.f90:299-395)
    void init_propagator() {
        sp_fock.assign(sys.nbasis + 1, 0.0);
        for (int i = 1; i <= sys.nbasis; ++i) sp_fock[i] = sys.bf[i].sp_eigv;
        if (sys.kind == SYS_UEG) {
            const double pi = 3.1415926535897931;
            for (int i = 1; i <= sys.nbasis; ++i) {
                double ex = 0.0;
                bool in_ref = false;
                for (int k = 0; k < sys.nel; ++k) {
                    const int o = occ_list0[k];
                    if ((o % 2) == (i % 2) && o != i) ex = ex - coulomb_int_ueg_3d(sys, o, i);
                    if (o == i) in_ref = true;
                }
                // the reference writes the Madelung constant as a default-kind (single precision) literal
                const double mad = in_ref ? (double)(-2.837297f) * std::pow(0.75 / (pi * (sys.ueg.rs * sys.ueg.rs * sys.ueg.rs) * (double)sys.nel), 1.0 / 3.0) : 0.0;
                sp_fock[i] = sp_fock[i] + ex + 0.5 * mad;
            }
        }
        ref_fock_sum = 0.0;
        for (int k = 0; k < sys.nel; ++k) ref_fock_sum = ref_fock_sum + sp_fock[occ_list0[k]];
        qn = in.quasi_newton;
        if (qn) {
            if (in.quasi_newton_threshold < 0.0) {
                qn_threshold = sp_fock[sys.nel + 1] - sp_fock[sys.nel];
                if (sys.kind == SYS_UEG) qn_threshold = 2.0 * qn_threshold;
            } else qn_threshold = in.quasi_newton_threshold;
        }
        qn_value = (in.quasi_newton_value < 0.0) ? qn_threshold : in.quasi_newton_value;
        if (qn) qn_pop_control = (in.quasi_newton_pop_control < 0.0) ? 1.0 / qn_threshold : in.quasi_newton_pop_control;
        else qn_pop_control = 1.0;
    }
    double fock_sum_of(const int* occ) const {   // sum_fock_values_occ_list - ref%fock_sum (src/fciqmc.f90:321-322)
        double fs = 0.0;
        for (int k = 0; k < sys.nel; ++k) fs = fs + sp_fock[occ[k]];
        return fs - ref_fock_sum;
    }
    // calc_qn_spawned_weighting / calc_qn_weighting (src/spawning.F90:2063-2137)
    double qn_spawned_weighting(double spawner_dfock, const Excit& c) const {
        if (!qn) return 1.0;
        double diagel = spawner_dfock;
        for (int k = 0; k < c.nexcit; ++k) diagel = diagel + sp_fock[c.to_orb[k]] - sp_fock[c.from_orb[k]];
        if (diagel < qn_threshold) diagel = qn_value;
        return 1.0 / diagel;
    }
    double qn_weighting(double dfock) const {
        if (!qn) return 1.0;
        return (dfock < qn_threshold) ? 1.0 / qn_value : 1.0 / dfock;
    }
    // end_report_loop (src/qmc_common.F90:1206-1231)
    void end_report_loop_pattempt() {
        if (!vary_psingles) return;
        if (vary_shift) vary_psingles = false;
        else update_pattempt();
    }
    // update_pattempt + communicate_pattempt_single_data + add_rep_accum_to_total + update_pattempt_single
    // (src/spawning.F90:2217-2372); every_attempts = 10000, every_min_attempts = 10 (src/excit_gens.f90:36-37)
    void update_pattempt() {
        PsColl sum;
        for (auto& a : ps_rep_accum) {
            sum.excit_gen_singles += a.excit_gen_singles; sum.excit_gen_doubles += a.excit_gen_doubles;
            sum.h_pgen_singles_sum += a.h_pgen_singles_sum; sum.h_pgen_doubles_sum += a.h_pgen_doubles_sum;
        }
        ps_total.excit_gen_singles = ps_total.excit_gen_singles + sum.excit_gen_singles;
        ps_total.excit_gen_doubles = ps_total.excit_gen_doubles + sum.excit_gen_doubles;
        ps_total.h_pgen_singles_sum = ps_total.h_pgen_singles_sum + sum.h_pgen_singles_sum;
        ps_total.h_pgen_doubles_sum = ps_total.h_pgen_doubles_sum + sum.h_pgen_doubles_sum;
        const double every_attempts = 10000.0, every_min_attempts = 10.0;
        if ((ps_total.excit_gen_singles + ps_total.excit_gen_doubles) > (ps_counter * every_attempts) &&
            ps_total.excit_gen_singles > (ps_counter * every_min_attempts) &&
            ps_total.excit_gen_doubles > (ps_counter * every_min_attempts)) {
            ps_counter = ps_counter + 1.0;
            double ps = (ps_total.h_pgen_singles_sum / ps_total.excit_gen_singles) /
                        ((ps_total.h_pgen_doubles_sum / ps_total.excit_gen_doubles) +
                         (ps_total.h_pgen_singles_sum / ps_total.excit_gen_singles));
            if (ps < (1.0 / every_attempts)) ps = 1.0 / every_attempts;
            double pd = 1.0 - ps;
            if (pd < (1.0 / every_attempts)) { pd = 1.0 / every_attempts; ps = 1.0 - pd; }
            eg.pattempt_single = ps; eg.pattempt_double = pd;
            pattempt_log.push_back(ps);
        }
        for (auto& a : ps_rep_accum) a = PsColl();
    }
    // reference_t
    std::vector<int> occ_list0;
    Det f0;
    // wall-Chebyshev propagator (cheb_t, src/qmc_data.f90:886-907; src/propagators.f90): order sub-cycles per MC cycle,
    // each with its own weight 1/(S_i - E_0) on the spawning amplitude and the death probability
    struct Cheby {
        bool on = false;
        int order = 1, icheb = 1;
        double range_lo = 0.0, range_hi = 0.0;
        std::vector<double> zeroes, weights = std::vector<double>(1, 1.0);
    } cheb;
    double shift_harmonic_forcing = 0.0;      // qmc_state%shift_harmonic_forcing (src/qmc.F90:195-212)
    double cheb_weight() const { return cheb.weights[cheb.icheb - 1]; }
    // update_chebyshev (src/propagators.f90:188-208)
    void update_chebyshev(double sh) {
        const double pi = 3.1415926535897931;
        cheb.range_lo = sh;
        for (int i = 1; i <= cheb.order; ++i) {
            cheb.zeroes[i - 1] = sh + (cheb.range_hi - cheb.range_lo) / 2 * (1 - std::cos(pi * i / (cheb.order + 0.5)));
            cheb.weights[i - 1] = 1 / (cheb.zeroes[i - 1] - sh);
        }
    }
    // <f1|H|f2> for two different determinants (get_hmatel, src/hamiltonian_molecular.f90:10-103: symmetry- and
    // spin-checked slater_condon1_mol / slater_condon2_mol)
    double offdiag_hmatel(const Det& f1, const Det& f2) const {
        DetInfo d;
        decode_det_occ(sys, f1, d);
        Excit ex = sys.get_excitation(f1, f2);
        if (ex.nexcit == 1) {
            if (sys.bf[ex.from_orb[0]].ms == sys.bf[ex.to_orb[0]].ms && sys.bf[ex.from_orb[0]].sym == sys.bf[ex.to_orb[0]].sym)
                return sys.slater_condon1_excit(d.occ, ex.from_orb[0], ex.to_orb[0], ex.perm);
        } else if (ex.nexcit == 2) {
            if (sys.bf[ex.from_orb[0]].ms + sys.bf[ex.from_orb[1]].ms == sys.bf[ex.to_orb[0]].ms + sys.bf[ex.to_orb[1]].ms &&
                sys.cross_product_basis(ex.from_orb[0], ex.from_orb[1]) == sys.cross_product_basis(ex.to_orb[0], ex.to_orb[1]))
                return sys.slater_condon2_excit(ex.from_orb[0], ex.from_orb[1], ex.to_orb[0], ex.to_orb[1], ex.perm);
        }
        return 0.0;
    }
    // init_chebyshev (src/propagators.f90:11-165): upper spectral bound from the Gershgorin circle of the highest
    // determinant (highest_det, :167-186) over its single and double excitations, or from its diagonal element only
    void init_chebyshev(int order, double cshift, double cscale, bool skip_gershgorin) {
        cheb.on = true; cheb.order = order; cheb.icheb = 1;
        cheb.zeroes.assign(order, 0.0); cheb.weights.assign(order, 1.0);
        std::vector<int> occ_max;
        for (int ia = 1; ia <= sys.nalpha; ++ia) occ_max.push_back(sys.nbasis - (ia * 2 - 1));
        for (int ib = 1; ib <= sys.nbeta; ++ib) occ_max.push_back(sys.nbasis - (ib - 1) * 2);
        Det f_max;
        for (int o : occ_max) det_set(f_max, o);
        const double hmm = diag_hmatel(sys, f_max);
        double e_max;
        if (!skip_gershgorin) {
            // enumerate_determinants(..., ex_level = 2, ref_sym = sys%symmetry, occ_list_max): every determinant of the
            // calculation's symmetry within two excitations of f_max, f_max included when it has that symmetry
            e_max = 0.0;
            std::vector<int> occ, virt;
            for (int o = 1; o <= sys.nbasis; ++o) (det_test(f_max, o) ? occ : virt).push_back(o);
            std::vector<int> so(occ_max);
            std::sort(so.begin(), so.end());
            const bool in_list = sys.symmetry_orb_list(so.data(), sys.nel) == sys.symmetry;
            if (in_list) {
                for (size_t i = 0; i < occ.size(); ++i)
                    for (int a : virt) {
                        Det g = f_max; det_clr(g, occ[i]); det_set(g, a);
                        e_max = e_max + std::fabs(offdiag_hmatel(f_max, g));
                        for (size_t j = i + 1; j < occ.size(); ++j)
                            for (int b : virt) {
                                if (b <= a) continue;
                                Det g2 = g; det_clr(g2, occ[j]); det_set(g2, b);
                                e_max = e_max + std::fabs(offdiag_hmatel(f_max, g2));
                            }
                    }
                e_max = e_max + std::fabs(hmm);
            }
            e_max = e_max - std::fabs(hmm) + hmm - H00;
        } else {
            e_max = hmm - H00;
        }
        e_max = (e_max + cshift) * cscale;
        cheb.range_lo = 0.0; cheb.range_hi = e_max;
        update_chebyshev(0.0);
    }

    double H00 = 0.0;
    int ref_ex_level = 0;
    // qmc_state_t
    double tau = 0.0, shift = 0.0;
    bool vary_shift = false;
    int64_t pop_real_factor = 1;
    int64_t spawn_cutoff = 0;
    std::vector<int> proc_map;
    int mc_cycles_done = 0;
    Estimators est;
    double rspawn_report = 0.0;
    double ntot_particles_old = 0.0;
    std::vector<RankState> ranks;
    std::vector<ReportRow> rows;
    bool error = false;
    uint32_t hash_seed = 7;  // src/qmc.F90:1497

    // ---------------------------------------------------------------- setup (init_qmc)
    void init() {
        // init_reference (src/qmc.F90:1162-1226)
        if (!in.ref_det.empty()) occ_list0 = in.ref_det;   // reference = { det = {...} } (src/reference_determinant.f90)
        else occ_list0 = set_reference_det(sys, sys.symmetry);
        f0 = sys.encode(occ_list0.data(), sys.nel);
        H00 = diag_hmatel(sys, f0);
        ref_ex_level = (in.ex_level < 0) ? sys.nel : in.ex_level;
        pop_real_factor = in.real_amplitudes ? (1ll << in.pop_real_bits) : 1;  // src/particle_t_utils.f90 (2^31; force_32: 2^11)
        double cutoff = in.real_amplitudes ? in.spawn_cutoff : 0.0;
        spawn_cutoff = (int64_t)std::ceil(cutoff * (double)pop_real_factor);  // src/spawn_data.F90:215
        if (in.spawned_walker_length % in.nprocs != 0)
            in.spawned_walker_length = (int64_t)std::ceil((float)in.spawned_walker_length / in.nprocs) * in.nprocs;
        proc_map.resize((size_t)in.nprocs * in.nslots);
        for (size_t i = 0; i < proc_map.size(); ++i) proc_map[i] = (int)(i % in.nprocs);  // src/load_balancing.F90:170
        tau = in.tau;
        shift = in.initial_shift;
        vary_shift = false;
        init_propagator();
        // init_excit_gen (src/qmc.F90:910-1010)
        eg.excit_gen = in.excit_gen;
        if (in.pattempt_single < 0 || in.pattempt_double < 0) {
            if (sys.kind == SYS_UEG) { eg.pattempt_single = 0.0; eg.pattempt_double = 1.0; }  // src/qmc_common.F90:178-182
            else find_single_double_prob(sys, occ_list0.data(), eg.pattempt_single, eg.pattempt_double);
        } else {
            eg.pattempt_single = in.pattempt_single / (in.pattempt_single + in.pattempt_double);
            eg.pattempt_double = 1.0 - in.pattempt_single;
        }
        if (in.excit_gen == EXCIT_GEN_RENORM_SPIN || in.excit_gen == EXCIT_GEN_NO_RENORM_SPIN)   // src/qmc.F90:974-988
            eg.pattempt_parallel = (in.pattempt_parallel < 0.0) ? find_parallel_spin_prob_mol(sys, in.nprocs) : in.pattempt_parallel;
        if (in.excit_gen == EXCIT_GEN_POWER_PITZER) {        // src/qmc.F90:994-1007
            if (sys.kind == SYS_UEG) init_excit_ueg_power_pitzer(sys, eg.ppn);
            else init_excit_mol_power_pitzer_occ_ref(sys, occ_list0.data(), eg.ppn);
        }
        if (in.excit_gen == EXCIT_GEN_POWER_PITZER_ORDERN)   // src/qmc.F90:1009-1018
            init_excit_mol_power_pitzer_orderN(sys, occ_list0.data(), eg.ppn);
        if (in.excit_gen == EXCIT_GEN_HEAT_BATH_UNIFORM || in.excit_gen == EXCIT_GEN_HEAT_BATH_SINGLE ||
            in.excit_gen == EXCIT_GEN_POWER_PITZER_OCC_IJ || in.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ)
            init_excit_mol_heat_bath(sys, eg.hb, false);
        if (in.excit_gen == EXCIT_GEN_HEAT_BATH) {
            if (!init_excit_mol_heat_bath(sys, eg.hb, true))
                throw std::runtime_error("heat_bath: not all single excitations can be accounted for");
        }
        ranks.clear();
        ranks.resize(in.nprocs);
        for (int ip = 0; ip < in.nprocs; ++ip) {
            RankState& r = ranks[ip];
            r.iproc = ip;
            r.send.assign(in.nprocs, {});
            if (in.rng_kind == 0) r.rng.reset(new DsfmtRng(in.seed + ip, 50000));
            else r.rng.reset(new PhiloxRng((uint32_t)in.seed));
        }
        initial_distribution();
        for (auto& r : ranks) recompute_nparticles(r);
        mc_cycles_done = 0;
        determ = Determ();
        semi_stoch_iter = std::max(in.ss_start_iter, mc_cycles_done + 1);   // src/fciqmc.f90:228
        rows.clear();
        est = Estimators();
    }

    int owner(const Det& f) const {
        return assign_particle_processor(f, sys.nbasis, hash_seed, 0, 0, in.nprocs, proc_map.data(), in.nslots);
    }

    // initial_distribution (src/qmc.F90:1507-1646), single reference, no spin-inverse
    void initial_distribution() {
        int D0_proc = owner(f0);
        for (auto& r : ranks) {
            r.states.clear(); r.pops.clear(); r.dat.clear();
            r.nstates = 0;
            if (r.iproc == D0_proc) {
                r.states.push_back(f0);
                r.pops.push_back((int64_t)std::llround(in.D0_population) * pop_real_factor);
                r.dat.push_back(0.0);
                r.nstates = 1;
            }
        }
    }
    void recompute_nparticles(RankState& r) {
        // init_estimators (src/qmc.F90:845-850)
        double s = 0.0;
        for (int64_t k = 0; k < r.nstates; ++k) s += std::fabs((double)r.pops[k] / (double)pop_real_factor);
        r.nparticles = s;
    }

    // ---------------------------------------------------------------- per-determinant pieces
    // update_proj_energy_mol (src/energy_evaluation.F90:906-986); returns contribution pieces.
    void update_proj_energy(const DetInfo& d, double pop, double& D0_acc, double& pe_acc) const {
        Excit ex = sys.get_excitation(d.f, f0);
        if (sys.kind == SYS_UEG) {
            // update_proj_energy_ueg (src/energy_evaluation.F90:1072-1127)
            if (ex.nexcit == 0) {
                bool same = true;
                for (int k = 0; k < sys.W; ++k) if (d.f.w[k] != f0.w[k]) same = false;
                if (same) D0_acc = D0_acc + pop;
            } else if (ex.nexcit == 2) {
                double h = slater_condon2_ueg(sys, ex.from_orb[0], ex.from_orb[1], ex.to_orb[0], ex.to_orb[1], ex.perm);
                pe_acc = pe_acc + h * pop;
            }
            return;
        }
        if (ex.nexcit == 0) {
            // any(f1/=f2) false => nexcit 0
            bool same = true;
            for (int k = 0; k < sys.W; ++k) if (d.f.w[k] != f0.w[k]) same = false;
            if (same) D0_acc = D0_acc + pop;
        } else if (ex.nexcit == 1) {
            if (sys.bf[ex.from_orb[0]].ms == sys.bf[ex.to_orb[0]].ms &&
                sys.bf[ex.from_orb[0]].sym == sys.bf[ex.to_orb[0]].sym) {
                double h = sys.slater_condon1_excit(d.occ, ex.from_orb[0], ex.to_orb[0], ex.perm);
                pe_acc = pe_acc + h * pop;
            }
        } else if (ex.nexcit == 2) {
            if (sys.bf[ex.from_orb[0]].ms + sys.bf[ex.from_orb[1]].ms ==
                sys.bf[ex.to_orb[0]].ms + sys.bf[ex.to_orb[1]].ms) {
                int ij_sym = sys.cross_product_basis(ex.from_orb[0], ex.from_orb[1]);
                int ab_sym = sys.cross_product_basis(ex.to_orb[0], ex.to_orb[1]);
                if (ij_sym == ab_sym) {
                    double h = sys.slater_condon2_excit(ex.from_orb[0], ex.from_orb[1], ex.to_orb[0], ex.to_orb[1],
                                                        ex.perm);
                    pe_acc = pe_acc + h * pop;
                }
            }
        }
    }

    // decide_nattempts (src/qmc_common.F90:379-406)
    static int decide_nattempts(Rng& rng, double population) {
        int nattempts = std::abs((int)population);
        double pextra = std::fabs(population) - nattempts;
        if (std::fabs(pextra) > depsilon) {
            if (pextra > rng.next()) nattempts++;
        }
        return nattempts;
    }

    // stochastic_round_spawned_particle (src/stoch_utils.f90:87-136)
    static int64_t stochastic_round_spawned_particle(int64_t cutoff, double pspawn, Rng& rng) {
        int64_t nspawn;
        if (pspawn < (double)cutoff) {
            if (pspawn > rng.next() * (double)cutoff) nspawn = cutoff; else nspawn = 0;
        } else {
            nspawn = (int64_t)pspawn;
            double padd = pspawn - (double)nspawn;
            if (padd > rng.next()) nspawn++;
        }
        return nspawn;
    }
    // attempt_to_spawn (src/spawning.F90:711-766)
    int64_t attempt_to_spawn(Rng& rng, double hmatel, double pgen, int64_t parent_sign) const {
        double pspawn = tau * std::fabs(hmatel) / pgen;
        pspawn = pspawn * (double)pop_real_factor;
        int64_t nspawn = stochastic_round_spawned_particle(spawn_cutoff, pspawn, rng);
        if (nspawn > 0) {
            int64_t mag = nspawn;
            int64_t signed_n = (parent_sign >= 0) ? mag : -mag;  // sign(nspawn, parent_sign)
            nspawn = (hmatel > 0.0) ? -signed_n : signed_n;
        }
        return nspawn;
    }
    // stochastic_round (src/stoch_utils.f90:51-85), one population
    static void stochastic_round(Rng& rng, int64_t& pop, int64_t cutoff) {
        int64_t ap = pop < 0 ? -pop : pop;
        if (ap < cutoff && pop != 0) {
            double r = rng.next() * (double)cutoff;
            if ((double)ap > r) pop = (pop < 0) ? -cutoff : cutoff; else pop = 0;
        }
    }

    // ---------------------------------------------------------------- one MC cycle, one rank: spawn+death
    void spawn_death_rank(RankState& r, uint32_t cycle_id) {
        Rng& rng = *r.rng;
        rng.set_cycle(cycle_id);
        // init_mc_cycle (src/qmc_common.F90:950-1017)
        for (auto& b : r.send) b.clear();
        r.ndeath = 0;
        r.nattempts = (int64_t)std::llround(2 * r.nparticles);
        const int64_t block_size = in.spawned_walker_length / in.nprocs;
        DetInfo d;
        int ideterm = 0;
        for (int64_t idet = 0; idet < r.nstates; ++idet) {
            decode_for(sys, EG(), r.states[idet], d);
            double real_population = (double)r.pops[idet] / (double)pop_real_factor;
            // set_determ_info (src/semi_stoch.F90:826-857)
            bool determ_parent = false;
            if (determ.doing && r.dflag[(size_t)idet] == 0) {
                r.dvector[(size_t)ideterm] = real_population;
                r.dindices[(size_t)ideterm] = idet;
                ideterm++;
                determ_parent = true;
            }
            // set_parent_flag (src/ifciqmc.f90:13-57), nspaces=1: deterministic states are always initiators
            d.initiator_flag = (std::fabs(real_population) > in.initiator_pop || determ_parent) ? 0 : 1;
            update_proj_energy(d, real_population, r.D0_population, r.proj_energy);
            const double dfock = qn ? fock_sum_of(d.occ) : 0.0;
            rng.begin(RNG_NATTEMPTS, d.f, sys.W, 0);
            int nattempts_det = decide_nattempts(rng, real_population);
            int64_t pop = r.pops[idet];
            for (int ip = 0; ip < nattempts_det; ++ip) {
                rng.begin(RNG_SPAWN, d.f, sys.W, (uint32_t)ip);
                GenResult g = gen_excit_sys(rng, sys, EG(), d);
                if (g.allowed && vary_psingles) {   // spawn_standard (src/spawning.F90:101-109): straight into rep_accum
                    if ((int)ps_rep_accum.size() != in.nprocs) ps_rep_accum.assign(in.nprocs, PsColl());
                    PsColl& a = ps_rep_accum[r.iproc];
                    if (g.conn.nexcit == 1) {
                        a.h_pgen_singles_sum = a.h_pgen_singles_sum + ((std::fabs(g.hmatel) * eg.pattempt_single) / g.pgen);
                        a.excit_gen_singles = a.excit_gen_singles + 1.0;
                    } else if (g.conn.nexcit == 2) {
                        a.h_pgen_doubles_sum = a.h_pgen_doubles_sum + ((std::fabs(g.hmatel) * eg.pattempt_double) / g.pgen);
                        a.excit_gen_doubles = a.excit_gen_doubles + 1.0;
                    }
                }
                const double qn_weight = g.allowed ? qn_spawned_weighting(dfock, g.conn) : 1.0;   // spawn_standard
                int64_t nspawned = attempt_to_spawn(rng, g.hmatel * qn_weight * cheb_weight(), g.pgen, pop);
                if (nspawned != 0) {
                    Det fnew = sys.create_excited_det(d.f, g.conn);
                    // spawning from the deterministic space into it is the projection's job (src/fciqmc.f90:726-736)
                    if (determ_parent && check_if_determ(fnew)) continue;
                    // create_spawned_particle[_initiator][_truncated] (src/spawning.F90:1074-1319)
                    if (in.ex_level >= 0 && sys.excitation_level(f0, fnew) > ref_ex_level) continue;
                    int dest = owner(fnew);
                    // add_[flagged_]spawned_particle (src/spawning.F90:907-1018)
                    if ((int64_t)r.send[dest].size() + 1 > block_size) {
                        r.spawn_error = true;
                    } else {
                        SpawnElem e;
                        e.f = fnew; e.pop = nspawned;
                        e.flag = in.initiator_approx ? d.initiator_flag : 0;
                        r.send[dest].push_back(e);
                    }
                }
            }
            // stochastic_death (src/death.f90:11-130); not for deterministic states (src/fciqmc.f90:368)
            if (!determ_parent) {
                rng.begin(RNG_DEATH, d.f, sys.W, 0);
                double Kii = r.dat[idet];
                double weight = qn_weighting(dfock);
                double pd = tau * ((Kii - est.proj_energy_old) * weight + (est.proj_energy_old - shift) * qn_pop_control) * 1.0;
                pd = pd * cheb_weight();
                int64_t& population = r.pops[idet];
                int64_t apop = population < 0 ? -population : population;
                pd = pd * (double)apop;
                int64_t kill = (int64_t)pd;
                pd = pd - (double)kill;
                double rr = rng.next();
                if (std::fabs(pd) > rr) {
                    if (pd > 0.0) kill++; else kill--;
                }
                int64_t old_population = population;
                if (population < 0) population += kill; else population -= kill;
                int64_t anew = population < 0 ? -population : population;
                int64_t aold = old_population < 0 ? -old_population : old_population;
                r.nparticles = r.nparticles + (double)(anew - aold) / (double)pop_real_factor;
                r.ndeath += (kill < 0 ? -kill : kill);
            }
        }
        // calc_events_spawn_t (src/spawn_data.F90:445-465)
        int ev = 0;
        for (auto& b : r.send) ev += (int)b.size();
        r.nspawn_events = ev;
    }

    // comm_spawn_t (src/spawn_data.F90:624-739): personalised all-to-all between emulated ranks
    void comm_spawn() {
        for (auto& r : ranks) r.recv.clear();
        for (int dst = 0; dst < in.nprocs; ++dst)
            for (int src = 0; src < in.nprocs; ++src)
                for (auto& e : ranks[src].send[dst]) ranks[dst].recv.push_back(e);
    }

    // annihilate_spawn_t (src/spawn_data.F90:859-935)
    void annihilate_spawn_t(std::vector<SpawnElem>& s) const {
        int head = (int)s.size();
        if (head == 0) return;
        int islot = 1, k = 1;
        const int upper = head;
        auto S = [&](int i) -> SpawnElem& { return s[i - 1]; };
        for (;;) {
            S(islot) = S(k);
            bool done = false;
            for (;;) {
                k++;
                if (k > upper) { done = true; break; }
                if (S(k).f == S(islot).f) {
                    S(islot).pop += S(k).pop;
                    S(islot).flag += S(k).flag;  // sdata(bit_str_len+1:) sums everything after the string
                } else break;
            }
            if (done) break;
            if (islot == upper) break;
            if (S(islot).pop != 0 || S(islot).flag != 0) islot++;
        }
        if (S(islot).pop == 0 && S(islot).flag == 0) islot--;
        s.resize(islot);
    }

    // annihilate_spawn_t_initiator (src/spawn_data.F90:937-1101), ntypes = 1
    void annihilate_spawn_t_initiator(std::vector<SpawnElem>& s) const {
        int head = (int)s.size();
        if (head == 0) return;
        int islot = 1, k = 1;
        const int upper = head;
        auto S = [&](int i) -> SpawnElem& { return s[i - 1]; };
        int64_t events = 0, initiator_pop = 0;
        for (;;) {
            S(islot) = S(k);
            if (!(S(k).flag & 1)) {
                initiator_pop = S(k).pop;
                events = 0;
            } else {
                initiator_pop = 0;
                if (in.literal_event_int32) {
                    int32_t t = (int32_t)(uint32_t)(uint64_t)S(k).pop;  // int(spawn_parts) default integer
                    events = (t >= 0) ? 1 : -1;                         // sign(1, .)
                } else {
                    events = (S(k).pop >= 0) ? 1 : -1;
                }
            }
            for (;;) {
                k++;
                bool same_slot = k <= head;
                if (same_slot) same_slot = (S(k).f == S(islot).f);
                if (same_slot) {
                    if (!(S(k).flag & 1)) initiator_pop += S(k).pop;
                    else if (S(k).pop < 0) events -= 1;
                    else if (S(k).pop > 0) events += 1;
                    S(islot).pop += S(k).pop;
                } else {
                    S(islot).flag = 0;
                    int64_t sgn_tot = (S(islot).pop >= 0) ? 1 : -1;
                    int64_t sgn_ini = (initiator_pop >= 0) ? 1 : -1;
                    if (initiator_pop != 0 && sgn_tot == sgn_ini) {
                        // keep
                    } else if ((events < 0 ? -events : events) > 1) {
                        // keep
                    } else {
                        S(islot).flag += 1;
                    }
                    break;
                }
            }
            if (islot == upper || k > upper) break;
            if (S(islot).pop != 0) islot++;
        }
        if (S(islot).pop == 0) islot--;
        s.resize(islot);
    }

    // annihilate_main_list[_initiator] (src/annihilation.f90:294-486)
    void annihilate_main_list(RankState& r, std::vector<SpawnElem>& s) const {
        int nannihilate = 0;
        int istart = 1, iend = (int)r.nstates;
        int head = (int)s.size();
        for (int i = 1; i <= head; ++i) {
            bool hit; int pos;
            binary_search(r.states, s[i - 1].f, istart, iend, sys.W, hit, pos);
            if (hit) {
                int64_t old_pop = r.pops[pos - 1];
                if (!in.initiator_approx) {
                    r.pops[pos - 1] += s[i - 1].pop;
                } else {
                    if (r.pops[pos - 1] != 0) r.pops[pos - 1] += s[i - 1].pop;
                    else if (!(s[i - 1].flag & 1)) r.pops[pos - 1] = s[i - 1].pop;
                }
                int64_t an = std::llabs(r.pops[pos - 1]), ao = std::llabs(old_pop);
                r.nparticles = r.nparticles + (double)(an - ao) / (double)pop_real_factor;
                nannihilate++;
                istart = pos + 1;
            } else {
                if (!in.initiator_approx) {
                    s[i - 1 - nannihilate] = s[i - 1];
                } else {
                    if (s[i - 1].flag & 1) {
                        nannihilate++;  // discard: spawned from a non-initiator onto an unoccupied det
                    } else {
                        SpawnElem e = s[i - 1];
                        s[i - 1 - nannihilate] = e;
                    }
                }
            }
        }
        s.resize(head - nannihilate);
    }

    // remove_unoccupied_dets (src/annihilation.f90:537-598)
    void remove_unoccupied_dets(RankState& r) const {
        Rng& rng = *r.rng;
        int64_t nzero = 0;
        const bool ss = determ.doing;
        for (int64_t i = 0; i < r.nstates; ++i) {
            const bool determ_det = ss && r.dflag[(size_t)i] == 0;
            if (in.real_amplitudes && !determ_det) {
                int64_t old_pop = r.pops[i];
                rng.begin(RNG_ROUND_MAIN, r.states[i], sys.W, 0);
                stochastic_round(rng, r.pops[i], pop_real_factor);
                r.nparticles = r.nparticles + (double)(std::llabs(r.pops[i]) - std::llabs(old_pop)) / (double)pop_real_factor;
            }
            if (r.pops[i] == 0 && !determ_det) {
                nzero++;
            } else if (nzero > 0) {
                int64_t k = i - nzero;
                r.states[k] = r.states[i]; r.pops[k] = r.pops[i]; r.dat[k] = r.dat[i];
                if (ss) r.dflag[(size_t)k] = r.dflag[(size_t)i];
            }
        }
        r.nstates -= nzero;
        r.states.resize(r.nstates); r.pops.resize(r.nstates); r.dat.resize(r.nstates);
        if (ss) r.dflag.resize((size_t)r.nstates);
    }

    // round_low_population_spawns (src/annihilation.f90:600-675)
    void round_low_population_spawns(RankState& r, std::vector<SpawnElem>& s) const {
        Rng& rng = *r.rng;
        int nremoved = 0;
        int head = (int)s.size();
        for (int i = 0; i < head; ++i) {
            rng.begin(RNG_ROUND_SPAWN, s[i].f, sys.W, 0);
            stochastic_round(rng, s[i].pop, pop_real_factor);
            if (s[i].pop == 0) nremoved++;
            else s[i - nremoved] = s[i];
        }
        s.resize(head - nremoved);
    }

    // insert_new_walkers (src/annihilation.f90:677-818)
    void insert_new_walkers(RankState& r, std::vector<SpawnElem>& s) const {
        int head = (int)s.size();
        if (!r.psip_error) {
            float fill = (float)(r.nstates + head) / (float)in.walker_length;
            if (fill > 1.00f) r.psip_error = true;
        }
        if (r.psip_error) return;
        int64_t nold = r.nstates;
        r.states.resize(nold + head); r.pops.resize(nold + head); r.dat.resize(nold + head);
        const bool ss = determ.doing;
        if (ss) r.dflag.resize((size_t)(nold + head), 1);
        int istart = 1, iend = (int)nold;
        for (int i = head; i >= 1; --i) {
            bool hit; int pos;
            binary_search(r.states, s[i - 1].f, istart, iend, sys.W, hit, pos);
            for (int j = iend; j >= pos; --j) {
                int k = j + i;
                r.states[k - 1] = r.states[j - 1]; r.pops[k - 1] = r.pops[j - 1]; r.dat[k - 1] = r.dat[j - 1];
                if (ss) r.dflag[(size_t)k - 1] = r.dflag[(size_t)j - 1];
            }
            int k = pos + i - 1;
            if (ss) r.dflag[(size_t)k - 1] = 1;   // a deterministic state never leaves the list, so a new state is not one
            // insert_new_walker (src/annihilation.f90:820-901)
            r.states[k - 1] = s[i - 1].f;
            r.pops[k - 1] = s[i - 1].pop;
            r.dat[k - 1] = diag_hmatel(sys, s[i - 1].f) - H00;
            double real_population = (double)s[i - 1].pop / (double)pop_real_factor;
            r.nparticles = r.nparticles + std::fabs(real_population);
            iend = pos - 1;
        }
        r.nstates = nold + head;
    }

    // direct_annihilation (src/annihilation.f90:9-79) for one rank, after comm
    void annihilate_rank(RankState& r) {
        std::vector<SpawnElem>& s = r.recv;
        if (determ.doing && in.ss_separate_annihilation) deterministic_annihilation(r);
        if (!s.empty()) {
            qsort_spawn(s, (int)s.size(), sys.W);
            if (in.initiator_approx) annihilate_spawn_t_initiator(s); else annihilate_spawn_t(s);
        }
        if (!s.empty()) {
            annihilate_main_list(r, s);
            remove_unoccupied_dets(r);
            if (in.real_amplitudes) round_low_population_spawns(r, s);
            insert_new_walkers(r, s);
        } else {
            remove_unoccupied_dets(r);
        }
    }

    // ---------------------------------------------------------------- load balancing
    // initialise_slot_pop (src/load_balancing.F90:624-654): population of rank r in every slot
    std::vector<double> slot_pop(const RankState& r) const {
        std::vector<double> sp(proc_map.size(), 0.0);
        for (int64_t i = 0; i < r.nstates; ++i) {
            int det_pos = 0;
            assign_particle_processor(r.states[i], sys.nbasis, hash_seed, 0, 0, in.nprocs, proc_map.data(), in.nslots, &det_pos);
            sp[(size_t)det_pos] = sp[(size_t)det_pos] + std::fabs((double)r.pops[i]);
        }
        for (auto& x : sp) x = x / (double)pop_real_factor;
        return sp;
    }
    // insertion_rank (lib/local/ranking.f90:81-117), 0-based
    static std::vector<int> insertion_rank(const std::vector<double>& arr, double tol) {
        std::vector<int> rank(arr.size());
        for (size_t i = 0; i < arr.size(); ++i) rank[i] = (int)i;
        for (int i = 1; i < (int)arr.size(); ++i) {
            int j = i - 1;
            const int tmp = rank[i];
            while (j >= 0) {
                if (arr[rank[j]] - arr[tmp] < tol) break;
                rank[j + 1] = rank[j];
                --j;
            }
            rank[j + 1] = tmp;
        }
        return rank;
    }
    // do_load_balancing (src/load_balancing.F90:209-323) on the slot list summed over the ranks (MPI_AllReduce):
    // check_imbalance (:353-381), find_processors (:520-601), reduce_slots (:478-518), redistribute_slots (:419-476).
    // Modifies proc_map; returns load%needed.
    bool do_load_balancing(double percent) {
        std::vector<double> slot_list(proc_map.size(), 0.0);
        for (auto& r : ranks) {
            std::vector<double> sp = slot_pop(r);
            for (size_t k = 0; k < sp.size(); ++k) slot_list[k] = slot_list[k] + sp[k];
        }
        const int np = in.nprocs;
        std::vector<double> procs_pop(np, 0.0);
        for (int pr = 0; pr < np; ++pr)
            for (size_t k = 0; k < proc_map.size(); ++k)
                if (proc_map[k] == pr) procs_pop[pr] = procs_pop[pr] + slot_list[k];
        double tot = 0.0;
        for (double x : procs_pop) tot = tot + x;
        const double pop_av = tot / np;
        bool needed = false;
        for (double x : procs_pop) if (x > pop_av + pop_av * percent) needed = true;
        if (!needed) return false;
        const double up_thresh = pop_av + (double)(int)(pop_av * percent), low_thresh = pop_av - (double)(int)(pop_av * percent);
        std::vector<int> donors;
        int nrecv = 0;
        for (int i = 0; i < np; ++i) {
            if (procs_pop[i] < low_thresh) nrecv++;
            else if (procs_pop[i] > up_thresh) donors.push_back(i);
        }
        std::vector<int> prank = insertion_rank(procs_pop, 1.0e-8);
        std::vector<int> receivers(prank.begin(), prank.begin() + nrecv);
        std::vector<int> d_index;
        std::vector<double> d_pop;
        for (int d : donors)
            for (size_t j = 0; j < slot_list.size(); ++j)
                if (proc_map[j] == d) { d_pop.push_back(slot_list[j]); d_index.push_back((int)j); }
        std::vector<int> d_rank = insertion_rank(d_pop, 1.0e-8);
        for (int pos : d_rank)
            for (int rcv : receivers) {
                const double new_pop = d_pop[pos] + procs_pop[rcv];
                const double donor_pop = procs_pop[proc_map[d_index[pos]]] - d_pop[pos];
                if (donor_pop >= low_thresh && new_pop <= up_thresh) {
                    procs_pop[proc_map[d_index[pos]]] = donor_pop;
                    procs_pop[rcv] = new_pop;
                    proc_map[d_index[pos]] = rcv;
                    break;
                }
            }
        return true;
    }
    // redistribute_load_balancing_dets (src/qmc_common.F90:1332-1390): redistribute_particles (:505-595) on every rank,
    // then direct_annihilation of the moved determinants.  The rounding streams of this extra annihilation are keyed
    // by cycle_id (the engine and the tests pass the cycle with the top bit set).
    void redistribute_fciqmc(uint32_t cycle_id) {
        // redistribute_load_balancing_dets (src/qmc_common.F90:1332-1390): the moved determinants are annihilated WITHOUT
        // the deterministic flags (direct_annihilation has no determ argument there), then redistribute_semi_stoch_t
        // (:597-650) rebuilds the semi-stochastic objects from determ%dets under the new proc_map
        const bool ss = determ.doing;
        const std::vector<Det> all_dets = determ.dets;
        determ.doing = false;
        const int64_t block_size = in.spawned_walker_length / in.nprocs;
        for (auto& r : ranks) {
            for (auto& b : r.send) b.clear();
            double nsent = 0.0;
            for (int64_t i = 0; i < r.nstates; ++i) {
                const int pproc = owner(r.states[i]);
                if (pproc != r.iproc) {
                    if ((int64_t)r.send[pproc].size() + 1 > block_size) r.spawn_error = true;
                    else { SpawnElem e; e.f = r.states[i]; e.pop = r.pops[i]; e.flag = 0; r.send[pproc].push_back(e); }
                    nsent = nsent + std::fabs((double)r.pops[i]);
                    r.pops[i] = 0;
                }
            }
            r.nparticles = r.nparticles - nsent / (double)pop_real_factor;
        }
        comm_spawn();
        for (auto& r : ranks) {
            r.rng->set_cycle(cycle_id);
            annihilate_rank(r);
        }
        if (ss) {   // recreate_determ_space (src/semi_stoch.F90:1725-1761): every rank keeps the determinants it now owns
            given_determ.assign((size_t)in.nprocs, {});
            for (const Det& f : all_dets) given_determ[(size_t)owner(f)].push_back(f);
            const int keep = in.ss_space;
            in.ss_space = 3;
            init_semi_stoch();
            in.ss_space = keep;
        }
    }


    // ---------------------------------------------------------------- semi-stochastic projection (src/semi_stoch.F90)
    struct Determ {
        bool doing = false;              // determ%doing_semi_stoch
        int tot_size = 0;
        std::vector<int> sizes;          // per rank
        std::vector<Det> dets;           // every deterministic state, rank by rank, each rank's sorted like its main list
        std::vector<Det> sorted;         // the same set sorted once, for check_if_determ (the reference hashes; membership only)
        std::vector<double> full_vector; // determ%full_vector
    } determ;
    int semi_stoch_iter = 1;
    std::vector<std::vector<Det>> given_determ;

    // get_hmatel (src/hamiltonian_molecular.f90:12-71, src/hamiltonian_ueg.f90:12-69)
    double get_hmatel(const Det& f1, const Det& f2) const {
        Excit ex = sys.get_excitation(f1, f2);
        if (ex.nexcit == 0) return diag_hmatel(sys, f1);
        if (sys.kind == SYS_UEG) {
            if (ex.nexcit == 2) return slater_condon2_ueg(sys, ex.from_orb[0], ex.from_orb[1], ex.to_orb[0], ex.to_orb[1], ex.perm);
            return 0.0;
        }
        if (ex.nexcit <= 2) return offdiag_hmatel(f1, f2);
        return 0.0;
    }
    // check_if_determ (src/semi_stoch.F90:795-824)
    bool check_if_determ(const Det& f) const {
        const int W = sys.W;
        auto it = std::lower_bound(determ.sorted.begin(), determ.sorted.end(), f,
                                   [W](const Det& a, const Det& b) { return det_less(a, b, W); });
        return it != determ.sorted.end() && *it == f;
    }
    // find_most_populated_dets (src/semi_stoch.F90:1318-1385)
    static void find_most_populated_dets(const std::vector<Det>& dets_in, const std::vector<int64_t>& pops_in, int ndets_in,
                                         int ndets_out, std::vector<Det>& dets_out, std::vector<int64_t>& pops_out) {
        dets_out.assign(dets_in.begin(), dets_in.begin() + ndets_out);
        pops_out.assign(ndets_out, 0);
        if (ndets_out == 0) return;
        pops_out[0] = std::llabs(pops_in[0]);
        int64_t min_pop = pops_out[0];
        int min_ind = 0;
        for (int i = 1; i < ndets_out; ++i) {
            pops_out[i] = std::llabs(pops_in[i]);
            if (pops_out[i] < min_pop) { min_pop = pops_out[i]; min_ind = i; }
        }
        for (int i = ndets_out; i < ndets_in; ++i) {
            if (std::llabs(pops_in[i]) > min_pop) {
                dets_out[min_ind] = dets_in[i];
                pops_out[min_ind] = std::llabs(pops_in[i]);
                min_pop = pops_out[0]; min_ind = 0;
                for (int j = 1; j < ndets_out; ++j)
                    if (pops_out[j] < min_pop) { min_pop = pops_out[j]; min_ind = j; }
            }
        }
    }
    // find_indices_of_most_populated_dets (src/semi_stoch.F90:1387-1448); 0-based indices, -1 = unused
    static void find_indices_of_most_populated_dets(int npops_in, int nind_out, const std::vector<int64_t>& pops,
                                                    std::vector<int>& indices) {
        indices.assign(nind_out, -1);
        if (nind_out == 0 || npops_in == 0) return;
        indices[0] = 0;
        int64_t min_pop = std::llabs(pops[0]);
        int min_ind = 0;
        for (int i = 1; i < std::min(nind_out, npops_in); ++i) {
            indices[i] = i;
            if (std::llabs(pops[i]) < min_pop) { min_pop = std::llabs(pops[i]); min_ind = i; }
        }
        for (int i = nind_out; i < npops_in; ++i) {
            if (std::llabs(pops[i]) > min_pop) {
                indices[min_ind] = i;
                min_pop = std::llabs(pops[indices[0]]); min_ind = 0;
                for (int j = 1; j < nind_out; ++j)
                    if (std::llabs(pops[indices[j]]) < min_pop) { min_pop = std::llabs(pops[indices[j]]); min_ind = j; }
            }
        }
    }
    // create_high_pop_space (src/semi_stoch.F90:1198-1316): every rank offers its min(target, nstates) most populated
    // determinants, the target_size most populated of the offers are kept
    void create_high_pop_space(std::vector<std::vector<Det>>& dets_this_proc) {
        const int np = in.nprocs;
        std::vector<int> all_ndets(np), displs(np + 1, 0);
        std::vector<std::vector<Det>> determ_dets(np);
        std::vector<int64_t> all_determ_pops;
        for (int r = 0; r < np; ++r) {
            const RankState& R = ranks[r];
            all_ndets[r] = (int)std::min<int64_t>(in.ss_target_size, R.nstates);
            std::vector<int64_t> pops;
            find_most_populated_dets(R.states, R.pops, (int)R.nstates, all_ndets[r], determ_dets[r], pops);
            displs[r + 1] = displs[r] + all_ndets[r];
            all_determ_pops.insert(all_determ_pops.end(), pops.begin(), pops.end());
        }
        const int ndets_tot = displs[np];
        const int determ_size = std::min(in.ss_target_size, ndets_tot);
        std::vector<int> indices;
        find_indices_of_most_populated_dets(ndets_tot, determ_size, all_determ_pops, indices);
        for (int r = 0; r < np; ++r) {
            dets_this_proc[r].clear();
            for (int i = 0; i < determ_size; ++i)
                if (indices[i] >= displs[r] && indices[i] < displs[r + 1])
                    dets_this_proc[r].push_back(determ_dets[r][(size_t)(indices[i] - displs[r])]);
        }
    }
    // create_ci_determ_space (src/semi_stoch.F90:1763-1823): every determinant of the calculation's symmetry and spin
    // within ci_space.ex_level excitations of the reference (enumerate_determinants); restated by stepping through the
    // excitations of the reference level by level and keeping those the Hamiltonian can connect to it by symmetry
    void create_ci_determ_space(std::vector<std::vector<Det>>& dets_this_proc) {
        std::vector<int> occ0 = occ_list0;
        const int nel = sys.nel, nb = sys.nbasis;
        std::vector<Det> out;
        const int maxl = std::min(in.ss_ci_ex_level, nel);
        // combinations of `l` occupied orbitals to remove and `l` virtual orbitals to add
        std::vector<int> virt;
        for (int o = 1; o <= nb; ++o) if (!det_test(f0, o)) virt.push_back(o);
        const int ref_sym = det_symmetry(f0);
        std::vector<int> hole(maxl), part(maxl);
        for (int l = 0; l <= maxl; ++l) {
            // choose l holes (ascending) then l particles (ascending)
            std::function<void(int, int)> holes = [&](int k, int start) {
                if (k == l) {
                    std::function<void(int, int)> parts = [&](int k2, int start2) {
                        if (k2 == l) {
                            Det f = f0;
                            int ms = 0;
                            for (int t = 0; t < l; ++t) { det_clr(f, hole[t]); ms -= sys.bf[hole[t]].ms; }
                            for (int t = 0; t < l; ++t) { det_set(f, part[t]); ms += sys.bf[part[t]].ms; }
                            if (ms != 0) return;
                            if (det_symmetry(f) != ref_sym) return;
                            out.push_back(f);
                            return;
                        }
                        for (int v = start2; v < (int)virt.size(); ++v) { part[k2] = virt[v]; parts(k2 + 1, v + 1); }
                    };
                    parts(0, 0);
                    return;
                }
                for (int o = start; o < nel; ++o) { hole[k] = occ0[o]; holes(k + 1, o + 1); }
            };
            holes(0, 0);
        }
        for (auto& v : dets_this_proc) v.clear();
        for (const Det& f : out) dets_this_proc[(size_t)owner(f)].push_back(f);   // add_det_to_determ_space(check_proc=.true.)
    }
    // symmetry label of a determinant: the product of its orbitals' symmetries (symmetry_orb_list); for the UEG the
    // total momentum, packed so that equal labels mean equal momenta
    int det_symmetry(const Det& f) const {
        if (sys.kind == SYS_UEG) {
            int kx = 0, ky = 0, kz = 0;
            for (int o = 1; o <= sys.nbasis; ++o)
                if (det_test(f, o)) { kx += sys.ueg.l[3 * o]; ky += sys.ueg.l[3 * o + 1]; kz += sys.ueg.l[3 * o + 2]; }
            return ((kx + 512) * 1024 + (ky + 512)) * 1024 + (kz + 512);
        }
        int sym = sys.gamma_sym;
        for (int o = 1; o <= sys.nbasis; ++o)
            if (det_test(f, o)) sym = sys.cross_product(sym, sys.bf[o].sym);
        return sym;
    }
    // init_semi_stoch_t (src/semi_stoch.F90:134-377) for all emulated ranks
    void init_semi_stoch() {
        const int np = in.nprocs, W = sys.W;
        std::vector<std::vector<Det>> dets_this_proc(np);
        if (in.ss_space == 1) create_high_pop_space(dets_this_proc);
        else if (in.ss_space == 2) create_ci_determ_space(dets_this_proc);
        else if (in.ss_space == 3) dets_this_proc = given_determ;   // read_determ_space: the caller's list (SEMI.STOCH file)
        determ.doing = true;
        determ.sizes.assign(np, 0);
        determ.dets.clear();
        std::vector<int> displs(np, 0);
        for (int r = 0; r < np; ++r) {
            std::sort(dets_this_proc[r].begin(), dets_this_proc[r].end(), [W](const Det& a, const Det& b) { return det_less(a, b, W); });
            determ.sizes[r] = (int)dets_this_proc[r].size();
            displs[r] = (int)determ.dets.size();
            determ.dets.insert(determ.dets.end(), dets_this_proc[r].begin(), dets_this_proc[r].end());
        }
        determ.tot_size = (int)determ.dets.size();
        determ.sorted = determ.dets;
        std::sort(determ.sorted.begin(), determ.sorted.end(), [W](const Det& a, const Det& b) { return det_less(a, b, W); });
        determ.full_vector.assign((size_t)determ.tot_size, 0.0);
        for (int r = 0; r < np; ++r) {
            RankState& R = ranks[r];
            const std::vector<Det>& mine = dets_this_proc[r];
            const bool separate = in.ss_separate_annihilation;
            const int n_spawnees = separate ? determ.sizes[r] : determ.tot_size;
            R.dvector.assign((size_t)determ.sizes[r], 0.0);
            R.dindices.assign((size_t)determ.sizes[r], 0);
            // create_determ_hamil (src/semi_stoch.F90:533-685)
            R.drho_minus_qn.assign((size_t)n_spawnees, 1.0);
            if (qn) {
                for (int j = 0; j < n_spawnees; ++j) {
                    DetInfo d;
                    decode_det_occ(sys, separate ? mine[(size_t)j] : determ.dets[(size_t)j], d);
                    R.drho_minus_qn[(size_t)j] = qn_weighting(fock_sum_of(d.occ));
                }
            }
            R.drow_ptr.assign((size_t)determ.tot_size + 1, 0);
            R.dcol_ind.clear(); R.dmat.clear();
            for (int i = 0; i < determ.tot_size; ++i) {
                R.drow_ptr[(size_t)i] = (int)R.dmat.size();
                for (int j = 0; j < determ.sizes[r]; ++j) {
                    double h = get_hmatel(determ.dets[(size_t)i], mine[(size_t)j]);
                    if (i == j + displs[r]) h = h - H00;
                    const double weight = separate ? R.drho_minus_qn[(size_t)j] : R.drho_minus_qn[(size_t)i];
                    h = weight * h;
                    if (std::fabs(h) > depsilon) { R.dmat.push_back(h); R.dcol_ind.push_back(j); }
                }
            }
            R.drow_ptr[(size_t)determ.tot_size] = (int)R.dmat.size();
            for (auto& x : R.drho_minus_qn) x = qn_pop_control - x;
            // add_determ_dets_to_psip_list (src/semi_stoch.F90:728-793)
            R.dflag.assign((size_t)R.nstates, 1);
            int istart = 1, iend = (int)R.nstates;
            for (int i = 0; i < determ.sizes[r]; ++i) {
                bool hit; int pos;
                binary_search(R.states, mine[(size_t)i], istart, iend, W, hit, pos);
                if (!hit) {
                    R.states.insert(R.states.begin() + (pos - 1), mine[(size_t)i]);
                    R.pops.insert(R.pops.begin() + (pos - 1), 0);
                    R.dat.insert(R.dat.begin() + (pos - 1), diag_hmatel(sys, mine[(size_t)i]) - H00);
                    R.dflag.insert(R.dflag.begin() + (pos - 1), 1);
                    R.nstates++;
                }
                R.dflag[(size_t)pos - 1] = 0;
                istart = pos + 1;
                iend = (int)R.nstates;
            }
        }
    }
    // determ_proj_separate_annihil (src/semi_stoch.F90:1009-1094) for all emulated ranks
    void determ_proj_separate() {
        size_t k = 0;
        for (auto& R : ranks)
            for (double v : R.dvector) determ.full_vector[k++] = v;   // mpi_allgatherv
        for (auto& R : ranks) determ_proj_rank(R);
    }
    // one rank's part, determ%full_vector already gathered
    void determ_proj_rank(RankState& R) {
        const double a = -1.0 * tau;
        for (size_t j = 0; j < R.dvector.size(); ++j)
            R.dvector[j] = (-tau * (est.proj_energy_old * R.drho_minus_qn[j] - shift * qn_pop_control)) * R.dvector[j];
        // csrpgemv(.true., .false., -tau, hamil, full_vector, vector) (lib/local/csr.f90:148-204)
        for (int irow = 0; irow < determ.tot_size; ++irow)
            for (int iz = R.drow_ptr[(size_t)irow]; iz < R.drow_ptr[(size_t)irow + 1]; ++iz) {
                const int icol = R.dcol_ind[(size_t)iz];
                R.dvector[(size_t)icol] = R.dvector[(size_t)icol] + a * R.dmat[(size_t)iz] * determ.full_vector[(size_t)irow];
            }
    }
    // determ_proj_combined_annihil (src/semi_stoch.F90:897-975) + create_spawned_particle_determ (:1096-1146); one rank only
    // (the np > 1 form also needs compress_determ_repeats, src/spawn_data.F90:1103-1170, which is not restated)
    void determ_proj_combined(RankState& R) {
        if (in.nprocs != 1) throw std::runtime_error("semi-stochastic: combined annihilation is restated for one rank only");
        Rng& rng = *R.rng;
        const int64_t block_size = in.spawned_walker_length / in.nprocs;
        for (int row = 0; row < determ.tot_size; ++row) {
            double out_vec = 0.0;   // csrpgemv_single_row
            for (int iz = R.drow_ptr[(size_t)row]; iz < R.drow_ptr[(size_t)row + 1]; ++iz)
                out_vec = out_vec + R.dmat[(size_t)iz] * R.dvector[(size_t)R.dcol_ind[(size_t)iz]];
            out_vec = -tau * (out_vec + (est.proj_energy_old * R.drho_minus_qn[(size_t)row] - qn_pop_control * shift) * R.dvector[(size_t)row]);
            const double scaled = out_vec * (double)pop_real_factor;
            const double sgn = (out_vec < 0.0) ? -1.0 : 1.0;
            int64_t nspawn = (int64_t)std::fabs(scaled);
            rng.begin(RNG_DETERM, determ.dets[(size_t)row], sys.W, 0);
            if (std::fabs(scaled) - (double)nspawn > rng.next()) nspawn++;
            nspawn = nspawn * (int64_t)std::llround(sgn);
            if ((int64_t)R.send[0].size() + 1 > block_size) { R.spawn_error = true; continue; }
            SpawnElem e; e.f = determ.dets[(size_t)row]; e.pop = nspawn; e.flag = 0;
            R.send[0].push_back(e);
        }
        R.nspawn_events = (int)R.send[0].size();   // calc_events_spawn_t runs after the projection (src/annihilation.f90:56)
    }
    // deterministic_annihilation (src/annihilation.f90:488-535)
    void deterministic_annihilation(RankState& R) {
        Rng& rng = *R.rng;
        for (size_t i = 0; i < R.dvector.size(); ++i) {
            const int64_t ind = R.dindices[i];
            double scaled_amp = R.dvector[i] * (double)pop_real_factor;
            const int64_t spawn_sign = (scaled_amp < 0.0) ? -1 : 1;
            scaled_amp = std::fabs(scaled_amp);
            int64_t nspawn = (int64_t)scaled_amp;
            scaled_amp = scaled_amp - (double)nspawn;
            rng.begin(RNG_DETERM, R.states[(size_t)ind], sys.W, 0);
            if (scaled_amp > rng.next()) nspawn++;
            const int64_t old_pop = R.pops[(size_t)ind];
            R.pops[(size_t)ind] = R.pops[(size_t)ind] + spawn_sign * nspawn;
            R.nparticles = R.nparticles + (double)(std::llabs(R.pops[(size_t)ind]) - std::llabs(old_pop)) / (double)pop_real_factor;
        }
    }

    void mc_cycle(uint32_t cycle_id) {
        for (auto& r : ranks) spawn_death_rank(r, cycle_id);
        if (determ.doing) {   // determ_projection (src/semi_stoch.F90:859-895)
            if (in.ss_separate_annihilation) determ_proj_separate();
            else for (auto& r : ranks) determ_proj_combined(r);
        }
        comm_spawn();
        for (auto& r : ranks) {
            annihilate_rank(r);
            // end_mc_cycle / spawning_rate (src/qmc_common.F90:1240-1304)
            double ndeath_real = (double)r.ndeath / (double)pop_real_factor;
            double rate = (r.nattempts > 0) ? (r.nspawn_events + ndeath_real) / (double)r.nattempts : 0.0;
            r.rspawn = r.rspawn + rate;
        }
    }

    // initial_ci_projected_energy (src/qmc_common.F90:697-797)
    void initial_ci_projected_energy() {
        double pe = 0.0, d0 = 0.0, ntot = 0.0;
        int64_t nst = 0;
        for (auto& r : ranks) {
            double pe_r = 0.0, d0_r = 0.0;
            DetInfo d;
            for (int64_t i = 0; i < r.nstates; ++i) {
                decode_det_occ(sys, r.states[i], d);
                update_proj_energy(d, (double)r.pops[i] / (double)pop_real_factor, d0_r, pe_r);
            }
            pe += pe_r; d0 += d0_r; ntot += r.nparticles; nst += r.nstates;
        }
        est.proj_energy = pe; est.D0_population = d0; est.tot_nstates = nst;
        ntot_particles_old = ntot;
    }

    // update_shift (src/energy_evaluation.F90:659-711), target_particles > 0 branch with
    // shift_harmonic_forcing = 0 (default)
    void update_shift(double nparticles_old, double nparticles, int nupdate_steps) {
        if (in.target_particles <= 0.0) {
            shift = shift - std::log(nparticles / nparticles_old) * in.shift_damping / (1.0 * tau * nupdate_steps);
        } else {
            shift = shift - std::log(nparticles / nparticles_old) * in.shift_damping / (1.0 * tau * nupdate_steps) -
                    std::log(nparticles / in.target_particles) * (shift_harmonic_forcing) / (1.0 * tau * nupdate_steps);
        }
    }

    // One report loop (src/fciqmc.f90:276-407 + end_report_loop)
    void report_loop(int ireport) {
        // get_sanitized_projected_energy (src/energy_evaluation.F90:1433-1456)
        est.proj_energy_old = (std::fabs(est.D0_population) < std::numeric_limits<double>::min())
                                  ? 0.0 : est.proj_energy / est.D0_population;
        // init_report_loop (src/qmc_common.F90:929-948)
        est.D0_population_old = est.D0_population;
        for (auto& r : ranks) { r.rspawn = 0.0; r.proj_energy = 0.0; r.D0_population = 0.0; }
        for (int icycle = 1; icycle <= in.ncycles; ++icycle) {
            int iter = mc_cycles_done + (ireport - 1) * in.ncycles + icycle;
            // the wall-Chebyshev propagator applies `order` linear projectors per cycle (src/fciqmc.f90:298-299); the
            // counter-based stream is keyed by the sub-cycle index
            for (int icheb = 1; icheb <= cheb.order; ++icheb) {
                cheb.icheb = icheb;
                // should the semi-stochastic projection start now? (src/fciqmc.f90:300-305)
                if (iter == semi_stoch_iter && in.ss_space != 0 && !determ.doing) init_semi_stoch();
                mc_cycle((uint32_t)(cheb.on ? (iter - 1) * cheb.order + icheb : iter));
            }
        }
        // update_energy_estimators (src/energy_evaluation.F90:126-201, 320-655)
        double pe = 0.0, d0 = 0.0, rsp = 0.0, ntot = 0.0;
        int64_t nst = 0, nev = 0;
        bool err = false;
        for (auto& r : ranks) {
            pe += r.proj_energy; d0 += r.D0_population; rsp += r.rspawn; ntot += r.nparticles;
            nst += r.nstates; nev += r.nspawn_events;
            err = err || r.spawn_error || r.psip_error;
        }
        est.proj_energy = pe / (in.ncycles * cheb.order);
        est.D0_population = d0 / (in.ncycles * cheb.order);
        rspawn_report = rsp / (in.ncycles * cheb.order * in.nprocs);
        est.tot_nstates = nst; est.tot_nspawn_events = nev;
        if (vary_shift) update_shift(ntot_particles_old, ntot, in.ncycles);
        est.D0_population_old = est.D0_population;
        ntot_particles_old = ntot;
        if (!vary_shift && ntot > in.target_particles) {
            vary_shift = true;
            // semi-stochastic start relative to the shift start (src/qmc_common.F90:1200-1204)
            if (in.ss_shift_iter != -1) semi_stoch_iter = in.ss_shift_iter + (mc_cycles_done + ireport * in.ncycles) + 1;
            if (in.vary_shift_from_proje) shift = est.proj_energy / est.D0_population;
            else shift = in.vary_shift_from;
        }
        error = err;
        if (cheb.on) update_chebyshev(shift);       // src/fciqmc.f90:427
        end_report_loop_pattempt();
        ReportRow row;
        row.iter = mc_cycles_done + ireport * in.ncycles;
        row.shift = shift; row.proj_energy = est.proj_energy; row.D0_population = est.D0_population;
        row.nparticles = ntot_particles_old; row.nstates = est.tot_nstates; row.nspawn_events = est.tot_nspawn_events;
        row.rspawn = rspawn_report;
        rows.push_back(row);
    }

    void run() {
        initial_ci_projected_energy();
        // initial_qmc_status row (iteration 0)
        ReportRow row0;
        row0.iter = mc_cycles_done; row0.shift = shift; row0.proj_energy = est.proj_energy;
        row0.D0_population = est.D0_population; row0.nparticles = ntot_particles_old;
        row0.nstates = est.tot_nstates; row0.nspawn_events = 0; row0.rspawn = 0.0;
        rows.push_back(row0);
        for (int ireport = 1; ireport <= in.nreport; ++ireport) {
            report_loop(ireport);
            if (error) break;
        }
        mc_cycles_done += in.ncycles * in.nreport;
    }
};

}  // namespace oracle
