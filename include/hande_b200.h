/* hande_b200 - B200-native FCIQMC walker-propagation engine behind HANDE's hot-path seams.
 *
 * C ABI of libhande_b200.so.  The reference (hande-qmc/hande) has no C plugin interface for this
 * path: its seams are Fortran procedure pointers bound in init_proc_pointers (src/qmc.F90:217-703;
 * interfaces src/proc_pointers.f90:15-343) and the particle_t / spawn_t layouts.  Those are
 * per-walker scalar callbacks; this ABI is the batch-granular equivalent a ~100-line
 * iso_c_binding shim in do_fciqmc would call (INTEGRATION.md shows the Fortran side).
 * Every entry point names the reference routine(s) it replaces.
 *
 * Conventions: plain pointers and sizes only; host memory unless noted; orbital indices 1-based
 * as in the reference; arrays are in the reference's (Fortran, column-major) order; all functions
 * return 0 on success and non-zero on CUDA/NCCL/usage errors (message via hb200_last_error()).
 * Soft errors of the algorithm (spawn list / main list overflow) never abort: they set the
 * spawn_error / psip_error flags of hb200_iter_out exactly like spawn%error / psip_list%error
 * (src/spawning.F90:938-947, src/annihilation.f90:750-771) and the host exits at the end of the
 * report loop.
 */
#ifndef HANDE_B200_H
#define HANDE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb200_engine hb200_engine;

/* Values of qmc_in%excit_gen (src/qmc_data.f90:31-69) that the engine implements. */
enum { HB200_EXCIT_GEN_RENORM = 0, HB200_EXCIT_GEN_RENORM_SPIN = 1, HB200_EXCIT_GEN_NO_RENORM_SPIN = 3, HB200_EXCIT_GEN_NO_RENORM = 2, HB200_EXCIT_GEN_POWER_PITZER = 4, HB200_EXCIT_GEN_POWER_PITZER_OCC = 5,
       HB200_EXCIT_GEN_POWER_PITZER_OCC_IJ = 6, HB200_EXCIT_GEN_POWER_PITZER_ORDERN = 7, HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC = 8, HB200_EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ = 9, HB200_EXCIT_GEN_HEAT_BATH = 10, HB200_EXCIT_GEN_HEAT_BATH_UNIFORM = 11,
       HB200_EXCIT_GEN_HEAT_BATH_SINGLE = 12 };

/* qmc_in_t / fciqmc options that size and configure the device state
 * (src/qmc_data.f90:121-512; lua keys src/lua_hande_calc.f90:1244-1503). */
typedef struct hb200_config {
    int32_t device;                /* CUDA device ordinal (one engine = one MPI rank = one GPU) */
    int32_t nbasis, nel;           /* spin-orbitals, electrons (after CAS freezing) */
    int32_t excit_gen;             /* HB200_EXCIT_GEN_* */
    double pattempt_single, pattempt_double; /* excit_gen_data%pattempt_* (find_single_double_prob) */
    int32_t real_amplitudes;       /* qmc_in%real_amplitudes: pops encoded * 2^31 (POP_SIZE=64) */
    double spawn_cutoff;           /* qmc_in%spawn_cutoff (only with real amplitudes) */
    int32_t initiator_approx;      /* qmc_in%initiator_approx */
    double initiator_pop;          /* qmc_in%initiator_pop (default 3.0) */
    int32_t trunc_level;           /* reference%ex_level when truncating the space, else -1 */
    int64_t walker_length;         /* main list capacity (states), particle_t */
    int64_t spawned_walker_length; /* spawn_t%array_len (elements, all destinations together) */
    uint32_t rng_seed;             /* qmc_in%seed: key of the Philox stream */
    int32_t nprocs, iproc, nslots; /* hash-owner sharding (src/spawning.F90:770-838; load_bal nslots) */
    int32_t hash_seed;             /* 7 (src/qmc.F90:1497) */
} hb200_config;

/* sys_t / basis_t / pg_sym / integral stores for read_in systems
 * (src/system.f90:366-457, src/basis_types.f90:58-130, src/point_group_symmetry.f90,
 *  src/molecular_integral_types.f90:19-72).  Basis arrays have nbasis+1 entries (entry 0 unused). */
typedef struct hb200_system_read_in {
    int32_t nbasis, nel, uhf;
    int32_t nsym_tot, sym0, sym_max, pg_mask, Lz_mask, Lz_offset, gamma_sym;
    int32_t nvirt, nvirt_alpha, nvirt_beta, max_nbss;
    double Ecore;
    const int32_t* bf_sym;         /* basis_fns(:)%sym */
    const int32_t* bf_ms;          /* basis_fns(:)%ms (+1 alpha / -1 beta) */
    const int32_t* bf_spatial;     /* basis_fns(:)%spatial_index */
    const int32_t* nbasis_sym_spin;    /* (2, nsym_tot) */
    const int32_t* sym_spin_basis_fns; /* (max_nbss, 2, nsym_tot) */
    const double* one_body;        /* dense (nbasis, nbasis) of get_one_body_int_mol_real */
    const double* two_body[4];     /* coulomb_integrals%integrals(chan)%v, 1 (RHF) or 4 (UHF) channels */
    int64_t nintgrls;              /* length of each two-body channel */
} hb200_system_read_in;

/* sys_t for the 3D uniform electron gas (sys = ueg{...}): plane-wave basis ordered by kinetic energy
 * (init_model_basis_fns, src/basis.f90:258-501), ueg_basis_t lookup (src/ueg.f90:43-85, src/ueg_types.f90) and
 * excit_gen_data_t%ueg_ternary_conserve (src/ueg.f90:87-140).  Basis arrays have nbasis+1 entries (entry 0
 * unused); odd index = alpha.  Integrals are analytic (coulomb_int_ueg_3d, src/ueg.f90:250-280). */
typedef struct hb200_system_ueg {
    int32_t nbasis, nel;
    double box_length;             /* sys%lattice%box_length(1) = r_s (4 pi N / 3)^(1/3) */
    const int32_t* kvec;           /* basis_fns(:)%l, (3, nbasis+1) */
    const double* sp_eigv;         /* basis_fns(:)%sp_eigv (kinetic energies) */
    int32_t kmax, offset, offset_inds[3];  /* ueg_basis_t */
    const int32_t* lookup;         /* ueg_basis_t%lookup, 1-based: n_lookup + 1 entries (entry 0 unused) */
    int64_t n_lookup;
    int32_t tern_kmax;             /* ternary_conserve(0:W, -tern_kmax:tern_kmax, ... x3), tern_kmax = 2 kmax */
    const uint64_t* ternary_conserve;
} hb200_system_ueg;

/* Per-call propagation inputs (qmc_state_t scalars read by the hot loop). */
typedef struct hb200_iter_in {
    double tau;                    /* qs%tau */
    double shift;                  /* qs%shift(1) */
    double proj_energy_old;        /* qs%estimators(1)%proj_energy_old (src/fciqmc.f90:278) */
    uint32_t first_cycle;          /* iteration number of the first cycle (Philox key) */
} hb200_iter_in;

/* Per-call outputs: what end_mc_cycle / update_energy_estimators need
 * (src/qmc_common.F90:1240-1304, src/energy_evaluation.F90:320-420). */
typedef struct hb200_iter_out {
    double proj_energy;            /* sum over cycles of sum_j H_0j N_j (this rank) */
    double D0_population;          /* sum over cycles of N_0 (this rank) */
    double nparticles;             /* psip_list%nparticles(1) after the last cycle */
    int64_t nstates;               /* psip_list%nstates after the last cycle */
    int64_t nspawn_events;         /* calc_events_spawn_t of the last cycle (this rank) */
    int64_t ndeath;                /* ndeath of the last cycle (encoded units) */
    int64_t nattempts;             /* nint(2*nparticles) at the start of the last cycle */
    double rspawn;                 /* sum over cycles of spawning_rate() */
    int64_t nattempts_spawn;       /* total spawning attempts made (all cycles; throughput metric) */
    int32_t spawn_error, psip_error;
    double walker_iterations;      /* sum over cycles of nparticles at the start of the cycle (throughput metric) */
} hb200_iter_out;

/* Per-cycle outputs of the CCMC cluster stage (what do_ccmc accumulates per cycle, src/ccmc.f90:625-880). */
typedef struct hb200_ccmc_out {
    double proj_energy;            /* proj_energy_cycle: sum of <D0|H|D_cluster> amplitude sign / pselect */
    double D0_population;          /* D0_population_cycle */
    double D0_normalisation;       /* population on the reference (get_D0_info) */
    double tot_abs_real_pop;       /* cumulative excip population without the reference */
    int64_t nattempts;             /* qs%estimators%nattempts = number of cluster selections */
    int64_t nattempts_spawn;       /* spawning attempts made (accepted clusters) */
    int64_t nspawn_events;         /* entries added to the spawn list (spawns and deaths) */
    int64_t ndeath;                /* sum |nkill| of stochastic_ccmc_death (spawned as anti-excips) */
    int64_t ndeath_nc;             /* sum |nkill| of stochastic_ccmc_death_nc (full_nc: in place) */
    int32_t spawn_error, psip_error;
} hb200_ccmc_out;

const char* hb200_last_error(void);

/* init_qmc sizing (src/qmc.F90:10-215): allocates all device state. NULL on failure. */
hb200_engine* hb200_create(const hb200_config* cfg);
void hb200_destroy(hb200_engine* e);

/* Copies the system tables to the device (read_in systems); builds the J/K diagonal tables. */
int hb200_set_system_read_in(hb200_engine* e, const hb200_system_read_in* sys);
/* Copies the UEG tables to the device: the decoder / gen_excit_ueg_no_renorm / slater_condon0_ueg /
 * update_proj_energy_ueg bindings of init_proc_pointers (src/qmc.F90:466-476) are selected by this call. */
int hb200_set_system_ueg(hb200_engine* e, const hb200_system_ueg* sys);
/* init_excit_mol_heat_bath (src/excit_gen_heat_bath_mol.F90:14-256) evaluated on the device. */
int hb200_build_heat_bath(hb200_engine* e);
/* The same tables handed over by a host that has already built them (excit_gen_heat_bath_t after
 * init_excit_mol_heat_bath, src/excit_gens.f90:143-153, src/excit_gen_heat_bath_mol.F90:14-256) instead of
 * hb200_build_heat_bath: uploaded as they are.  Arrays in the reference's column-major order, aliasK 1-based. */
typedef struct hb200_heat_bath_tables {
    const double* i_weights;          /* (nbasis) */
    const double* ij_weights;         /* (nbasis, nbasis) */
    const double* ija_weights;        /* hb_ija%weights(nbasis, nbasis, nbasis) */
    const double* ija_aliasU;
    const int32_t* ija_aliasK;
    const double* ija_weights_tot;    /* (nbasis, nbasis) */
    const double* ijab_weights;       /* hb_ijab%weights(nbasis, nbasis, nbasis, nbasis) */
    const double* ijab_aliasU;
    const int32_t* ijab_aliasK;
    const double* ijab_weights_tot;   /* (nbasis, nbasis, nbasis) */
} hb200_heat_bath_tables;
int hb200_set_excit_tables(hb200_engine* e, const hb200_heat_bath_tables* tables);
/* Test/inspection: copy heat-bath table `which` to the host (0 i_w,1 ij_w,2 ija_w,3 ija_U,4 ija_tot,
 * 5 ijab_w,6 ijab_U,7 ijab_tot as double; 8 ija_K, 9 ijab_K as int32). n = element count expected. */
int hb200_download_heat_bath(hb200_engine* e, int which, void* out, int64_t n);

/* reference_t: f0 (W words) and H00 (init_reference src/qmc.F90:1162-1226). */
int hb200_set_reference(hb200_engine* e, const uint64_t* f0, double H00);
/* proc_map%map(0:nprocs*nslots-1) (src/spawn_data.F90:13-24, src/load_balancing.F90:141-172). */
int hb200_set_proc_map(hb200_engine* e, const int32_t* map, int32_t n);
/* Load balancing (src/load_balancing.F90:209-323, src/qmc_common.F90:1019-1078, 1332-1390).  The policy stays on the
 * host; the engine supplies what it needs from the resident list and moves the determinants:
 *   hb200_slot_populations      initialise_slot_pop: this rank's population in each of the nprocs*nslots slots
 *   (host)                      MPI_AllReduce of the slots, check_imbalance / find_processors / redistribute_slots
 *   hb200_set_proc_map          the modified proc_map%map
 *   hb200_redistribute_particles  redistribute_particles: determinants now owned elsewhere -> spawn blocks
 *   hb200_comm_spawn + hb200_annihilate_spawn + hb200_annihilate_main   direct_annihilation of the moved list */
int hb200_slot_populations(hb200_engine* e, double* slot_pop, int32_t n);
int hb200_redistribute_particles(hb200_engine* e, double* nsent);

/* particle_t upload/download: states(W,N), pops(1,N), dat(1,N), sorted ascending
 * (src/qmc_data.f90:615-682); used at init, restart read/write (src/restart_hdf5.F90:307,539). */
int hb200_upload_psips(hb200_engine* e, const uint64_t* states, const int64_t* pops, const double* dat, int64_t nstates);
/* The same upload split in two so that it overlaps the propagation of the list that is already resident (for hosts
 * that keep particle_t on the CPU and push a list every report loop): _begin starts the host-to-device copy on a
 * second stream into a staging buffer and returns; _commit waits for it and makes that list the current one. */
int hb200_upload_psips_begin(hb200_engine* e, const uint64_t* states, const int64_t* pops, const double* dat, int64_t nstates);
int hb200_upload_psips_commit(hb200_engine* e);
int hb200_download_psips(hb200_engine* e, uint64_t* states, int64_t* pops, double* dat, int64_t capacity, int64_t* nstates);
int64_t hb200_nstates(hb200_engine* e);

/* ncycles full MC cycles: the body of do_fciqmc's icycle loop (src/fciqmc.f90:293-398):
 * init_mc_cycle, the idet loop (spawn + death + projected energy), direct_annihilation, end_mc_cycle. */
int hb200_iterate(hb200_engine* e, int32_t ncycles, const hb200_iter_in* in, hb200_iter_out* out);

/* CCMC (ccmc{...}; src/ccmc.f90:603-896), stochastic cluster selection: one thread per selection attempt runs
 * select_cluster (src/ccmc_selection.f90:93-392), do_ccmc_accumulation, spawner_ccmc and stochastic_ccmc_death
 * (src/ccmc_death_spawning.f90:11-441); spawned and killed excips enter the spawn list and are annihilated by the same
 * stages as FCIQMC.  ex_level = reference%ex_level (2 = CCSD, 3 = CCSDT); the engine must have been created with
 * trunc_level = ex_level. */
int hb200_ccmc_spawn(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, int32_t ex_level, hb200_ccmc_out* out);
int hb200_ccmc_iterate(hb200_engine* e, int32_t ncycles, const hb200_iter_in* in, int32_t ex_level, hb200_iter_out* out);
/* spawn%hash_shift (number of CCMC cycles done: the engine adds one per cycle, src/ccmc.f90:540,625) and
 * spawn%move_freq (ccmc_in%move_freq, default 5) of the time-varying owner rule (src/spawning.F90:812-836); call after
 * a restart.  With nprocs > 1 every cycle ends with redistribute_particles (src/qmc_common.F90:505-595) through the
 * same NCCL exchange as the spawned excips, and the reference population is broadcast from its owner (get_D0_info). */
int hb200_ccmc_set_hash_shift(hb200_engine* e, int32_t hash_shift, int32_t move_freq);
/* ccmc = { full_non_composite = true } (ccmc_in%full_nc): every excitor is propagated as a non-composite cluster of its
 * own (select_nc_cluster, do_nc_ccmc_propagation) and dies in place (stochastic_ccmc_death_nc), composite clusters of
 * size >= 2 are selected stochastically and the reference is selected nint(|N_0|) times. */
int hb200_ccmc_set_full_nc(hb200_engine* e, int32_t full_nc);

/* qmc = { quasi_newton = true } (src/qmc_data.f90:227-241,866-884): spawning amplitudes are scaled by
 * calc_qn_spawned_weighting and the death probability by calc_qn_weighting / quasi_newton_pop_control
 * (src/spawning.F90:2063-2137, src/death.f90:87).  sp_fock[0..nbasis] (entry 0 unused) is propagator%sp_fock as
 * init_sp_fock makes it (src/qmc.F90:1064-1090: sp_eigv, plus exchange and Madelung terms for the 3D UEG), ref_fock_sum
 * its sum over the reference; threshold / value / pop_control as init_quasi_newton resolves them (:1092-1160).
 * sp_fock = NULL switches the propagator off.  FCIQMC only in this version. */
int hb200_set_quasi_newton(hb200_engine* e, const double* sp_fock, double ref_fock_sum, double threshold, double value,
                           double pop_control);
/* qmc = { pattempt_update = true } (qmc_in%pattempt_update, src/qmc.F90:1049-1060).  hb200_set_pattempt replaces
 * excit_gen_data%pattempt_single / pattempt_double; while accumulate != 0 every allowed excitation generated by
 * hb200_iterate / hb200_ccmc_iterate adds |H_ij| pattempt_{single,double} / pgen and 1 to the p_single_double_coll_t sums
 * (update_p_single_double_data, src/spawning.F90:2139-2215; call sites src/spawning.F90:104-109 and
 * src/ccmc_death_spawning.f90:150-157).  hb200_get_ps_stats returns this rank's rep_accum
 * {h_pgen_singles_sum, excit_gen_singles, h_pgen_doubles_sum, excit_gen_doubles} and optionally zeroes it; the host does
 * communicate_pattempt_single_data + update_pattempt_single (src/spawning.F90:2217-2372) once per report loop.  As in
 * src/check_input.F90:192-197 accumulation is refused for the UEG and for excit_gen = heat_bath. */
int hb200_set_pattempt(hb200_engine* e, double pattempt_single, double pattempt_double, int32_t accumulate);
/* excit_gen = renorm_spin / no_renorm_spin (gen_excit_mol_spin, gen_excit_mol_no_renorm_spin, src/excit_gen_mol.f90:103-193,
 * 286-380; choose_ij_spin_mol :684-800): qmc_in%pattempt_parallel, the probability that i and j have parallel spins.  A
 * negative value computes it as find_parallel_spin_prob_mol does (src/qmc_common.F90:262-377, src/qmc.F90:974-988); call
 * after hb200_set_system_read_in and before the first iteration. */
int hb200_set_pattempt_parallel(hb200_engine* e, double pattempt_parallel);
/* excit_gen = power_pitzer_orderN ('heat_bath_power_pitzer_ref'; gen_excit_mol_power_pitzer_orderN,
 * src/excit_gen_power_pitzer_mol.F90:941-1258): builds the reference-mapped alias tables ppn_i_s, ppn_ia_s, ppn_i_d,
 * ppn_ij_d, ppn_ia_d, ppn_jb_d on the device as init_excit_mol_power_pitzer_orderN does (:215-572, with
 * check_min_weight_ratio :140-213; min_weight = qmc_in%power_pitzer_min_weight, default 0.01).  Call after
 * hb200_set_reference (the tables depend on the reference determinant). */
int hb200_build_power_pitzer_orderN(hb200_engine* e, double min_weight);
/* excit_gen = power_pitzer (gen_excit_mol_power_pitzer_occ_ref, src/excit_gen_power_pitzer_mol.F90:650-939): the
 * reference's pp_ia_d / pp_jb_d alias tables (init_excit_mol_power_pitzer_occ_ref, :19-138); after hb200_set_reference. */
int hb200_build_power_pitzer(hb200_engine* e, double min_weight);
double hb200_get_pattempt_parallel(hb200_engine* e);
int hb200_get_ps_stats(hb200_engine* e, double* out4, int32_t reset);

/* Stage-level entry points (same state machine as hb200_iterate, one stage per call). */
/* do idet loop: decoder_ptr, set_parent_flag, update_proj_energy_ptr, decide_nattempts,
 * do_fciqmc_spawning_attempt, stochastic_death (src/fciqmc.f90:315-371). */
int hb200_spawn_death(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, hb200_iter_out* out);
/* comm_spawn_t (src/spawn_data.F90:624-739): all-to-all of the per-destination blocks over NCCL. */
int hb200_comm_spawn(hb200_engine* e);
/* annihilate_wrapper_spawn_t (sort + annihilate_spawn_t[_initiator], src/spawn_data.F90:359-418,859-1101) */
int hb200_annihilate_spawn(hb200_engine* e);
/* annihilate_main_list_wrapper (src/annihilation.f90:214-290): annihilate_main_list[_initiator],
 * remove_unoccupied_dets, round_low_population_spawns, insert_new_walkers. */
int hb200_annihilate_main(hb200_engine* e, uint32_t cycle, hb200_iter_out* out);
/* spawn%sdata(element_len, n) of the current stage (element = W string words, population, flag). */
int hb200_download_spawn(hb200_engine* e, int64_t* sdata, int64_t capacity, int64_t* n);
int hb200_upload_spawn(hb200_engine* e, const int64_t* sdata, int64_t n);
/* spawn%head after the spawning stage: counts[d] = number of elements in the block destined for rank d
 * (the send counts of comm_spawn_t, src/spawn_data.F90:686-693); hb200_download_spawn returns the blocks
 * concatenated in this order.  Lets a host stage the exchange itself (hb200_upload_spawn on the receiver). */
int hb200_spawn_counts(hb200_engine* e, int64_t* counts, int32_t nprocs);

/* Pure-function batches evaluated on the device (parity probes; also used by the host for init):
 * sc0_ptr over a list of determinants (src/hamiltonian_molecular.f90:73-139). */
int hb200_sc0_batch(hb200_engine* e, const uint64_t* states, int64_t n, double* out);
/* One spawning attempt per (determinant, attempt index) with the engine's Philox stream:
 * gen_excit_ptr%full + attempt_to_spawn (src/spawning.F90:34-124).  iout[n][8] = nexcit, from1, from2,
 * to1, to2, perm, allowed, owner; dout[n][2] = pgen, hmatel; nspawn[n]. */
int hb200_gen_excit_batch(hb200_engine* e, const uint64_t* states, const int64_t* pops, const uint32_t* attempt,
                          int64_t n, uint32_t cycle, double tau, int32_t* iout, double* dout, int64_t* nspawn);

/* The same with INJECTED random numbers (level 1 of the correctness contract: "kernels fed identical injected random
 * numbers"): attempt k draws rn[k][0], rn[k][1], ... (nrn per attempt, uniform on [0,1)) in the order the reference's
 * gen_excit_ptr%full and then attempt_to_spawn draw from dSFMT, instead of the engine's Philox stream; nused[k] = how
 * many were drawn.  A Fortran host feeds the numbers its own generator consumed and compares choice, pgen, H_ij, nspawn. */
int hb200_gen_excit_batch_rn(hb200_engine* e, const uint64_t* states, const int64_t* pops, const double* rn, int32_t nrn,
                             int64_t n, double tau, int32_t* iout, double* dout, int64_t* nspawn, int32_t* nused);

/* Wall-Chebyshev propagator (qmc = { chebyshev = {...} }; src/propagators.f90:11-208, src/fciqmc.f90:298-299,427): the
 * host keeps the spectral range and the weights 1/(S_i - E_0) (init_chebyshev, update_chebyshev) and runs the `order`
 * sub-cycles of an MC cycle as separate hb200_iterate(1) calls; this sets the weight applied to the spawning amplitude
 * (src/spawning.F90:117-118) and to the death probability (src/death.f90:89) until the next call.  Default 1. */
int hb200_set_propagator_weight(hb200_engine* e, double weight);

/* Semi-stochastic projection (semi_stoch = {...}; src/semi_stoch.F90, deterministic_annihilation of
 * src/annihilation.f90:488-535), the reference's default projection mode (separate_annihilation = true).
 *
 * hb200_set_determ_space replaces init_semi_stoch_t (src/semi_stoch.F90:134-377) once the host has chosen the space
 * (create_high_pop_space / create_ci_determ_space / read_determ_from_file are host logic; hande_b200.semi_stoch mirrors the
 * first): dets = determ%dets, all determ%tot_size determinants rank by rank (ceil(nbasis/64) words each), every rank's
 * part in ascending list order; sizes = determ%sizes(0:nprocs-1).  It builds the membership table check_if_determ uses
 * (:795-824), this rank's slice of the deterministic Hamiltonian on the device (create_determ_hamil, :533-685,
 * <D_i|H|D_j> - H00 delta_ij times the quasi-Newton weight of D_j, elements above depsilon), rho_minus_qn_weight, and
 * adds the rank's deterministic states that its main list lacks with zero population (add_determ_dets_to_psip_list,
 * :728-793).  From then on hb200_iterate and the staged calls treat deterministic states as the reference does: always
 * initiators (set_parent_flag), no death step, spawning onto another deterministic state cancelled
 * (src/fciqmc.f90:368,726-736), never rounded nor removed (remove_unoccupied_dets), and once per cycle
 * v <- -tau (H - S) v over the space (determ_proj_separate_annihil, :1009-1094; the amplitudes of all ranks are
 * all-gathered with NCCL) is rounded stochastically to the amplitude resolution and added to their populations.
 * All sizes zero switches the projection off.  hb200_upload_psips with a space set requires the new list to hold every
 * deterministic state of the rank. */
int hb200_set_determ_space(hb200_engine* e, const uint64_t* dets, const int32_t* sizes);
/* determ%hamil of this rank, stored by column (column j = deterministic state j of this rank, rows in determ%dets
 * order - the order csrpgemv's transposed product accumulates in, lib/local/csr.f90:188-194).  Returns nnz; with
 * non-null arrays fills col_ptr[sizes(iproc)+1], row[nnz], val[nnz]. */
int64_t hb200_determ_hamil(hb200_engine* e, int64_t* col_ptr, int32_t* row, double* val);
/* determ%vector of this rank: which = 0 the amplitudes of its deterministic states now (set_determ_info, :826-857);
 * which = 1 the result of the last projection. */
int hb200_determ_vector(hb200_engine* e, int32_t which, double* vec);
/* Staged form of the projection for hosts that drive the cycle themselves (between hb200_spawn_death and
 * hb200_annihilate_main): full_vector = determ%full_vector as the host's mpi_allgatherv delivers it (tot_size doubles,
 * rank by rank), or NULL to gather on the device (one rank, or NCCL).  Does determ_proj_separate_annihil and
 * deterministic_annihilation for MC cycle `cycle`. */
int hb200_determ_project(hb200_engine* e, const hb200_iter_in* in, uint32_t cycle, const double* full_vector);

/* NCCL bootstrap for nprocs > 1: rank 0 calls hb200_get_unique_id, the host broadcasts the 128 bytes
 * (MPI_Bcast in the Fortran host, torch.distributed in this repo's harness), every rank calls
 * hb200_comm_init.  Replaces MPI_COMM_WORLD use in comm_spawn_t. */
int hb200_get_unique_id(uint8_t id[128]);
int hb200_comm_init(hb200_engine* e, const uint8_t id[128]);

/* Peer-to-peer exchange of the spawn blocks inside hb200_iterate (all ranks on one node, one process per GPU).
 * Replaces comm_spawn_t's MPI_Alltoall + MPI_Alltoallv (src/spawn_data.F90:624-739) by stores into the owner rank's
 * receive buffer over NVLink that overlap the spawning step (the north-star's "exchange overlaps the next spawning
 * chunk"): the spawning kernel runs in chunks of tiles, and while chunk c+1 spawns, chunk c's part of every
 * per-destination block is pushed - one system-scope atomic on the destination's head counter reserves the room, a
 * copy kernel writes the elements through the peer mapping.  hb200_p2p_export allocates this rank's receive block and
 * returns its 64-byte CUDA IPC handle; the host all-gathers the handles (MPI_Allgather in the Fortran host) and every
 * rank calls hb200_p2p_import(handles[nprocs][64]).  Call after hb200_comm_init (the NCCL communicator is still used
 * for the collective that ends an exchange).  Without these calls hb200_iterate uses the NCCL send/recv path. */
int hb200_p2p_export(hb200_engine* e, uint8_t handle[64]);
int hb200_p2p_import(hb200_engine* e, const uint8_t* handles, int32_t nprocs);
/* All ranks must run the same exchange: when hb200_p2p_import fails on any rank (GPUs without peer access), the host
 * calls hb200_p2p_enable(h, 0) on every rank and hb200_iterate falls back to the NCCL send/recv path. */
int hb200_p2p_enable(hb200_engine* e, int32_t on);
/* A host without NCCL (plain MPI; or several ranks sharing one GPU, which NCCL refuses) ends an exchange with its own
 * barrier instead: fn(arg) must return once every rank has called it (MPI_Barrier).  With a host barrier set,
 * hb200_comm_init is not needed for hb200_iterate. */
typedef void (*hb200_barrier_fn)(void* arg);
int hb200_set_host_barrier(hb200_engine* e, hb200_barrier_fn fn, void* arg);

/* Timing of the stages of the last hb200_iterate call, milliseconds (CUDA events on the engine stream):
 * ms[0] spawn+death kernel, ms[1] exchange, ms[2] sort+annihilate_spawn, ms[3] main-list annihilation+merge,
 * ms[4] total, ms[5] the k_spawn_death kernel alone.  Counters: cnt[0] = spawn kernel launches, cnt[1] = all kernel launches. */
int hb200_last_timing(hb200_engine* e, double ms[8], int64_t cnt[4]);

#ifdef __cplusplus
}
#endif
#endif
