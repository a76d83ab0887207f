// ORACLE (test infrastructure, NOT product code): C API over the C++ restatement so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can drive it through ctypes.
#include <cstring>
#include <string>
#include <thread>
#include <chrono>
#include "system.hpp"
#include "rng.hpp"
#include "excit_gen.hpp"
#include "fciqmc.hpp"
#include "ccmc.hpp"

using namespace oracle;

static thread_local std::string g_err;

#define ORC_TRY try {
#define ORC_CATCH(rv)                     \
    }                                     \
    catch (const std::exception& e) {     \
        g_err = e.what();                 \
        return rv;                        \
    }

extern "C" {

struct orc_qmc_in {
    double tau;
    int seed;
    double D0_population;
    int ncycles, nreport;
    double target_particles, initial_shift, shift_damping, vary_shift_from;
    int vary_shift_from_proje;
    int initiator_approx;
    double initiator_pop;
    int real_amplitudes;
    double spawn_cutoff;
    int excit_gen;
    double pattempt_single, pattempt_double;
    int64_t walker_length, spawned_walker_length;
    int ex_level;
    int nprocs, nslots;
    int rng_kind;
    int literal_event_int32;
};

const char* orc_last_error() { return g_err.c_str(); }

int orc_set_ref_lib(const char* path) {
    ORC_TRY
    DsfmtLib::get(path);
    return 0;
    ORC_CATCH(-1)
}

void* orc_create() { return (Oracle*)new OracleCcmc(); }
void orc_destroy(void* h) { delete (OracleCcmc*)(Oracle*)h; }

int orc_read_fcidump(void* h, const char* path, int nel, int ms, int sym, int cas_nel, int cas_norb) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    ReadInOpts opt;
    opt.nel = nel; opt.ms = ms; opt.sym = sym; opt.cas[0] = cas_nel; opt.cas[1] = cas_norb;
    o->sys = System();
    read_in_file(o->sys, path, opt);
    return 0;
    ORC_CATCH(-1)
}

// sys = ueg { electrons, ms, dim = 3, cutoff, rs } (lua_hande_system / init_system + init_model_basis_fns)
int orc_init_ueg(void* h, int nel, int ms, double rs, double ecutoff) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    o->sys = System();
    init_ueg_system(o->sys, nel, ms, rs, ecutoff);
    return 0;
    ORC_CATCH(-1)
}
// reference = { det = {...} }: explicit reference determinant used by orc_init (n = 0 clears it)
void orc_set_ref_det_list(void* h, const int* occ, int n) {
    Oracle* o = (Oracle*)h;
    o->in.ref_det.assign(occ, occ + n);
}
// UEG tables for the host side of the engine: kvec[3*nbasis] (0-based storage of the 1-based functions),
// info[0..5] = kmax, offset, offset_inds[3], n_lookup; dinfo[0..2] = L, rs, ecutoff
void orc_ueg_info(void* h, int* kvec, int64_t* info, double* dinfo) {
    const System& s = ((Oracle*)h)->sys;
    const UegData& u = s.ueg;
    for (int i = 1; i <= s.nbasis; ++i)
        for (int d = 0; d < 3; ++d) kvec[3 * (i - 1) + d] = u.l[3 * i + d];
    info[0] = u.kmax; info[1] = u.offset; info[2] = u.offset_inds[0]; info[3] = u.offset_inds[1]; info[4] = u.offset_inds[2];
    info[5] = (int64_t)u.lookup.size(); info[6] = u.tK; info[7] = u.tD; info[8] = (int64_t)u.ternary.size();
    dinfo[0] = u.L; dinfo[1] = u.rs; dinfo[2] = u.ecutoff;
}
const int* orc_ueg_lookup(void* h) { return ((Oracle*)h)->sys.ueg.lookup.data(); }
const uint64_t* orc_ueg_ternary(void* h) { return ((Oracle*)h)->sys.ueg.ternary.data(); }

// info[0..]: nbasis, nel, W, nsym_tot, sym0, sym_max, nalpha, nbeta, uhf, pg_mask, Lz_mask, Lz_offset,
//            gamma_sym, max_nbss, nvirt, nvirt_alpha, nvirt_beta, symmetry, int_err, n_two_body_channels
void orc_sys_info(void* h, int64_t* info) {
    const System& s = ((Oracle*)h)->sys;
    int64_t v[] = {s.nbasis, s.nel, s.W, s.nsym_tot, s.sym0, s.sym_max, s.nalpha, s.nbeta, s.uhf, s.pg_mask,
                   s.Lz_mask, s.Lz_offset, s.gamma_sym, s.max_nbss, s.nvirt, s.nvirt_alpha, s.nvirt_beta,
                   s.symmetry, s.int_err, (int64_t)s.two_body.size(), s.nintgrls};
    memcpy(info, v, sizeof(v));
}
double orc_ecore(void* h) { return ((Oracle*)h)->sys.Ecore; }

// arrays of length nbasis (0-based storage of the 1-based functions)
void orc_basis(void* h, int* sym, int* ms, int* spatial, int* sym_index, int* sym_spin_index, double* sp_eigv) {
    const System& s = ((Oracle*)h)->sys;
    for (int i = 1; i <= s.nbasis; ++i) {
        sym[i - 1] = s.bf[i].sym; ms[i - 1] = s.bf[i].ms; spatial[i - 1] = s.bf[i].spatial_index;
        sym_index[i - 1] = s.bf[i].sym_index; sym_spin_index[i - 1] = s.bf[i].sym_spin_index;
        sp_eigv[i - 1] = s.bf[i].sp_eigv;
    }
}
// nbasis_sym_spin: [2*nsym_tot]; sym_spin_basis_fns: [max_nbss*2*nsym_tot]
void orc_sym_tables(void* h, int* nbss, int* ssbf) {
    const System& s = ((Oracle*)h)->sys;
    memcpy(nbss, s.nbasis_sym_spin.data(), s.nbasis_sym_spin.size() * sizeof(int));
    memcpy(ssbf, s.sym_spin_basis_fns.data(), s.sym_spin_basis_fns.size() * sizeof(int));
}
double orc_get_one_body(void* h, int i, int j) { return ((Oracle*)h)->sys.get_one_body_real(i, j); }
double orc_get_two_body(void* h, int i, int j, int a, int b) { return ((Oracle*)h)->sys.get_two_body_real(i, j, a, b); }
const double* orc_two_body_store(void* h, int chan) { return ((Oracle*)h)->sys.two_body[chan].data(); }

static Det mkdet(const System& s, const uint64_t* f) {
    Det d;
    for (int k = 0; k < s.W; ++k) d.w[k] = f[k];
    return d;
}

double orc_sc0(void* h, const uint64_t* f) {
    const System& s = ((Oracle*)h)->sys;
    return diag_hmatel(s, mkdet(s, f));
}
// update_proj_energy_ptr for unit population: returns <D|H|D0> as accumulated into proj_energy (0 for the reference)
double orc_proj_hmatel(void* h, const uint64_t* f) {
    Oracle* o = (Oracle*)h;
    DetInfo d;
    decode_det_occ(o->sys, mkdet(o->sys, f), d);
    double d0 = 0.0, pe = 0.0;
    o->update_proj_energy(d, 1.0, d0, pe);
    return pe;
}
// get_hmatel(f1, f2) (src/hamiltonian_molecular.f90:12-71, src/hamiltonian_ueg.f90:12-69), diagonal included
double orc_get_hmatel(void* h, const uint64_t* f1, const uint64_t* f2) {
    Oracle* o = (Oracle*)h;
    return o->get_hmatel(mkdet(o->sys, f1), mkdet(o->sys, f2));
}
// Slater-Condon single: matrix element <D|H|D_i^a> incl. permutation sign
double orc_sc1(void* h, const uint64_t* f, int i, int a) {
    const System& s = ((Oracle*)h)->sys;
    Det d = mkdet(s, f);
    int occ[MAXNEL];
    s.decode(d, occ);
    Excit e; e.nexcit = 1; e.from_orb[0] = i; e.to_orb[0] = a;
    s.find_excitation_permutation1(d, e);
    return s.slater_condon1_excit(occ, i, a, e.perm);
}
double orc_sc2(void* h, const uint64_t* f, int i, int j, int a, int b) {
    const System& s = ((Oracle*)h)->sys;
    Det d = mkdet(s, f);
    Excit e; e.nexcit = 2; e.from_orb[0] = i; e.from_orb[1] = j; e.to_orb[0] = a; e.to_orb[1] = b;
    s.find_excitation_permutation2(d, e);
    return s.slater_condon2_excit(i, j, a, b, e.perm);
}
// symmetry/spin-checked variants used to enumerate the full Hamiltonian row in tests
double orc_sc1_checked(void* h, const uint64_t* f, int i, int a) {
    const System& s = ((Oracle*)h)->sys;
    Det d = mkdet(s, f);
    if (s.bf[i].ms != s.bf[a].ms) return 0.0;
    int occ[MAXNEL];
    s.decode(d, occ);
    Excit e; e.nexcit = 1; e.from_orb[0] = i; e.to_orb[0] = a;
    s.find_excitation_permutation1(d, e);
    return s.slater_condon1(occ, i, a, e.perm);
}
double orc_sc2_checked(void* h, const uint64_t* f, int i, int j, int a, int b) {
    const System& s = ((Oracle*)h)->sys;
    Det d = mkdet(s, f);
    Excit e; e.nexcit = 2; e.from_orb[0] = i; e.from_orb[1] = j; e.to_orb[0] = a; e.to_orb[1] = b;
    s.find_excitation_permutation2(d, e);
    return s.slater_condon2(i, j, a, b, e.perm);
}

int32_t orc_murmur_bit_string(void* h, const uint64_t* f, uint32_t seed) {
    const System& s = ((Oracle*)h)->sys;
    return murmurhash_bit_string(mkdet(s, f), s.nbasis, seed);
}
uint32_t orc_murmur2(const void* key, int len, uint32_t seed) { return murmurhash2(key, len, seed); }
uint32_t orc_ref_murmur2(const void* key, int len, uint32_t seed) {
    ORC_TRY
    return DsfmtLib::get().murmur2(key, len, seed);
    ORC_CATCH(0)
}
int orc_owner(void* h, const uint64_t* f) {
    Oracle* o = (Oracle*)h;
    return o->owner(mkdet(o->sys, f));
}

// Fill n doubles from the reference dSFMT stream through the 50,000-element buffer.
int orc_dsfmt_stream(int seed, int n, double* out) {
    ORC_TRY
    DsfmtRng r(seed, 50000);
    for (int i = 0; i < n; ++i) out[i] = r.next();
    return 0;
    ORC_CATCH(-1)
}
// Philox stream sample: n draws for (seed, cycle, purpose, f, attempt)
void orc_philox_stream(uint32_t seed, uint32_t cycle, uint32_t purpose, const uint64_t* f, int W, uint32_t attempt,
                       int n, double* out) {
    PhiloxRng r(seed);
    r.set_cycle(cycle);
    Det d;
    for (int k = 0; k < W; ++k) d.w[k] = f[k];
    r.begin(purpose, d, W, attempt);
    for (int i = 0; i < n; ++i) out[i] = r.next();
}

int orc_set_qmc(void* h, const orc_qmc_in* q) {
    Oracle* o = (Oracle*)h;
    QmcIn& in = o->in;
    in.tau = q->tau; in.seed = q->seed; in.D0_population = q->D0_population;
    in.ncycles = q->ncycles; in.nreport = q->nreport; in.target_particles = q->target_particles;
    in.initial_shift = q->initial_shift; in.shift_damping = q->shift_damping;
    in.vary_shift_from = q->vary_shift_from; in.vary_shift_from_proje = q->vary_shift_from_proje;
    in.initiator_approx = q->initiator_approx; in.initiator_pop = q->initiator_pop;
    in.real_amplitudes = q->real_amplitudes; in.spawn_cutoff = q->spawn_cutoff;
    in.excit_gen = q->excit_gen; in.pattempt_single = q->pattempt_single; in.pattempt_double = q->pattempt_double;
    in.walker_length = q->walker_length; in.spawned_walker_length = q->spawned_walker_length;
    in.ex_level = q->ex_level; in.nprocs = q->nprocs; in.nslots = q->nslots; in.rng_kind = q->rng_kind;
    in.literal_event_int32 = q->literal_event_int32;
    return 0;
}
int orc_init(void* h) {
    ORC_TRY
    ((Oracle*)h)->init();
    return 0;
    ORC_CATCH(-1)
}
int orc_run(void* h) {
    ORC_TRY
    ((Oracle*)h)->run();
    return 0;
    ORC_CATCH(-1)
}
// ccmc{...}: stochastic-selection CCMC on the current system (single rank)
int orc_run_ccmc(void* h) {
    ORC_TRY
    ((OracleCcmc*)(Oracle*)h)->run_ccmc();
    return 0;
    ORC_CATCH(-1)
}
void orc_get_nattempts_rows(void* h, int64_t* out) {
    OracleCcmc* o = (OracleCcmc*)(Oracle*)h;
    for (size_t i = 0; i < o->nattempts_rows.size(); ++i) out[i] = o->nattempts_rows[i];
}
int orc_nrows(void* h) { return (int)((Oracle*)h)->rows.size(); }
// rows[nrows][8]: iter, shift, proj_energy, D0, nparticles, nstates, nspawn_events, rspawn
void orc_get_rows(void* h, double* out) {
    Oracle* o = (Oracle*)h;
    for (size_t i = 0; i < o->rows.size(); ++i) {
        const ReportRow& r = o->rows[i];
        double* p = out + 8 * i;
        p[0] = r.iter; p[1] = r.shift; p[2] = r.proj_energy; p[3] = r.D0_population; p[4] = r.nparticles;
        p[5] = (double)r.nstates; p[6] = (double)r.nspawn_events; p[7] = r.rspawn;
    }
}
// ref[0]=H00, ref[1]=pattempt_single, ref[2]=pattempt_double; occ0[nel]; f0[W]
void orc_reference(void* h, double* ref, int* occ0, uint64_t* f0) {
    Oracle* o = (Oracle*)h;
    ref[0] = o->H00; ref[1] = o->eg.pattempt_single; ref[2] = o->eg.pattempt_double;
    for (int i = 0; i < o->sys.nel; ++i) occ0[i] = o->occ_list0[i];
    for (int k = 0; k < o->sys.W; ++k) f0[k] = o->f0.w[k];
}
int64_t orc_nstates(void* h, int rank) { return ((Oracle*)h)->ranks[rank].nstates; }
double orc_nparticles(void* h, int rank) { return ((Oracle*)h)->ranks[rank].nparticles; }
void orc_get_psips(void* h, int rank, uint64_t* states, int64_t* pops, double* dat) {
    Oracle* o = (Oracle*)h;
    RankState& r = o->ranks[rank];
    int W = o->sys.W;
    for (int64_t i = 0; i < r.nstates; ++i) {
        for (int k = 0; k < W; ++k) states[i * W + k] = r.states[i].w[k];
        pops[i] = r.pops[i];
        dat[i] = r.dat[i];
    }
}
void orc_set_psips(void* h, int rank, int64_t n, const uint64_t* states, const int64_t* pops, const double* dat) {
    Oracle* o = (Oracle*)h;
    RankState& r = o->ranks[rank];
    int W = o->sys.W;
    r.states.resize(n); r.pops.resize(n); r.dat.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        r.states[i] = Det();
        for (int k = 0; k < W; ++k) r.states[i].w[k] = states[i * W + k];
        r.pops[i] = pops[i];
        r.dat[i] = dat[i];
    }
    r.nstates = n;
    o->recompute_nparticles(r);
    o->determ.doing = false;   // a deterministic space is built on the list it is given (orc_init_semi_stoch afterwards)
}
void orc_set_reference_det(void* h, const uint64_t* f0) {
    Oracle* o = (Oracle*)h;
    o->f0 = mkdet(o->sys, f0);
    o->occ_list0.resize(o->sys.nel);
    o->sys.decode(o->f0, o->occ_list0.data());
    o->H00 = diag_hmatel(o->sys, o->f0);
}

// Run ncycles MC cycles with fixed shift / proj_energy_old / tau (what hb200_iterate does).
// out[0]=sum proj_energy, out[1]=sum D0, out[2]=nparticles(total), out[3]=nstates(total),
// out[4]=nspawn_events(last cycle, total), out[5]=ndeath(last cycle, total), out[6]=rspawn sum,
// out[7]=error flag, out[8]=nattempts(last cycle), out[9]=total rng draws
int orc_iterate(void* h, int ncycles, uint32_t first_cycle_id, double tau, double shift, double proj_energy_old,
                double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    o->tau = tau; o->shift = shift; o->est.proj_energy_old = proj_energy_old;
    for (auto& r : o->ranks) { r.rspawn = 0.0; r.proj_energy = 0.0; r.D0_population = 0.0; }
    for (int c = 0; c < ncycles; ++c) o->mc_cycle(first_cycle_id + c);
    double pe = 0, d0 = 0, np = 0, rs = 0;
    int64_t ns = 0, nev = 0, nd = 0, na = 0;
    uint64_t draws = 0;
    bool err = false;
    for (auto& r : o->ranks) {
        pe += r.proj_energy; d0 += r.D0_population; np += r.nparticles; ns += r.nstates;
        nev += r.nspawn_events; nd += r.ndeath; rs += r.rspawn; na += r.nattempts;
        err = err || r.spawn_error || r.psip_error;
        draws += r.rng->ndraws;
    }
    out[0] = pe; out[1] = d0; out[2] = np; out[3] = (double)ns; out[4] = (double)nev; out[5] = (double)nd;
    out[6] = rs; out[7] = err ? 1.0 : 0.0; out[8] = (double)na; out[9] = (double)draws;
    return 0;
    ORC_CATCH(-1)
}

// Stage probes: run only the spawn/death loop of one cycle on every emulated rank (no comm, no annihilation)
// and read back the per-destination send blocks concatenated by destination (spawn%sdata before comm_spawn_t).
int orc_stage_spawn(void* h, uint32_t cycle_id, double tau, double shift, double proj_energy_old, double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    o->tau = tau; o->shift = shift; o->est.proj_energy_old = proj_energy_old;
    double pe = 0, d0 = 0; int64_t nev = 0, nd = 0;
    for (auto& r : o->ranks) {
        r.proj_energy = 0.0; r.D0_population = 0.0;
        o->spawn_death_rank(r, cycle_id);
        pe += r.proj_energy; d0 += r.D0_population; nev += r.nspawn_events; nd += r.ndeath;
    }
    out[0] = pe; out[1] = d0; out[2] = (double)nev; out[3] = (double)nd;
    return 0;
    ORC_CATCH(-1)
}
// CCMC stage probe: cluster selection + spawning + death of one cycle on rank 0 (no annihilation); the spawn list is
// read with orc_get_spawn and the cycle finished with orc_stage_annihilate.
// out: proj_energy_cycle, D0_population_cycle, D0_normalisation, nattempts, nattempts_spawn, nspawn_events, ndeath
int orc_ccmc_stage_spawn(void* h, uint32_t cycle_id, double tau, double shift, double proj_energy_old, double* out) {
    ORC_TRY
    OracleCcmc* o = (OracleCcmc*)(Oracle*)h;
    o->tau = tau; o->shift = shift; o->est.proj_energy_old = proj_energy_old;
    RankState& r = o->ranks[0];
    for (auto& rr : o->ranks) { rr.proj_energy = 0.0; rr.D0_population = 0.0; }
    o->ccmc_spawn_stage(cycle_id);
    out[0] = o->last.proj_energy; out[1] = o->last.D0_population; out[2] = o->last.D0_normalisation;
    out[3] = (double)o->last.nattempts; out[4] = (double)o->last.nattempts_spawn; out[5] = (double)o->last.nspawn_events;
    out[6] = (double)o->last.ndeath; out[7] = (double)o->last.ndeath_nc;
    return 0;
    ORC_CATCH(-1)
}
void orc_ccmc_set_full_nc(void* h, int full_nc) { ((OracleCcmc*)(Oracle*)h)->full_nc = full_nc != 0; }
// qmc = { pattempt_update = true }; orc_ccmc_get_pattempt_log copies up to n values of pattempt_single after each change
void orc_set_pattempt_update(void* h, int on) { ((Oracle*)h)->vary_psingles = on != 0; }
// rep_accum of one rank: h_pgen_singles_sum, excit_gen_singles, h_pgen_doubles_sum, excit_gen_doubles
void orc_get_ps_stats(void* h, int rank, double* out, int reset) {
    Oracle* o = (Oracle*)h;
    for (int k = 0; k < 4; ++k) out[k] = 0.0;
    if (rank < (int)o->ps_rep_accum.size()) {
        auto& a = o->ps_rep_accum[rank];
        out[0] = a.h_pgen_singles_sum; out[1] = a.excit_gen_singles; out[2] = a.h_pgen_doubles_sum; out[3] = a.excit_gen_doubles;
        if (reset) a = Oracle::PsColl();
    }
}
// power_pitzer_orderN tables: which = 0..5 (i_s, ia_s, i_d, ij_d, ia_d, jb_d); part = 0 w, 1 U, 2 tot (doubles), K via _i
const double* orc_ppn_ptr_d(void* h, int which, int part, int64_t* n) {
    Oracle* o = (Oracle*)h;
    AliasCols* t[6] = {&o->eg.ppn.i_s, &o->eg.ppn.ia_s, &o->eg.ppn.i_d, &o->eg.ppn.ij_d, &o->eg.ppn.ia_d, &o->eg.ppn.jb_d};
    std::vector<double>& v = part == 0 ? t[which]->w : (part == 1 ? t[which]->U : t[which]->tot);
    *n = (int64_t)v.size();
    return v.data();
}
const int* orc_ppn_ptr_i(void* h, int which, int64_t* n) {
    Oracle* o = (Oracle*)h;
    AliasCols* t[6] = {&o->eg.ppn.i_s, &o->eg.ppn.ia_s, &o->eg.ppn.i_d, &o->eg.ppn.ij_d, &o->eg.ppn.ia_d, &o->eg.ppn.jb_d};
    *n = (int64_t)t[which]->K.size();
    return t[which]->K.data();
}
// power_pitzer (occ_ref) tables: which = 0 pp_ia_d, 1 pp_jb_d; virt lists: spin 0 beta, 1 alpha
const double* orc_pp_ptr_d(void* h, int which, int part, int64_t* n) {
    Oracle* o = (Oracle*)h;
    AliasCols& t = which == 0 ? o->eg.ppn.pp_ia_d : o->eg.ppn.pp_jb_d;
    std::vector<double>& v = part == 0 ? t.w : (part == 1 ? t.U : t.tot);
    *n = (int64_t)v.size();
    return v.data();
}
const int* orc_pp_ptr_i(void* h, int which, int64_t* n) {
    Oracle* o = (Oracle*)h;
    AliasCols& t = which == 0 ? o->eg.ppn.pp_ia_d : o->eg.ppn.pp_jb_d;
    *n = (int64_t)t.K.size();
    return t.K.data();
}
const int* orc_pp_virt(void* h, int spin, int* n) {
    Oracle* o = (Oracle*)h;
    std::vector<int>& v = spin ? o->eg.ppn.virt_list_alpha : o->eg.ppn.virt_list_beta;
    *n = (int)v.size();
    return v.data();
}
int orc_pp_stride(void* h) { return ((Oracle*)h)->eg.ppn.pp_ia_d.stride; }
const int* orc_ppn_occ(void* h) { return ((Oracle*)h)->eg.ppn.occ_list.data(); }
// qmc = { quasi_newton = true, quasi_newton_threshold/value/pop_control (negative: defaults) }; call before orc_init
void orc_set_quasi_newton(void* h, int on, double threshold, double value, double pop_control) {
    Oracle* o = (Oracle*)h;
    o->in.quasi_newton = on != 0; o->in.quasi_newton_threshold = threshold; o->in.quasi_newton_value = value;
    o->in.quasi_newton_pop_control = pop_control;
}
// out: ref_fock_sum, threshold, value, pop_control; sp_fock[nbasis+1]
void orc_get_quasi_newton(void* h, double* out, double* sp_fock) {
    Oracle* o = (Oracle*)h;
    out[0] = o->ref_fock_sum; out[1] = o->qn_threshold; out[2] = o->qn_value; out[3] = o->qn_pop_control;
    for (size_t i = 0; i < o->sp_fock.size(); ++i) sp_fock[i] = o->sp_fock[i];
}
void orc_set_pattempt_parallel(void* h, double pp) { ((Oracle*)h)->in.pattempt_parallel = pp; ((Oracle*)h)->eg.pattempt_parallel = pp; }
double orc_get_pattempt_parallel(void* h) { return ((Oracle*)h)->eg.pattempt_parallel; }
void orc_set_pattempt(void* h, double ps, double pd) { ((Oracle*)h)->eg.pattempt_single = ps; ((Oracle*)h)->eg.pattempt_double = pd; }
int orc_get_pattempt_log(void* h, double* out, int n) {
    Oracle* o = (Oracle*)h;
    const int m = (int)o->pattempt_log.size();
    for (int i = 0; i < m && i < n; ++i) out[i] = o->pattempt_log[i];
    return m;
}
int orc_ccmc_get_hash_shift(void* h) { return ((OracleCcmc*)(Oracle*)h)->hash_shift; }
void orc_ccmc_set_hash_shift(void* h, int shift, int move_freq) {
    OracleCcmc* o = (OracleCcmc*)(Oracle*)h;
    o->hash_shift = shift; o->move_freq = move_freq;
}
int64_t orc_spawn_count(void* h, int rank) {
    int64_t n = 0;
    for (auto& b : ((Oracle*)h)->ranks[rank].send) n += (int64_t)b.size();
    return n;
}
// sdata[n][W+2]: string words, population, flag
void orc_get_spawn(void* h, int rank, int64_t* sdata) {
    Oracle* o = (Oracle*)h;
    int W = o->sys.W, E = W + 2;
    int64_t k = 0;
    for (auto& b : o->ranks[rank].send)
        for (auto& e : b) {
            for (int w = 0; w < W; ++w) sdata[k * E + w] = (int64_t)e.f.w[w];
            sdata[k * E + W] = e.pop;
            sdata[k * E + W + 1] = e.flag;
            ++k;
        }
}
// qmc = { shift_harmonic_forcing = ... } on its own (src/qmc.F90:195-212); call after orc_init
void orc_set_harmonic_forcing(void* h, double v) { ((Oracle*)h)->shift_harmonic_forcing = v; }
// wall-Chebyshev propagator: init_chebyshev (call after orc_init), optional harmonic forcing of the shift
int orc_init_chebyshev(void* h, int order, double cshift, double cscale, int skip_gershgorin, double harmonic_forcing,
                       double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    o->init_chebyshev(order, cshift, cscale, skip_gershgorin != 0);
    o->shift_harmonic_forcing = harmonic_forcing;
    out[0] = o->cheb.range_hi;
    for (int i = 0; i < order; ++i) { out[1 + i] = o->cheb.zeroes[i]; out[1 + order + i] = o->cheb.weights[i]; }
    return 0;
    ORC_CATCH(-1)
}
// staged use (the host owns the Chebyshev state and passes the weight of the sub-cycle, as hb200_set_propagator_weight)
int orc_set_propagator_weight(void* h, double w) {
    Oracle* o = (Oracle*)h;
    o->cheb.weights.assign(1, w);
    o->cheb.icheb = 1;
    return 0;
}
int orc_set_chebyshev_step(void* h, int icheb, double shift_for_update, int do_update) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    if (do_update) o->update_chebyshev(shift_for_update);
    o->cheb.icheb = icheb;
    return 0;
    ORC_CATCH(-1)
}
// load balancing: slot populations of one rank, the policy on the summed slots, the proc_map, the redistribution
int orc_slot_pop(void* h, int rank, double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    std::vector<double> sp = o->slot_pop(o->ranks[rank]);
    for (size_t k = 0; k < sp.size(); ++k) out[k] = sp[k];
    return (int)sp.size();
    ORC_CATCH(-1)
}
int orc_do_load_balancing(void* h, double percent) {
    ORC_TRY
    return ((Oracle*)h)->do_load_balancing(percent) ? 1 : 0;
    ORC_CATCH(-1)
}
int orc_get_proc_map(void* h, int* out) {
    Oracle* o = (Oracle*)h;
    for (size_t k = 0; k < o->proc_map.size(); ++k) out[k] = o->proc_map[k];
    return (int)o->proc_map.size();
}
int orc_set_proc_map(void* h, const int* map, int n) {
    Oracle* o = (Oracle*)h;
    if ((size_t)n != o->proc_map.size()) return -1;
    for (int k = 0; k < n; ++k) o->proc_map[k] = map[k];
    return 0;
}
int orc_redistribute(void* h, uint32_t cycle_id) {
    ORC_TRY
    ((Oracle*)h)->redistribute_fciqmc(cycle_id);
    return 0;
    ORC_CATCH(-1)
}
// finish the cycle started by orc_stage_spawn: comm + annihilation on every rank
int orc_stage_annihilate(void* h) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    o->comm_spawn();
    for (auto& r : o->ranks) o->annihilate_rank(r);
    return 0;
    ORC_CATCH(-1)
}

// One excitation-generation + spawn attempt with the Philox stream (kernel parity probe).
// iout: nexcit, from1, from2, to1, to2, perm, allowed ; dout: pgen, hmatel ; nspawn: encoded spawn
int orc_gen_excit_philox(void* h, const uint64_t* f, uint32_t cycle, uint32_t attempt, int64_t parent_pop,
                         double tau, int* iout, double* dout, int64_t* nspawn) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    PhiloxRng rng((uint32_t)o->in.seed);
    rng.set_cycle(cycle);
    DetInfo d;
    Det fd = mkdet(o->sys, f);
    decode_for(o->sys, o->eg, fd, d);
    rng.begin(RNG_SPAWN, fd, o->sys.W, attempt);
    GenResult g = gen_excit_sys(rng, o->sys, o->eg, d);
    double save_tau = o->tau;
    o->tau = tau;
    *nspawn = o->attempt_to_spawn(rng, g.hmatel, g.pgen, parent_pop);
    o->tau = save_tau;
    iout[0] = g.conn.nexcit; iout[1] = g.conn.from_orb[0]; iout[2] = g.conn.from_orb[1];
    iout[3] = g.conn.to_orb[0]; iout[4] = g.conn.to_orb[1]; iout[5] = g.conn.perm; iout[6] = g.allowed;
    dout[0] = g.pgen; dout[1] = g.hmatel;
    return 0;
    ORC_CATCH(-1)
}

// Exact pgen check helper: sample one excitation with an arbitrary caller-provided stream
// (list of doubles in [0,1)); returns number of randoms consumed.
struct ListRng : Rng {
    const double* v; int n; int k = 0;
    ListRng(const double* v_, int n_) : v(v_), n(n_) {}
    void begin(uint32_t, const Det&, int, uint32_t) override {}
    void set_cycle(uint32_t) override {}
    bool continue_past_end = false;      // like the engine's ListStream: an equidistributed continuation, k > n flags it
    double next() override {
        ndraws++;
        if (k < n) return v[k++];
        if (!continue_past_end) throw std::runtime_error("ListRng exhausted");
        const double x = 0.5 + 0.6180339887498949 * (double)(++k);
        return x - std::floor(x);
    }
};
int orc_gen_excit_list(void* h, const uint64_t* f, const double* rn, int nrn, int* iout, double* dout) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    ListRng rng(rn, nrn);
    DetInfo d;
    Det fd = mkdet(o->sys, f);
    decode_for(o->sys, o->eg, fd, d);
    GenResult g = gen_excit_sys(rng, o->sys, o->eg, d);
    iout[0] = g.conn.nexcit; iout[1] = g.conn.from_orb[0]; iout[2] = g.conn.from_orb[1];
    iout[3] = g.conn.to_orb[0]; iout[4] = g.conn.to_orb[1]; iout[5] = g.conn.perm; iout[6] = g.allowed;
    dout[0] = g.pgen; dout[1] = g.hmatel;
    return rng.k;
    ORC_CATCH(-1)
}

// the same followed by attempt_to_spawn on the next number of the list (tau from orc_set_qmc / the last staged cycle)
int orc_gen_excit_spawn_list(void* h, const uint64_t* f, int64_t parent_pop, double tau, const double* rn, int nrn, int* iout,
                             double* dout, int64_t* nspawn) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    ListRng rng(rn, nrn);
    rng.continue_past_end = true;
    DetInfo d;
    Det fd = mkdet(o->sys, f);
    decode_for(o->sys, o->eg, fd, d);
    GenResult g = gen_excit_sys(rng, o->sys, o->eg, d);
    iout[0] = g.conn.nexcit; iout[1] = g.conn.from_orb[0]; iout[2] = g.conn.from_orb[1];
    iout[3] = g.conn.to_orb[0]; iout[4] = g.conn.to_orb[1]; iout[5] = g.conn.perm; iout[6] = g.allowed;
    dout[0] = g.pgen; dout[1] = g.hmatel;
    const double tau_keep = o->tau;
    o->tau = tau;
    *nspawn = o->attempt_to_spawn(rng, g.hmatel, g.pgen, parent_pop);
    o->tau = tau_keep;
    return rng.k;
    ORC_CATCH(-1)
}

// Heat-bath tables (column-major, Fortran order), for comparison with the device builder.
int64_t orc_hb_nb(void* h) { return ((Oracle*)h)->eg.hb.nb; }
const double* orc_hb_ptr_d(void* h, int which) {
    HeatBath& hb = ((Oracle*)h)->eg.hb;
    switch (which) {
        case 0: return hb.i_weights.data();
        case 1: return hb.ij_weights.data();
        case 2: return hb.ija_w.data();
        case 3: return hb.ija_U.data();
        case 4: return hb.ija_tot.data();
        case 5: return hb.ijab_w.data();
        case 6: return hb.ijab_U.data();
        case 7: return hb.ijab_tot.data();
    }
    return nullptr;
}
const int* orc_hb_ptr_i(void* h, int which) {
    HeatBath& hb = ((Oracle*)h)->eg.hb;
    return which == 0 ? hb.ija_K.data() : hb.ijab_K.data();
}

// CPU baseline: run `ncycles` MC cycles on `nthreads` independent replicas of the current state
// (one emulated MPI rank per core, as the reference would be run with mpiexec -np <cores>), each
// with its own stream; returns wall seconds and the attempts/walker-iterations processed.
int orc_cpu_baseline(void* h, int nthreads, int ncycles, double tau, double shift, double proj_energy_old,
                     double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    std::vector<std::unique_ptr<Oracle>> reps(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        reps[t].reset(new Oracle());
        Oracle& r = *reps[t];
        r.sys = o->sys; r.in = o->in; r.eg_shared = &o->eg; r.occ_list0 = o->occ_list0; r.f0 = o->f0; r.H00 = o->H00;
        r.ref_ex_level = o->ref_ex_level; r.pop_real_factor = o->pop_real_factor; r.spawn_cutoff = o->spawn_cutoff;
        r.proc_map = o->proc_map; r.tau = tau; r.shift = shift; r.est.proj_energy_old = proj_energy_old;
        r.ranks.resize(1);
        r.ranks[0].iproc = 0;
        r.ranks[0].send.assign(1, {});
        r.in.nprocs = 1; r.in.nslots = 1; r.proc_map.assign(1, 0);
        r.ranks[0].states = o->ranks[0].states; r.ranks[0].pops = o->ranks[0].pops; r.ranks[0].dat = o->ranks[0].dat;
        r.ranks[0].nstates = o->ranks[0].nstates; r.ranks[0].nparticles = o->ranks[0].nparticles;
        r.ranks[0].rng.reset(new PhiloxRng((uint32_t)(o->in.seed + 1000 * (t + 1))));
    }
    std::vector<double> walker_iters(nthreads, 0.0), attempts(nthreads, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            Oracle& r = *reps[t];
            for (int c = 0; c < ncycles; ++c) {
                walker_iters[t] += r.ranks[0].nparticles;
                r.mc_cycle(1 + c);
                attempts[t] += (double)r.ranks[0].nattempts / 2.0;
            }
        });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    double wi = 0, at = 0;
    for (int t = 0; t < nthreads; ++t) { wi += walker_iters[t]; at += attempts[t]; }
    out[0] = std::chrono::duration<double>(t1 - t0).count();
    out[1] = wi; out[2] = at;
    return 0;
    ORC_CATCH(-1)
}

}  // extern "C"

// ---- per-rank stage API: lets a test drive ONE emulated rank per process (exchange done by the caller),
// used by the world_size-2 gloo test of the host-side driver.
extern "C" {
int orc_rank_spawn(void* h, int rank, uint32_t cycle_id, double tau, double shift, double proj_energy_old, double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    o->tau = tau; o->shift = shift; o->est.proj_energy_old = proj_energy_old;
    RankState& r = o->ranks[rank];
    r.proj_energy = 0.0; r.D0_population = 0.0;
    o->spawn_death_rank(r, cycle_id);
    // one rank per process: with nprocs > 1 the deterministic projection needs every rank's amplitudes - the caller
    // gathers them and calls orc_rank_determ_project before orc_rank_annihilate
    if (o->determ.doing && o->in.nprocs == 1) o->determ_proj_separate();
    out[0] = r.proj_energy; out[1] = r.D0_population; out[2] = (double)r.nspawn_events; out[3] = (double)r.ndeath;
    out[4] = (double)r.nattempts;
    return 0;
    ORC_CATCH(-1)
}
int64_t orc_rank_send_count(void* h, int rank, int dest) { return (int64_t)((Oracle*)h)->ranks[rank].send[dest].size(); }
void orc_rank_get_send(void* h, int rank, int dest, int64_t* sdata) {
    Oracle* o = (Oracle*)h;
    int W = o->sys.W, E = W + 2;
    int64_t k = 0;
    for (auto& e : o->ranks[rank].send[dest]) {
        for (int w = 0; w < W; ++w) sdata[k * E + w] = (int64_t)e.f.w[w];
        sdata[k * E + W] = e.pop; sdata[k * E + W + 1] = e.flag;
        ++k;
    }
}
int orc_rank_annihilate(void* h, int rank, const int64_t* sdata, int64_t n, double* out) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    RankState& r = o->ranks[rank];
    int W = o->sys.W, E = W + 2;
    r.recv.clear();
    for (int64_t k = 0; k < n; ++k) {
        SpawnElem e;
        for (int w = 0; w < W; ++w) e.f.w[w] = (uint64_t)sdata[k * E + w];
        e.pop = sdata[k * E + W]; e.flag = sdata[k * E + W + 1];
        r.recv.push_back(e);
    }
    o->annihilate_rank(r);
    out[0] = r.nparticles; out[1] = (double)r.nstates; out[2] = (r.spawn_error || r.psip_error) ? 1.0 : 0.0;
    return 0;
    ORC_CATCH(-1)
}

// semi_stoch = { space = "high" | "ci", size, start_iteration, shift_start_iteration, ci_space = { ex_level },
// separate_annihilation }; pop_real_bits: 31, or 11 for real_amplitude_force_32.  Call before orc_init.
void orc_set_semi_stoch(void* h, int space, int target_size, int start_iter, int shift_iter, int ci_ex_level, int separate,
                        int pop_real_bits) {
    Oracle* o = (Oracle*)h;
    o->in.ss_space = space; o->in.ss_target_size = target_size; o->in.ss_start_iter = start_iter;
    o->in.ss_shift_iter = shift_iter; o->in.ss_ci_ex_level = ci_ex_level; o->in.ss_separate_annihilation = separate != 0;
    o->in.pop_real_bits = pop_real_bits;
}
void orc_set_vary_shift(void* h, int on) { ((Oracle*)h)->vary_shift = on != 0; }
// init_semi_stoch_t on the current lists of all emulated ranks (what do_fciqmc does at iteration semi_stoch_iter)
int orc_init_semi_stoch(void* h) {
    ORC_TRY
    ((Oracle*)h)->init_semi_stoch();
    return 0;
    ORC_CATCH(-1)
}
// the same with the space given by the caller: dets = every deterministic state, rank by rank (sizes[nprocs])
int orc_init_semi_stoch_dets(void* h, const uint64_t* dets, const int* sizes) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    // reuse init_semi_stoch through a "ci"-like hook: temporarily mark the given states as the most populated
    const int W = o->sys.W;
    std::vector<std::vector<Det>> mine((size_t)o->in.nprocs);
    size_t k = 0;
    for (int r = 0; r < o->in.nprocs; ++r)
        for (int i = 0; i < sizes[r]; ++i, ++k) {
            Det f;
            for (int w = 0; w < W; ++w) f.w[w] = dets[k * W + w];
            mine[(size_t)r].push_back(f);
        }
    o->given_determ = mine;
    const int keep = o->in.ss_space;
    o->in.ss_space = 3;
    o->init_semi_stoch();
    o->in.ss_space = keep;
    return 0;
    ORC_CATCH(-1)
}
int orc_determ_sizes(void* h, int* sizes) {
    Oracle* o = (Oracle*)h;
    for (size_t r = 0; r < o->determ.sizes.size(); ++r) sizes[r] = o->determ.sizes[r];
    return o->determ.tot_size;
}
void orc_get_determ_dets(void* h, uint64_t* out) {
    Oracle* o = (Oracle*)h;
    const int W = o->sys.W;
    for (size_t i = 0; i < o->determ.dets.size(); ++i)
        for (int w = 0; w < W; ++w) out[i * W + w] = o->determ.dets[i].w[w];
}
// determ%hamil of one rank: returns nnz; row_ptr[tot_size+1] (0-based), col_ind[nnz] (0-based), mat[nnz]
int64_t orc_get_determ_hamil(void* h, int rank, int* row_ptr, int* col_ind, double* mat) {
    Oracle* o = (Oracle*)h;
    RankState& R = o->ranks[rank];
    if (row_ptr) {
        std::copy(R.drow_ptr.begin(), R.drow_ptr.end(), row_ptr);
        std::copy(R.dcol_ind.begin(), R.dcol_ind.end(), col_ind);
        std::copy(R.dmat.begin(), R.dmat.end(), mat);
    }
    return (int64_t)R.dmat.size();
}
// determ%vector of one rank (after a cycle: -tau (H - S) v restricted to the rank) and determ%flags (0 = deterministic)
void orc_get_determ_vector(void* h, int rank, double* vec, uint8_t* flags) {
    Oracle* o = (Oracle*)h;
    RankState& R = o->ranks[rank];
    if (vec) std::copy(R.dvector.begin(), R.dvector.end(), vec);
    if (flags) std::copy(R.dflag.begin(), R.dflag.end(), flags);
}
// one rank per process (tests/oracle_engine.py): this rank's determ%vector as set_determ_info left it, and the
// projection of this rank given determ%full_vector gathered by the caller (tot_size doubles, rank by rank)
void orc_rank_get_dvector(void* h, int rank, double* out) {
    Oracle* o = (Oracle*)h;
    std::copy(o->ranks[rank].dvector.begin(), o->ranks[rank].dvector.end(), out);
}
int orc_rank_determ_project(void* h, int rank, const double* full) {
    ORC_TRY
    Oracle* o = (Oracle*)h;
    for (size_t i = 0; i < o->determ.full_vector.size(); ++i) o->determ.full_vector[i] = full[i];
    o->determ_proj_rank(o->ranks[rank]);
    return 0;
    ORC_CATCH(-1)
}
}
