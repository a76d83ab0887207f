// TEST HARNESS (not product code): compiles the __host__ __device__ core of the CUDA engine
// (hande_b200/csrc/hb_core.cuh) with g++ so that the per-attempt logic - excitation generators,
// Slater-Condon rules, Philox stream, owner hash, death/spawn arithmetic - can be compared with the
// oracle on a machine without a GPU.  The product library never uses these host instantiations.
#include <type_traits>
#include <cstring>
#include <vector>
#include "../../hande_b200/csrc/hb_core.cuh"

using namespace hb;

static Sys g_sys;
static Params g_par;
static std::vector<uint8_t> g_sym;
static std::vector<int8_t> g_ms;
static std::vector<uint16_t> g_spatial;
static std::vector<double> g_J, g_K;
static std::vector<D2> g_CX;

template <int W, class R>
static void gen_one(R& rng, const uint64_t* f, int64_t parent_pop, int* iout, double* dout, int64_t* nspawn) {
    uint8_t occ[HB_MAXNEL], su[64];
    decode_det<W>(f, occ);
    if (g_sys.kind != SYS_UEG) {
        uint8_t su2[64];
        build_symunocc(g_sys, occ, su);
        build_symunocc_masks<W>(g_sys, f, su2);
        for (int c = 0; c < 2 * g_sys.nsym_tot; ++c) if (su[c] != su2[c]) { iout[6] = -99; return; }   // mask-based == loop-based
    }
    Gen g;
    gen_excit<W>(rng, g_sys, g_par, f, occ, su, g);
    *nspawn = attempt_to_spawn(rng, g_par, g.hmatel, g.pgen, parent_pop);
    iout[0] = g.nexcit; iout[1] = g.from1; iout[2] = g.from2; iout[3] = g.to1; iout[4] = g.to2; iout[5] = g.perm;
    iout[6] = g.allowed;
    dout[0] = g.pgen; dout[1] = g.hmatel;
}

extern "C" {

void hd_set_sys(int nbasis, int nel, int nsym_tot, int sym0, int sym_max, int pg_mask, int Lz_mask, int Lz_offset,
                int gamma_sym, int uhf, int nvirt, int nvirt_alpha, int nvirt_beta, int max_nbss, double Ecore,
                const int* sym, const int* ms, const int* spatial, const int* nbss, const int* ssbf, const double* h1,
                const double* const* v2) {
    Sys& s = g_sys;
    memset(&s, 0, sizeof(s));
    s.nbasis = nbasis; s.nel = nel; s.W = (nbasis + 63) / 64;
    s.nsym_tot = nsym_tot; s.sym0 = sym0; s.sym_max = sym_max; s.pg_mask = pg_mask; s.Lz_mask = Lz_mask;
    s.Lz_offset = Lz_offset; s.gamma_sym = gamma_sym; s.uhf = uhf; s.nvirt = nvirt; s.nvirt_alpha = nvirt_alpha;
    s.nvirt_beta = nvirt_beta; s.max_nbss = max_nbss; s.Ecore = Ecore;
    g_sym.assign(nbasis + 1, 0); g_ms.assign(nbasis + 1, 0); g_spatial.assign(nbasis + 1, 0);
    for (int i = 1; i <= nbasis; ++i) { g_sym[i] = (uint8_t)sym[i]; g_ms[i] = (int8_t)ms[i]; g_spatial[i] = (uint16_t)spatial[i]; }
    s.bf_sym = g_sym.data(); s.bf_ms = g_ms.data(); s.bf_spatial = g_spatial.data();
    s.nbss = nbss; s.ssbf = ssbf; s.h1 = h1;
    for (int c = 0; c < (uhf ? 4 : 1); ++c) s.v2[c] = v2[c];
    g_J.assign((size_t)nbasis * nbasis, 0.0); g_K.assign((size_t)nbasis * nbasis, 0.0);
    for (int i = 1; i <= nbasis; ++i)
        for (int j = 1; j <= nbasis; ++j) {
            g_J[(size_t)(i - 1) * nbasis + (j - 1)] = two_body(s, i, j, i, j);
            g_K[(size_t)(i - 1) * nbasis + (j - 1)] = two_body(s, i, j, j, i);
        }
    s.Jd = g_J.data(); s.Kd = g_K.data();
    static std::vector<uint64_t> g_sumask;
    g_sumask.assign((size_t)2 * nsym_tot * s.W, 0);
    for (int i = 1; i <= nbasis; ++i) {
        const int c = (ms[i] > 0 ? 1 : 0) + 2 * sym[i];
        g_sumask[(size_t)c * s.W + ((i - 1) >> 6)] |= 1ull << ((i - 1) & 63);
    }
    s.su_mask = g_sumask.data();
    const int NT = uhf ? nbasis : nbasis / 2;
    s.NT = NT;
    g_CX.assign((size_t)NT * NT * NT, D2{0.0, 0.0});
    for (int ti = 0; ti < NT; ++ti)
        for (int ta = 0; ta < NT; ++ta)
            for (int tj = 0; tj < NT; ++tj) {
                const int i = uhf ? ti + 1 : 2 * ti + 1, a = uhf ? ta + 1 : 2 * ta + 1, j = uhf ? tj + 1 : 2 * tj + 1;
                g_CX[((size_t)ti * NT + ta) * NT + tj] = D2{two_body(s, i, j, a, j), two_body(s, i, j, j, a)};
            }
    s.sc1CX = g_CX.data();
}
// uniform electron gas (hb200_set_system_ueg on the host side of the engine)
static std::vector<K4> g_kv;
void hd_set_sys_ueg(int nbasis, int nel, double box_length, const int* kvec, const double* sp_eigv, int kmax, int offset,
                    const int* offset_inds, const int* lookup, int tern_kmax, const uint64_t* tern) {
    Sys& s = g_sys;
    memset(&s, 0, sizeof(s));
    s.kind = SYS_UEG;
    s.nbasis = nbasis; s.nel = nel; s.W = (nbasis + 63) / 64; s.nsym_tot = 1;
    g_sym.assign(nbasis + 1, 0); g_ms.assign(nbasis + 1, 0); g_spatial.assign(nbasis + 1, 0); g_kv.assign(nbasis + 1, K4{0, 0, 0, 0});
    for (int i = 1; i <= nbasis; ++i) {
        g_ms[i] = (i & 1) ? 1 : -1; g_spatial[i] = (uint16_t)((i + 1) / 2);
        g_kv[i] = K4{kvec[3 * i], kvec[3 * i + 1], kvec[3 * i + 2], 0};
    }
    s.bf_sym = g_sym.data(); s.bf_ms = g_ms.data(); s.bf_spatial = g_spatial.data();
    s.ueg_k = g_kv.data(); s.sp_eigv = sp_eigv; s.ueg_lookup = lookup; s.ueg_tern = tern;
    s.ueg_piL = 3.1415926535897931 * box_length;
    s.ueg_kmax = kmax; s.ueg_offset = offset;
    for (int d = 0; d < 3; ++d) s.ueg_oi[d] = offset_inds[d];
    s.ueg_tK = tern_kmax; s.ueg_tD = 2 * tern_kmax + 1;
}
void hd_set_heat_bath(const double* i_w, const double* ij_w, const double* ija_w, const double* ija_U, const int* ija_K,
                      const double* ija_tot, const double* ijab_w, const double* ijab_U, const int* ijab_K,
                      const double* ijab_tot) {
    Sys& s = g_sys;
    s.hb_i_w = i_w; s.hb_ij_w = ij_w; s.hb_ija_w = ija_w; s.hb_ija_U = ija_U; s.hb_ija_K = ija_K; s.hb_ija_tot = ija_tot;
    s.hb_ijab_w = ijab_w; s.hb_ijab_U = ijab_U; s.hb_ijab_K = ijab_K; s.hb_ijab_tot = ijab_tot;
}
void hd_set_params(int excit_gen, double ps, double pd, double tau, double shift, double pe_old, int64_t real_factor,
                   int64_t spawn_cutoff, uint32_t seed, const uint64_t* f0, double H00) {
    Params& p = g_par;
    memset(&p, 0, sizeof(p));
    p.excit_gen = excit_gen; p.pattempt_single = ps; p.pattempt_double = pd; p.tau = tau; p.shift = shift;
    p.proj_energy_old = pe_old; p.real_factor = real_factor; p.spawn_cutoff = spawn_cutoff; p.seed = seed;
    p.hash_seed = 7; p.nprocs = 1; p.nslots = 1; p.H00 = H00; p.trunc_level = -1;
    for (int k = 0; k < g_sys.W; ++k) p.f0[k] = f0[k];
}

void hd_set_ppn(int which, const double* w, const double* U, const int* K, const double* tot) {
    g_sys.ppn[which].w = w; g_sys.ppn[which].U = U; g_sys.ppn[which].K = K; g_sys.ppn[which].tot = tot;
}
void hd_set_ppn_occ(const int* occ) { g_sys.ppn_occ = occ; }
void hd_set_pp(int which, const double* w, const double* U, const int* K, const double* tot) {
    Sys::AliasTab& t = which == 0 ? g_sys.pp_ia : g_sys.pp_jb;
    t.w = w; t.U = U; t.K = K; t.tot = tot;
}
void hd_set_pp_virt(const int* vb, int nb, const int* va, int na, int stride) {
    g_sys.pp_virt[0] = vb; g_sys.pp_virt[1] = va; g_sys.pp_nvirt[0] = nb; g_sys.pp_nvirt[1] = na; g_sys.pp_sia = stride;
}
void hd_set_pattempt_parallel(double pp) { g_par.pattempt_parallel = pp; }

void hd_gen_excit_philox(const uint64_t* f, uint32_t cycle, uint32_t attempt, int64_t parent_pop, int* iout,
                         double* dout, int64_t* nspawn) {
    PhiloxStream rng;
    switch (g_sys.W) {
        case 1: rng.begin(g_par.seed, cycle, RNG_SPAWN, det_hash64<1>(f), attempt); gen_one<1>(rng, f, parent_pop, iout, dout, nspawn); break;
        case 2: rng.begin(g_par.seed, cycle, RNG_SPAWN, det_hash64<2>(f), attempt); gen_one<2>(rng, f, parent_pop, iout, dout, nspawn); break;
        case 3: rng.begin(g_par.seed, cycle, RNG_SPAWN, det_hash64<3>(f), attempt); gen_one<3>(rng, f, parent_pop, iout, dout, nspawn); break;
        default: rng.begin(g_par.seed, cycle, RNG_SPAWN, det_hash64<4>(f), attempt); gen_one<4>(rng, f, parent_pop, iout, dout, nspawn); break;
    }
}
int hd_gen_excit_list(const uint64_t* f, const double* rn, int nrn, int* iout, double* dout) {
    ListStream rng{rn, nrn, 0};
    int64_t ns;
    switch (g_sys.W) {
        case 1: gen_one<1>(rng, f, 1, iout, dout, &ns); break;
        case 2: gen_one<2>(rng, f, 1, iout, dout, &ns); break;
        case 3: gen_one<3>(rng, f, 1, iout, dout, &ns); break;
        default: gen_one<4>(rng, f, 1, iout, dout, &ns); break;
    }
    return rng.k - 1;  // attempt_to_spawn consumed one more
}
double hd_sc0(const uint64_t* f) {
    uint8_t occ[HB_MAXNEL];
    switch (g_sys.W) { case 1: decode_det<1>(f, occ); break; case 2: decode_det<2>(f, occ); break;
                       case 3: decode_det<3>(f, occ); break; default: decode_det<4>(f, occ); }
    return g_sys.kind == SYS_UEG ? slater_condon0_ueg(g_sys, occ) : slater_condon0(g_sys, occ);
}
double hd_sc1(const uint64_t* f, int i, int a) {
    uint8_t occ[HB_MAXNEL];
    bool pm;
    switch (g_sys.W) { case 1: decode_det<1>(f, occ); pm = excit_perm1<1>(f, i, a); break;
                       case 2: decode_det<2>(f, occ); pm = excit_perm1<2>(f, i, a); break;
                       case 3: decode_det<3>(f, occ); pm = excit_perm1<3>(f, i, a); break;
                       default: decode_det<4>(f, occ); pm = excit_perm1<4>(f, i, a); }
    return slater_condon1_excit(g_sys, occ, i, a, pm);
}
double hd_sc2(const uint64_t* f, int i, int j, int a, int b) {
    bool pm;
    switch (g_sys.W) { case 1: pm = excit_perm2<1>(f, i, j, a, b); break; case 2: pm = excit_perm2<2>(f, i, j, a, b); break;
                       case 3: pm = excit_perm2<3>(f, i, j, a, b); break; default: pm = excit_perm2<4>(f, i, j, a, b); }
    return slater_condon2_excit(g_sys, i, j, a, b, pm);
}
int32_t hd_murmur(const uint64_t* f, uint32_t seed) { return (int32_t)murmur2_words(f, g_sys.nbasis, seed); }
int hd_owner(const uint64_t* f, int nprocs, int nslots) { return owner_slot(f, g_sys.nbasis, 7, nprocs, nslots); }
void hd_philox_stream(uint32_t seed, uint32_t cycle, uint32_t purpose, const uint64_t* f, int W, uint32_t attempt, int n,
                      double* out) {
    PhiloxStream r;
    uint64_t h = (W == 1) ? det_hash64<1>(f) : (W == 2) ? det_hash64<2>(f) : (W == 3) ? det_hash64<3>(f) : det_hash64<4>(f);
    r.begin(seed, cycle, purpose, h, attempt);
    for (int i = 0; i < n; ++i) out[i] = r.next();
}
// proj-energy contribution and death for one determinant (Philox stream)
double hd_proj_hmatel(const uint64_t* f, int* is_ref) {
    uint8_t occ[HB_MAXNEL];
    bool r; double h;
    switch (g_sys.W) { case 1: decode_det<1>(f, occ); h = proj_energy_hmatel<1>(g_sys, g_par, f, occ, r); break;
                       case 2: decode_det<2>(f, occ); h = proj_energy_hmatel<2>(g_sys, g_par, f, occ, r); break;
                       case 3: decode_det<3>(f, occ); h = proj_energy_hmatel<3>(g_sys, g_par, f, occ, r); break;
                       default: decode_det<4>(f, occ); h = proj_energy_hmatel<4>(g_sys, g_par, f, occ, r); }
    *is_ref = r;
    return h;
}
// <f1|H|f2> - H00 delta as the deterministic Hamiltonian of the semi-stochastic projection uses it (hb_semistoch.cuh
// ss_hmatel: get_hmatel of src/semi_stoch.F90:620): popcount reject, diagonal from slater_condon0, else hmatel_pair
double hd_hmatel_pair(const uint64_t* f1, const uint64_t* f2) {
    uint8_t occ[HB_MAXNEL];
    bool same;
    auto run = [&](auto wtag) {
        constexpr int W = decltype(wtag)::value;
        int nx = 0;
        for (int k = 0; k < W; ++k) nx += popc64(f1[k] ^ f2[k]);
        if (nx > 4) return 0.0;
        decode_det<W>(f1, occ);
        if (nx == 0) return ((g_sys.kind == SYS_UEG) ? slater_condon0_ueg(g_sys, occ) : slater_condon0(g_sys, occ)) - g_par.H00;
        return hmatel_pair<W>(g_sys, f1, occ, f2, same);
    };
    switch (g_sys.W) {
        case 1: return run(std::integral_constant<int, 1>());
        case 2: return run(std::integral_constant<int, 2>());
        case 3: return run(std::integral_constant<int, 3>());
        default: return run(std::integral_constant<int, 4>());
    }
}
// check_if_determ by bisection of the sorted space (hb_core.cuh ss_check_if_determ)
int hd_check_if_determ(const uint64_t* sorted, int n, const uint64_t* f) {
    switch (g_sys.W) {
        case 1: return ss_check_if_determ<1>(sorted, n, f);
        case 2: return ss_check_if_determ<2>(sorted, n, f);
        case 3: return ss_check_if_determ<3>(sorted, n, f);
        default: return ss_check_if_determ<4>(sorted, n, f);
    }
}
// Brute-force check of the guarded prefix-sum alias selection against the reference's table construction
// (generate_alias_tables + select_weighted_value_precalc).  Returns the number of mismatches; out[0] = calls,
// out[1] = fallbacks to the exact walk.
long long hd_alias_selftest_mode(long long ntrials, unsigned long long seed, long long* out, int only_mode);
long long hd_alias_selftest(long long ntrials, unsigned long long seed, long long* out) { return hd_alias_selftest_mode(ntrials, seed, out, -1); }
long long hd_alias_selftest_mode(long long ntrials, unsigned long long seed, long long* out, int only_mode) {
    unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 12345;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
    long long bad = 0, nfallback = 0;
    for (long long t = 0; t < ntrials; ++t) {
        const int N = 2 + (int)(rnd() * 39);
        double w[HB_MAXNEL], U[HB_MAXNEL];
        int K[HB_MAXNEL], un[HB_MAXNEL], ov[HB_MAXNEL];
        const bool aimed = only_mode < 100;
        const int mode = only_mode >= 100 ? only_mode - 100 : (only_mode >= 0 ? only_mode : (int)(rnd() * 7));
        double tot = 0.0;
        for (int q = 0; q < N; ++q) {
            double v;
            switch (mode) {
                case 0: v = rnd(); break;                                   // uniform
                case 1: v = -log(1.0 - rnd()); break;                        // exponential
                case 2: v = (rnd() < 0.3) ? 0.0 : rnd(); break;              // many zeros
                case 3: v = 1.0 + 1e-10 * (rnd() - 0.5); break;              // nearly equal
                case 4: v = 1.0 + 4e-16 * (double)((int)(rnd() * 5) - 2); break;  // equal up to a few ulps
                case 5: v = (q == 0) ? 50.0 * rnd() : rnd() * rnd() * rnd(); break;  // one dominant entry
                default: v = (double)(1 + (int)(rnd() * 4)); break;         // small integers (exact ties)
            }
            w[q] = v; 
        }
        for (int q = 0; q < N; ++q) tot = tot + w[q];
        if (!(tot > 0.0)) continue;
        generate_alias_tables(N, w, tot, U, K, un, ov);
        for (int rep = 0; rep < 8; ++rep) {
            double r = rnd();
            if (rep == 7 && aimed) {   // aim at a table boundary to exercise the guard on x
                const int kk = (int)(rnd() * N);
                r = ((double)kk + fmin(fmax(U[kk], 0.0), 0.999999) * (1.0 + 1e-12 * (rnd() - 0.5))) / N;
                if (!(r >= 0.0 && r < 1.0)) r = rnd();
            }
            ListStream a{&r, 1, 0}, b{&r, 1, 0};
            const int ref = select_precalc(a, N, U, K);
            const int got = select_alias_staged(b, N, w, 1, tot);
            if (ref != got) bad++;
            // the fixed-trip-count scan (clobbers its copy of the weights; 0 = "run the exact walk")
            {
                double wc[HB_MAXNEL];
                for (int q = 0; q < N; ++q) wc[q] = w[q];
                double xx = r * N;
                const int kk = (int)xx;
                xx = xx - kk;
                int got2 = (N <= 32) ? alias_select_scan<uint32_t>(N, wc, 1, N / tot, kk, xx) : alias_select_scan<uint64_t>(N, wc, 1, N / tot, kk, xx);
                if (got2 == 0) { nfallback++; got2 = (N <= 32) ? alias_walk_exact<uint32_t>(N, w, 1, N / tot, kk, xx) : alias_walk_exact<uint64_t>(N, w, 1, N / tot, kk, xx); }
                if (ref != got2) bad++;
            }
        }
    }
    out[0] = 8 * ntrials; out[1] = nfallback;
    return bad;
}
}  // extern "C"
