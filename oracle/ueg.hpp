// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the reference's uniform-electron-gas path (3D, no twist):
//   basis            src/basis.f90:258-501 (init_model_basis_fns), src/kpoints.f90:9-53 (calc_kinetic),
//                    src/system.f90:486-509 (box length), lib/local/ranking.f90 (stable rank, tolerance depsilon)
//   indexing         src/ueg.f90:43-85 (init_ueg_indexing), :142-172 (ueg_basis_index)
//   ternary_conserve src/ueg.f90:87-140
//   integrals        src/ueg.f90:176-280 (get_two_e_int_ueg, coulomb_int_ueg_3d)
//   Slater-Condon    src/hamiltonian_ueg.f90:71-223, src/determinants.f90:403-425
//   generator        src/excit_gen_ueg.f90:27-360 (gen_excit_ueg_no_renorm, choose_ij_k, find_ab_ueg,
//                    calc_pgen_ueg_no_renorm)
// Random numbers are consumed in the reference's order: one for (i,j), one for a (only if max_na > 0).
#pragma once
#include "system.hpp"
#include "rng.hpp"
#include "excit_gen.hpp"

namespace oracle {

// src/system.f90:486-509 + src/basis.f90:258-501 + src/ueg.f90:43-140
inline void init_ueg_system(System& sys, int nel, int ms, double rs, double ecutoff) {
    const double pi = 3.1415926535897931;  // lib/local/const.F90
    sys.kind = SYS_UEG;
    sys.nel = nel; sys.Ms = ms;
    sys.nalpha = (nel + ms) / 2; sys.nbeta = (nel - ms) / 2;
    sys.uhf = false;
    UegData& u = sys.ueg;
    u.rs = rs; u.ecutoff = ecutoff;
    u.L = rs * std::pow((4 * pi * nel) / 3, 1.0 / 3.0);
    const double rl = 1.0 / u.L;
    const int nmax = (int)std::ceil(std::sqrt(2 * ecutoff));
    struct Tmp { int l[3]; double e; };
    std::vector<Tmp> tmp;
    for (int k = -nmax; k <= nmax; ++k)
        for (int j = -nmax; j <= nmax; ++j)
            for (int i = -nmax; i <= nmax; ++i) {
                if ((double)(i * i + j * j + k * k) / 2 > ecutoff) continue;
                Tmp t;
                t.l[0] = i; t.l[1] = j; t.l[2] = k;
                // calc_kinetic: kc(d) = sum((k+ktwist)*rlattice(d,:)); kinetic = 2*pi**2*dot_product(kc,kc)
                double kc[3];
                for (int d = 0; d < 3; ++d) {
                    double s = 0.0;
                    for (int e = 0; e < 3; ++e) s += (t.l[e] + 0.0) * (e == d ? rl : 0.0);
                    kc[d] = s;
                }
                double dot = 0.0;
                for (int d = 0; d < 3; ++d) dot += kc[d] * kc[d];
                t.e = 2 * pi * pi * dot;
                tmp.push_back(t);
            }
    const int nsp = (int)tmp.size();
    std::vector<double> eig(nsp + 1, 0.0);
    for (int i = 1; i <= nsp; ++i) eig[i] = tmp[i - 1].e;
    std::vector<int> rank;
    insertion_rank(eig, nsp, rank, depsilon);
    sys.nbasis = 2 * nsp;
    sys.nvirt = sys.nbasis - nel;
    sys.nvirt_alpha = nsp - sys.nalpha; sys.nvirt_beta = nsp - sys.nbeta;
    sys.W = (sys.nbasis + 63) / 64;
    if (sys.W > MAXW) throw std::runtime_error("ueg: basis too large for the oracle's MAXW");
    sys.bf.assign(sys.nbasis + 1, BasisFn());
    u.l.assign((size_t)3 * (sys.nbasis + 1), 0);
    for (int i = 1; i <= nsp; ++i) {
        const Tmp& t = tmp[rank[i] - 1];
        for (int s = 0; s < 2; ++s) {
            const int o = 2 * i - 1 + s;
            sys.bf[o].ms = s == 0 ? 1 : -1;
            sys.bf[o].sp_eigv = t.e;
            sys.bf[o].spatial_index = i;
            for (int d = 0; d < 3; ++d) u.l[3 * o + d] = t.l[d];
        }
    }
    // trivial point-group data so that the generic (symmetry-blind) parts of the loop work
    sys.pg_mask = 0; sys.Lz_mask = 0; sys.Lz_offset = 0; sys.gamma_sym = 0;
    sys.sym0 = 0; sys.sym_max = 0; sys.nsym = 1; sys.nsym_tot = 1; sys.sym_max_tot = 0;
    sys.nbasis_sym_spin.assign(2, nsp);
    sys.max_nbss = nsp;
    sys.sym_spin_basis_fns.assign((size_t)2 * nsp, 0);
    for (int i = 1; i <= nsp; ++i) { sys.sym_spin_basis_fns[(i - 1) + nsp * 0] = 2 * i; sys.sym_spin_basis_fns[(i - 1) + nsp * 1] = 2 * i - 1; }
    // init_ueg_indexing (src/ueg.f90:43-85)
    u.kmax = (int)std::ceil(std::sqrt(2 * ecutoff));
    const int Nk = 2 * u.kmax + 1;
    u.offset_inds[0] = 1; u.offset_inds[1] = Nk; u.offset_inds[2] = Nk * Nk;
    u.offset = (u.offset_inds[0] + u.offset_inds[1] + u.offset_inds[2]) * u.kmax + 1;
    u.lookup.assign((size_t)Nk * Nk * Nk + 1, -1);
    for (int i = 1; i <= sys.nbasis; i += 2) {
        int idx = u.offset;
        for (int d = 0; d < 3; ++d) idx += u.l[3 * i + d] * u.offset_inds[d];
        u.lookup[idx] = i;
    }
    // init_ternary_conserve (src/ueg.f90:87-140)
    const int K = 2 * u.kmax, D = 2 * K + 1;
    u.tK = K; u.tD = D;
    u.ternary.assign((size_t)(sys.W + 1) * D * D * D, 0);
    for (int k3 = -K; k3 <= K; ++k3)
        for (int k2 = -K; k2 <= K; ++k2)
            for (int k1 = -K; k1 <= K; ++k1) {
                uint64_t* t = &u.ternary[(size_t)(sys.W + 1) * ((k1 + K) + (size_t)D * ((k2 + K) + (size_t)D * (k3 + K)))];
                const int kt[3] = {k1, k2, k3};
                for (int a = 1; a <= sys.nbasis - 1; a += 2) {
                    int dot = 0;
                    for (int d = 0; d < 3; ++d) { const int x = kt[d] - u.l[3 * a + d]; dot += x * x; }
                    if ((double)dot / 2 - ecutoff < 1.e-8) {
                        t[0] += 1;
                        t[1 + ((a - 1) >> 6)] |= (1ull << ((a - 1) & 63));
                    }
                }
            }
}

// src/ueg.f90:142-172
inline int ueg_basis_index(const System& sys, const int* k, int spin) {
    const UegData& u = sys.ueg;
    int mn = std::min(k[0], std::min(k[1], k[2])), mx = std::max(k[0], std::max(k[1], k[2]));
    if (mn < -u.kmax || mx > u.kmax) return -1;
    int indx = u.lookup[k[0] * u.offset_inds[0] + k[1] * u.offset_inds[1] + k[2] * u.offset_inds[2] + u.offset];
    if (spin < 0) indx = indx + 1;   // (as in the reference: -1 + 1 = 0 for a wavevector outside the basis)
    return indx;
}

// src/ueg.f90:250-280
inline double coulomb_int_ueg_3d(const System& sys, int i, int a) {
    const double pi = 3.1415926535897931;
    const UegData& u = sys.ueg;
    int qq = 0;
    for (int d = 0; d < 3; ++d) { const int q = u.l[3 * i + d] - u.l[3 * a + d]; qq += q * q; }
    return 1.0 / (pi * u.L * qq);
}

// src/ueg.f90:176-215
inline double get_two_e_int_ueg(const System& sys, int i, int j, int a, int b) {
    const UegData& u = sys.ueg;
    double intgrl = 0.0;
    bool cons = true;
    for (int d = 0; d < 3; ++d)
        if (u.l[3 * i + d] + u.l[3 * j + d] - u.l[3 * a + d] - u.l[3 * b + d] != 0) cons = false;
    if (cons) {
        if (sys.bf[i].ms == sys.bf[a].ms && sys.bf[j].ms == sys.bf[b].ms) intgrl = intgrl + coulomb_int_ueg_3d(sys, i, a);
        if (sys.bf[i].ms == sys.bf[b].ms && sys.bf[j].ms == sys.bf[a].ms) intgrl = intgrl - coulomb_int_ueg_3d(sys, i, b);
    }
    return intgrl;
}

// src/hamiltonian_ueg.f90:71-99,127-156; src/determinants.f90:403-425
inline double slater_condon0_ueg_orb_list(const System& sys, const int* occ) {
    double spe = 0.0;
    for (int i = 0; i < sys.nel; ++i) spe = spe + sys.bf[occ[i]].sp_eigv;
    double ex = 0.0;
    for (int i = 0; i < sys.nel; ++i)
        for (int j = i + 1; j < sys.nel; ++j)
            if ((occ[i] % 2) == (occ[j] % 2)) ex = ex - coulomb_int_ueg_3d(sys, occ[i], occ[j]);
    return spe + ex;
}
// src/hamiltonian_ueg.f90:158-184
inline double slater_condon2_ueg(const System& sys, int i, int j, int a, int b, bool perm) {
    double h = get_two_e_int_ueg(sys, i, j, a, b);
    return perm ? -h : h;
}
// src/hamiltonian_ueg.f90:186-223
inline double slater_condon2_ueg_excit(const System& sys, int i, int a, int b, bool perm) {
    double h = 0.0;
    if (sys.bf[i].ms == sys.bf[a].ms) h = coulomb_int_ueg_3d(sys, i, a);
    if (sys.bf[i].ms == sys.bf[b].ms) h = h - coulomb_int_ueg_3d(sys, i, b);
    return perm ? -h : h;
}

// gen_excit_ueg_no_renorm (src/excit_gen_ueg.f90:27-101) with choose_ij_k (:105-190), find_ab_ueg (:194-306),
// calc_pgen_ueg_no_renorm (:310-360)
inline GenResult gen_excit_ueg_no_renorm(Rng& rng, const System& sys, const DetInfo& d) {
    const UegData& u = sys.ueg;
    GenResult g;
    g.conn.nexcit = 2;
    const int nel = sys.nel;
    // choose_ij_k
    double r = rng.next();
    int ind = (int)(r * nel * (nel - 1) / 2) + 1;
    int jj = (int)(1.50 + std::sqrt(2 * ind - 1.750));
    int ii = ind - ((jj - 1) * (jj - 2)) / 2;
    const int i = d.occ[ii - 1], j = d.occ[jj - 1];
    g.conn.from_orb[0] = i; g.conn.from_orb[1] = j;
    const int ij_spin = sys.bf[i].ms + sys.bf[j].ms;
    int ij_k[3];
    for (int dd = 0; dd < 3; ++dd) ij_k[dd] = u.l[3 * i + dd] + u.l[3 * j + dd];
    // find_ab_ueg
    const uint64_t* t = &u.ternary[(size_t)(sys.W + 1) * ((ij_k[0] + u.tK) + (size_t)u.tD * ((ij_k[1] + u.tK) + (size_t)u.tD * (ij_k[2] + u.tK)))];
    uint64_t poss[MAXW];
    int nposs[MAXW], max_na = 0;
    for (int w = 0; w < sys.W; ++w) {
        const uint64_t tc = (ij_spin == -2) ? (t[1 + w] << 1) : t[1 + w];
        poss[w] = ~d.f.w[w] & tc;
        nposs[w] = __builtin_popcountll(poss[w]);
        max_na += nposs[w];
    }
    if (max_na > 0) {
        int a = (int)(max_na * rng.next()) + 1;
        int n = 0;
        for (int w = 0; w < sys.W; ++w) {
            if (n + nposs[w] >= a) {
                uint64_t x = poss[w];
                for (int s = 1; s < a - n; ++s) x &= x - 1;
                a = w * 64 + __builtin_ctzll(x) + 1;
                break;
            } else {
                n += nposs[w];
            }
        }
        int kb[3];
        for (int dd = 0; dd < 3; ++dd) kb[dd] = ij_k[dd] - u.l[3 * a + dd];
        int b = ueg_basis_index(sys, kb, ij_spin == 2 ? 1 : -1);
        g.allowed = !det_test(d.f, b);
        if (a > b) { int tmp = a; a = b; b = tmp; }
        g.conn.to_orb[0] = a; g.conn.to_orb[1] = b;
    } else {
        g.allowed = false;
    }
    if (g.allowed) {
        g.pgen = 2.0 / (nel * (nel - 1) * max_na);
        if (ij_spin != 0) g.pgen = g.pgen * 2;
        sys.find_excitation_permutation2(d.f, g.conn);
        g.hmatel = slater_condon2_ueg_excit(sys, g.conn.from_orb[0], g.conn.to_orb[0], g.conn.to_orb[1], g.conn.perm);
    } else {
        g.hmatel = 0.0;
        g.pgen = 1.0;
    }
    return g;
}

// init_excit_ueg_power_pitzer (src/excit_gen_ueg.f90:362-408) with create_weighted_excitation_list_ueg
// (src/hamiltonian_ueg.f90): for every orbital i an alias table over all orbitals a of its spin, weights |<ia|ai>|
inline void init_excit_ueg_power_pitzer(const System& sys, PowerPitzerN& pp) {
    const int nbas = sys.nbasis, maxv = nbas / 2;
    pp.pp_ia_d.alloc(maxv, nbas + 1);
    for (int i = 1; i <= nbas; ++i) {
        double* w = &pp.pp_ia_d.w[(size_t)maxv * i];
        double tot = 0.0;
        for (int j = 1; j <= maxv; ++j) {
            const int a = j * 2 - (i % 2);
            const double weight = (a != i) ? std::fabs(coulomb_int_ueg_3d(sys, i, a)) : 0.0;
            w[j - 1] = weight;
            tot = tot + weight;
        }
        pp.pp_ia_d.tot[i] = tot;
        generate_alias_tables(maxv, w, tot, &pp.pp_ia_d.U[(size_t)maxv * i], &pp.pp_ia_d.K[(size_t)maxv * i]);
    }
}
// gen_excit_ueg_power_pitzer (src/excit_gen_ueg.f90:410-566)
inline GenResult gen_excit_ueg_power_pitzer(Rng& rng, const System& sys, const ExcitGenData& eg, const DetInfo& d) {
    const UegData& u = sys.ueg;
    const PowerPitzerN& pp = eg.ppn;
    GenResult g;
    g.conn.nexcit = 2;
    const int nel = sys.nel, maxv = sys.nbasis / 2;
    double r = rng.next();
    int ind = (int)(r * nel * (nel - 1) / 2) + 1;
    int jj = (int)(1.50 + std::sqrt(2 * ind - 1.750));
    int ii = ind - ((jj - 1) * (jj - 2)) / 2;
    const int i = d.occ[ii - 1], j = d.occ[jj - 1];
    const int ij_spin = sys.bf[i].ms + sys.bf[j].ms;
    int ij_k[3];
    for (int dd = 0; dd < 3; ++dd) ij_k[dd] = u.l[3 * i + dd] + u.l[3 * j + dd];
    const int a_ind = select_weighted_value_precalc(rng, maxv, &pp.pp_ia_d.U[(size_t)maxv * i], &pp.pp_ia_d.K[(size_t)maxv * i]);
    const int a = 2 * a_ind - (i % 2);
    int b = 0, b_ind = 0;
    g.allowed = !det_test(d.f, a);
    if (g.allowed) {
        int kb[3];
        for (int dd = 0; dd < 3; ++dd) kb[dd] = ij_k[dd] - u.l[3 * a + dd];
        if (ij_spin == 2) b = ueg_basis_index(sys, kb, 1);
        else if (ij_spin == 0) b = ueg_basis_index(sys, kb, -sys.bf[a].ms);
        else b = ueg_basis_index(sys, kb, -1);
        if (b <= 0) g.allowed = false;
        else { b_ind = (b + 1) / 2; g.allowed = !det_test(d.f, b); }
    }
    if (g.allowed) {
        const double* w = &pp.pp_ia_d.w[(size_t)maxv * i];
        if (ij_spin == 0) g.pgen = w[a_ind - 1] / pp.pp_ia_d.tot[i];
        else g.pgen = (w[a_ind - 1] + w[b_ind - 1]) / pp.pp_ia_d.tot[i];
        g.pgen = g.pgen * 2.0 / (nel * (nel - 1));
        g.allowed = (a != b);
    }
    if (g.allowed) {
        g.conn.from_orb[0] = i; g.conn.from_orb[1] = j;
        g.conn.to_orb[0] = std::min(a, b); g.conn.to_orb[1] = std::max(a, b);
        sys.find_excitation_permutation2(d.f, g.conn);
        g.hmatel = slater_condon2_ueg_excit(sys, g.conn.from_orb[0], g.conn.to_orb[0], g.conn.to_orb[1], g.conn.perm);
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
    return g;
}

// sc0_ptr / gen_excit_ptr%full dispatch over the system kind (init_proc_pointers, src/qmc.F90:217-703)
inline double diag_hmatel(const System& sys, const Det& f) {
    if (sys.kind == SYS_UEG) {
        int occ[256];
        sys.decode(f, occ);
        return slater_condon0_ueg_orb_list(sys, occ);
    }
    return sys.slater_condon0(f);
}
inline GenResult gen_excit_sys(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    if (sys.kind == SYS_UEG)
        return eg.excit_gen == EXCIT_GEN_POWER_PITZER ? gen_excit_ueg_power_pitzer(rng, sys, eg, d) : gen_excit_ueg_no_renorm(rng, sys, d);
    return gen_excit(rng, sys, eg, d);
}

}  // namespace oracle
