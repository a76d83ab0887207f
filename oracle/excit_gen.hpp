// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the reference's molecular excitation generators:
//   renorm      src/excit_gen_mol.f90:16-101,384-448,521-682,802-946,1140-1327
//   no_renorm   src/excit_gen_mol.f90:195-284,450-517,950-1136,1329-1445
//   heat_bath   src/excit_gen_heat_bath_mol.F90:14-548, src/excit_gen_utils.f90:9-66,142-160,
//               lib/local/alias.f90:13-184
// Random numbers are consumed in exactly the reference's order.
#pragma once
#include "system.hpp"
#include "rng.hpp"

namespace oracle {

constexpr int MAXNEL = 64;

enum ExcitGenKind {  // values of src/qmc_data.f90:31-69
    EXCIT_GEN_RENORM = 0,
    EXCIT_GEN_RENORM_SPIN = 1,
    EXCIT_GEN_NO_RENORM = 2,
    EXCIT_GEN_NO_RENORM_SPIN = 3,
    EXCIT_GEN_POWER_PITZER = 4,
    EXCIT_GEN_POWER_PITZER_OCC = 5,
    EXCIT_GEN_POWER_PITZER_OCC_IJ = 6,
    EXCIT_GEN_POWER_PITZER_ORDERN = 7,
    EXCIT_GEN_CAUCHY_SCHWARZ_OCC = 8,
    EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ = 9,
    EXCIT_GEN_HEAT_BATH = 10,
    EXCIT_GEN_HEAT_BATH_UNIFORM = 11,
    EXCIT_GEN_HEAT_BATH_SINGLE = 12,
};

struct DetInfo {
    Det f;
    int occ[MAXNEL];
    std::vector<int> symunocc;  // [(ims-1) + 2*sym]
    int initiator_flag = 0;
    // heat-bath per-determinant cache (det_info_t%i_d_occ, src/determinant_data.f90:22-70)
    bool double_precalc = false;
    double i_d_weights[MAXNEL];
    double i_d_weights_tot = 0.0;
    // decode_det_spinocc_spinsymunocc (src/determinant_decoders.f90:371-495): unoccupied orbitals per spin, ascending
    std::vector<int> unocc_alpha, unocc_beta;
    // heat_bath_single per-determinant cache (det_info_t%i_s_occ, ia_s_weights_occ, unocc_list)
    bool single_precalc = false;
    std::vector<int> unocc;
    std::vector<double> i_s_weights, ia_s_weights;   // ia_s_weights[(a_ind-1) + nvirt*(i_ind-1)]
    double i_s_weights_tot = 0.0;
    // power_pitzer_orderN: det_info_t%ref_cdet_occ_list (find_diff_ref_cdet, src/excit_gen_utils.f90:220-269)
    int ref_cdet_occ[MAXNEL];
    inline int su(int ims, int sym) const { return symunocc[(ims - 1) + 2 * sym]; }
};

// decode_det_occ / decode_det_occ_symunocc (src/determinant_decoders.f90:15-35,166-206)
inline void decode_det_occ(const System& sys, const Det& f, DetInfo& d) {
    d.f = f;
    sys.decode(f, d.occ);
    d.double_precalc = false;
    d.single_precalc = false;
}
inline void decode_det_occ_symunocc(const System& sys, const Det& f, DetInfo& d) {
    decode_det_occ(sys, f, d);
    d.symunocc = sys.nbasis_sym_spin;
    for (int i = 0; i < sys.nel; ++i) {
        const BasisFn& b = sys.bf[d.occ[i]];
        d.symunocc[((b.ms + 3) / 2 - 1) + 2 * b.sym]--;
    }
}

struct GenResult {
    Excit conn;
    double pgen = 1.0;
    double hmatel = 0.0;
    bool allowed = false;
};

// ------------------------------------------------------------------------ alias method
// lib/local/alias.f90:101-184.  Arrays are 0-based here; returned/stored indices 1-based.
inline void generate_alias_tables(int N, const double* weights, double totweight, double* aliasU, int* aliasK) {
    std::vector<int> underfull(N), overfull(N);
    int nunder = 0, nover = 0;
    double scale = N / totweight;
    for (int i = 0; i < N; ++i) aliasU[i] = weights[i] * scale;
    for (int i = 0; i < N; ++i) {
        if (aliasU[i] <= 1.0) underfull[nunder++] = i; else overfull[nover++] = i;
        aliasK[i] = i + 1;
    }
    while (nover > 0 && nunder > 0) {
        int ov = overfull[nover - 1];
        int un = underfull[nunder - 1];
        aliasK[un] = ov + 1;
        nunder--;
        aliasU[ov] = aliasU[ov] - (1 - aliasU[un]);
        if (aliasU[ov] < 1.0) {
            underfull[nunder++] = overfull[nover - 1];
            nover--;
        }
    }
}
// lib/local/alias.f90:13-66
inline int select_weighted_value_precalc(Rng& rng, int N, const double* aliasU, const int* aliasK) {
    double x = rng.next() * N;
    int K = (int)std::floor(x);
    x = x - K;
    if (x < aliasU[K]) return K + 1;
    return aliasK[K];
}
// lib/local/alias.f90:68-99
inline int select_weighted_value(Rng& rng, int N, const double* weights, double totweight) {
    double aliasU[MAXNEL * 4];
    int aliasK[MAXNEL * 4];
    generate_alias_tables(N, weights, totweight, aliasU, aliasK);
    return select_weighted_value_precalc(rng, N, aliasU, aliasK);
}

// ------------------------------------------------------------------------ heat-bath tables
struct HeatBath {
    int nb = 0;
    std::vector<double> i_weights;      // (nb)
    std::vector<double> ij_weights;     // (j,i)
    std::vector<double> ija_w, ija_U;   // (a,j,i)
    std::vector<int> ija_K;
    std::vector<double> ija_tot;        // (j,i)
    std::vector<double> ijab_w, ijab_U; // (b,a,j,i)
    std::vector<int> ijab_K;
    std::vector<double> ijab_tot;       // (a,j,i)
    inline size_t i2(int j, int i) const { return (size_t)(j - 1) + (size_t)nb * (i - 1); }
    inline size_t i3(int a, int j, int i) const { return (size_t)(a - 1) + (size_t)nb * ((j - 1) + (size_t)nb * (i - 1)); }
    inline size_t i4(int b, int a, int j, int i) const {
        return (size_t)(b - 1) + (size_t)nb * ((a - 1) + (size_t)nb * ((j - 1) + (size_t)nb * (i - 1)));
    }
};

// init_excit_mol_heat_bath (src/excit_gen_heat_bath_mol.F90:14-256).  Returns false if the
// "not all single excitations can be accounted for" check (:171-182) fails for original=true.
inline bool init_excit_mol_heat_bath(const System& sys, HeatBath& hb, bool original) {
    int nb = sys.nbasis;
    hb.nb = nb;
    size_t n2 = (size_t)nb * nb, n3 = n2 * nb, n4 = n3 * nb;
    hb.i_weights.assign(nb, 0.0);
    hb.ij_weights.assign(n2, 0.0);
    hb.ija_w.assign(n3, 0.0); hb.ija_U.assign(n3, 0.0); hb.ija_K.assign(n3, 0);
    hb.ija_tot.assign(n2, 0.0);
    hb.ijab_w.assign(n4, 0.0); hb.ijab_U.assign(n4, 0.0); hb.ijab_K.assign(n4, 0);
    hb.ijab_tot.assign(n3, 0.0);
    std::vector<int> j_nonzero(n2, 0);  // (a,i)
    bool ok = true;
    for (int i = 1; i <= nb; ++i) {
        double i_weight = 0.0;
        for (int j = 1; j <= nb; ++j) {
            double ij_weight = 0.0;
            hb.ija_tot[hb.i2(j, i)] = 0.0;
            if (i != j) {
                int i_tmp = std::min(i, j), j_tmp = std::max(i, j);
                int ij_sym = sys.sym_conj(sys.cross_product_basis(i_tmp, j_tmp));
                double ija_weights_tot = 0.0, i_weight_extra = 0.0;
                for (int a = 1; a <= nb; ++a) {
                    double ija_weight = 0.0;
                    hb.ijab_tot[hb.i3(a, j, i)] = 0.0;
                    if (a != i && a != j) {
                        int isymb = sys.sym_conj(sys.cross_product(ij_sym, sys.bf[a].sym));
                        for (int b = 1; b <= nb; ++b) {
                            double ijab_weight = 0.0;
                            bool spin_ok = (sys.bf[i_tmp].ms == sys.bf[a].ms && sys.bf[j_tmp].ms == sys.bf[b].ms) ||
                                           (sys.bf[i_tmp].ms == sys.bf[b].ms && sys.bf[j_tmp].ms == sys.bf[a].ms);
                            if (spin_ok && sys.bf[b].sym == isymb && b != a && b != i && b != j) {
                                int a_tmp = std::min(a, b), b_tmp = std::max(a, b);
                                double h = std::fabs(sys.slater_condon2_excit(i_tmp, j_tmp, a_tmp, b_tmp, false));
                                i_weight_extra = i_weight_extra + h;
                                ij_weight = ij_weight + h;
                                ija_weight = ija_weight + h;
                                ijab_weight = ijab_weight + h;
                            }
                            hb.ijab_w[hb.i4(b, a, j, i)] = ijab_weight;
                            hb.ijab_tot[hb.i3(a, j, i)] = hb.ijab_tot[hb.i3(a, j, i)] + ijab_weight;
                        }
                    }
                    hb.ija_w[hb.i3(a, j, i)] = ija_weight;
                    ija_weights_tot = ija_weights_tot + ija_weight;
                    if (ija_weight > depsilon) j_nonzero[(size_t)(a - 1) + (size_t)nb * (i - 1)]++;
                }
                i_weight = i_weight + i_weight_extra;
                hb.ija_tot[hb.i2(j, i)] = hb.ija_tot[hb.i2(j, i)] + ija_weights_tot;
            }
            hb.ij_weights[hb.i2(j, i)] = ij_weight;
        }
        hb.i_weights[i - 1] = i_weight;
        for (int a = 1; a <= nb; ++a) {
            if (j_nonzero[(size_t)(a - 1) + (size_t)nb * (i - 1)] < (nb - sys.nel) && original && i != a) {
                int isyma = sys.cross_product(sys.bf[i].sym, sys.gamma_sym);
                if (sys.bf[a].sym == isyma && sys.bf[a].ms == sys.bf[i].ms) ok = false;
            }
        }
    }
    for (int i = 1; i <= nb; ++i)
        for (int j = 1; j <= nb; ++j)
            if (std::fabs(hb.ija_tot[hb.i2(j, i)]) > 0.0) {
                size_t o3 = hb.i3(1, j, i);
                generate_alias_tables(nb, &hb.ija_w[o3], hb.ija_tot[hb.i2(j, i)], &hb.ija_U[o3], &hb.ija_K[o3]);
                for (int a = 1; a <= nb; ++a)
                    if (std::fabs(hb.ijab_tot[hb.i3(a, j, i)]) > 0.0) {
                        size_t o4 = hb.i4(1, a, j, i);
                        generate_alias_tables(nb, &hb.ijab_w[o4], hb.ijab_tot[hb.i3(a, j, i)], &hb.ijab_U[o4],
                                              &hb.ijab_K[o4]);
                    }
            }
    return ok;
}

// alias_table_data_*_t (src/excit_gens.f90:43-100): columns of `stride` entries
struct AliasCols {
    int stride = 0;
    std::vector<double> w, U, tot;
    std::vector<int> K;
    void alloc(int stride_, size_t ncols) {
        stride = stride_;
        w.assign((size_t)stride * ncols, 0.0); U.assign((size_t)stride * ncols, 0.0); K.assign((size_t)stride * ncols, 0);
        tot.assign(ncols, 0.0);
    }
};
// excit_gen_power_pitzer_t, the ppn_* members (src/excit_gens.f90:102-141): reference-mapped O(N) Power-Pitzer tables
struct PowerPitzerN {
    std::vector<int> occ_list, all_list_alpha, all_list_beta;
    int n_all_alpha = 0, n_all_beta = 0;
    double min_weight = 0.01;        // qmc_in%power_pitzer_min_weight (src/qmc_data.f90:224)
    AliasCols i_s, ia_s, i_d, ij_d, ia_d, jb_d;   // columns: [1], [nbasis], [1], [nbasis], [nbasis], [nsym_tot * nbasis]
    // excit_gen = power_pitzer (pp_ia_d, pp_jb_d of init_excit_mol_power_pitzer_occ_ref): columns [nel], [nsym_tot * nel]
    std::vector<int> virt_list_alpha, virt_list_beta;
    AliasCols pp_ia_d, pp_jb_d;
};

struct ExcitGenData {
    int excit_gen = EXCIT_GEN_RENORM;
    PowerPitzerN ppn;
    double pattempt_single = 0.0, pattempt_double = 1.0;
    double pattempt_parallel = 0.0;   // renorm_spin / no_renorm_spin: probability that i and j have parallel spins
    HeatBath hb;
};

// find_single_double_prob (src/qmc_common.F90:152-260), read_in branch
inline void find_single_double_prob(const System& sys, const int* occ, double& psingle, double& pdouble) {
    std::vector<int> virt = sys.nbasis_sym_spin;
    auto V = [&](int ims, int sym) -> int& { return virt[(ims - 1) + 2 * sym]; };
    for (int i = 0; i < sys.nel; ++i) V((sys.bf[occ[i]].ms + 3) / 2, sys.bf[occ[i]].sym)--;
    int nsingles = 0;
    for (int i = 0; i < sys.nel; ++i) nsingles += V((sys.bf[occ[i]].ms + 3) / 2, sys.bf[occ[i]].sym);
    int ndoubles = 0;
    for (int i = 0; i < sys.nel; ++i) {
        int ims1 = (sys.bf[occ[i]].ms + 3) / 2;
        for (int j = i + 1; j < sys.nel; ++j) {
            int ims2 = (sys.bf[occ[j]].ms + 3) / 2;
            for (int isyma = sys.sym0; isyma <= sys.sym_max; ++isyma) {
                int isymb = sys.cross_product(sys.bf[occ[i]].sym, sys.bf[occ[j]].sym);
                isymb = sys.cross_product(isyma, isymb);
                if (isyma == isymb) {
                    if (ims1 == ims2) ndoubles += (V(ims1, isyma) * (V(ims2, isymb) - 1)) / 2;
                    else ndoubles += V(ims1, isyma) * V(ims2, isymb);
                } else if (isyma < isymb) {
                    ndoubles += V(ims1, isyma) * V(ims2, isymb);
                    if (ims1 != ims2) ndoubles += V(ims2, isyma) * V(ims1, isymb);
                }
            }
        }
    }
    psingle = (double)nsingles / (nsingles + ndoubles);
    pdouble = (double)ndoubles / (nsingles + ndoubles);
}

// ------------------------------------------------------------------------ renorm pieces
// choose_ij_mol (src/excit_gen_mol.f90:620-682)
inline void choose_ij_mol(Rng& rng, const System& sys, const int* occ, int& i, int& j, int& ij_sym, int& ij_spin,
                          double& pgen_ij) {
    int nel = sys.nel;
    int ind = (int)(rng.next() * nel * (nel - 1) / 2) + 1;
    int j_ind = (int)(1.50 + std::sqrt(2 * ind - 1.750));
    int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
    i = occ[i_ind - 1];
    j = occ[j_ind - 1];
    ij_sym = sys.sym_conj(sys.cross_product_basis(i, j));
    ij_spin = sys.bf[i].ms + sys.bf[j].ms;
    pgen_ij = 2.0 / (nel * (nel - 1));
}

// choose_ia_mol (src/excit_gen_mol.f90:521-616)
inline void choose_ia_mol(Rng& rng, const System& sys, int op_sym, const DetInfo& d, int& i, int& a, bool& allowed) {
    allowed = false;
    for (int k = 0; k < sys.nel; ++k) {
        int imsa = (sys.bf[d.occ[k]].ms + 3) / 2;
        int isyma = sys.cross_product(sys.bf[d.occ[k]].sym, op_sym);
        if (d.su(imsa, isyma) != 0) { allowed = true; break; }
    }
    if (allowed) {
        for (;;) {
            i = d.occ[(int)(rng.next() * sys.nel)];
            int imsa = (sys.bf[i].ms + 3) / 2;
            int isyma = sys.cross_product(sys.bf[i].sym, op_sym);
            if (d.su(imsa, isyma) != 0) {
                for (;;) {
                    int ind = (int)(sys.nbss(imsa, isyma) * rng.next()) + 1;
                    a = sys.ssbf(ind, imsa, isyma);
                    if (!det_test(d.f, a)) break;
                }
                break;
            }
        }
    }
}

// choose_ab_mol (src/excit_gen_mol.f90:802-946)
inline void choose_ab_mol(Rng& rng, const System& sys, const DetInfo& d, int sym, int spin, int& a, int& b,
                          bool& allowed) {
    allowed = false;
    int fac = 1, shift = 0, na = sys.nbasis;
    if (spin == -2) {
        for (int isyma = sys.sym0; isyma <= sys.sym_max; ++isyma) {
            int isymb = sys.sym_conj(sys.cross_product(isyma, sym));
            if (d.su(1, isyma) > 0 && (d.su(1, isymb) > 1 || (d.su(1, isymb) == 1 && isyma != isymb))) {
                allowed = true; break;
            }
        }
        fac = 2; shift = 0; na = sys.nbasis / 2;
    } else if (spin == 0) {
        for (int isyma = sys.sym0; isyma <= sys.sym_max; ++isyma) {
            int isymb = sys.sym_conj(sys.cross_product(isyma, sym));
            if ((d.su(1, isyma) > 0 && d.su(2, isymb) > 0) || (d.su(2, isyma) > 0 && d.su(1, isymb) > 0)) {
                allowed = true; break;
            }
        }
        fac = 1; shift = 0; na = sys.nbasis;
    } else {
        for (int isyma = sys.sym0; isyma <= sys.sym_max; ++isyma) {
            int isymb = sys.sym_conj(sys.cross_product(isyma, sym));
            if (d.su(2, isyma) > 0 && (d.su(2, isymb) > 1 || (d.su(2, isymb) == 1 && isyma != isymb))) {
                allowed = true; break;
            }
        }
        fac = 2; shift = 1; na = sys.nbasis / 2;
    }
    if (allowed) {
        for (;;) {
            a = (int)(rng.next() * na) + 1;
            a = fac * a - shift;
            if (!det_test(d.f, a)) {
                int imsb = (spin - sys.bf[a].ms + 3) / 2;
                int isymb = sys.sym_conj(sys.cross_product(sym, sys.bf[a].sym));
                if (d.su(imsb, isymb) > 1 || (d.su(imsb, isymb) == 1 && (isymb != sys.bf[a].sym || spin == 0))) {
                    for (;;) {
                        int ind = (int)(sys.nbss(imsb, isymb) * rng.next()) + 1;
                        b = sys.ssbf(ind, imsb, isymb);
                        if (b != a && !det_test(d.f, b)) break;
                    }
                    break;
                }
            }
        }
        if (a > b) std::swap(a, b);
    }
}

// calc_pgen_single_mol (src/excit_gen_mol.f90:1140-1190)
inline double calc_pgen_single_mol(const System& sys, int op_sym, const DetInfo& d, int a) {
    int ni = sys.nel;
    for (int k = 0; k < sys.nel; ++k) {
        int imsa = (sys.bf[d.occ[k]].ms + 3) / 2;
        int isyma = sys.cross_product(sys.bf[d.occ[k]].sym, op_sym);
        if (d.su(imsa, isyma) == 0) ni--;
    }
    int imsi = (sys.bf[a].ms + 3) / 2, isymi = sys.bf[a].sym;
    return 1.0 / (ni * d.su(imsi, isymi));
}

// calc_pgen_double_mol (src/excit_gen_mol.f90:1192-1327)
inline double calc_pgen_double_mol(const System& sys, int ij_sym, int a, int b, int spin, const DetInfo& d) {
    int imsa = (sys.bf[a].ms + 3) / 2, imsb = (sys.bf[b].ms + 3) / 2;
    int n_aij;
    double p_aijb, p_bija;
    if (spin == -2 || spin == 2) {
        int s = (spin == -2) ? 1 : 2;
        n_aij = (spin == -2) ? sys.nvirt_beta : sys.nvirt_alpha;
        for (int isyma = sys.sym0; isyma <= sys.sym_max; ++isyma) {
            int isymb = sys.sym_conj(sys.cross_product(isyma, ij_sym));
            if (d.su(s, isymb) == 0) n_aij -= d.su(s, isyma);
            else if (isyma == isymb && d.su(s, isymb) == 1) n_aij -= d.su(s, isyma);
        }
        if (sys.bf[a].sym == sys.bf[b].sym) {
            p_aijb = 1.0 / (d.su(imsa, sys.bf[a].sym) - 1);
            p_bija = 1.0 / (d.su(imsb, sys.bf[b].sym) - 1);
        } else {
            p_aijb = 1.0 / d.su(imsa, sys.bf[a].sym);
            p_bija = 1.0 / d.su(imsb, sys.bf[b].sym);
        }
    } else {
        n_aij = sys.nvirt;
        for (int isyma = sys.sym0; isyma <= sys.sym_max; ++isyma) {
            int isymb = sys.sym_conj(sys.cross_product(isyma, ij_sym));
            if (d.su(1, isymb) == 0) n_aij -= d.su(2, isyma);
            if (d.su(2, isymb) == 0) n_aij -= d.su(1, isyma);
        }
        p_aijb = 1.0 / d.su(imsa, sys.bf[a].sym);
        p_bija = 1.0 / d.su(imsb, sys.bf[b].sym);
    }
    return (1.0 / n_aij) * (p_bija + p_aijb);
}

// gen_excit_mol (src/excit_gen_mol.f90:16-101)
inline GenResult gen_excit_mol(Rng& rng, const System& sys, const ExcitGenData& eg, const DetInfo& d) {
    GenResult r;
    if (rng.next() < eg.pattempt_single) {
        choose_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
        r.conn.nexcit = 1;
        if (r.allowed) {
            r.pgen = eg.pattempt_single * calc_pgen_single_mol(sys, sys.gamma_sym, d, r.conn.to_orb[0]);
            sys.find_excitation_permutation1(d.f, r.conn);
            r.hmatel = sys.slater_condon1_excit(d.occ, r.conn.from_orb[0], r.conn.to_orb[0], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
    } else {
        int ij_sym, ij_spin;
        double pgen_ij;
        choose_ij_mol(rng, sys, d.occ, r.conn.from_orb[0], r.conn.from_orb[1], ij_sym, ij_spin, pgen_ij);
        choose_ab_mol(rng, sys, d, ij_sym, ij_spin, r.conn.to_orb[0], r.conn.to_orb[1], r.allowed);
        r.conn.nexcit = 2;
        if (r.allowed) {
            r.pgen = eg.pattempt_double * pgen_ij *
                     calc_pgen_double_mol(sys, ij_sym, r.conn.to_orb[0], r.conn.to_orb[1], ij_spin, d);
            sys.find_excitation_permutation2(d.f, r.conn);
            r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0],
                                                r.conn.to_orb[1], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
    }
    return r;
}

// ------------------------------------------------------------------------ no_renorm
// find_ia_mol (src/excit_gen_mol.f90:950-1010)
inline void find_ia_mol(Rng& rng, const System& sys, int op_sym, const DetInfo& d, int& i, int& a, bool& allowed) {
    i = d.occ[(int)(rng.next() * sys.nel)];
    int imsa = (sys.bf[i].ms + 3) / 2;
    int isyma = sys.cross_product(sys.bf[i].sym, op_sym);
    int ind = (int)(sys.nbss(imsa, isyma) * rng.next()) + 1;
    if (sys.nbss(imsa, isyma) == 0) {
        allowed = false;
    } else {
        a = sys.ssbf(ind, imsa, isyma);
        allowed = !det_test(d.f, a);
    }
}
// find_ab_mol (src/excit_gen_mol.f90:1012-1136)
inline void find_ab_mol(Rng& rng, const System& sys, const DetInfo& d, int sym, int spin, int& a, int& b,
                        bool& allowed) {
    int fac = 1, shift = 0, na = sys.nbasis;
    if (spin == -2) { fac = 2; shift = 0; na = sys.nbasis / 2; }
    else if (spin == 2) { fac = 2; shift = 1; na = sys.nbasis / 2; }
    for (;;) {
        a = (int)(rng.next() * na) + 1;
        a = fac * a - shift;
        if (!det_test(d.f, a)) break;
    }
    int imsb = (spin - sys.bf[a].ms + 3) / 2;
    int isymb = sys.sym_conj(sys.cross_product(sym, sys.bf[a].sym));
    if (sys.nbss(imsb, isymb) == 0) {
        allowed = false;
    } else if (spin != 0 && isymb == sys.bf[a].sym && sys.nbss(imsb, isymb) == 1) {
        allowed = false;
    } else {
        for (;;) {
            int ind = (int)(sys.nbss(imsb, isymb) * rng.next()) + 1;
            b = sys.ssbf(ind, imsb, isymb);
            if (b != a) break;
        }
        allowed = !det_test(d.f, b);
        if (a > b) std::swap(a, b);
    }
}
// src/excit_gen_mol.f90:1329-1445
inline double calc_pgen_single_mol_no_renorm(const System& sys, int a) {
    return 1.0 / (sys.nel * sys.nbss((sys.bf[a].ms + 3) / 2, sys.bf[a].sym));
}
inline double calc_pgen_double_mol_no_renorm(const System& sys, int a, int b, int spin) {
    int n_aij = (spin == -2) ? sys.nvirt_beta : (spin == 0 ? sys.nvirt : sys.nvirt_alpha);
    int imsa = (sys.bf[a].ms + 3) / 2, isyma = sys.bf[a].sym;
    int imsb = (sys.bf[b].ms + 3) / 2, isymb = sys.bf[b].sym;
    double p_aijb, p_bija;
    if (isyma == isymb && imsa == imsb) {
        p_aijb = 1.0 / (sys.nbss(imsa, isyma) - 1);
        p_bija = 1.0 / (sys.nbss(imsb, isymb) - 1);
    } else {
        p_aijb = 1.0 / sys.nbss(imsa, isyma);
        p_bija = 1.0 / sys.nbss(imsb, isymb);
    }
    return (1.0 / n_aij) * (p_bija + p_aijb);
}
// gen_excit_mol_no_renorm (src/excit_gen_mol.f90:195-284)
inline GenResult gen_excit_mol_no_renorm(Rng& rng, const System& sys, const ExcitGenData& eg, const DetInfo& d) {
    GenResult r;
    if (rng.next() < eg.pattempt_single) {
        find_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
        r.conn.nexcit = 1;
        if (r.allowed) {
            r.pgen = eg.pattempt_single * calc_pgen_single_mol_no_renorm(sys, r.conn.to_orb[0]);
            sys.find_excitation_permutation1(d.f, r.conn);
            r.hmatel = sys.slater_condon1_excit(d.occ, r.conn.from_orb[0], r.conn.to_orb[0], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
    } else {
        int ij_sym, ij_spin;
        double pgen_ij;
        choose_ij_mol(rng, sys, d.occ, r.conn.from_orb[0], r.conn.from_orb[1], ij_sym, ij_spin, pgen_ij);
        find_ab_mol(rng, sys, d, ij_sym, ij_spin, r.conn.to_orb[0], r.conn.to_orb[1], r.allowed);
        r.conn.nexcit = 2;
        if (r.allowed) {
            r.pgen = eg.pattempt_double * pgen_ij *
                     calc_pgen_double_mol_no_renorm(sys, r.conn.to_orb[0], r.conn.to_orb[1], ij_spin);
            sys.find_excitation_permutation2(d.f, r.conn);
            r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0],
                                                r.conn.to_orb[1], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
    }
    return r;
}

// ------------------------------------------------------------------------ renorm_spin / no_renorm_spin
// find_parallel_spin_prob_mol (src/qmc_common.F90:262-377): ratio of sum |H_ij->ab| over parallel-spin ij to the sum over
// all ij; the i loop is split over the MPI ranks (get_proc_loop_range, lib/local/parallel.F90:227-269) and the partial
// sums are reduced, which fixes the summation order reproduced here.
inline double find_parallel_spin_prob_mol(const System& sys, int nprocs) {
    const int nb = sys.nbasis;
    double par_tot = 0.0, ortho_tot = 0.0;
    int i_end = 0;
    for (int ip = 0; ip < nprocs; ++ip) {
        const int i_start = i_end + 1;
        i_end = i_start + nb / nprocs - 1;
        if (ip < nb % nprocs) i_end = i_end + 1;
        double par = 0.0, ortho = 0.0;
        for (int i = i_start; i <= i_end; ++i)
            for (int j = 1; j <= nb; ++j) {
                if (i == j) continue;
                const int it = std::min(i, j), jt = std::max(i, j);
                const int ij_sym = sys.sym_conj(sys.cross_product_basis(it, jt));
                for (int a = 1; a <= nb; ++a) {
                    if (a == i || a == j) continue;
                    const int isymb = sys.sym_conj(sys.cross_product(ij_sym, sys.bf[a].sym));
                    for (int b = 1; b <= nb; ++b) {
                        const bool spin_ok = (sys.bf[it].ms == sys.bf[a].ms && sys.bf[jt].ms == sys.bf[b].ms) ||
                                             (sys.bf[it].ms == sys.bf[b].ms && sys.bf[jt].ms == sys.bf[a].ms);
                        if (!(spin_ok && sys.bf[b].sym == isymb && b != a && b != i && b != j)) continue;
                        const int at = std::min(a, b), bt = std::max(a, b);
                        const double h = std::fabs(sys.slater_condon2_excit(it, jt, at, bt, false));
                        if (sys.bf[it].ms == sys.bf[jt].ms) par = par + h;
                        else ortho = ortho + h;
                    }
                }
            }
        par_tot += par; ortho_tot += ortho;
    }
    return par_tot / (par_tot + ortho_tot);
}

// choose_ij_spin_mol (src/excit_gen_mol.f90:684-800): decide parallel / anti-parallel first, then pick the pair
inline void choose_ij_spin_mol(Rng& rng, const System& sys, const DetInfo& d, double pattempt_parallel, int& i, int& j,
                               int& ij_sym, int& ij_spin, double& pgen_ij, bool& allowed) {
    // decode_det_spinocc_symunocc (src/determinant_decoders.f90:208-262): alpha = ms +1
    int occ_a[MAXNEL], occ_b[MAXNEL], na = 0, nb = 0;
    for (int q = 0; q < sys.nel; ++q) {
        if (sys.bf[d.occ[q]].ms > 0) occ_a[na++] = d.occ[q];
        else occ_b[nb++] = d.occ[q];
    }
    const int nalpha = sys.nalpha, nbeta = sys.nbeta;
    allowed = true;
    if (rng.next() < pattempt_parallel) {
        if (rng.next() < ((double)nalpha / (double)(nalpha + nbeta))) {
            if (nalpha < 2) allowed = false;
            else {
                const int ind = (int)(rng.next() * (double)(nalpha * (nalpha - 1)) / 2.0) + 1;
                const int j_ind = (int)(1.50 + std::sqrt(2 * ind - 1.750));
                const int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
                i = occ_a[i_ind - 1]; j = occ_a[j_ind - 1];
                pgen_ij = (pattempt_parallel * ((double)nalpha / (double)(nalpha + nbeta)) * 2.0 * (1.0 / nalpha) *
                           (1.0 / (nalpha - 1)));
            }
        } else {
            if (nbeta < 2) allowed = false;
            else {
                const int ind = (int)(rng.next() * (double)(nbeta * (nbeta - 1)) / 2.0) + 1;
                const int j_ind = (int)(1.50 + std::sqrt(2 * ind - 1.750));
                const int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
                i = occ_b[i_ind - 1]; j = occ_b[j_ind - 1];
                pgen_ij = (pattempt_parallel * ((double)nbeta / (double)(nalpha + nbeta)) * 2.0 * (1.0 / nbeta) *
                           (1.0 / (nbeta - 1)));
            }
        }
    } else {
        if (nbeta < 1 || nalpha < 1) allowed = false;
        else {
            const int i_ind = (int)(rng.next() * nalpha) + 1;
            const int j_ind = (int)(rng.next() * nbeta) + 1;
            i = occ_a[i_ind - 1]; j = occ_b[j_ind - 1];
            if (j < i) std::swap(i, j);
            pgen_ij = (1.0 - pattempt_parallel) * (1.0 / (nalpha * nbeta));
        }
    }
    if (allowed) {
        ij_sym = sys.sym_conj(sys.cross_product_basis(i, j));
        ij_spin = sys.bf[i].ms + sys.bf[j].ms;
    } else { pgen_ij = 1.0; i = -1; j = -1; ij_sym = -1; ij_spin = -1; }
}

// gen_excit_mol_spin / gen_excit_mol_no_renorm_spin (src/excit_gen_mol.f90:103-193,286-380)
inline GenResult gen_excit_mol_spin(Rng& rng, const System& sys, const ExcitGenData& eg, const DetInfo& d, bool renorm) {
    GenResult r;
    if (rng.next() < eg.pattempt_single) {
        if (renorm) choose_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
        else find_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
        r.conn.nexcit = 1;
        if (r.allowed) {
            r.pgen = eg.pattempt_single * (renorm ? calc_pgen_single_mol(sys, sys.gamma_sym, d, r.conn.to_orb[0])
                                                  : calc_pgen_single_mol_no_renorm(sys, r.conn.to_orb[0]));
            sys.find_excitation_permutation1(d.f, r.conn);
            r.hmatel = sys.slater_condon1_excit(d.occ, r.conn.from_orb[0], r.conn.to_orb[0], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
    } else {
        int ij_sym, ij_spin;
        double pgen_ij;
        choose_ij_spin_mol(rng, sys, d, eg.pattempt_parallel, r.conn.from_orb[0], r.conn.from_orb[1], ij_sym, ij_spin,
                           pgen_ij, r.allowed);
        if (r.allowed) {
            if (renorm) choose_ab_mol(rng, sys, d, ij_sym, ij_spin, r.conn.to_orb[0], r.conn.to_orb[1], r.allowed);
            else find_ab_mol(rng, sys, d, ij_sym, ij_spin, r.conn.to_orb[0], r.conn.to_orb[1], r.allowed);
            r.conn.nexcit = 2;
        }
        if (r.allowed) {
            r.pgen = eg.pattempt_double * pgen_ij *
                     (renorm ? calc_pgen_double_mol(sys, ij_sym, r.conn.to_orb[0], r.conn.to_orb[1], ij_spin, d)
                             : calc_pgen_double_mol_no_renorm(sys, r.conn.to_orb[0], r.conn.to_orb[1], ij_spin));
            sys.find_excitation_permutation2(d.f, r.conn);
            r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0],
                                                r.conn.to_orb[1], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
    }
    return r;
}

// ------------------------------------------------------------------------ heat bath (original)
// gen_excit_mol_heat_bath (src/excit_gen_heat_bath_mol.F90:258-548)
inline GenResult gen_excit_mol_heat_bath(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    GenResult r;
    const HeatBath& hb = eg.hb;
    const int nel = sys.nel, nb = sys.nbasis;
    double ij_w[MAXNEL], ji_w[MAXNEL];
    double ij_tot = 0.0, ji_tot = 0.0;
    if (!d.double_precalc) {
        // find_i_d_weights (src/excit_gen_utils.f90:142-160)
        d.i_d_weights_tot = 0.0;
        for (int p = 0; p < nel; ++p) {
            d.i_d_weights[p] = hb.i_weights[d.occ[p] - 1];
            d.i_d_weights_tot = d.i_d_weights_tot + d.i_d_weights[p];
        }
        d.double_precalc = true;
    }
    // select_ij_heat_bath (src/excit_gen_utils.f90:9-66)
    int i_ind = select_weighted_value(rng, nel, d.i_d_weights, d.i_d_weights_tot);
    int i = d.occ[i_ind - 1];
    int j_ind = 0, j = 0;
    for (int p = 0; p < nel; ++p) {
        ij_w[p] = hb.ij_weights[hb.i2(d.occ[p], i)];
        ij_tot = ij_tot + ij_w[p];
    }
    bool allowed;
    if (ij_tot > 0.0) {
        j_ind = select_weighted_value(rng, nel, ij_w, ij_tot);
        j = d.occ[j_ind - 1];
        for (int p = 0; p < nel; ++p) {
            ji_w[p] = hb.ij_weights[hb.i2(d.occ[p], j)];
            ji_tot = ji_tot + ji_w[p];
        }
        allowed = true;
    } else {
        allowed = false;
    }
    allowed = allowed && (std::fabs(hb.ija_tot[hb.i2(j > 0 ? j : 1, i)]) > 0.0) && j > 0;

    int a = 0, b = 0;
    bool dbl = true;
    double psingle = 0.0, hmatel_mod_ia = 0.0;
    if (allowed) {
        size_t o3 = hb.i3(1, j, i);
        a = select_weighted_value_precalc(rng, nb, &hb.ija_U[o3], &hb.ija_K[o3]);
        if (!det_test(d.f, a)) {
            int ims = sys.bf[i].ms;
            int isyma = sys.cross_product(sys.bf[i].sym, sys.gamma_sym);
            if (sys.bf[a].sym == isyma && sys.bf[a].ms == ims) {
                r.conn.from_orb[0] = i; r.conn.to_orb[0] = a; r.conn.nexcit = 1;
                sys.find_excitation_permutation1(d.f, r.conn);
                double hs = sys.slater_condon1_excit(d.occ, i, a, r.conn.perm);
                hmatel_mod_ia = std::fabs(hs);
                double x = rng.next();
                double wt = hb.ijab_tot[hb.i3(a, j, i)];
                if (hmatel_mod_ia < wt) psingle = hmatel_mod_ia / (wt + hmatel_mod_ia);
                else psingle = 0.5;
                dbl = !(x < psingle);
            } else {
                dbl = true;
                psingle = 0.0;
            }
        } else {
            allowed = false;
        }
    }
    if (allowed) {
        if (dbl) {
            size_t o4 = hb.i4(1, a, j, i);
            b = select_weighted_value_precalc(rng, nb, &hb.ijab_U[o4], &hb.ijab_K[o4]);
            if (!det_test(d.f, b)) {
                const double pi_ = d.i_d_weights[i_ind - 1] / d.i_d_weights_tot;
                const double pj_ = d.i_d_weights[j_ind - 1] / d.i_d_weights_tot;
                auto psingle_of = [&](int from, int to, int other) -> double {
                    // single-vs-double coin for ordering (from -> to | other)
                    int ims = sys.bf[from].ms;
                    int isyma = sys.cross_product(sys.bf[from].sym, sys.gamma_sym);
                    if (sys.bf[to].sym == isyma && sys.bf[to].ms == ims) {
                        Excit e; e.from_orb[0] = from; e.to_orb[0] = to; e.nexcit = 1;
                        sys.find_excitation_permutation1(d.f, e);
                        double hm = std::fabs(sys.slater_condon1_excit(d.occ, from, to, e.perm));
                        double wt = hb.ijab_tot[hb.i3(to, other, from)];
                        if (hm < wt) return hm / (wt + hm);
                        return 0.5;
                    }
                    return 0.0;
                };
                double pgen_ija = ((pi_) * (ij_w[j_ind - 1] / ij_tot)) *
                                  (hb.ija_w[hb.i3(a, j, i)] / hb.ija_tot[hb.i2(j, i)]) * (1.0 - psingle) *
                                  (hb.ijab_w[hb.i4(b, a, j, i)] / hb.ijab_tot[hb.i3(a, j, i)]);
                double ps = psingle_of(i, b, j);
                double pgen_ijb = ((pi_) * (ij_w[j_ind - 1] / ij_tot)) *
                                  (hb.ija_w[hb.i3(b, j, i)] / hb.ija_tot[hb.i2(j, i)]) * (1.0 - ps) *
                                  (hb.ijab_w[hb.i4(a, b, j, i)] / hb.ijab_tot[hb.i3(b, j, i)]);
                ps = psingle_of(j, a, i);
                double pgen_jia = ((pj_) * (ji_w[i_ind - 1] / ji_tot)) *
                                  (hb.ija_w[hb.i3(a, i, j)] / hb.ija_tot[hb.i2(i, j)]) * (1.0 - ps) *
                                  (hb.ijab_w[hb.i4(b, a, i, j)] / hb.ijab_tot[hb.i3(a, i, j)]);
                ps = psingle_of(j, b, i);
                double pgen_jib = ((pj_) * (ji_w[i_ind - 1] / ji_tot)) *
                                  (hb.ija_w[hb.i3(b, i, j)] / hb.ija_tot[hb.i2(i, j)]) * (1.0 - ps) *
                                  (hb.ijab_w[hb.i4(a, b, i, j)] / hb.ijab_tot[hb.i3(b, i, j)]);
                r.pgen = pgen_ija + pgen_ijb + pgen_jia + pgen_jib;
                r.conn.from_orb[0] = std::min(i, j); r.conn.from_orb[1] = std::max(i, j);
                r.conn.to_orb[0] = std::min(a, b); r.conn.to_orb[1] = std::max(a, b);
                r.conn.nexcit = 2;
                sys.find_excitation_permutation2(d.f, r.conn);
                r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0],
                                                    r.conn.to_orb[1], r.conn.perm);
                r.allowed = true;
            } else {
                r.allowed = false; r.hmatel = 0.0; r.pgen = 1.0;
                r.conn.nexcit = 2;
            }
        } else {
            r.conn.from_orb[0] = i; r.conn.to_orb[0] = a; r.conn.nexcit = 1;
            r.hmatel = sys.slater_condon1_excit(d.occ, i, a, r.conn.perm);
            double pgen = 0.0;
            for (int q = 0; q < nel; ++q) {
                int oq = d.occ[q];
                if (i != oq && a != oq) {
                    double wt = hb.ijab_tot[hb.i3(a, oq, i)];
                    double psq;
                    if (hmatel_mod_ia < wt) psq = hmatel_mod_ia / (wt + hmatel_mod_ia);
                    else psq = 0.5;
                    pgen = pgen + (psq * (ij_w[q] / ij_tot) * (hb.ija_w[hb.i3(a, oq, i)] / hb.ija_tot[hb.i2(oq, i)]));
                }
            }
            r.pgen = pgen * (d.i_d_weights[i_ind - 1] / d.i_d_weights_tot);
            r.allowed = true;
        }
    } else {
        r.allowed = false; r.hmatel = 0.0; r.pgen = 1.0;
    }
    return r;
}

// gen_excit_mol_heat_bath_uniform (src/excit_gen_heat_bath_mol.F90:550-718): heat_bath_uniform (singles from the
// renormalised uniform generator) and heat_bath_single (singles with exact weights,
// gen_single_excit_heat_bath_exact :720-805 + find_ia_single_weights src/excit_gen_utils.f90:162-218).
inline GenResult gen_excit_mol_heat_bath_uniform(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    GenResult r;
    const HeatBath& hb = eg.hb;
    const int nel = sys.nel, nb = sys.nbasis;
    if (rng.next() < eg.pattempt_single) {
        r.conn.nexcit = 1;
        if (eg.excit_gen == EXCIT_GEN_HEAT_BATH_UNIFORM) {
            // gen_single_excit_mol (src/excit_gen_mol.f90:384-448)
            choose_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
            if (r.allowed) {
                r.pgen = eg.pattempt_single * calc_pgen_single_mol(sys, sys.gamma_sym, d, r.conn.to_orb[0]);
                sys.find_excitation_permutation1(d.f, r.conn);
                r.hmatel = sys.slater_condon1_excit(d.occ, r.conn.from_orb[0], r.conn.to_orb[0], r.conn.perm);
            } else { r.hmatel = 0.0; r.pgen = 1.0; }
        } else {
            if (!d.single_precalc) {
                // decode_det_occ_unocc: unocc_list ascending; find_ia_single_weights
                d.unocc.clear();
                for (int o = 1; o <= nb; ++o) if (!det_test(d.f, o)) d.unocc.push_back(o);
                const int nvirt = (int)d.unocc.size();
                d.ia_s_weights.assign((size_t)nvirt * nel, 0.0);
                d.i_s_weights.assign(nel, 0.0);
                d.i_s_weights_tot = 0.0;
                for (int ii = 0; ii < nel; ++ii) {
                    d.i_s_weights[ii] = 0.0;
                    for (int aa = 0; aa < nvirt; ++aa) {
                        double h = sys.slater_condon1(d.occ, d.occ[ii], d.unocc[aa], false);
                        d.ia_s_weights[(size_t)aa + (size_t)nvirt * ii] = std::fabs(h);
                        d.i_s_weights[ii] = d.i_s_weights[ii] + std::fabs(h);
                    }
                    d.i_s_weights_tot = d.i_s_weights_tot + d.i_s_weights[ii];
                }
                d.single_precalc = true;
            }
            const int nvirt = (int)d.unocc.size();
            r.allowed = true;
            if (d.i_s_weights_tot < depsilon) {
                r.allowed = false; r.hmatel = 0.0; r.pgen = 1.0;
            } else {
                int i_ind = select_weighted_value(rng, nel, d.i_s_weights.data(), d.i_s_weights_tot);
                int i = d.occ[i_ind - 1];
                const double* col = &d.ia_s_weights[(size_t)nvirt * (i_ind - 1)];
                int a_ind = select_weighted_value(rng, nvirt, col, d.i_s_weights[i_ind - 1]);
                int a = d.unocc[a_ind - 1];
                r.conn.from_orb[0] = i; r.conn.to_orb[0] = a;
                r.pgen = eg.pattempt_single * (d.i_s_weights[i_ind - 1] / d.i_s_weights_tot) *
                         (col[a_ind - 1] / d.i_s_weights[i_ind - 1]);
                sys.find_excitation_permutation1(d.f, r.conn);
                r.hmatel = sys.slater_condon1_excit(d.occ, i, a, r.conn.perm);
            }
        }
        return r;
    }
    double ij_w[MAXNEL], ji_w[MAXNEL];
    double ij_tot = 0.0, ji_tot = 0.0;
    if (!d.double_precalc) {
        d.i_d_weights_tot = 0.0;
        for (int p = 0; p < nel; ++p) {
            d.i_d_weights[p] = hb.i_weights[d.occ[p] - 1];
            d.i_d_weights_tot = d.i_d_weights_tot + d.i_d_weights[p];
        }
        d.double_precalc = true;
    }
    // select_ij_heat_bath (src/excit_gen_utils.f90:9-66)
    int i_ind = select_weighted_value(rng, nel, d.i_d_weights, d.i_d_weights_tot);
    int i = d.occ[i_ind - 1];
    int j_ind = 0, j = 0;
    for (int p = 0; p < nel; ++p) {
        ij_w[p] = hb.ij_weights[hb.i2(d.occ[p], i)];
        ij_tot = ij_tot + ij_w[p];
    }
    bool allowed = false;
    if (ij_tot > 0.0) {
        j_ind = select_weighted_value(rng, nel, ij_w, ij_tot);
        j = d.occ[j_ind - 1];
        for (int p = 0; p < nel; ++p) {
            ji_w[p] = hb.ij_weights[hb.i2(d.occ[p], j)];
            ji_tot = ji_tot + ji_w[p];
        }
        allowed = std::fabs(hb.ija_tot[hb.i2(j, i)]) > 0.0;
    }
    double pgen = 0.0;
    if (allowed)
        pgen = ((d.i_d_weights[i_ind - 1] / d.i_d_weights_tot) * (ij_w[j_ind - 1] / ij_tot)) +
               ((d.i_d_weights[j_ind - 1] / d.i_d_weights_tot) * (ji_w[i_ind - 1] / ji_tot));
    if (j < i) std::swap(i, j);
    int a = 0, b = 0;
    if (allowed) {
        size_t o3 = hb.i3(1, j, i);
        a = select_weighted_value_precalc(rng, nb, &hb.ija_U[o3], &hb.ija_K[o3]);
        if (std::fabs(hb.ijab_tot[hb.i3(a, j, i)]) > 0.0 && !det_test(d.f, a)) {
            size_t o4 = hb.i4(1, a, j, i);
            b = select_weighted_value_precalc(rng, nb, &hb.ijab_U[o4], &hb.ijab_K[o4]);
            allowed = !det_test(d.f, b);
        } else {
            allowed = false;
        }
    }
    r.allowed = allowed;
    if (allowed) {
        r.conn.from_orb[0] = i; r.conn.from_orb[1] = j;
        r.conn.to_orb[0] = std::min(a, b); r.conn.to_orb[1] = std::max(a, b);
        r.conn.nexcit = 2;
        sys.find_excitation_permutation2(d.f, r.conn);
        r.hmatel = sys.slater_condon2_excit(i, j, r.conn.to_orb[0], r.conn.to_orb[1], r.conn.perm);
        r.pgen = (1.0 - eg.pattempt_single) * pgen *
                 (((hb.ija_w[hb.i3(a, j, i)] / hb.ija_tot[hb.i2(j, i)]) *
                   (hb.ijab_w[hb.i4(b, a, j, i)] / hb.ijab_tot[hb.i3(a, j, i)])) +
                  ((hb.ija_w[hb.i3(b, j, i)] / hb.ija_tot[hb.i2(j, i)]) *
                   (hb.ijab_w[hb.i4(a, b, j, i)] / hb.ijab_tot[hb.i3(b, j, i)])));
    } else {
        r.conn.nexcit = 2;
        r.hmatel = 0.0; r.pgen = 1.0;
    }
    return r;
}

// decode_det_spinocc_spinsymunocc (src/determinant_decoders.f90:371-495)
inline void decode_det_spinocc_spinsymunocc(const System& sys, const Det& f, DetInfo& d) {
    decode_det_occ_symunocc(sys, f, d);
    d.unocc_alpha.clear(); d.unocc_beta.clear();
    for (int o = 1; o <= sys.nbasis; ++o)
        if (!det_test(f, o)) ((o % 2 == 1) ? d.unocc_alpha : d.unocc_beta).push_back(o);
}
// create_weighted_excitation_list_mol (src/hamiltonian_molecular.f90:348-390): weights sqrt|<ia|ai>| (Power-Pitzer) or
// sqrt|<ia|ia>| (Cauchy-Schwarz; get_two_body_int_cou_mol_real, src/qmc.F90:437-441) over a_list, zero for a == b
inline double create_weighted_excitation_list_mol(const System& sys, bool cauchy_schwarz, int i, int b, const int* a_list,
                                                  int n, double* weights) {
    double tot = 0.0;
    for (int k = 0; k < n; ++k) {
        if (a_list[k] != b) {
            const int a = a_list[k];
            const double w = cauchy_schwarz ? sys.get_two_body_real(i, a, i, a) : sys.get_two_body_real(i, a, a, i);
            weights[k] = std::sqrt(std::fabs(w));
            tot = tot + weights[k];
        } else {
            weights[k] = 0.0;
        }
    }
    return tot;
}
// gen_excit_mol_power_pitzer_occ (src/excit_gen_power_pitzer_mol.F90:1260-1549): ij selected uniformly
// (excit_gen_power_pitzer_occ, excit_gen_cauchy_schwarz_occ) or with the heat-bath-like weights ppm_i_d_weights /
// ppm_ij_d_weights (the _occ_ij variants).  init_excit_mol_power_pitzer_orderM_ij (:585-647) builds those tables with
// init_double_weights_ab (src/excit_gen_utils.f90:68-140): the same sums, in the same order, as i_weights / ij_weights
// of init_excit_mol_heat_bath, so the heat-bath tables are used for them.
inline GenResult gen_excit_mol_power_pitzer_occ(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    GenResult r;
    const bool cs = eg.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC || eg.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ;
    const bool weighted_ij = eg.excit_gen == EXCIT_GEN_POWER_PITZER_OCC_IJ || eg.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ;
    if (rng.next() < eg.pattempt_single) {
        choose_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
        r.conn.nexcit = 1;
        if (r.allowed) {
            r.pgen = eg.pattempt_single * calc_pgen_single_mol(sys, sys.gamma_sym, d, r.conn.to_orb[0]);
            sys.find_excitation_permutation1(d.f, r.conn);
            r.hmatel = sys.slater_condon1_excit(d.occ, r.conn.from_orb[0], r.conn.to_orb[0], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
        return r;
    }
    int i, j, ij_sym, ij_spin;
    double pgen_ij;
    bool allowed = true, a_found = false;
    if (weighted_ij) {
        const HeatBath& hb = eg.hb;
        const int nel = sys.nel;
        double ij_w[MAXNEL], ji_w[MAXNEL], ij_tot = 0.0, ji_tot = 0.0;
        if (!d.double_precalc) {
            d.i_d_weights_tot = 0.0;
            for (int p = 0; p < nel; ++p) {
                d.i_d_weights[p] = hb.i_weights[d.occ[p] - 1];
                d.i_d_weights_tot = d.i_d_weights_tot + d.i_d_weights[p];
            }
            d.double_precalc = true;
        }
        int i_ind = select_weighted_value(rng, nel, d.i_d_weights, d.i_d_weights_tot);
        i = d.occ[i_ind - 1];
        int j_ind = 1;
        j = i;
        for (int p = 0; p < nel; ++p) { ij_w[p] = hb.ij_weights[hb.i2(d.occ[p], i)]; ij_tot = ij_tot + ij_w[p]; }
        if (ij_tot > 0.0) {
            j_ind = select_weighted_value(rng, nel, ij_w, ij_tot);
            j = d.occ[j_ind - 1];
            for (int p = 0; p < nel; ++p) { ji_w[p] = hb.ij_weights[hb.i2(d.occ[p], j)]; ji_tot = ji_tot + ji_w[p]; }
            allowed = true;
        } else {
            allowed = false;
        }
        pgen_ij = 0.0;
        ij_spin = 0; ij_sym = 0;
        if (allowed) {
            ij_spin = sys.bf[i].ms + sys.bf[j].ms;
            ij_sym = sys.sym_conj(sys.cross_product_basis(i, j));
            pgen_ij = ((d.i_d_weights[i_ind - 1] / d.i_d_weights_tot) * (ij_w[j_ind - 1] / ij_tot)) +
                      ((d.i_d_weights[j_ind - 1] / d.i_d_weights_tot) * (ji_w[i_ind - 1] / ji_tot));
            if (j < i) std::swap(i, j);
        }
    } else {
        choose_ij_mol(rng, sys, d.occ, i, j, ij_sym, ij_spin, pgen_ij);
    }
    const std::vector<int>& ilist = (sys.bf[i].ms < 0) ? d.unocc_beta : d.unocc_alpha;
    const int ni = (int)ilist.size();
    std::vector<double> ia_w(std::max(ni, 1)), jb_w, ja_w;
    double ia_tot = 0.0, jb_tot = 0.0, ja_tot = 0.0;
    int a = 0, b = 0, a_ind = 0, b_ind = 0;
    if (ni > 0 && allowed) {
        ia_tot = create_weighted_excitation_list_mol(sys, cs, i, 0, ilist.data(), ni, ia_w.data());
        if (ia_tot > 0.0) {
            a_ind = select_weighted_value(rng, ni, ia_w.data(), ia_tot);
            a = ilist[a_ind - 1];
            a_found = true;
        }
    }
    int isymb = 0, imsb = 0, nb_list = 0;
    const int* blist = nullptr;
    if (a_found && allowed) {
        isymb = sys.sym_conj(sys.cross_product(ij_sym, sys.bf[a].sym));
        imsb = (ij_spin - sys.bf[a].ms + 3) / 2;
        nb_list = sys.nbss(imsb, isymb);
        if (nb_list > 0) {
            blist = &sys.sym_spin_basis_fns[(size_t)sys.max_nbss * ((imsb - 1) + 2 * isymb)];
            jb_w.resize(nb_list);
            jb_tot = create_weighted_excitation_list_mol(sys, cs, j, a, blist, nb_list, jb_w.data());
        } else {
            jb_tot = 0.0;
        }
    }
    r.allowed = false;
    if (a_found && jb_tot > 0.0 && allowed) {
        b_ind = select_weighted_value(rng, nb_list, jb_w.data(), jb_tot);
        b = blist[b_ind - 1];
        if (!det_test(d.f, b)) {
            double pgen;
            if (ij_spin == 0) {
                pgen = ia_w[a_ind - 1] / ia_tot * jb_w[b_ind - 1] / jb_tot;
            } else {
                const std::vector<int>& ul = (imsb == 1) ? d.unocc_beta : d.unocc_alpha;
                const int b_ind_rev = (int)(std::lower_bound(ul.begin(), ul.end(), b) - ul.begin()) + 1;
                const int isyma = sys.sym_conj(sys.cross_product(ij_sym, isymb));
                const int na_list = sys.nbss(imsb, isyma);
                const int* alist = &sys.sym_spin_basis_fns[(size_t)sys.max_nbss * ((imsb - 1) + 2 * isyma)];
                ja_w.resize(na_list);
                ja_tot = create_weighted_excitation_list_mol(sys, cs, j, b, alist, na_list, ja_w.data());
                const int a_ind_rev = (int)(std::lower_bound(alist, alist + na_list, a) - alist) + 1;
                pgen = (ia_w[a_ind - 1] * jb_w[b_ind - 1]) / (ia_tot * jb_tot) +
                       (ia_w[b_ind_rev - 1] * ja_w[a_ind_rev - 1]) / (ia_tot * ja_tot);
            }
            r.pgen = eg.pattempt_double * pgen * pgen_ij;
            r.conn.nexcit = 2;
            r.conn.from_orb[0] = std::min(i, j); r.conn.from_orb[1] = std::max(i, j);
            r.conn.to_orb[0] = std::min(a, b); r.conn.to_orb[1] = std::max(a, b);
            r.allowed = true;
        }
    }
    if (r.allowed) {
        sys.find_excitation_permutation2(d.f, r.conn);
        r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0], r.conn.to_orb[1],
                                            r.conn.perm);
    } else {
        r.conn.nexcit = 2;
        r.hmatel = 0.0; r.pgen = 1.0;
    }
    return r;
}

// ------------------------------------------------------------------------ power_pitzer_orderN
// check_min_weight_ratio (src/excit_gen_power_pitzer_mol.F90:140-213): raise the non-zero weights below
// min_ratio/(number of non-zero weights) of the total to that floor
inline void check_min_weight_ratio(double* weights, double& weights_tot, int n, double min_ratio) {
    double min_weight_tmp = 0.0;
    int nonzero = 0;
    if (!(weights_tot > 0.0 && min_ratio > 0.0)) return;
    for (int i = 0; i < n; ++i) if (weights[i] > 0.0) nonzero++;
    double min_weight = (min_ratio / (float)nonzero) * weights_tot;
    while (std::fabs(min_weight_tmp - min_weight) > depsilon) {
        min_weight_tmp = min_weight;
        double keep = 0.0;
        int cnt = 0;
        for (int k = 0; k < n; ++k) {
            if (weights[k] > 0.0 && weights[k] < min_weight) cnt++;
            else keep = keep + weights[k];
        }
        if (cnt == nonzero) break;
        min_weight = (min_ratio / (float)nonzero) * (keep / (1 - (min_ratio * (float)cnt / (float)nonzero)));
    }
    weights_tot = 0.0;
    for (int j = 0; j < n; ++j) {
        if (weights[j] > 0.0 && weights[j] < min_weight) weights[j] = min_weight;
        weights_tot = weights_tot + weights[j];
    }
}
// single_excitation_weight_mol (src/hamiltonian_molecular.f90:444-523)
inline double single_excitation_weight_mol(const System& sys, const int* occ0, int i, int a) {
    const int nel = sys.nel, nvirt = sys.nbasis - sys.nel;
    std::vector<int> virt(nvirt);
    int vp = 0, op = 0;
    for (int pos = 1; pos <= sys.nbasis; ++pos) {
        if (op >= nel) virt[vp++] = pos;
        else if (occ0[op] == pos) op++;
        else virt[vp++] = pos;
    }
    int n_jb = 0;
    double weight = 0.0;
    for (int j = 0; j < nel; ++j)
        for (int b = 0; b < nvirt; ++b) {
            n_jb++;
            weight = weight + std::fabs(sys.get_two_body_real(i, occ0[j], occ0[j], a) - sys.get_two_body_real(i, occ0[j], a, occ0[j]) +
                                        sys.get_two_body_real(i, virt[b], a, virt[b]) - sys.get_two_body_real(i, virt[b], virt[b], a));
        }
    return weight / (double)n_jb;
}
// init_double_weights_ab (src/excit_gen_utils.f90:68-140)
inline void init_double_weights_ab(const System& sys, int i, int j, double& weight) {
    const int it = std::min(i, j), jt = std::max(i, j);
    const int ij_sym = sys.sym_conj(sys.cross_product_basis(it, jt));
    for (int a = 1; a <= sys.nbasis; ++a) {
        if (a == it || a == jt) continue;
        const int isymb = sys.sym_conj(sys.cross_product(ij_sym, sys.bf[a].sym));
        for (int b = 1; b <= sys.nbasis; ++b) {
            const bool spin_ok = (sys.bf[it].ms == sys.bf[a].ms && sys.bf[jt].ms == sys.bf[b].ms) ||
                                 (sys.bf[it].ms == sys.bf[b].ms && sys.bf[jt].ms == sys.bf[a].ms);
            if (spin_ok && sys.bf[b].sym == isymb && a != b && b != it && b != jt)
                weight = weight + std::fabs(sys.slater_condon2_excit(it, jt, std::min(a, b), std::max(a, b), false));
        }
    }
}
// init_excit_mol_power_pitzer_orderN (src/excit_gen_power_pitzer_mol.F90:215-572).  Every table entry is computed by one
// rank and gathered, so the result does not depend on the number of ranks.
inline void init_excit_mol_power_pitzer_orderN(const System& sys, const int* occ0, PowerPitzerN& pp) {
    const int nel = sys.nel, nb = sys.nbasis, mv = sys.max_nbss, nsym = sys.nsym_tot;
    pp.occ_list.assign(occ0, occ0 + nel);
    std::sort(pp.occ_list.begin(), pp.occ_list.end());
    pp.all_list_alpha.clear(); pp.all_list_beta.clear();
    for (int i = 1; i <= nb; ++i) (sys.bf[i].ms == -1 ? pp.all_list_beta : pp.all_list_alpha).push_back(i);
    pp.n_all_alpha = (int)pp.all_list_alpha.size(); pp.n_all_beta = (int)pp.all_list_beta.size();
    const int nall_max = std::max(pp.n_all_alpha, pp.n_all_beta);
    pp.i_s.alloc(nel, 1); pp.i_d.alloc(nel, 1);
    pp.ia_s.alloc(mv, nb + 1); pp.ij_d.alloc(nel, nb + 1); pp.ia_d.alloc(nall_max, nb + 1);
    pp.jb_d.alloc(mv, (size_t)nsym * (nb + 1));
    const int* occ = pp.occ_list.data();
    // i in a single excitation
    for (int i = 0; i < nel; ++i) {
        double i_weight = 0.0;
        const int ims = sys.bf[occ[i]].ms;
        const int isyma = sys.cross_product(sys.bf[occ[i]].sym, sys.gamma_sym);
        for (int a = 1; a <= nb; ++a)
            if (a != occ[i] && sys.bf[a].sym == isyma && sys.bf[a].ms == ims)
                i_weight = i_weight + single_excitation_weight_mol(sys, occ0, occ[i], a);
        if (i_weight < depsilon) i_weight = 10.0 * depsilon;
        pp.i_s.w[i] = i_weight;
    }
    pp.i_s.tot[0] = 0.0;
    for (int i = 0; i < nel; ++i) pp.i_s.tot[0] += pp.i_s.w[i];
    check_min_weight_ratio(pp.i_s.w.data(), pp.i_s.tot[0], nel, pp.min_weight);
    generate_alias_tables(nel, pp.i_s.w.data(), pp.i_s.tot[0], pp.i_s.U.data(), pp.i_s.K.data());
    // a given i in a single excitation (i: any basis function)
    for (int i = 1; i <= nb; ++i) {
        double tot = 0.0;
        const int imsa = (3 + sys.bf[i].ms) / 2;
        const int isyma = sys.cross_product(sys.bf[i].sym, sys.gamma_sym);
        double* w = &pp.ia_s.w[(size_t)mv * i];
        const int n = sys.nbss(imsa, isyma);
        for (int a = 1; a <= n; ++a) {
            w[a - 1] = 0.0;
            if (sys.ssbf(a, imsa, isyma) != i) {
                w[a - 1] = single_excitation_weight_mol(sys, occ0, i, sys.ssbf(a, imsa, isyma));
                if (w[a - 1] < depsilon) w[a - 1] = 10.0 * depsilon;
            }
            tot = tot + w[a - 1];
        }
        pp.ia_s.tot[i] = tot;
        check_min_weight_ratio(w, pp.ia_s.tot[i], n, pp.min_weight);
        generate_alias_tables(n, w, pp.ia_s.tot[i], &pp.ia_s.U[(size_t)mv * i], &pp.ia_s.K[(size_t)mv * i]);
    }
    // i in a double excitation
    for (int i = 0; i < nel; ++i) {
        double i_weight = 0.0;
        for (int j = 0; j < nel; ++j)
            if (i != j) init_double_weights_ab(sys, occ[i], occ[j], i_weight);
        if (i_weight < depsilon) i_weight = 10.0 * depsilon;
        pp.i_d.w[i] = i_weight;
    }
    pp.i_d.tot[0] = 0.0;
    for (int i = 0; i < nel; ++i) pp.i_d.tot[0] += pp.i_d.w[i];
    check_min_weight_ratio(pp.i_d.w.data(), pp.i_d.tot[0], nel, pp.min_weight);
    generate_alias_tables(nel, pp.i_d.w.data(), pp.i_d.tot[0], pp.i_d.U.data(), pp.i_d.K.data());
    // j given i (i: any basis function, j: position in the reference)
    for (int i = 1; i <= nb; ++i) {
        double tot = 0.0;
        double* w = &pp.ij_d.w[(size_t)nel * i];
        for (int j = 0; j < nel; ++j) {
            double ij_weight = 0.0;
            if (occ[j] != i) init_double_weights_ab(sys, i, occ[j], ij_weight);
            if (ij_weight < depsilon) ij_weight = 10.0 * depsilon;
            w[j] = ij_weight;
            tot = tot + ij_weight;
        }
        pp.ij_d.tot[i] = tot;
        check_min_weight_ratio(w, pp.ij_d.tot[i], nel, pp.min_weight);
        generate_alias_tables(nel, w, pp.ij_d.tot[i], &pp.ij_d.U[(size_t)nel * i], &pp.ij_d.K[(size_t)nel * i]);
    }
    // a given i and b given j: sqrt|<ia|ai>| over all orbitals of the spin of i / over each symmetry class
    for (int i = 1; i <= nb; ++i) {
        const bool beta = sys.bf[i].ms == -1;
        const std::vector<int>& all = beta ? pp.all_list_beta : pp.all_list_alpha;
        const int nall = (int)all.size();
        if (nall > 0) {
            double* w = &pp.ia_d.w[(size_t)nall_max * i];
            pp.ia_d.tot[i] = create_weighted_excitation_list_mol(sys, false, i, i, all.data(), nall, w);
            generate_alias_tables(nall, w, pp.ia_d.tot[i], &pp.ia_d.U[(size_t)nall_max * i], &pp.ia_d.K[(size_t)nall_max * i]);
        }
        const int ims = beta ? 1 : 2;
        for (int bsym = 0; bsym <= sys.sym_max_tot; ++bsym) {
            const int n = sys.nbss(ims, bsym);
            if (n <= 0) continue;
            std::vector<int> list(n);
            for (int k = 1; k <= n; ++k) list[k - 1] = sys.ssbf(k, ims, bsym);
            const size_t col = (size_t)bsym + (size_t)nsym * i;
            double* w = &pp.jb_d.w[(size_t)mv * col];
            pp.jb_d.tot[col] = create_weighted_excitation_list_mol(sys, false, i, i, list.data(), n, w);
            generate_alias_tables(n, w, pp.jb_d.tot[col], &pp.jb_d.U[(size_t)mv * col], &pp.jb_d.K[(size_t)mv * col]);
        }
    }
}
// get_excitation_locations (src/excitations.F90) + find_diff_ref_cdet (src/excit_gen_utils.f90:220-269): the occupied
// orbitals of cdet in the order of the reference's (same-spin replacements for the orbitals that differ)
inline int ref_cdet_locations(const System& sys, const PowerPitzerN& pp, const DetInfo& d, int* ref_store, int* det_store) {
    const int nel = sys.nel;
    const int* ref_list = pp.occ_list.data();
    const int* det_list = d.occ;
    int j = 1, det_sind = 0, ref_sind = 0;
    bool done = false;
    for (int i = 1; i <= nel && !done; ++i) {
        while (det_list[j - 1] < ref_list[i - 1]) {
            det_store[det_sind++] = j;
            j++;
            if (j > nel) { done = true; break; }
        }
        if (done) break;
        if (det_list[j - 1] > ref_list[i - 1]) ref_store[ref_sind++] = i;
        else j++;
        if (j > nel) break;
    }
    while (j <= nel) { det_store[det_sind++] = j; j++; }
    int i_back = nel, i_back_pos = det_sind;
    while (ref_sind < det_sind) {
        ref_store[i_back_pos - 1] = i_back;
        i_back--; i_back_pos--; ref_sind++;
    }
    const int nex = ref_sind;
    // pair every differing reference orbital with a determinant orbital of the same spin
    for (int ii = 0; ii < nex; ++ii) {
        if (sys.bf[ref_list[ref_store[ii] - 1]].ms != sys.bf[det_list[det_store[ii] - 1]].ms) {
            int jj = ii + 1;
            while (sys.bf[ref_list[ref_store[ii] - 1]].ms != sys.bf[det_list[det_store[jj] - 1]].ms) jj++;
            std::swap(det_store[ii], det_store[jj]);
        }
    }
    return nex;
}
inline void find_diff_ref_cdet(const System& sys, const PowerPitzerN& pp, DetInfo& d) {
    int ref_store[MAXNEL], det_store[MAXNEL];
    const int nex = ref_cdet_locations(sys, pp, d, ref_store, det_store);
    for (int k = 0; k < sys.nel; ++k) d.ref_cdet_occ[k] = pp.occ_list[k];
    for (int ii = 0; ii < nex; ++ii) d.ref_cdet_occ[ref_store[ii] - 1] = d.occ[det_store[ii] - 1];
}
inline int binary_search_int(const int* list, int n, int item) {   // 1-based position, list ascending
    int lo = 1, hi = n;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (list[mid - 1] == item) return mid;
        if (list[mid - 1] < item) lo = mid + 1; else hi = mid - 1;
    }
    return 0;
}
// gen_excit_mol_power_pitzer_orderN (src/excit_gen_power_pitzer_mol.F90:941-1258)
inline GenResult gen_excit_mol_power_pitzer_orderN(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    GenResult r;
    const PowerPitzerN& pp = eg.ppn;
    const int nel = sys.nel, mv = sys.max_nbss, nsym = sys.nsym_tot;
    const int nall_max = std::max(pp.n_all_alpha, pp.n_all_beta);
    if (!d.single_precalc) {
        find_diff_ref_cdet(sys, pp, d);
        d.single_precalc = true; d.double_precalc = true;
    }
    if (rng.next() < eg.pattempt_single) {
        const int i_ind_ref = select_weighted_value_precalc(rng, nel, pp.i_s.U.data(), pp.i_s.K.data());
        const int i_cdet = d.ref_cdet_occ[i_ind_ref - 1];
        const int imsa = (3 + sys.bf[i_cdet].ms) / 2;
        const int isyma = sys.cross_product(sys.bf[i_cdet].sym, sys.gamma_sym);
        int a_ind = 0, a_cdet = 0;
        r.conn.nexcit = 1;
        if (sys.nbss(imsa, isyma) > 0) {
            a_ind = select_weighted_value_precalc(rng, sys.nbss(imsa, isyma), &pp.ia_s.U[(size_t)mv * i_cdet],
                                                  &pp.ia_s.K[(size_t)mv * i_cdet]);
            a_cdet = sys.ssbf(a_ind, imsa, isyma);
            r.allowed = !det_test(d.f, a_cdet);
        } else r.allowed = false;
        if (r.allowed) {
            r.pgen = (pp.i_s.w[i_ind_ref - 1] / pp.i_s.tot[0]) *
                     (pp.ia_s.w[(size_t)mv * i_cdet + a_ind - 1] / pp.ia_s.tot[i_cdet]);
            r.pgen = eg.pattempt_single * r.pgen;
            r.conn.from_orb[0] = i_cdet; r.conn.to_orb[0] = a_cdet;
            sys.find_excitation_permutation1(d.f, r.conn);
            r.hmatel = sys.slater_condon1_excit(d.occ, i_cdet, a_cdet, r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
        return r;
    }
    r.conn.nexcit = 2;
    const int i_ind_ref = select_weighted_value_precalc(rng, nel, pp.i_d.U.data(), pp.i_d.K.data());
    int i_cdet = d.ref_cdet_occ[i_ind_ref - 1];
    const int j_ind_ref = select_weighted_value_precalc(rng, nel, &pp.ij_d.U[(size_t)nel * i_cdet], &pp.ij_d.K[(size_t)nel * i_cdet]);
    int j_cdet = d.ref_cdet_occ[j_ind_ref - 1];
    double pgen = 1.0;
    int ij_spin = 0;
    if (j_cdet != i_cdet) {
        r.allowed = true;
        pgen = (pp.i_d.w[i_ind_ref - 1] / pp.i_d.tot[0]) * (pp.ij_d.w[(size_t)nel * i_cdet + j_ind_ref - 1] / pp.ij_d.tot[i_cdet]);
        ij_spin = sys.bf[i_cdet].ms + sys.bf[j_cdet].ms;
        pgen = pgen + ((pp.i_d.w[j_ind_ref - 1] / pp.i_d.tot[0]) * (pp.ij_d.w[(size_t)nel * j_cdet + i_ind_ref - 1] / pp.ij_d.tot[j_cdet]));
        if (j_cdet < i_cdet) std::swap(i_cdet, j_cdet);
    } else r.allowed = false;
    int a_ind = 0, a_cdet = 0, b_ind = 0, b_cdet = 0, ij_sym = 0, isymb = 0, imsb = 1;
    if (r.allowed) {
        if (sys.bf[i_cdet].ms < 0) {
            a_ind = select_weighted_value_precalc(rng, pp.n_all_beta, &pp.ia_d.U[(size_t)nall_max * i_cdet], &pp.ia_d.K[(size_t)nall_max * i_cdet]);
            a_cdet = pp.all_list_beta[a_ind - 1];
        } else {
            a_ind = select_weighted_value_precalc(rng, pp.n_all_alpha, &pp.ia_d.U[(size_t)nall_max * i_cdet], &pp.ia_d.K[(size_t)nall_max * i_cdet]);
            a_cdet = pp.all_list_alpha[a_ind - 1];
        }
        if (det_test(d.f, a_cdet)) r.allowed = false;
    }
    if (r.allowed) {
        ij_sym = sys.sym_conj(sys.cross_product_basis(i_cdet, j_cdet));
        isymb = sys.sym_conj(sys.cross_product(ij_sym, sys.bf[a_cdet].sym));
        imsb = (sys.bf[j_cdet].ms + 3) / 2;
        if (sys.nbss(imsb, isymb) == 0) r.allowed = false;
    }
    if (r.allowed) {
        const size_t colb = (size_t)isymb + (size_t)nsym * j_cdet;
        b_ind = select_weighted_value_precalc(rng, sys.nbss(imsb, isymb), &pp.jb_d.U[(size_t)mv * colb], &pp.jb_d.K[(size_t)mv * colb]);
        b_cdet = sys.ssbf(b_ind, imsb, isymb);
        if (a_cdet != b_cdet && !det_test(d.f, b_cdet)) {
            const double pa = pp.ia_d.w[(size_t)nall_max * i_cdet + a_ind - 1] / pp.ia_d.tot[i_cdet];
            if (ij_spin == 0) {
                pgen = pgen * (pa * pp.jb_d.w[(size_t)mv * colb + b_ind - 1] / pp.jb_d.tot[colb]);
            } else {
                const int b_rev = (imsb == 1) ? binary_search_int(pp.all_list_beta.data(), pp.n_all_beta, b_cdet)
                                              : binary_search_int(pp.all_list_alpha.data(), pp.n_all_alpha, b_cdet);
                const int isyma = sys.sym_conj(sys.cross_product(ij_sym, isymb));
                int a_rev = 0;
                for (int k = 1; k <= sys.nbss(imsb, isyma); ++k)
                    if (sys.ssbf(k, imsb, isyma) == a_cdet) { a_rev = k; break; }
                const size_t cola = (size_t)isyma + (size_t)nsym * j_cdet;
                pgen = pgen * (pa * pp.jb_d.w[(size_t)mv * colb + b_ind - 1] / pp.jb_d.tot[colb] +
                               pp.ia_d.w[(size_t)nall_max * i_cdet + b_rev - 1] / pp.ia_d.tot[i_cdet] *
                                   pp.jb_d.w[(size_t)mv * cola + a_rev - 1] / pp.jb_d.tot[cola]);
            }
            pgen = eg.pattempt_double * pgen;
        } else r.allowed = false;
    }
    if (r.allowed) {
        r.pgen = pgen;
        r.conn.from_orb[0] = i_cdet; r.conn.from_orb[1] = j_cdet;
        r.conn.to_orb[0] = std::min(a_cdet, b_cdet); r.conn.to_orb[1] = std::max(a_cdet, b_cdet);
        sys.find_excitation_permutation2(d.f, r.conn);
        r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0], r.conn.to_orb[1], r.conn.perm);
    } else { r.hmatel = 0.0; r.pgen = 1.0; }
    return r;
}

// ------------------------------------------------------------------------ power_pitzer (reference-mapped, O(N))
// init_excit_mol_power_pitzer_occ_ref (src/excit_gen_power_pitzer_mol.F90:19-138)
inline void init_excit_mol_power_pitzer_occ_ref(const System& sys, const int* occ0, PowerPitzerN& pp) {
    const int nel = sys.nel, nb = sys.nbasis, mv = sys.max_nbss, nsym = sys.nsym_tot;
    const int maxv = std::max(sys.nvirt_alpha, sys.nvirt_beta);
    pp.occ_list.assign(occ0, occ0 + nel);
    std::sort(pp.occ_list.begin(), pp.occ_list.end());
    pp.virt_list_alpha.clear(); pp.virt_list_beta.clear();
    int j = 0;
    for (int i = 1; i <= nb; ++i) {
        if (i == pp.occ_list[j]) { if (j < nel - 1) j++; }
        else (sys.bf[i].ms == -1 ? pp.virt_list_beta : pp.virt_list_alpha).push_back(i);
    }
    pp.pp_ia_d.alloc(std::max(maxv, 1), nel);
    pp.pp_jb_d.alloc(mv, (size_t)nsym * nel);
    for (int i = 0; i < nel; ++i) {
        const int oj = pp.occ_list[i];
        const bool beta = sys.bf[oj].ms == -1;
        const std::vector<int>& virt = beta ? pp.virt_list_beta : pp.virt_list_alpha;
        const int nv = (int)virt.size();
        if (nv > 0) {
            double* w = &pp.pp_ia_d.w[(size_t)pp.pp_ia_d.stride * i];
            pp.pp_ia_d.tot[i] = create_weighted_excitation_list_mol(sys, false, oj, 0, virt.data(), nv, w);
            check_min_weight_ratio(w, pp.pp_ia_d.tot[i], nv, pp.min_weight);
            generate_alias_tables(nv, w, pp.pp_ia_d.tot[i], &pp.pp_ia_d.U[(size_t)pp.pp_ia_d.stride * i],
                                  &pp.pp_ia_d.K[(size_t)pp.pp_ia_d.stride * i]);
        }
        const int ims = beta ? 1 : 2;
        for (int bsym = 0; bsym <= sys.sym_max_tot; ++bsym) {
            const int n = sys.nbss(ims, bsym);
            if (n <= 0) continue;
            std::vector<int> list(n);
            for (int k = 1; k <= n; ++k) list[k - 1] = sys.ssbf(k, ims, bsym);
            const size_t col = (size_t)bsym + (size_t)nsym * i;
            double* w = &pp.pp_jb_d.w[(size_t)mv * col];
            pp.pp_jb_d.tot[col] = create_weighted_excitation_list_mol(sys, false, oj, 0, list.data(), n, w);
            check_min_weight_ratio(w, pp.pp_jb_d.tot[col], n, pp.min_weight);
            generate_alias_tables(n, w, pp.pp_jb_d.tot[col], &pp.pp_jb_d.U[(size_t)mv * col], &pp.pp_jb_d.K[(size_t)mv * col]);
        }
    }
}
// gen_excit_mol_power_pitzer_occ_ref (src/excit_gen_power_pitzer_mol.F90:650-939): ij uniform among the REFERENCE's
// occupied orbitals, a and b from the reference's alias tables, all three mapped onto the determinant
inline GenResult gen_excit_mol_power_pitzer_occ_ref(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    GenResult r;
    const PowerPitzerN& pp = eg.ppn;
    const int nel = sys.nel, mv = sys.max_nbss, nsym = sys.nsym_tot;
    if (rng.next() < eg.pattempt_single) {   // gen_single_excit_mol_no_renorm
        find_ia_mol(rng, sys, sys.gamma_sym, d, r.conn.from_orb[0], r.conn.to_orb[0], r.allowed);
        r.conn.nexcit = 1;
        if (r.allowed) {
            r.pgen = eg.pattempt_single * calc_pgen_single_mol_no_renorm(sys, r.conn.to_orb[0]);
            sys.find_excitation_permutation1(d.f, r.conn);
            r.hmatel = sys.slater_condon1_excit(d.occ, r.conn.from_orb[0], r.conn.to_orb[0], r.conn.perm);
        } else { r.hmatel = 0.0; r.pgen = 1.0; }
        return r;
    }
    r.conn.nexcit = 2;
    // choose_ij_mol on the reference's list
    const int ind = (int)(rng.next() * nel * (nel - 1) / 2) + 1;
    const int j_ind_ref = (int)(1.50 + std::sqrt(2 * ind - 1.750));
    const int i_ind_ref = ind - ((j_ind_ref - 1) * (j_ind_ref - 2)) / 2;
    const double pgen_ij = 2.0 / (nel * (nel - 1));
    const int i_ref = pp.occ_list[i_ind_ref - 1], j_ref = pp.occ_list[j_ind_ref - 1];
    const int ij_spin = sys.bf[i_ref].ms + sys.bf[j_ref].ms;
    const std::vector<int>& virt = (sys.bf[i_ref].ms < 0) ? pp.virt_list_beta : pp.virt_list_alpha;
    const int nv = (int)virt.size();
    bool a_found = false;
    int a_ind_ref = 0, a_ref = 0;
    const size_t sia = (size_t)pp.pp_ia_d.stride;
    if (nv > 0) {
        a_ind_ref = select_weighted_value_precalc(rng, nv, &pp.pp_ia_d.U[sia * (i_ind_ref - 1)], &pp.pp_ia_d.K[sia * (i_ind_ref - 1)]);
        a_ref = virt[a_ind_ref - 1];
        a_found = true;
    }
    int ref_store[MAXNEL], det_store[MAXNEL], nex = 0;
    int i_cdet = i_ref, j_cdet = j_ref, a_cdet = a_ref, ij_sym = 0, isymb = 0, imsb = 1;
    if (a_found) {
        nex = ref_cdet_locations(sys, pp, d, ref_store, det_store);
        for (int ii = 0; ii < nex; ++ii) {
            if (ref_store[ii] == i_ind_ref) i_cdet = d.occ[det_store[ii] - 1];
            else if (ref_store[ii] == j_ind_ref) j_cdet = d.occ[det_store[ii] - 1];
            if (d.occ[det_store[ii] - 1] == a_ref) a_cdet = pp.occ_list[ref_store[ii] - 1];
        }
        ij_sym = sys.sym_conj(sys.cross_product_basis(i_cdet, j_cdet));
        isymb = sys.sym_conj(sys.cross_product(ij_sym, sys.bf[a_cdet].sym));
        imsb = (sys.bf[j_ref].ms + 3) / 2;
    }
    r.allowed = false;
    int b_cdet = 0;
    if (a_found && sys.nbss(imsb, isymb) > 0) {
        const size_t colb = (size_t)isymb + (size_t)nsym * (j_ind_ref - 1);
        const int b_ind = select_weighted_value_precalc(rng, sys.nbss(imsb, isymb), &pp.pp_jb_d.U[(size_t)mv * colb], &pp.pp_jb_d.K[(size_t)mv * colb]);
        b_cdet = sys.ssbf(b_ind, imsb, isymb);
        if (a_cdet != b_cdet && !det_test(d.f, b_cdet)) {
            double pgen;
            const double pa = pp.pp_ia_d.w[sia * (i_ind_ref - 1) + a_ind_ref - 1] / pp.pp_ia_d.tot[i_ind_ref - 1];
            if (ij_spin == 0) {
                pgen = pa * pp.pp_jb_d.w[(size_t)mv * colb + b_ind - 1] / pp.pp_jb_d.tot[colb];
            } else {
                int b_ref = b_cdet;
                for (int ii = 0; ii < nex; ++ii)
                    if (pp.occ_list[ref_store[ii] - 1] == b_cdet) { b_ref = d.occ[det_store[ii] - 1]; break; }
                const int b_rev = binary_search_int(virt.data(), nv, b_ref);
                const int isyma = sys.sym_conj(sys.cross_product(ij_sym, isymb));
                int a_rev = 0;
                for (int k = 1; k <= sys.nbss(imsb, isyma); ++k)
                    if (sys.ssbf(k, imsb, isyma) == a_cdet) { a_rev = k; break; }
                const size_t cola = (size_t)isyma + (size_t)nsym * (j_ind_ref - 1);
                pgen = pa * pp.pp_jb_d.w[(size_t)mv * colb + b_ind - 1] / pp.pp_jb_d.tot[colb] +
                       pp.pp_ia_d.w[sia * (i_ind_ref - 1) + b_rev - 1] / pp.pp_ia_d.tot[i_ind_ref - 1] *
                           pp.pp_jb_d.w[(size_t)mv * cola + a_rev - 1] / pp.pp_jb_d.tot[cola];
            }
            r.pgen = eg.pattempt_double * pgen * pgen_ij;
            r.allowed = true;
        }
    }
    if (r.allowed) {
        r.conn.from_orb[0] = std::min(i_cdet, j_cdet); r.conn.from_orb[1] = std::max(i_cdet, j_cdet);
        r.conn.to_orb[0] = std::min(a_cdet, b_cdet); r.conn.to_orb[1] = std::max(a_cdet, b_cdet);
        sys.find_excitation_permutation2(d.f, r.conn);
        r.hmatel = sys.slater_condon2_excit(r.conn.from_orb[0], r.conn.from_orb[1], r.conn.to_orb[0], r.conn.to_orb[1], r.conn.perm);
    } else { r.hmatel = 0.0; r.pgen = 1.0; }
    return r;
}

inline GenResult gen_excit(Rng& rng, const System& sys, const ExcitGenData& eg, DetInfo& d) {
    switch (eg.excit_gen) {
        case EXCIT_GEN_RENORM: return gen_excit_mol(rng, sys, eg, d);
        case EXCIT_GEN_NO_RENORM: return gen_excit_mol_no_renorm(rng, sys, eg, d);
        case EXCIT_GEN_RENORM_SPIN: return gen_excit_mol_spin(rng, sys, eg, d, true);
        case EXCIT_GEN_NO_RENORM_SPIN: return gen_excit_mol_spin(rng, sys, eg, d, false);
        case EXCIT_GEN_HEAT_BATH: return gen_excit_mol_heat_bath(rng, sys, eg, d);
        case EXCIT_GEN_POWER_PITZER_OCC:
        case EXCIT_GEN_POWER_PITZER_OCC_IJ:
        case EXCIT_GEN_CAUCHY_SCHWARZ_OCC:
        case EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ: return gen_excit_mol_power_pitzer_occ(rng, sys, eg, d);
        case EXCIT_GEN_HEAT_BATH_UNIFORM:
        case EXCIT_GEN_HEAT_BATH_SINGLE: return gen_excit_mol_heat_bath_uniform(rng, sys, eg, d);
        case EXCIT_GEN_POWER_PITZER_ORDERN: return gen_excit_mol_power_pitzer_orderN(rng, sys, eg, d);
        case EXCIT_GEN_POWER_PITZER: return gen_excit_mol_power_pitzer_occ_ref(rng, sys, eg, d);
        default: throw std::runtime_error("oracle: excitation generator not implemented");
    }
}
inline void decode_for(const System& sys, const ExcitGenData& eg, const Det& f, DetInfo& d) {
    if (eg.excit_gen == EXCIT_GEN_POWER_PITZER_OCC || eg.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC ||
        eg.excit_gen == EXCIT_GEN_POWER_PITZER_OCC_IJ || eg.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ)
        decode_det_spinocc_spinsymunocc(sys, f, d);
    else if (eg.excit_gen == EXCIT_GEN_RENORM || eg.excit_gen == EXCIT_GEN_RENORM_SPIN ||
             eg.excit_gen == EXCIT_GEN_HEAT_BATH_UNIFORM) decode_det_occ_symunocc(sys, f, d);
    else decode_det_occ(sys, f, d);
}

}  // namespace oracle
