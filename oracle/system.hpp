// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the reference's system description for the FCIQMC hot path:
// FCIDUMP reader, basis ordering, Abelian point-group symmetry, integral stores,
// Slater-Condon rules and determinant bit-string utilities.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  All citations are file:line relative to the
// reference tree (hande-qmc/hande).
//
// Conventions kept from the reference so the restatement can be checked line by line:
//  * orbital indices are 1-based; odd index = alpha (ms=+1), even = beta (ms=-1)
//    (src/read_in.F90:862-920);
//  * orbital i lives in bit (i-1)%64 of word (i-1)/64 (src/basis_types.f90:134-185);
//  * ims = (ms+3)/2 : beta -> 1, alpha -> 2 (src/point_group_symmetry.f90 init_pg_symmetry).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include <fstream>
#include <sstream>

namespace oracle {

constexpr double depsilon = 1.e-12;  // lib/local/const.F90:91
#ifndef ORACLE_MAXW
#define ORACLE_MAXW 4
#endif
constexpr int MAXW = ORACLE_MAXW;    // max 64-bit words per determinant: 4 in liboracle.so, 32 in liboracle_wide.so

struct BasisFn {
    int sym = 0;            // point-group irrep (0-based, incl. Lz field if used)
    int ms = 0;             // +1 alpha / -1 beta
    int spatial_index = 0;  // 1-based spatial orbital
    int sym_index = 0;      // index among functions of the same sym
    int sym_spin_index = 0; // index among functions of the same (spin, sym)
    int lz = 0;
    double sp_eigv = 0.0;
};

struct Det {
    uint64_t w[MAXW];
    Det() { for (int i = 0; i < MAXW; ++i) w[i] = 0; }
    bool operator==(const Det& o) const {
        for (int i = 0; i < MAXW; ++i) if (w[i] != o.w[i]) return false;
        return true;
    }
    bool operator!=(const Det& o) const { return !(*this == o); }
};

// bit_str_cmp semantics (src/bit_utils.F90:452-479): unsigned compare, last word most
// significant.  Returns true if a < b in the reference's ascending list order.
inline bool det_less(const Det& a, const Det& b, int W) {
    for (int i = W - 1; i >= 0; --i) {
        if (a.w[i] < b.w[i]) return true;
        if (a.w[i] > b.w[i]) return false;
    }
    return false;
}

inline bool det_test(const Det& f, int orb /*1-based*/) {
    return (f.w[(orb - 1) >> 6] >> ((orb - 1) & 63)) & 1ull;
}
inline void det_set(Det& f, int orb) { f.w[(orb - 1) >> 6] |= (1ull << ((orb - 1) & 63)); }
inline void det_clr(Det& f, int orb) { f.w[(orb - 1) >> 6] &= ~(1ull << ((orb - 1) & 63)); }

// lib/local/utils.F90:449-480  tri_ind(i,j) = i(i-1)/2 + j  (i >= j, 1-based)
inline int64_t tri_ind(int64_t i, int64_t j) { return (i * (i - 1)) / 2 + j; }
inline int64_t tri_ind_reorder(int64_t i, int64_t j) { return i >= j ? tri_ind(i, j) : tri_ind(j, i); }

struct Excit {
    int nexcit = 0;
    int from_orb[2] = {0, 0};
    int to_orb[2] = {0, 0};
    bool perm = false;
};

enum SystemKind { SYS_READ_IN = 0, SYS_UEG = 1 };

// sys_ueg_t / ueg_basis_t (src/system.f90:238-264, src/ueg_types.f90) + excit_gen_data_t%ueg_ternary_conserve
struct UegData {
    double rs = 0.0, ecutoff = 0.0, L = 0.0;
    int kmax = 0, offset = 0, offset_inds[3] = {0, 0, 0};
    std::vector<int> l;        // wavevector of basis function o: l[3*o + d] (1-based o)
    std::vector<int> lookup;   // 1-based flat index -> alpha basis function or -1
    int tK = 0, tD = 0;        // ternary_conserve(0:W, -tK:tK, -tK:tK, -tK:tK), tD = 2*tK+1
    std::vector<uint64_t> ternary;
};

struct System {
    int kind = SYS_READ_IN;
    UegData ueg;
    // --- sizes
    int nbasis = 0, nel = 0, Ms = 0, nalpha = 0, nbeta = 0;
    int nvirt = 0, nvirt_alpha = 0, nvirt_beta = 0;
    int W = 1;            // tot_string_len
    bool uhf = false;
    int cas[2] = {-1, -1};
    std::vector<BasisFn> bf;  // 1-based (bf[0] unused)
    // --- symmetry (src/point_group_symmetry.f90:82-229)
    int pg_mask = 0, Lz_mask = 0, Lz_offset = 0, Lz_divisor = 1, gamma_sym = 0;
    int sym0 = 0, sym_max = 0, nsym = 0, nsym_tot = 1, sym_max_tot = 0;
    int symmetry = INT32_MAX;  // requested reference symmetry (huge => Aufbau)
    std::vector<int> nbasis_sym_spin;     // [(ims-1) + 2*sym]
    int max_nbss = 0;
    std::vector<int> sym_spin_basis_fns;  // [(ind-1) + max_nbss*((ims-1) + 2*sym)]
    // --- integrals
    double Ecore = 0.0;
    std::vector<std::vector<double>> one_body;  // [(spin-1)*nsym_tot + sym] -> triangular (1-based tri_ind - 1)
    std::vector<std::vector<double>> two_body;  // [channel-1] (1 RHF / 4 UHF)
    int64_t nintgrls = 0;
    int int_err = 0;

    inline int nbss(int ims, int sym) const { return nbasis_sym_spin[(ims - 1) + 2 * sym]; }
    inline int ssbf(int ind, int ims, int sym) const {
        return sym_spin_basis_fns[(ind - 1) + max_nbss * ((ims - 1) + 2 * sym)];
    }
    // src/point_group_symmetry.f90:297-316
    inline int cross_product(int si, int sj) const {
        return ((si ^ sj) & pg_mask) | ((si & Lz_mask) + (sj & Lz_mask) - Lz_offset);
    }
    // src/point_group_symmetry.f90:318-333
    inline int sym_conj(int s) const {
        return (s & pg_mask) | ((2 * Lz_offset - (s & Lz_mask)) & Lz_mask);
    }
    inline int cross_product_basis(int i, int j) const { return cross_product(bf[i].sym, bf[j].sym); }

    // ---- one-body store (src/molecular_integrals.F90:560-732)
    inline double& one_body_ref(int i, int j) {
        int spin = uhf ? (bf[i].ms + 3) / 2 : 1;
        int ii = bf[i].sym_spin_index, jj = bf[j].sym_spin_index;
        if (ii >= jj) return one_body[(spin - 1) * nsym_tot + bf[i].sym][tri_ind(ii, jj) - 1];
        return one_body[(spin - 1) * nsym_tot + bf[j].sym][tri_ind(jj, ii) - 1];
    }
    inline double get_one_body_nonzero(int i, int j) const {
        return const_cast<System*>(this)->one_body_ref(i, j);
    }
    inline bool check_one_body_sym(int i, int j) const {
        return bf[i].sym == cross_product(gamma_sym, bf[j].sym);
    }
    inline double get_one_body_real(int i, int j) const {
        if (check_one_body_sym(i, j) && bf[i].ms == bf[j].ms) return get_one_body_nonzero(i, j);
        return 0.0;
    }
    void store_one_body(int i, int j, double x) {
        if (check_one_body_sym(i, j) && bf[i].ms == bf[j].ms) {
            one_body_ref(i, j) = x;
        } else if (std::fabs(x) > depsilon) {
            int_err++;
        }
    }
    // ---- two-body store (src/molecular_integrals.F90:736-845)
    inline void two_body_indx(int i, int j, int a, int b, int& chan, int64_t& indx) const {
        int ii, jj, aa, bb;
        if (i < a) { ii = a; aa = i; } else { ii = i; aa = a; }
        if (j < b) { jj = b; bb = j; } else { jj = j; bb = b; }
        int64_t ia = tri_ind(bf[ii].spatial_index, bf[aa].spatial_index);
        int64_t jb = tri_ind(bf[jj].spatial_index, bf[bb].spatial_index);
        indx = (ia < jb) ? tri_ind(jb, ia) : tri_ind(ia, jb);
        if (uhf) {
            if (ia < jb || (ia == jb && ii < jj)) { int t = ii; ii = jj; jj = t; }
            if (bf[ii].ms == -1) chan = (bf[jj].ms == -1) ? 1 : 3;
            else chan = (bf[jj].ms == 1) ? 2 : 4;
        } else {
            chan = 1;
        }
    }
    inline double get_two_body_nonzero(int i, int j, int a, int b) const {
        int chan; int64_t indx;
        two_body_indx(i, j, a, b, chan, indx);
        return two_body[chan - 1][indx - 1];
    }
    inline bool check_two_body_sym(int i, int j, int a, int b) const {
        int sij = cross_product_basis(i, j), sab = cross_product_basis(a, b);
        return sij == cross_product(sab, gamma_sym);
    }
    inline double get_two_body_real(int i, int j, int a, int b) const {
        if (check_two_body_sym(i, j, a, b) && bf[i].ms == bf[a].ms && bf[j].ms == bf[b].ms)
            return get_two_body_nonzero(i, j, a, b);
        return 0.0;
    }
    void store_two_body(int i, int j, int a, int b, double x) {
        if (check_two_body_sym(i, j, a, b) && bf[i].ms == bf[a].ms && bf[j].ms == bf[b].ms) {
            int chan; int64_t indx;
            two_body_indx(i, j, a, b, chan, indx);
            two_body[chan - 1][indx - 1] = x;
        } else if (std::fabs(x) > depsilon) {
            int_err++;
        }
    }

    // ---- determinants
    // decode_det (src/determinants.f90:243-297): ascending orbital order.
    inline void decode(const Det& f, int* occ) const {
        int n = 0;
        for (int iw = 0; iw < W; ++iw) {
            uint64_t x = f.w[iw];
            while (x) {
                int b = __builtin_ctzll(x);
                occ[n++] = iw * 64 + b + 1;
                x &= x - 1;
            }
        }
    }
    inline Det encode(const int* occ, int n) const {
        Det f;
        for (int i = 0; i < n; ++i) det_set(f, occ[i]);
        return f;
    }
    int symmetry_orb_list(const int* occ, int n) const {
        int s = gamma_sym;
        for (int i = 0; i < n; ++i) s = cross_product(s, bf[occ[i]].sym);
        return s;
    }

    // ---- Slater-Condon (src/hamiltonian_molecular.f90)
    // :99-139
    double slater_condon0_orb_list(const int* occ) const {
        double h = Ecore;
        for (int iel = 0; iel < nel; ++iel) {
            int i = occ[iel];
            h = h + get_one_body_nonzero(i, i);
            for (int jel = iel + 1; jel < nel; ++jel) {
                int j = occ[jel];
                h = h + get_two_body_nonzero(i, j, i, j);
                if (bf[i].ms == bf[j].ms) h = h - get_two_body_nonzero(i, j, j, i);
            }
        }
        return h;
    }
    double slater_condon0(const Det& f) const {
        int occ[256];
        decode(f, occ);
        return slater_condon0_orb_list(occ);
    }
    // :199-259
    double slater_condon1_excit(const int* occ, int i, int a, bool perm) const {
        double h = get_one_body_nonzero(i, a);
        for (int iel = 0; iel < nel; ++iel) {
            int j = occ[iel];
            if (j != i) {
                h = h + get_two_body_nonzero(i, j, a, j);
                if (bf[j].ms == bf[i].ms) h = h - get_two_body_nonzero(i, j, j, a);
            }
        }
        return perm ? -h : h;
    }
    // :141-197 (symmetry/spin-checked variant)
    double slater_condon1(const int* occ, int i, int a, bool perm) const {
        if (bf[i].sym != bf[a].sym) return 0.0;
        double h = get_one_body_real(i, a);
        for (int iel = 0; iel < nel; ++iel) {
            int j = occ[iel];
            if (j != i) {
                h = h + get_two_body_real(i, j, a, j);
                h = h - get_two_body_real(i, j, j, a);
            }
        }
        return perm ? -h : h;
    }
    // :300-346
    double slater_condon2_excit(int i, int j, int a, int b, bool perm) const {
        double h = 0.0;
        if (bf[i].ms == bf[a].ms) h = get_two_body_nonzero(i, j, a, b);
        if (bf[i].ms == bf[b].ms) h = h - get_two_body_nonzero(i, j, b, a);
        return perm ? -h : h;
    }
    // :261-298
    double slater_condon2(int i, int j, int a, int b, bool perm) const {
        double h = get_two_body_real(i, j, a, b) - get_two_body_real(i, j, b, a);
        return perm ? -h : h;
    }

    // ---- excitations (src/excitations.F90)
    inline int popcount_above_masked(const Det& f, int orb_lo, int orb_hi) const {
        // number of set bits of f strictly between positions of orb_lo and orb_hi is not what
        // the reference computes; see find_excitation_permutation* below which restate the masks.
        (void)f; (void)orb_lo; (void)orb_hi; return 0;
    }
    // excit_mask(:,orb) = all bits strictly above orb (src/excitations.F90:25-58)
    inline void excit_mask(int orb, uint64_t* m) const {
        int iw = (orb - 1) >> 6, ib = (orb - 1) & 63;
        for (int k = 0; k < W; ++k) {
            if (k < iw) m[k] = 0;
            else if (k > iw) m[k] = ~0ull;
            else m[k] = (ib == 63) ? 0ull : (~0ull << (ib + 1));
        }
        // bits beyond nbasis in the last word are never set in f, so leaving them set in the
        // mask is harmless (the reference only sets bits of existing basis functions).
    }
    // :247-280
    void find_excitation_permutation1(const Det& f, Excit& e) const {
        uint64_t mi[MAXW], ma[MAXW];
        excit_mask(e.from_orb[0], mi);
        excit_mask(e.to_orb[0], ma);
        int perm = 0;
        for (int k = 0; k < W; ++k) perm += __builtin_popcountll(f.w[k] & (mi[k] ^ ma[k]));
        if (e.from_orb[0] > e.to_orb[0]) perm -= 1;
        e.perm = ((perm % 2) + 2) % 2 == 1;  // mod(perm,2)==1; perm>=0 whenever the -1 applies
    }
    // :282-363
    void find_excitation_permutation2(const Det& f, Excit& e) const {
        uint64_t mi[MAXW], ma[MAXW], mj[MAXW], mb[MAXW];
        excit_mask(e.from_orb[0], mi);
        excit_mask(e.to_orb[0], ma);
        excit_mask(e.from_orb[1], mj);
        excit_mask(e.to_orb[1], mb);
        int perm = 0;
        for (int k = 0; k < W; ++k) {
            uint64_t ia = mi[k] ^ ma[k], jb = mj[k] ^ mb[k];
            perm += __builtin_popcountll((f.w[k] & ia) ^ (f.w[k] & jb));
        }
        if (e.from_orb[0] > e.to_orb[0]) perm += 1;
        if (e.from_orb[0] > e.to_orb[1]) perm += 1;
        if (e.from_orb[1] > e.to_orb[1] || e.from_orb[1] < e.to_orb[0]) perm += 1;
        e.perm = (perm % 2) == 1;
    }
    // :365-406
    Det create_excited_det(const Det& f, const Excit& e) const {
        Det g = f;
        for (int k = 0; k < e.nexcit; ++k) { det_clr(g, e.from_orb[k]); det_set(g, e.to_orb[k]); }
        return g;
    }
    // get_excitation (src/excitations.F90:75-200): f1 -> f2 (f1 = from, f2 = to)
    Excit get_excitation(const Det& f1, const Det& f2) const {
        Excit ex;
        bool same = true;
        for (int k = 0; k < W; ++k) if (f1.w[k] != f2.w[k]) same = false;
        if (same) return ex;
        int nx = 0;
        for (int k = 0; k < W; ++k) nx += __builtin_popcountll(f1.w[k] ^ f2.w[k]);
        ex.nexcit = nx / 2;
        int shift = nel - ex.nexcit;
        if (ex.nexcit <= 2) {
            int iexcit1 = 0, iexcit2 = 0, iel1 = 0, iel2 = 0, perm = 0;
            for (int i = 0; i < W; ++i) {
                if (f1.w[i] == f2.w[i]) {
                    if ((((iexcit1 - iexcit2) % 2) + 2) % 2 == 1) {
                        int n = __builtin_popcountll(f1.w[i]);
                        iel1 += n; iel2 += n;
                    }
                    continue;
                }
                for (int j = 0; j < 64; ++j) {
                    bool t1 = (f1.w[i] >> j) & 1ull, t2 = (f2.w[i] >> j) & 1ull;
                    if (t2) iel2++;
                    if (t1) {
                        iel1++;
                        if (!t2) {
                            iexcit1++;
                            ex.from_orb[iexcit1 - 1] = i * 64 + j + 1;
                            perm += (shift - iel1 + iexcit1);
                        }
                    } else if (t2) {
                        iexcit2++;
                        ex.to_orb[iexcit2 - 1] = i * 64 + j + 1;
                        perm += (shift - iel2 + iexcit2);
                    }
                }
            }
            ex.perm = (((perm % 2) + 2) % 2) == 1;
        }
        return ex;
    }
    int excitation_level(const Det& f1, const Det& f2) const {
        int nx = 0;
        for (int k = 0; k < W; ++k) nx += __builtin_popcountll(f1.w[k] ^ f2.w[k]);
        return nx / 2;
    }
};

// ---------------------------------------------------------------------------------------
// FCIDUMP reader (src/read_in.F90:12-860).  Real orbitals, RHF or UHF, optional CAS.
// ---------------------------------------------------------------------------------------
struct FcidumpHeader {
    int norb = 0, nelec = 0, ms2 = INT32_MAX, isym = 0;
    bool uhf = false;
    std::vector<int64_t> orbsym;
    std::vector<int> syml, symlz;
    size_t body_offset = 0;
};

inline std::string upper(std::string s) { for (auto& c : s) c = toupper(c); return s; }

inline FcidumpHeader parse_fcidump_header(const std::string& text) {
    // Fortran namelist &FCI ... &END (or '/').  Tolerant tokenizer.
    FcidumpHeader h;
    size_t pos_end = std::string::npos;
    std::string up = upper(text.substr(0, std::min<size_t>(text.size(), 1 << 16)));
    size_t p1 = up.find("&END");
    size_t p2 = up.find("$END");
    size_t p3 = up.find("\n /");
    size_t p4 = up.find("/\n");
    for (size_t p : {p1, p2}) if (p != std::string::npos) pos_end = std::min(pos_end, p + 4);
    if (pos_end == std::string::npos) {
        for (size_t p : {p3, p4}) if (p != std::string::npos) pos_end = std::min(pos_end, p + 2);
    }
    if (pos_end == std::string::npos) throw std::runtime_error("FCIDUMP: namelist terminator not found");
    std::string nl = up.substr(0, pos_end);
    // find end of that line
    size_t eol = text.find('\n', pos_end - 1);
    h.body_offset = (eol == std::string::npos) ? text.size() : eol + 1;
    for (auto& c : nl) if (c == ',' || c == '\n' || c == '\r' || c == '\t') c = ' ';
    // insert spaces around '='
    std::string s2;
    for (char c : nl) { if (c == '=') s2 += " = "; else s2 += c; }
    std::istringstream iss(s2);
    std::vector<std::string> tok;
    std::string t;
    while (iss >> t) tok.push_back(t);
    auto is_key = [&](size_t i) { return i + 1 < tok.size() && tok[i + 1] == "="; };
    for (size_t i = 0; i < tok.size(); ++i) {
        if (!is_key(i)) continue;
        std::string key = tok[i];
        std::vector<std::string> vals;
        size_t j = i + 2;
        while (j < tok.size() && !is_key(j) && tok[j] != "&END" && tok[j] != "$END" && tok[j] != "/") {
            vals.push_back(tok[j]); ++j;
        }
        if (key == "NORB") h.norb = atoi(vals.at(0).c_str());
        else if (key == "NELEC") h.nelec = atoi(vals.at(0).c_str());
        else if (key == "MS2") h.ms2 = atoi(vals.at(0).c_str());
        else if (key == "ISYM") h.isym = atoi(vals.at(0).c_str());
        else if (key == "UHF") h.uhf = (vals.at(0).find('T') != std::string::npos);
        else if (key == "ORBSYM") for (auto& v : vals) h.orbsym.push_back(atoll(v.c_str()));
        else if (key == "SYML") for (auto& v : vals) h.syml.push_back(atoi(v.c_str()));
        else if (key == "SYMLZ") for (auto& v : vals) h.symlz.push_back(atoi(v.c_str()));
        i = j - 1;
    }
    if (h.norb == 0) throw std::runtime_error("FCIDUMP: norb not provided");
    h.orbsym.resize(1000, 0);
    h.syml.resize(1000, 0);
    h.symlz.resize(1000, 0);
    return h;
}

struct IntLine { double x; int i, a, j, b; };

inline std::vector<IntLine> parse_fcidump_body(const std::string& text, size_t off) {
    std::vector<IntLine> out;
    const char* p = text.c_str() + off;
    const char* end = text.c_str() + text.size();
    while (p < end) {
        char* q;
        // Fortran list-directed reads accept 'D' exponents; handle by strtod fallback.
        while (p < end && isspace((unsigned char)*p)) ++p;
        if (p >= end) break;
        double x = strtod(p, &q);
        if (q == p) break;
        if (*q == 'D' || *q == 'd') {
            std::string tmp(p, q - p);
            tmp += 'e';
            const char* r = q + 1;
            while (r < end && (isdigit((unsigned char)*r) || *r == '+' || *r == '-')) tmp += *r++;
            x = strtod(tmp.c_str(), nullptr);
            q = const_cast<char*>(r);
        }
        p = q;
        long v[4];
        bool ok = true;
        for (int k = 0; k < 4; ++k) {
            v[k] = strtol(p, &q, 10);
            if (q == p) { ok = false; break; }
            p = q;
        }
        if (!ok) break;
        out.push_back({x, (int)v[0], (int)v[1], (int)v[2], (int)v[3]});
    }
    return out;
}

// lib/local/ranking.f90 insertion_rank_dp (stable, with tolerance); 1-based ranks.
inline void insertion_rank(const std::vector<double>& arr /*1-based*/, int n, std::vector<int>& rank, double tol) {
    rank.assign(n + 1, 0);
    for (int i = 1; i <= n; ++i) rank[i] = i;
    for (int i = 2; i <= n; ++i) {
        int j = i - 1;
        int tmp = rank[i];
        while (j >= 1) {
            if (arr[rank[j]] - arr[tmp] < tol) break;
            rank[j + 1] = rank[j];
            --j;
        }
        rank[j + 1] = tmp;
    }
}

struct ReadInOpts {
    int nel = 0;              // 0 => from FCIDUMP
    int ms = INT32_MAX;       // huge => from FCIDUMP
    int sym = INT32_MAX;      // huge => Aufbau
    int cas[2] = {-1, -1};
};

// src/read_in.F90:862-920 (point-group, non-momentum branch)
inline void init_basis_fns_read_in(int norb, bool uhf, const std::vector<int64_t>& orbsym,
                                   const std::vector<int>& lz, const std::vector<double>& sp_eigv,
                                   const int* sp_eigv_rank /*1-based: rank[1..norb]*/, std::vector<BasisFn>& arr) {
    for (int i = 1; i <= norb; ++i) {
        int rank = sp_eigv_rank[i];
        if (uhf) {
            BasisFn& b = arr[i];
            b.sym = (int)(orbsym[rank - 1] - 1);
            b.lz = lz[rank - 1];
            b.ms = (i % 2 == 0) ? -1 : 1;
            b.spatial_index = (i + 1) / 2;
            b.sp_eigv = sp_eigv[rank];
        } else {
            for (int s = 0; s < 2; ++s) {
                BasisFn& b = arr[2 * i - 1 + s];
                b.sym = (int)(orbsym[rank - 1] - 1);
                b.lz = lz[rank - 1];
                b.ms = (s == 0) ? 1 : -1;
                b.spatial_index = i;
                b.sp_eigv = sp_eigv[rank];
            }
        }
    }
}

// src/point_group_symmetry.f90:82-229.  Lz symmetry is not used (useLz = .false.).
inline void init_pg_symmetry(System& sys) {
    int maxv = 0;
    for (int i = 1; i <= sys.nbasis; ++i) maxv = std::max(maxv, sys.bf[i].sym);
    // maxsym = 2**ceiling(log(real(maxval+1))/log(2.0)) in single precision.
    float r = std::log((float)(maxv + 1)) / std::log(2.0f);
    int maxsym = 1 << (int)std::ceil(r);
    int maxLz = 0;
    sys.pg_mask = maxsym - 1;
    sys.Lz_divisor = maxsym;
    sys.Lz_mask = ((1 << (int)std::ceil(std::log((float)(6 * maxLz + 1)) / std::log(2.0f))) - 1) * sys.Lz_divisor;
    sys.Lz_offset = 3 * maxLz * sys.Lz_divisor;
    sys.gamma_sym = sys.Lz_offset;
    if (sys.symmetry < INT32_MAX) sys.symmetry += sys.Lz_offset;
    sys.sym0 = (-maxLz * sys.Lz_divisor + sys.Lz_offset) & sys.Lz_mask;
    sys.sym_max = maxLz * sys.Lz_divisor + sys.Lz_offset + maxsym - 1;
    sys.nsym = sys.sym_max - sys.sym0;
    sys.nsym_tot = (6 * maxLz + 1) * sys.Lz_divisor;
    sys.sym_max_tot = sys.nsym_tot - 1;
    sys.nbasis_sym_spin.assign(2 * sys.nsym_tot, 0);
    std::vector<int> nbasis_sym(sys.nsym_tot, 0);
    for (int i = 1; i <= sys.nbasis; ++i) {
        BasisFn& b = sys.bf[i];
        nbasis_sym[b.sym]++;
        b.sym_index = nbasis_sym[b.sym];
        int ims = (b.ms + 3) / 2;
        sys.nbasis_sym_spin[(ims - 1) + 2 * b.sym]++;
        b.sym_spin_index = sys.nbasis_sym_spin[(ims - 1) + 2 * b.sym];
    }
    sys.max_nbss = 0;
    for (int v : sys.nbasis_sym_spin) sys.max_nbss = std::max(sys.max_nbss, v);
    sys.sym_spin_basis_fns.assign((size_t)sys.max_nbss * 2 * sys.nsym_tot, 0);
    for (int i = 1; i <= sys.nbasis; ++i) {
        const BasisFn& b = sys.bf[i];
        int ims = (b.ms + 3) / 2;
        // minloc over the column: first zero element (entries are >=0, filled entries >0)
        for (int ind = 1; ind <= sys.max_nbss; ++ind) {
            int& slot = sys.sym_spin_basis_fns[(ind - 1) + sys.max_nbss * ((ims - 1) + 2 * b.sym)];
            if (slot == 0) { slot = i; break; }
        }
    }
}

// src/read_in.F90:1129-1216
inline void get_sp_eigv(const std::vector<IntLine>& lines, int norb, int nel_in, std::vector<double>& sp_eigv,
                        bool& found) {
    int nocc = nel_in / 2;
    found = false;
    sp_eigv.assign(norb + 1, 0.0);
    std::vector<char> seen_ijij((size_t)(norb + 1) * (norb + 1), 0), seen_ijji((size_t)(norb + 1) * (norb + 1), 0);
    auto S = [&](int i, int j) { return (size_t)i * (norb + 1) + j; };
    for (const auto& L : lines) {
        int i = L.i, a = L.a, j = L.j, b = L.b;
        double x = L.x;
        if (i > 0 && j == 0 && a == 0 && b == 0) {
            found = true;
            sp_eigv[i] = x;
        } else if (!found) {
            if (i == j && a == b && i == a && i > 0) {
                if (i <= nocc) sp_eigv[i] += x;
            } else if (i == a && j == b && b > 0) {
                if (!seen_ijij[S(i, j)]) {
                    seen_ijij[S(i, j)] = 1; seen_ijij[S(j, i)] = 1;
                    if (i <= nocc) sp_eigv[j] += 2 * x;
                    if (j <= nocc) sp_eigv[i] += 2 * x;
                }
            } else if (((i == b && j == a) || (i == j && a == b)) && b > 0) {
                if (!seen_ijji[S(i, a)]) {
                    seen_ijji[S(i, a)] = 1; seen_ijji[S(a, i)] = 1;
                    if (i <= nocc) sp_eigv[a] -= x;
                    if (a <= nocc) sp_eigv[i] -= x;
                }
            } else if (i == a && j == 0 && b == 0 && i > 0) {
                sp_eigv[i] += x;
            }
        }
    }
}

inline void read_in_integrals(System& sys, const std::string& text, const ReadInOpts& opt) {
    FcidumpHeader h = parse_fcidump_header(text);
    std::vector<IntLine> lines = parse_fcidump_body(text, h.body_offset);
    int norb = h.norb;
    sys.uhf = h.uhf;
    sys.nel = opt.nel;
    sys.Ms = opt.ms;
    sys.symmetry = opt.sym;
    sys.cas[0] = opt.cas[0]; sys.cas[1] = opt.cas[1];
    int rhf_fac;
    if (sys.uhf) { sys.nbasis = norb; rhf_fac = 1; } else { sys.nbasis = 2 * norb; rhf_fac = 2; }
    const int self_coulomb_fac = 1;
    if (sys.nel == 0 && sys.Ms == INT32_MAX) {
        if (h.nelec == 0 || h.ms2 == INT32_MAX) throw std::runtime_error("nel/ms not provided");
        sys.nel = h.nelec; sys.Ms = h.ms2;
    } else if (sys.Ms == INT32_MAX || sys.nel == 0) {
        throw std::runtime_error("provide both nel and ms or neither");
    }
    std::vector<double> sp_eigv;
    bool found;
    get_sp_eigv(lines, norb, sys.nel, sp_eigv, found);

    std::vector<int> sp_eigv_rank(norb + 1, 0), sp_fcidump_rank(norb + 1, 0);
    if (sys.uhf) {
        int nh_a = (norb + 1) / 2, nh_b = norb / 2;
        std::vector<double> ea(nh_a + 1), eb(nh_b + 1);
        for (int k = 1; k <= nh_a; ++k) ea[k] = sp_eigv[2 * k - 1];
        for (int k = 1; k <= nh_b; ++k) eb[k] = sp_eigv[2 * k];
        std::vector<int> ra, rb;
        insertion_rank(ea, nh_a, ra, depsilon);
        insertion_rank(eb, nh_b, rb, depsilon);
        for (int k = 1; k <= nh_a; ++k) sp_eigv_rank[2 * k - 1] = 2 * ra[k] - 1;
        for (int k = 1; k <= nh_b; ++k) sp_eigv_rank[2 * k] = 2 * rb[k];
    } else {
        std::vector<int> r;
        insertion_rank(sp_eigv, norb, r, depsilon);
        for (int k = 1; k <= norb; ++k) sp_eigv_rank[k] = r[k];
    }
    sp_eigv_rank[0] = 0;
    for (int i = 0; i <= norb; ++i)
        for (int j = 0; j <= norb; ++j)
            if (sp_eigv_rank[j] == i) { sp_fcidump_rank[i] = j; break; }

    int active_basis_offset = 0;
    if (sys.cas[0] > 0 && sys.cas[1] > 0) {
        active_basis_offset = sys.nel - sys.cas[0];
        sys.nbasis = 2 * sys.cas[1];
        sys.nel = sys.cas[0];
    }
    sys.nvirt = sys.nbasis - sys.nel;
    norb = sys.uhf ? sys.nbasis : sys.nbasis / 2;
    sys.bf.assign(sys.nbasis + 1, BasisFn());
    init_basis_fns_read_in(norb, sys.uhf, h.orbsym, h.symlz, sp_eigv,
                           sp_eigv_rank.data() + active_basis_offset / rhf_fac, sys.bf);
    int minsym = 0;
    for (int i = 1; i <= sys.nbasis; ++i) minsym = std::min(minsym, sys.bf[i].sym);
    if (minsym < 0) for (int i = 1; i <= sys.nbasis; ++i) sys.bf[i].sym = 0;
    sys.W = (sys.nbasis + 63) / 64;
    if (sys.W > MAXW) throw std::runtime_error("oracle: too many basis functions for MAXW");
    init_pg_symmetry(sys);

    // init_one_body_t / init_two_body_t (src/molecular_integrals.F90:60-190)
    int nspin1 = sys.uhf ? 2 : 1;
    sys.one_body.assign((size_t)nspin1 * sys.nsym_tot, {});
    for (int sp = 1; sp <= nspin1; ++sp)
        for (int s = 0; s < sys.nsym_tot; ++s) {
            int n = sys.nbss(sp, s);
            sys.one_body[(sp - 1) * sys.nsym_tot + s].assign((size_t)(n * (n + 1)) / 2, 0.0);
        }
    int64_t npairs = ((int64_t)(sys.nbasis / 2) * (sys.nbasis / 2 + 1)) / 2;
    sys.nintgrls = (npairs * (npairs + 1)) / 2;
    sys.two_body.assign(sys.uhf ? 4 : 1, std::vector<double>((size_t)sys.nintgrls, 0.0));

    sys.Ecore = 0.0;
    sys.int_err = 0;
    int nb = sys.nbasis;
    std::vector<char> seen_iha((size_t)(nb * (nb + 1)) / 2 + 1, 0);
    int abo = active_basis_offset;
    std::vector<int> seen_ijij((size_t)(abo * (abo + 1)) / 2 + 1, 0);
    // seen_iaib(-abo+1:0, 1:nb(nb+1)/2)
    size_t ntri = (size_t)(nb * (nb + 1)) / 2;
    std::vector<int> seen_iaib((size_t)std::max(abo, 1) * (ntri + 1), 0);
    auto IAIB = [&](int core, int64_t t) -> int& { return seen_iaib[(size_t)(core + abo - 1) * (ntri + 1) + t]; };

    for (const auto& L : lines) {
        double x = L.x;
        if (L.i > h.norb || L.a > h.norb || L.j > h.norb || L.b > h.norb) continue;
        int i = rhf_fac * sp_fcidump_rank[L.i];
        int j = rhf_fac * sp_fcidump_rank[L.j];
        int a = rhf_fac * sp_fcidump_rank[L.a];
        int b = rhf_fac * sp_fcidump_rank[L.b];
        int ii = i - abo, jj = j - abo, aa = a - abo, bb = b - abo;
        if (std::max(std::max(ii, jj), std::max(aa, bb)) > sys.nbasis) continue;
        if (i == 0 && j == 0 && a == 0 && b == 0) {
            sys.Ecore += x;
        } else if (i > 0 && j == 0 && a == 0 && b == 0) {
            // single-particle eigenvalue line: already handled
        } else if (j == 0 && b == 0) {
            // <i|h|a>
            if (ii < 1 && ii == aa) {
                sys.Ecore += x * rhf_fac;
            } else if (ii > 0 && aa > 0) {
                if (!seen_iha[tri_ind_reorder(ii, aa)]) {
                    x = x + sys.get_one_body_real(ii, aa);
                    sys.store_one_body(ii, aa, x);
                    seen_iha[tri_ind_reorder(ii, aa)] = 1;
                }
            }
        } else {
            int orbs[4] = {ii, jj, aa, bb};
            int nact = 0;
            for (int k = 0; k < 4; ++k) if (orbs[k] > 0) nact++;
            if (nact == 0) {
                if (ii == aa && jj == bb && ii == jj) {
                    if (!sys.uhf && seen_ijij[tri_ind_reorder(i, j)] % 2 == 0) {
                        sys.Ecore += x * self_coulomb_fac;
                        seen_ijij[tri_ind_reorder(i, j)] += 1;
                    }
                } else if (ii == aa && jj == bb && ii != jj) {
                    if (seen_ijij[tri_ind_reorder(i, j)] % 2 == 0) {
                        sys.Ecore += x * rhf_fac * rhf_fac;
                        seen_ijij[tri_ind_reorder(i, j)] += 1;
                    }
                } else if ((ii == bb && jj == aa && ii != jj) || (ii == jj && aa == bb && ii != aa)) {
                    int64_t ti = (ii == jj) ? tri_ind_reorder(i, a) : tri_ind_reorder(i, j);
                    if (seen_ijij[ti] < 2) {
                        sys.Ecore -= rhf_fac * x;
                        seen_ijij[ti] += 2;
                    }
                }
            } else if (nact == 2) {
                int active[2], core[2], ia = 0, ic = 0;
                for (int k = 0; k < 4; ++k) {
                    if (orbs[k] > 0) active[ia++] = orbs[k]; else core[ic++] = orbs[k];
                }
                if (core[0] == core[1]) {
                    int64_t t = tri_ind_reorder(active[0], active[1]);
                    if ((ii == core[0] && aa == core[0]) || (jj == core[0] && bb == core[0])) {
                        // <ij|aj> with j core: Coulomb contribution to <i|h|a>
                        if (IAIB(core[0], t) % 2 == 0) {
                            x = x * rhf_fac + sys.get_one_body_real(active[0], active[1]);
                            sys.store_one_body(active[0], active[1], x);
                            IAIB(core[0], t) += 1;
                        }
                    } else {
                        // exchange contribution
                        bool gam = (sys.cross_product(sys.sym_conj(sys.bf[active[0]].sym), sys.bf[active[1]].sym)
                                    == sys.gamma_sym);
                        if (IAIB(core[0], t) < 2 && gam) {
                            x = sys.get_one_body_real(active[0], active[1]) - x;
                            sys.store_one_body(active[0], active[1], x);
                            IAIB(core[0], t) += 2;
                        }
                    }
                }
            } else if (nact == 4) {
                sys.store_two_body(ii, jj, aa, bb, x);
            }
        }
    }

    // set_spin_polarisation (src/calc_system_init.f90:11-93), default branch
    sys.nbeta = (sys.nel - sys.Ms) / 2;
    sys.nalpha = (sys.nel + sys.Ms) / 2;
    sys.nvirt_alpha = sys.nbasis / 2 - sys.nalpha;
    sys.nvirt_beta = sys.nbasis / 2 - sys.nbeta;
}

inline void read_in_file(System& sys, const std::string& path, const ReadInOpts& opt) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open FCIDUMP: " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    read_in_integrals(sys, ss.str(), opt);
}

// src/reference_determinant.f90:46-301 (read_in branch).  occ is 1-based orbital list, sorted.
inline std::vector<int> set_reference_det(const System& sys, int ref_sym) {
    int nel = sys.nel;
    std::vector<int> occ(nel);
    for (int i = 1; i <= sys.nalpha; ++i) occ[i - 1] = 2 * i - 1;
    for (int i = 1; i <= sys.nbeta; ++i) occ[i - 1 + sys.nalpha] = 2 * i;
    if (ref_sym != INT32_MAX && ref_sym >= sys.sym0 && ref_sym <= sys.sym_max) {
        Det f = sys.encode(occ.data(), nel);
        int sym = sys.symmetry_orb_list(occ.data(), nel);
        if (sym != ref_sym) {
            double eigv_sum = std::numeric_limits<double>::max();
            std::vector<int> tmp(nel), curr(nel);
            for (int icore = 0; icore < nel; ++icore) {
                int i = occ[icore];
                for (int ivirt = 1; ivirt <= sys.nbasis; ++ivirt) {
                    if (!det_test(f, ivirt) && sys.bf[i].ms == sys.bf[ivirt].ms) {
                        tmp = occ; tmp[icore] = ivirt;
                        if (sys.symmetry_orb_list(tmp.data(), nel) == ref_sym) {
                            double s = 0.0;
                            for (int iel = 0; iel < nel; ++iel) s += sys.bf[tmp[iel]].sp_eigv;
                            if (s + depsilon < eigv_sum) { curr = tmp; eigv_sum = s; }
                        }
                    }
                }
            }
            for (int icore = 0; icore < nel; ++icore) {
                int i = occ[icore];
                for (int jcore = icore + 1; jcore < nel; ++jcore) {
                    int j = occ[jcore];
                    for (int ivirt = 1; ivirt <= sys.nbasis; ++ivirt) {
                        if (det_test(f, ivirt)) continue;
                        for (int jvirt = ivirt + 1; jvirt <= sys.nbasis; ++jvirt) {
                            if (!det_test(f, jvirt) &&
                                (sys.bf[i].ms + sys.bf[j].ms) == (sys.bf[ivirt].ms + sys.bf[jvirt].ms)) {
                                tmp = occ; tmp[icore] = ivirt; tmp[jcore] = jvirt;
                                if (sys.symmetry_orb_list(tmp.data(), nel) == ref_sym) {
                                    double s = 0.0;
                                    for (int iel = 0; iel < nel; ++iel) s += sys.bf[tmp[iel]].sp_eigv;
                                    if (s + depsilon < eigv_sum) { curr = tmp; eigv_sum = s; }
                                }
                            }
                        }
                    }
                }
            }
            if (!(eigv_sum < std::numeric_limits<double>::max()))
                throw std::runtime_error("Could not find determinant of required symmetry.");
            occ = curr;
        }
    }
    std::sort(occ.begin(), occ.end());
    return occ;
}

// MurmurHash2 (Austin Appleby, public domain; lib/external/MurmurHash2.c:16) restated, and the
// bit-string wrapper lib/local/hash.f90:29-60: hashes ceil(nbits/32)*4 bytes of f, little-endian.
inline uint32_t murmurhash2(const void* key, int len, uint32_t seed) {
    const uint32_t m = 0x5bd1e995u;
    const int r = 24;
    uint32_t h = seed ^ (uint32_t)len;
    const unsigned char* data = (const unsigned char*)key;
    while (len >= 4) {
        uint32_t k;
        memcpy(&k, data, 4);
        k *= m; k ^= k >> r; k *= m;
        h *= m; h ^= k;
        data += 4; len -= 4;
    }
    switch (len) {
        case 3: h ^= (uint32_t)data[2] << 16; /* fallthrough */
        case 2: h ^= (uint32_t)data[1] << 8;  /* fallthrough */
        case 1: h ^= (uint32_t)data[0]; h *= m;
    }
    h ^= h >> 13; h *= m; h ^= h >> 15;
    return h;
}
inline int32_t murmurhash_bit_string(const Det& f, int nbits, uint32_t seed) {
    int nbytes = ((nbits + 31) / 32) * 4;
    return (int32_t)murmurhash2(f.w, nbytes, seed);
}
// Fortran modulo(a, p) for p > 0: non-negative result.
inline int fmodulo(int64_t a, int64_t p) { int64_t r = a % p; if (r < 0) r += p; return (int)r; }

// assign_particle_processor (src/spawning.F90:770-838), shift == 0 branch and the CCMC
// time-varying branch.
inline int assign_particle_processor(const Det& f, int nbits, uint32_t seed, int shift, int freq, int np,
                                     const int* proc_map, int nslots, int* slot_pos = nullptr) {
    int32_t hash = murmurhash_bit_string(f, nbits, seed);
    int slot;
    if (shift == 0) {
        slot = fmodulo(hash, (int64_t)np * nslots);
    } else {
        // offset = ishft(hash+shift, -freq) on a default (32-bit) integer: logical shift right.
        uint32_t hs = (uint32_t)(hash + shift);
        uint64_t offset = (uint64_t)(hs >> freq);
        Det g = f;
        g.w[0] ^= offset;
        hash = murmurhash_bit_string(g, nbits, seed);
        slot = fmodulo(hash, (int64_t)np * nslots);
    }
    if (slot_pos) *slot_pos = slot;
    return proc_map[slot];
}

}  // namespace oracle
