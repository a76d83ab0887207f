// hande_b200: per-determinant / per-attempt device logic of the FCIQMC propagation hot path.
//
// Everything here is __host__ __device__ so that the same source is compiled (a) by nvcc into the
// sm_100a kernels of libhande_b200.so and (b) by g++ into tests/hdcheck (a CPU harness that checks
// these functions against the oracle without a GPU).  The product never runs the host versions.
//
// Semantics follow the reference (hande-qmc/hande); each function cites the routine it replaces.
// Orbital indices are 1-based as in the reference; orbital i is bit (i-1)%64 of word (i-1)/64
// (src/basis_types.f90:134-185); odd = alpha (ms=+1), even = beta (ms=-1); ims = (ms+3)/2.
//
// Floating point: compile with -fmad=false (nvcc) / -ffp-contract=off (g++): sums must be evaluated
// in the reference's order without FMA contraction so that excitation choice, nspawn and death
// counts are bit-exact against the oracle.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_HDN inline __host__ __device__
#define HB_HDNI inline __host__ __device__ __noinline__
#else
#define HB_HD inline
#define HB_HDN inline
#define HB_HDNI inline
#endif

#ifndef HB_MAXW
#define HB_MAXW 32          // widest bit string (words): wide lists are stored with 32 words per determinant
#endif
// Occupied-orbital lists hold 1-based spin-orbital indices: bytes while nbasis <= 254 (W <= 4); the translation units
// of the wide layout (W = 32, up to 2048 spin-orbitals: the plane-wave bases of the UEG) are compiled with -DHB_OCC16.
#ifdef HB_OCC16
typedef uint16_t occ_t;
#else
typedef uint8_t occ_t;
#endif
#define HB_MAXNEL 64
#define HB_MAX_CLUSTER 8    // largest CCMC cluster (ex_level + 2 <= 8): size of the selection buffers
#define HB_MAXLIST 128      // longest per-spin / per-symmetry-class orbital list (nbasis <= 254)

// Streaming (evict-first) loads for data with no reuse on the SM - the 2 GB heat-bath row tables and the walker list -
// so that they do not displace the small reused tables (hb_ij_w, single-excitation rows) from L1.
#if defined(__CUDA_ARCH__)
#define HB_LDCS(p) __ldcs(p)
#else
#define HB_LDCS(p) (*(p))
#endif

namespace hb {

struct alignas(16) D2 { double x, y; };
// One entry of a heat-bath alias row, packed so that a selection (aliasU, aliasK of the drawn slot) and the weight of
// the entry it usually returns come from ONE 32-byte sector: {aliasU, weight, aliasK} of hb_ija(a,j,i) / hb_ijab(b,a,j,i)
// plus p = weight / weights_tot of its row (the conditional probability the pgen formulas divide out at run time; the
// same IEEE division of the same operands, done once when the tables are built)
struct alignas(32) HbRec { double U, w; int K, pad0; double p; };
struct alignas(16) K4 { int x, y, z, w; };   // wavevector of a plane-wave basis function (w unused)

enum { SYS_READ_IN = 0, SYS_UEG = 1 };

// ------------------------------------------------------------------------------------------------
// System tables resident in HBM/L2 (plain pointers; filled by hb200_set_system_read_in).
// ------------------------------------------------------------------------------------------------
struct Sys {
    int kind;                   // SYS_READ_IN (molecular, FCIDUMP) or SYS_UEG (uniform electron gas, 3D)
    int nbasis, nel, W;
    int nsym_tot, sym0, sym_max, pg_mask, Lz_mask, Lz_offset, gamma_sym;
    int uhf, nvirt, nvirt_alpha, nvirt_beta, max_nbss;
    double Ecore;
    // basis (1-based: entry 0 unused)
    const uint8_t* bf_sym;      // [nbasis+1]
    const int8_t* bf_ms;        // [nbasis+1]
    const uint16_t* bf_spatial; // [nbasis+1]
    // symmetry tables (src/point_group_symmetry.f90:82-229)
    const int* nbss;            // nbasis_sym_spin[(ims-1) + 2*sym]
    const int* ssbf;            // sym_spin_basis_fns[(ind-1) + max_nbss*((ims-1)+2*sym)]
    const uint64_t* su_mask;    // [2*nsym_tot][W] bit strings of the basis functions of each (spin, sym) class
    // integrals (src/molecular_integrals.F90): dense one-body <i|h|j> (0 where symmetry forbids),
    // two-body store v[indx-1] per spin channel exactly as two_body_t%integrals(chan)%v
    const double* h1;           // [(i-1)*nbasis + (j-1)]
    const double* v2[4];
    // diagonal helper tables built on device from v2: J(i,j)=<ij|ij>, K(i,j)=<ij|ji>, spin-orbital indexed
    const double* Jd;           // [(i-1)*nbasis + (j-1)]
    const double* Kd;
    // single-excitation tables built on device from v2: C(i,a,j)=<ij|aj>, X(i,a,j)=<ij|ja>, stored
    // [(t(i)*NT + t(a))*NT + t(j)] with t() = spatial index-1 (RHF, NT = nbasis/2) or orbital-1 (UHF, NT = nbasis):
    // the occupied-orbital sum of slater_condon1_mol_excit then walks two contiguous rows instead of
    // gathering ~30 scattered entries of the 8-fold store (same values, same summation order).
    const D2* sc1CX;            // {C, X} pairs: one 16-byte load per occupied orbital
    int NT;
    // the same integrals laid out for a branch-free sum over an occupied list (FCIQMC heat-bath kernel): row (i, a) =
    // sc1T + ((i-1)*sc1A + ta)*(nbasis+1), ta = a-1 (UHF, sc1A = nbasis) or (a-1)/2 (RHF, sc1A = nbasis/2; a has the spin
    // of i); entry j (1-based spin-orbital) = {<ij|aj>, <ij|ja> if j has the spin of i else 0}; entries 0 and j = i are
    // {0, 0}, so list padding and the excited orbital itself add exact zeros
    const D2* sc1T;
    int sc1A;
    // Power-Pitzer / Cauchy-Schwarz weights sqrt|<ia|ai>| and sqrt|<ia|ia>| for every orbital pair: [2][nbasis][nbasis]
    // (null: computed on the fly)
    const double* ppw;
    // heat-bath tables (src/excit_gens.f90:143-153), column-major as in the reference
    const double* hb_i_w;       // (nb)
    const double* hb_ij_w;      // (j,i)
    const double* hb_ija_w;     // (a,j,i)
    const double* hb_ija_U;
    const int* hb_ija_K;
    const double* hb_ija_tot;   // (j,i)
    const double* hb_ijab_w;    // (b,a,j,i)
    const double* hb_ijab_U;
    const int* hb_ijab_K;
    const double* hb_ijab_tot;  // (a,j,i)
    // the same rows as packed records (built by hb200_build_heat_bath; read by the FCIQMC heat-bath kernel)
    const HbRec* hb_ija_rec;    // (a,j,i)
    const HbRec* hb_ijab_rec;   // (b,a,j,i)
    // power_pitzer_orderN tables (excit_gen_power_pitzer_t ppn_*, src/excit_gens.f90:102-141), column-major:
    // PPN_IS (nel), PPN_IAS (max_nbss, i), PPN_ID (nel), PPN_IJD (nel, i), PPN_IAD (nbasis/2, i), PPN_JBD (max_nbss, sym, i)
    // with i = 0..nbasis (column 0 unused)
    struct AliasTab { const double* w; const double* U; const int* K; const double* tot; } ppn[6];
    const int* ppn_occ;         // the reference's occupied orbitals, ascending
    // power_pitzer tables (pp_ia_d (max nvirt per spin, nel), pp_jb_d (max_nbss, sym, nel)) and the reference's virtual
    // orbitals per spin ([0] beta, [1] alpha), ascending
    AliasTab pp_ia, pp_jb;
    const int* pp_virt[2];
    int pp_nvirt[2], pp_sia;
    // uniform electron gas (src/ueg.f90, src/ueg_types.f90): plane-wave basis, analytic integrals
    const K4* ueg_k;            // [nbasis+1] wavevectors in units of 2 pi / L
    const double* sp_eigv;      // [nbasis+1] kinetic energies
    double ueg_piL;             // pi * box_length (coulomb_int_ueg_3d = 1 / (pi L |q|^2))
    const int* ueg_lookup;      // ueg_basis_t%lookup (1-based flat index -> alpha basis function, -1 if outside)
    int ueg_kmax, ueg_offset, ueg_oi[3];
    const uint64_t* ueg_tern;   // ternary_conserve(0:W, -tK:tK, -tK:tK, -tK:tK)
    int ueg_tK, ueg_tD;
};

// Per-calculation parameters (qmc_in_t / qmc_state_t scalars the kernels need).
struct Params {
    int excit_gen;              // EXCIT_GEN_* (the reference's enumerator values)
    double pattempt_single, pattempt_double;
    double pattempt_parallel;   // renorm_spin / no_renorm_spin: probability that i and j have parallel spins
    double tau, shift, proj_energy_old;
    int64_t real_factor;        // pop_real_factor (1 or 2^31)
    int64_t spawn_cutoff;       // encoded (src/spawn_data.F90:215)
    int initiator;              // initiator approximation on
    double initiator_pop;
    int real_amplitudes;
    int trunc_level;            // max excitation level from f0 allowed (-1: none)
    uint32_t seed, cycle;
    uint32_t hash_seed;         // 7 (src/qmc.F90:1497)
    int nprocs, iproc, nslots;
    int ccmc_shift, ccmc_freq;  // spawn%hash_shift / spawn%move_freq (CCMC only)
    uint64_t f0[HB_MAXW];
    int we;                     // words of a bit string in the host's layout (ceil(nbasis/64)); the wide layout pads to 32
    double H00;
    struct PsPartials* ps_part; // per-block sums for qmc_in%pattempt_update (null: not accumulating)
    // quasi-Newton propagator (propagator_t, src/qmc_data.f90:866-884); sp_fock is 1-based
    int qn;
    double qn_threshold, qn_value, qn_pop_control, ref_fock_sum;
    const double* sp_fock;
    // wall-Chebyshev propagator (src/propagators.f90): weight 1/(S_i - E_0) of the current sub-cycle on the spawning
    // amplitude (src/spawning.F90:117-118) and the death probability (src/death.f90:89); 1 for the linear projector
    double cheby_weight;
    // semi-stochastic projection (semi_stoch_t, src/semi_stoch.F90:40-108): one bit per state of the main list (set:
    // deterministic, i.e. determ%flags == 0) and every deterministic determinant of every rank, sorted, for
    // check_if_determ; null / 0 while the projection is off
    const uint32_t* ss_bits;
    const uint64_t* ss_sorted;
    int ss_tot;
};
// p_single_double_coll_t (src/excit_gens.f90:13-27): sums of |H_ij| pattempt_{single,double} / pgen over the allowed
// single / double excitations generated, and how many there were
struct PsPartials { double h_pgen_singles_sum, h_pgen_doubles_sum; long long excit_gen_singles, excit_gen_doubles; };

// the reference's enumerator values (src/qmc_data.f90:31-69: renorm, renorm_spin, no_renorm, no_renorm_spin, power_pitzer,
// power_pitzer_occ, power_pitzer_occ_ij, power_pitzer_orderN, cauchy_schwarz_occ, cauchy_schwarz_occ_ij, heat_bath,
// heat_bath_uniform, heat_bath_single)
enum { EXCIT_GEN_RENORM = 0, EXCIT_GEN_RENORM_SPIN = 1, EXCIT_GEN_NO_RENORM = 2, EXCIT_GEN_NO_RENORM_SPIN = 3, EXCIT_GEN_POWER_PITZER = 4, EXCIT_GEN_POWER_PITZER_OCC = 5, EXCIT_GEN_POWER_PITZER_OCC_IJ = 6,
       EXCIT_GEN_POWER_PITZER_ORDERN = 7,
       EXCIT_GEN_CAUCHY_SCHWARZ_OCC = 8, EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ = 9,
       EXCIT_GEN_HEAT_BATH = 10, EXCIT_GEN_HEAT_BATH_UNIFORM = 11, EXCIT_GEN_HEAT_BATH_SINGLE = 12 };
enum { PPN_IS = 0, PPN_IAS = 1, PPN_ID = 2, PPN_IJD = 3, PPN_IAD = 4, PPN_JBD = 5 };
#define HB_NW(p) ((W > 4) ? (p).we : W)
enum { RNG_NATTEMPTS = 0, RNG_SPAWN = 1, RNG_DEATH = 2, RNG_ROUND_MAIN = 3, RNG_ROUND_SPAWN = 4, RNG_DETERM = 5 };

// ------------------------------------------------------------------------------------------------
// Counter-based random stream: Philox4x32-10 keyed by (seed, cycle); counter
// (purpose<<24 | draw/2, attempt, hash64(det)).  Identical to oracle/rng.hpp PhiloxRng.
// ------------------------------------------------------------------------------------------------
HB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

HB_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// nw: words hashed (the host's width; the wide layout pads a determinant with zero words that must not enter the key)
template <int W>
HB_HD uint64_t det_hash64(const uint64_t* f, int nw = W) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (W > 4 && i >= nw) break;
        uint64_t z = f[i] + h;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        h = z + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    }
    return h;
}

struct PhiloxStream {
    uint32_t k0, k1, c1, c2, c3, purpose, draw;
    uint32_t buf[4];
    uint32_t pre[8];     // blocks 0 and 1 (draws 0..3) when prefetched
    bool have_pre;
    HB_HD void begin(uint32_t seed, uint32_t cycle, uint32_t purpose_, uint64_t dethash, uint32_t attempt) {
        k0 = seed; k1 = cycle; purpose = purpose_; c1 = attempt;
        c2 = (uint32_t)dethash; c3 = (uint32_t)(dethash >> 32); draw = 0; have_pre = false;
    }
    // Evaluate the first two blocks (draws 0..3) now - called where the warp is converged, so that the draws made
    // inside divergent generator code (rejection loops, single/double branches) cost a register read instead of a
    // serialised 10-round Philox evaluation per divergent path.  The stream itself is unchanged.
    HB_HD void prefetch() {
        philox4x32_10((purpose << 24) | 0u, c1, c2, c3, k0, k1, pre);
        philox4x32_10((purpose << 24) | 1u, c1, c2, c3, k0, k1, pre + 4);
        have_pre = true;
    }
    HB_HD double next() {
        uint32_t lo, hi;
        if (have_pre && draw < 4u) {
            const uint32_t* q = pre;
            lo = (draw == 0) ? q[0] : (draw == 1) ? q[2] : (draw == 2) ? q[4] : q[6];
            hi = (draw == 0) ? q[1] : (draw == 1) ? q[3] : (draw == 2) ? q[5] : q[7];
        } else {
            if ((draw & 1u) == 0) philox4x32_10((purpose << 24) | (draw >> 1), c1, c2, c3, k0, k1, buf);
            lo = (draw & 1u) ? buf[2] : buf[0];
            hi = (draw & 1u) ? buf[3] : buf[1];
        }
        draw++;
        uint64_t u = ((uint64_t)hi << 32) | lo;
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
};

// Test stream: a caller-provided list of uniforms (hdcheck only).
// injected list of uniform numbers; past its end the stream continues with an equidistributed sequence (so that a
// rejection loop always terminates) and k > n tells the caller the list was too short
struct ListStream {
    const double* v; int n; int k;
    HB_HD double next() {
        if (k < n) return v[k++];
        const double x = 0.5 + 0.6180339887498949 * (double)(++k);
        return x - floor(x);
    }
};

// ------------------------------------------------------------------------------------------------
// MurmurHash2 (public-domain algorithm; lib/external/MurmurHash2.c:16) over the first
// ceil(nbits/32)*4 bytes of the bit string (lib/local/hash.f90:29-60), and the owner rule
// (assign_particle_processor, src/spawning.F90:770-838, shift == 0 branch).
// ------------------------------------------------------------------------------------------------
HB_HD uint32_t murmur2_words(const uint64_t* f, int nbits, uint32_t seed) {
    const uint32_t m = 0x5bd1e995u;
    int nwords32 = (nbits + 31) / 32;
    uint32_t h = seed ^ (uint32_t)(nwords32 * 4);
    for (int i = 0; i < nwords32; ++i) {
        uint32_t k = (uint32_t)(f[i >> 1] >> ((i & 1) * 32));
        k *= m; k ^= k >> 24; k *= m;
        h *= m; h ^= k;
    }
    h ^= h >> 13; h *= m; h ^= h >> 15;
    return h;
}
HB_HD int owner_slot(const uint64_t* f, int nbits, uint32_t seed, int nprocs, int nslots) {
    int32_t hash = (int32_t)murmur2_words(f, nbits, seed);
    int64_t p = (int64_t)nprocs * nslots;
    int64_t r = (int64_t)hash % p;
    if (r < 0) r += p;  // Fortran modulo
    return (int)r;
}

// assign_particle_processor with a time-varying shift (CCMC; src/spawning.F90:812-836): the hash of the label is
// re-hashed after xor-ing ishft(hash + shift, -freq) (logical shift of a default integer) into its first word
template <int W>
HB_HD int owner_slot_shift(const uint64_t* f, int nbits, uint32_t seed, int shift, int freq, int nprocs, int nslots) {
    if (shift == 0) return owner_slot(f, nbits, seed, nprocs, nslots);
    const uint32_t hs = murmur2_words(f, nbits, seed) + (uint32_t)shift;   // 32-bit wrap-around as in the reference build
    uint64_t g[W];
#pragma unroll
    for (int k = 0; k < W; ++k) g[k] = f[k];
    g[0] ^= (uint64_t)(hs >> freq);
    return owner_slot(g, nbits, seed, nprocs, nslots);
}

// ------------------------------------------------------------------------------------------------
// Bit-string helpers
// ------------------------------------------------------------------------------------------------
HB_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
HB_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
// basis_fns(:)%ms and the ims index (ms+3)/2: spin-orbitals are stored alpha (odd index, ms=+1), beta (even, ms=-1)
// for every system the reference builds (src/read_in.F90:862-920, src/basis.f90:465-486); spatial_index = (i+1)/2.
HB_HD int ms_of(int o) { return (o & 1) ? 1 : -1; }
HB_HD int ims_of(int o) { return (o & 1) ? 2 : 1; }
HB_HD int spatial_of(int o) { return (o + 1) >> 1; }
HB_HD bool det_test(const uint64_t* f, int orb) { return (f[(orb - 1) >> 6] >> ((orb - 1) & 63)) & 1ull; }

// decode_det (src/determinants.f90:243-297): occupied orbitals ascending.
template <int W>
HB_HD int decode_det(const uint64_t* f, occ_t* occ) {
    int n = 0;
#pragma unroll
    for (int iw = 0; iw < 2 * W; ++iw) {
        uint32_t x = (uint32_t)(f[iw >> 1] >> ((iw & 1) * 32));
        while (x) {
#if defined(__CUDA_ARCH__)
            const int b = __ffs((int)x) - 1;
#else
            const int b = __builtin_ctz(x);
#endif
            occ[n++] = (occ_t)(iw * 32 + b + 1);
            x &= x - 1;
        }
    }
    return n;
}
// bit_str_cmp order (src/bit_utils.F90:452-479): unsigned, last word most significant.
template <int W>
HB_HD bool det_less(const uint64_t* a, const uint64_t* b) {
#pragma unroll
    for (int i = W - 1; i >= 0; --i) {
        if (a[i] < b[i]) return true;
        if (a[i] > b[i]) return false;
    }
    return false;
}
template <int W>
HB_HD bool det_eq(const uint64_t* a, const uint64_t* b) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < W; ++i) e = e && (a[i] == b[i]);
    return e;
}

// Number of set bits of f strictly above orbital `orb` (excit_mask, src/excitations.F90:25-58).
template <int W>
HB_HD int popc_above(const uint64_t* f, int orb) {
    int iw = (orb - 1) >> 6, ib = (orb - 1) & 63;
    int n = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        uint64_t m = (k < iw) ? 0ull : ((k > iw) ? ~0ull : ((ib == 63) ? 0ull : (~0ull << (ib + 1))));
        n += popc64(f[k] & m);
    }
    return n;
}
// popcount( f & (mask_i ^ mask_a) ): occupied orbitals between the two positions (incl. the lower...)
template <int W>
HB_HD int popc_between(const uint64_t* f, int i, int a) {
    int n = 0;
    int iw = (i - 1) >> 6, ib = (i - 1) & 63, aw = (a - 1) >> 6, ab = (a - 1) & 63;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        uint64_t mi = (k < iw) ? 0ull : ((k > iw) ? ~0ull : ((ib == 63) ? 0ull : (~0ull << (ib + 1))));
        uint64_t ma = (k < aw) ? 0ull : ((k > aw) ? ~0ull : ((ab == 63) ? 0ull : (~0ull << (ab + 1))));
        n += popc64(f[k] & (mi ^ ma));
    }
    return n;
}
// find_excitation_permutation1 (src/excitations.F90:247-280)
template <int W>
HB_HD bool excit_perm1(const uint64_t* f, int i, int a) {
    int perm = popc_between<W>(f, i, a);
    if (i > a) perm -= 1;
    return (perm & 1) != 0;
}
// find_excitation_permutation2 (src/excitations.F90:282-363); i<j, a<b as produced by the generators
template <int W>
HB_HD bool excit_perm2(const uint64_t* f, int i, int j, int a, int b) {
    int perm = 0;
    int iw = (i - 1) >> 6, ib = (i - 1) & 63, aw = (a - 1) >> 6, ab = (a - 1) & 63;
    int jw = (j - 1) >> 6, jb = (j - 1) & 63, bw = (b - 1) >> 6, bb = (b - 1) & 63;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        uint64_t mi = (k < iw) ? 0ull : ((k > iw) ? ~0ull : ((ib == 63) ? 0ull : (~0ull << (ib + 1))));
        uint64_t ma = (k < aw) ? 0ull : ((k > aw) ? ~0ull : ((ab == 63) ? 0ull : (~0ull << (ab + 1))));
        uint64_t mj = (k < jw) ? 0ull : ((k > jw) ? ~0ull : ((jb == 63) ? 0ull : (~0ull << (jb + 1))));
        uint64_t mb = (k < bw) ? 0ull : ((k > bw) ? ~0ull : ((bb == 63) ? 0ull : (~0ull << (bb + 1))));
        perm += popc64((f[k] & (mi ^ ma)) ^ (f[k] & (mj ^ mb)));
    }
    if (i > a) perm += 1;
    if (i > b) perm += 1;
    if (j > b || j < a) perm += 1;
    return (perm & 1) != 0;
}

// ------------------------------------------------------------------------------------------------
// Symmetry algebra (src/point_group_symmetry.f90:297-346)
// ------------------------------------------------------------------------------------------------
HB_HD int cross_product(const Sys& s, int a, int b) {
    return ((a ^ b) & s.pg_mask) | ((a & s.Lz_mask) + (b & s.Lz_mask) - s.Lz_offset);
}
HB_HD int sym_conj(const Sys& s, int a) {
    return (a & s.pg_mask) | ((2 * s.Lz_offset - (a & s.Lz_mask)) & s.Lz_mask);
}
HB_HD int nbss(const Sys& s, int ims, int sym) { return s.nbss[(ims - 1) + 2 * sym]; }
HB_HD int ssbf(const Sys& s, int ind, int ims, int sym) {
    return s.ssbf[(ind - 1) + s.max_nbss * ((ims - 1) + 2 * sym)];
}

// ------------------------------------------------------------------------------------------------
// Integral store lookups (src/molecular_integrals.F90:685-845,1217-1257; lib/local/utils.F90:449-480)
// ------------------------------------------------------------------------------------------------
HB_HD int64_t tri_ind(int64_t i, int64_t j) { return (i * (i - 1)) / 2 + j; }

HB_HD int tri_ind32(int i, int j) { return (i * (i - 1)) / 2 + j; }
HB_HD double two_body(const Sys& s, int i, int j, int a, int b) {
    // 32-bit index arithmetic: nbasis <= 255 (engine limit) => < 2^14 pairs => index < 2^27
    int ii, jj, aa, bb;
    if (i < a) { ii = a; aa = i; } else { ii = i; aa = a; }
    if (j < b) { jj = b; bb = j; } else { jj = j; bb = b; }
    const int ia = tri_ind32(spatial_of(ii), spatial_of(aa));
    const int jb = tri_ind32(spatial_of(jj), spatial_of(bb));
    const int indx = (ia < jb) ? tri_ind32(jb, ia) : tri_ind32(ia, jb);
    int chan = 0;
    if (s.uhf) {
        if (ia < jb || (ia == jb && ii < jj)) { int t = ii; ii = jj; jj = t; }
        if (ms_of(ii) == -1) chan = (ms_of(jj) == -1) ? 0 : 2;
        else chan = (ms_of(jj) == 1) ? 1 : 3;
    }
    return s.v2[chan][indx - 1];
}
// get_two_body_int_mol_real (src/molecular_integrals.F90:1177-1215): zero unless spin and symmetry allow <ij|ab>
HB_HD double two_body_real(const Sys& s, int i, int j, int a, int b) {
    const int sij = cross_product(s, s.bf_sym[i], s.bf_sym[j]), sab = cross_product(s, s.bf_sym[a], s.bf_sym[b]);
    if (sij == cross_product(s, sab, s.gamma_sym) && ms_of(i) == ms_of(a) && ms_of(j) == ms_of(b)) return two_body(s, i, j, a, b);
    return 0.0;
}
HB_HD double one_body(const Sys& s, int i, int j) { return s.h1[(i - 1) * s.nbasis + (j - 1)]; }

// slater_condon0_mol_orb_list (src/hamiltonian_molecular.f90:99-139) through the J/K tables.
HB_HDN double slater_condon0(const Sys& s, const occ_t* occ) {
    double h = s.Ecore;
    const int nb = s.nbasis;
    for (int iel = 0; iel < s.nel; ++iel) {
        int i = occ[iel];
        h = h + one_body(s, i, i);
        for (int jel = iel + 1; jel < s.nel; ++jel) {
            int j = occ[jel];
            h = h + s.Jd[(i - 1) * nb + (j - 1)];
            if (ms_of(i) == ms_of(j)) h = h - s.Kd[(i - 1) * nb + (j - 1)];
        }
    }
    return h;
}
// slater_condon1_mol_excit (src/hamiltonian_molecular.f90:199-259); integrals through the C/X row tables
HB_HD int tix(const Sys& s, int i) { return s.uhf ? (i - 1) : ((i - 1) >> 1); }
HB_HDN double slater_condon1_excit(const Sys& s, const occ_t* occ, int i, int a, bool perm) {
    double h = one_body(s, i, a);
    const D2* __restrict__ row = s.sc1CX + ((long long)tix(s, i) * s.NT + tix(s, a)) * s.NT;
    const int nel = s.nel;
    // chunks of 4 occupied orbitals: issue the (independent) loads first, then add in occ_list order
    // The gathers of this kernel are bound by the L1 tag rate (one 32-byte sector per lane per cycle), not by
    // instructions: the two spin-orbitals of a doubly occupied spatial orbital are adjacent in occ_list and share their
    // RHF table entry, which is then loaded once.
    int tprev = -1;
    D2 vprev = {0.0, 0.0};
    for (int q0 = 0; q0 < nel; q0 += 4) {
        int jj[4];
        D2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            jj[k] = (q0 + k < nel) ? occ[q0 + k] : i;
            const int t = tix(s, jj[k]);
            const int tp = (k == 0) ? tprev : tix(s, jj[k - 1]);
            if (t != tp) v[k] = row[t];
            else v[k] = (k == 0) ? vprev : v[k - 1];
        }
        tprev = tix(s, jj[3]); vprev = v[3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = jj[k];
            if (j != i) {
                h = h + v[k].x;
                const bool same = s.uhf ? (ms_of(j) == ms_of(i)) : (((j ^ i) & 1) == 0);
                if (same) h = h - v[k].y;
            }
        }
    }
    return perm ? -h : h;
}
// slater_condon2_mol_excit (src/hamiltonian_molecular.f90:300-346)
HB_HD double slater_condon2_excit(const Sys& s, int i, int j, int a, int b, bool perm) {
    double h = 0.0;
    if (ms_of(i) == ms_of(a)) h = two_body(s, i, j, a, b);
    if (ms_of(i) == ms_of(b)) h = h - two_body(s, i, j, b, a);
    return perm ? -h : h;
}

// ------------------------------------------------------------------------------------------------
// Excitation result
// ------------------------------------------------------------------------------------------------
struct Gen {
    int nexcit, from1, from2, to1, to2;
    bool perm, allowed;
    double pgen, hmatel;
};

// symunocc(ims, sym) = nbasis_sym_spin - occupied (decode_det_occ_symunocc,
// src/determinant_decoders.f90:166-206); stored as uint8 [(ims-1)+2*sym]
HB_HD void build_symunocc(const Sys& s, const occ_t* occ, uint8_t* su) {
    for (int k = 0; k < 2 * s.nsym_tot; ++k) su[k] = (uint8_t)s.nbss[k];
    for (int i = 0; i < s.nel; ++i) {
        int o = occ[i];
        su[(ims_of(o) - 1) + 2 * s.bf_sym[o]]--;
    }
}
// the same through per-class occupation masks: symunocc(c) = nbasis_sym_spin(c) - popcount(f & mask(c)),
// su_mask[c*W + w] = bit string of the basis functions of class c = (ims-1) + 2*sym
template <int W>
HB_HD void build_symunocc_masks(const Sys& s, const uint64_t* f, uint8_t* su) {
    for (int c = 0; c < 2 * s.nsym_tot; ++c) {
        int n = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) n += popc64(f[w] & s.su_mask[c * W + w]);
        su[c] = (uint8_t)(s.nbss[c] - n);
    }
}
#define HB_SU(ims, sym) ((int)su[((ims) - 1) + 2 * (sym)])

// choose_ij_mol (src/excit_gen_mol.f90:620-682)
template <class R>
HB_HD void choose_ij(R& rng, const Sys& s, const occ_t* occ, int& i, int& j, int& ij_sym, int& ij_spin) {
    int nel = s.nel;
    int ind = (int)(rng.next() * nel * (nel - 1) / 2) + 1;
    int j_ind = (int)(1.50 + sqrt(2 * ind - 1.750));
    int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
    i = occ[i_ind - 1];
    j = occ[j_ind - 1];
    ij_sym = sym_conj(s, cross_product(s, s.bf_sym[i], s.bf_sym[j]));
    ij_spin = ms_of(i) + ms_of(j);
}

// choose_ij_spin_mol (src/excit_gen_mol.f90:684-800): parallel or anti-parallel pair first (pattempt_parallel), then the
// pair from occ_list_alpha / occ_list_beta (decode_det_spinocc_symunocc: alpha = ms +1)
HB_HD int nth_occ_of_spin(const occ_t* occ, int nel, int ms, int k) {
    for (int q = 0; q < nel; ++q)
        if (ms_of(occ[q]) == ms && --k == 0) return occ[q];
    return 0;
}
template <class R>
HB_HD void choose_ij_spin(R& rng, const Sys& s, const Params& p, const occ_t* occ, int& i, int& j, int& ij_sym,
                          int& ij_spin, double& pgen_ij, bool& allowed) {
    const int nalpha = s.nbasis / 2 - s.nvirt_alpha, nbeta = s.nbasis / 2 - s.nvirt_beta;
    allowed = true;
    if (rng.next() < p.pattempt_parallel) {
        const bool alpha = rng.next() < ((double)nalpha / (double)(nalpha + nbeta));
        const int n = alpha ? nalpha : nbeta;
        if (n < 2) {
            allowed = false;
        } else {
            const int ind = (int)(rng.next() * (double)(n * (n - 1)) / 2.0) + 1;
            const int j_ind = (int)(1.50 + sqrt(2 * ind - 1.750));
            const int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
            i = nth_occ_of_spin(occ, s.nel, alpha ? 1 : -1, i_ind);
            j = nth_occ_of_spin(occ, s.nel, alpha ? 1 : -1, j_ind);
            pgen_ij = (p.pattempt_parallel * ((double)n / (double)(nalpha + nbeta)) * 2.0 * (1.0 / n) * (1.0 / (n - 1)));
        }
    } else {
        if (nbeta < 1 || nalpha < 1) {
            allowed = false;
        } else {
            const int i_ind = (int)(rng.next() * nalpha) + 1;
            const int j_ind = (int)(rng.next() * nbeta) + 1;
            i = nth_occ_of_spin(occ, s.nel, 1, i_ind);
            j = nth_occ_of_spin(occ, s.nel, -1, j_ind);
            if (j < i) { const int t = i; i = j; j = t; }
            pgen_ij = (1.0 - p.pattempt_parallel) * (1.0 / (nalpha * nbeta));
        }
    }
    if (allowed) {
        ij_sym = sym_conj(s, cross_product(s, s.bf_sym[i], s.bf_sym[j]));
        ij_spin = ms_of(i) + ms_of(j);
    } else {
        pgen_ij = 1.0; i = 0; j = 0; ij_sym = 0; ij_spin = 0;
    }
}

// gen_single_excit_mol (src/excit_gen_mol.f90:384-448): choose_ia_mol (:802-870) + calc_pgen_single_mol (:1140-1190)
template <int W, class R>
HB_HDN void gen_single_renorm(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                              const uint8_t* su, Gen& g) {
    const int nel = s.nel;
    g.from2 = 0; g.to2 = 0; g.perm = false;
        // choose_ia_mol
        g.nexcit = 1;
        bool allowed = false;
        int ni = nel;
        for (int k = 0; k < nel; ++k) {
            int o = occ[k];
            int imsa = ims_of(o);
            int isyma = cross_product(s, s.bf_sym[o], s.gamma_sym);
            if (HB_SU(imsa, isyma) != 0) allowed = true; else ni--;
        }
        g.allowed = allowed;
        if (allowed) {
            int i, a;
            for (;;) {
                i = occ[(int)(rng.next() * nel)];
                int imsa = ims_of(i);
                int isyma = cross_product(s, s.bf_sym[i], s.gamma_sym);
                if (HB_SU(imsa, isyma) != 0) {
                    int n = nbss(s, imsa, isyma);
                    for (;;) {
                        int ind = (int)(n * rng.next()) + 1;
                        a = ssbf(s, ind, imsa, isyma);
                        if (!det_test(f, a)) break;
                    }
                    break;
                }
            }
            g.from1 = i; g.to1 = a;
            // calc_pgen_single_mol
            g.pgen = p.pattempt_single * (1.0 / (ni * HB_SU(ims_of(a), s.bf_sym[a])));
            g.perm = excit_perm1<W>(f, i, a);
            g.hmatel = slater_condon1_excit(s, occ, i, a, g.perm);
        } else {
            g.hmatel = 0.0; g.pgen = 1.0; g.from1 = 0; g.to1 = 0;
        }
}

// gen_excit_mol: renormalised uniform generator (src/excit_gen_mol.f90:16-101,384-448,521-616,
// 802-946,1140-1327)
// SPIN: gen_excit_mol_spin (src/excit_gen_mol.f90:103-193), excit_gen = renorm_spin - ij from choose_ij_spin_mol
template <int W, bool SPIN = false, class R>
HB_HDN void gen_excit_renorm(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                             const uint8_t* su, Gen& g) {
    const int nel = s.nel;
    g.from2 = 0; g.to2 = 0; g.perm = false;
    if (rng.next() < p.pattempt_single) {
        gen_single_renorm<W>(rng, s, p, f, occ, su, g);
    } else {
        g.nexcit = 2;
        int i, j, ij_sym, spin;
        double pgen_ij = 2.0 / (nel * (nel - 1));
        bool ij_ok = true;
        if (SPIN) choose_ij_spin(rng, s, p, occ, i, j, ij_sym, spin, pgen_ij, ij_ok);
        else choose_ij(rng, s, occ, i, j, ij_sym, spin);
        g.from1 = i; g.from2 = j;
        // choose_ab_mol
        bool allowed = false;
        int fac = 1, shift = 0, na = s.nbasis;
        if (!ij_ok) {
        } else if (spin == 0) {
            for (int isyma = s.sym0; isyma <= s.sym_max; ++isyma) {
                int isymb = sym_conj(s, cross_product(s, isyma, ij_sym));
                if ((HB_SU(1, isyma) > 0 && HB_SU(2, isymb) > 0) || (HB_SU(2, isyma) > 0 && HB_SU(1, isymb) > 0)) {
                    allowed = true; break;
                }
            }
        } else {
            const int sp = (spin == -2) ? 1 : 2;
            for (int isyma = s.sym0; isyma <= s.sym_max; ++isyma) {
                int isymb = sym_conj(s, cross_product(s, isyma, ij_sym));
                if (HB_SU(sp, isyma) > 0 && (HB_SU(sp, isymb) > 1 || (HB_SU(sp, isymb) == 1 && isyma != isymb))) {
                    allowed = true; break;
                }
            }
            fac = 2; shift = (spin == -2) ? 0 : 1; na = s.nbasis / 2;
        }
        g.allowed = allowed;
        if (allowed) {
            int a, b;
            for (;;) {
                a = (int)(rng.next() * na) + 1;
                a = fac * a - shift;
                if (!det_test(f, a)) {
                    int imsb = (spin - ms_of(a) + 3) / 2;
                    int isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
                    int nu = HB_SU(imsb, isymb);
                    if (nu > 1 || (nu == 1 && (isymb != s.bf_sym[a] || spin == 0))) {
                        int n = nbss(s, imsb, isymb);
                        for (;;) {
                            int ind = (int)(n * rng.next()) + 1;
                            b = ssbf(s, ind, imsb, isymb);
                            if (b != a && !det_test(f, b)) break;
                        }
                        break;
                    }
                }
            }
            if (a > b) { int t = a; a = b; b = t; }
            g.to1 = a; g.to2 = b;
            // calc_pgen_double_mol
            int imsa = ims_of(a), imsb = ims_of(b);
            int n_aij;
            double p_aijb, p_bija;
            if (spin != 0) {
                const int sp = (spin == -2) ? 1 : 2;
                n_aij = (spin == -2) ? s.nvirt_beta : s.nvirt_alpha;
                for (int isyma = s.sym0; isyma <= s.sym_max; ++isyma) {
                    int isymb = sym_conj(s, cross_product(s, isyma, ij_sym));
                    if (HB_SU(sp, isymb) == 0) n_aij -= HB_SU(sp, isyma);
                    else if (isyma == isymb && HB_SU(sp, isymb) == 1) n_aij -= HB_SU(sp, isyma);
                }
                if (s.bf_sym[a] == s.bf_sym[b]) {
                    p_aijb = 1.0 / (HB_SU(imsa, s.bf_sym[a]) - 1);
                    p_bija = 1.0 / (HB_SU(imsb, s.bf_sym[b]) - 1);
                } else {
                    p_aijb = 1.0 / HB_SU(imsa, s.bf_sym[a]);
                    p_bija = 1.0 / HB_SU(imsb, s.bf_sym[b]);
                }
            } else {
                n_aij = s.nvirt;
                for (int isyma = s.sym0; isyma <= s.sym_max; ++isyma) {
                    int isymb = sym_conj(s, cross_product(s, isyma, ij_sym));
                    if (HB_SU(1, isymb) == 0) n_aij -= HB_SU(2, isyma);
                    if (HB_SU(2, isymb) == 0) n_aij -= HB_SU(1, isyma);
                }
                p_aijb = 1.0 / HB_SU(imsa, s.bf_sym[a]);
                p_bija = 1.0 / HB_SU(imsb, s.bf_sym[b]);
            }
            g.pgen = p.pattempt_double * pgen_ij * ((1.0 / n_aij) * (p_bija + p_aijb));
            g.perm = excit_perm2<W>(f, i, j, a, b);
            g.hmatel = slater_condon2_excit(s, i, j, a, b, g.perm);
        } else {
            g.hmatel = 0.0; g.pgen = 1.0; g.to1 = 0; g.to2 = 0;
        }
    }
}

// gen_excit_mol_no_renorm (src/excit_gen_mol.f90:195-284,450-517,950-1136,1329-1445)
// SPIN: gen_excit_mol_no_renorm_spin (src/excit_gen_mol.f90:286-380), excit_gen = no_renorm_spin
template <int W, bool SPIN = false, class R>
HB_HDN void gen_excit_no_renorm(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                Gen& g) {
    const int nel = s.nel;
    g.from2 = 0; g.to2 = 0; g.perm = false; g.to1 = 0;
    if (rng.next() < p.pattempt_single) {
        g.nexcit = 1;
        int i = occ[(int)(rng.next() * nel)];
        int imsa = ims_of(i);
        int isyma = cross_product(s, s.bf_sym[i], s.gamma_sym);
        int n = nbss(s, imsa, isyma);
        int ind = (int)(n * rng.next()) + 1;
        g.from1 = i;
        int a = 0;
        if (n == 0) {
            g.allowed = false;
        } else {
            a = ssbf(s, ind, imsa, isyma);
            g.allowed = !det_test(f, a);
        }
        if (g.allowed) {
            g.to1 = a;
            g.pgen = p.pattempt_single * (1.0 / (nel * nbss(s, ims_of(a), s.bf_sym[a])));
            g.perm = excit_perm1<W>(f, i, a);
            g.hmatel = slater_condon1_excit(s, occ, i, a, g.perm);
        } else {
            g.hmatel = 0.0; g.pgen = 1.0;
        }
    } else {
        g.nexcit = 2;
        int i, j, ij_sym, spin;
        double pgen_ij = 2.0 / (nel * (nel - 1));
        bool ij_ok = true;
        if (SPIN) choose_ij_spin(rng, s, p, occ, i, j, ij_sym, spin, pgen_ij, ij_ok);
        else choose_ij(rng, s, occ, i, j, ij_sym, spin);
        g.from1 = i; g.from2 = j;
        int fac = 1, shift = 0, na = s.nbasis;
        if (spin == -2) { fac = 2; shift = 0; na = s.nbasis / 2; }
        else if (spin == 2) { fac = 2; shift = 1; na = s.nbasis / 2; }
        int a = 0, b = 0;
        int imsb = 1, isymb = 0, n = 0;
        if (ij_ok) {
            for (;;) {
                a = (int)(rng.next() * na) + 1;
                a = fac * a - shift;
                if (!det_test(f, a)) break;
            }
            imsb = (spin - ms_of(a) + 3) / 2;
            isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
            n = nbss(s, imsb, isymb);
        }
        if (n == 0) {
            g.allowed = false;
        } else if (spin != 0 && isymb == s.bf_sym[a] && n == 1) {
            g.allowed = false;
        } else {
            for (;;) {
                int ind = (int)(n * rng.next()) + 1;
                b = ssbf(s, ind, imsb, isymb);
                if (b != a) break;
            }
            g.allowed = !det_test(f, b);
            if (a > b) { int t = a; a = b; b = t; }
        }
        if (g.allowed) {
            g.to1 = a; g.to2 = b;
            int n_aij = (spin == -2) ? s.nvirt_beta : (spin == 0 ? s.nvirt : s.nvirt_alpha);
            int imsa = ims_of(a), isyma = s.bf_sym[a];
            int imsb2 = ims_of(b), isymb2 = s.bf_sym[b];
            double p_aijb, p_bija;
            if (isyma == isymb2 && imsa == imsb2) {
                p_aijb = 1.0 / (nbss(s, imsa, isyma) - 1);
                p_bija = 1.0 / (nbss(s, imsb2, isymb2) - 1);
            } else {
                p_aijb = 1.0 / nbss(s, imsa, isyma);
                p_bija = 1.0 / nbss(s, imsb2, isymb2);
            }
            g.pgen = p.pattempt_double * pgen_ij * ((1.0 / n_aij) * (p_bija + p_aijb));
            g.perm = excit_perm2<W>(f, i, j, a, b);
            g.hmatel = slater_condon2_excit(s, i, j, a, b, g.perm);
        } else {
            g.hmatel = 0.0; g.pgen = 1.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Alias method (lib/local/alias.f90:13-184).  N <= HB_MAXNEL for the on-the-fly tables.
// ------------------------------------------------------------------------------------------------
HB_HDN void generate_alias_tables(int N, const double* weights, double totweight, double* aliasU, int* aliasK,
                                  int* underfull, int* overfull) {
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
// STREAM: evict-first loads for tables with no reuse (the 2 GB hb_ijab rows); the 30 MB hb_ija rows stay L2-resident
template <bool STREAM = false, class R>
HB_HD int select_precalc(R& rng, int N, const double* aliasU, const int* aliasK) {
    double x = rng.next() * N;
    int K = (int)floor(x);
    x = x - K;
    // both loads are issued together: one memory round trip instead of two
    const double u = STREAM ? HB_LDCS(aliasU + K) : aliasU[K];
    const int alias = STREAM ? HB_LDCS(aliasK + K) : aliasK[K];
    return (x < u) ? K + 1 : alias;
}
template <class R>
HB_HDN int select_weighted_value(R& rng, int N, const double* weights, double totweight) {
    double aliasU[HB_MAXNEL];
    int aliasK[HB_MAXNEL], under[HB_MAXNEL], over[HB_MAXNEL];
    generate_alias_tables(N, weights, totweight, aliasU, aliasK, under, over);
    return select_precalc(rng, N, aliasU, aliasK);
}

// On-the-fly alias table over the occupied orbitals (select_weighted_value, lib/local/alias.f90:68-184) with
// weights w(q) = tab[occ[q]-1], N = nel <= 64.  Equivalent to generate_alias_tables: the overfull stack is only
// ever popped, so it is a bitmask read from its highest bit; the underfull stack is its initial ascending content
// (a bitmask) plus at most one element pushed on top (the overfull entry that just dropped below one).
HB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
// sum_q tab[occ[q]-1] in occ_list order; loads issued four at a time (adding the 0.0 padding is exact)
// byte lists whose length is a multiple of four and that start on a word boundary (the spawn kernel's staged lists) are
// read four orbitals to a 32-bit word
HB_HD bool occ_words_ok(const occ_t* occ, int nel) {
    return sizeof(occ_t) == 1 && (nel & 3) == 0 && (((size_t)occ) & 3) == 0;
}
HB_HD double sum_occ(const double* __restrict__ tab, const occ_t* occ, int nel) {
    double tot = 0.0;
    if (occ_words_ok(occ, nel)) {
        const uint32_t* ow = reinterpret_cast<const uint32_t*>(occ);
        const double* __restrict__ tab1 = tab - 1;
        for (int q0 = 0; q0 < nel; q0 += 4) {
            const uint32_t o4 = ow[q0 >> 2];
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = tab1[(o4 >> (8 * k)) & 0xffu];
#pragma unroll
            for (int k = 0; k < 4; ++k) tot = tot + v[k];
        }
        return tot;
    }
    for (int q0 = 0; q0 < nel; q0 += 4) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (q0 + k < nel) ? tab[occ[q0 + k] - 1] : 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) tot = tot + v[k];
    }
    return tot;
}
// select_weighted_value (lib/local/alias.f90:68-184) for the on-the-fly tables over the occupied orbitals, N <= 64.
// wq[q*stride] holds the UNSCALED weight of list position q (staged by the caller, who also formed their total in list
// order); scale = N / totweight.  The alias construction consumes no random numbers, so the single draw of
// select_weighted_value_precalc is taken first and only slot k's final (aliasU, aliasK) are tracked - no table is stored.
// Values evolve exactly as in generate_alias_tables: underfull entries are never modified; an overfull entry is modified
// only while it is on top of the overfull stack (kept in a register) and, once it drops below one, it is pushed on the
// underfull stack and therefore popped by the very next step (the "carry").  Both stacks start in ascending index order
// and are only ever popped, so they are bitmasks read from their highest bit.
template <class Mask>
HB_HD int mask_top(Mask x);
template <>
HB_HD int mask_top<uint32_t>(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
template <>
HB_HD int mask_top<uint64_t>(uint64_t x) { return 63 - clz64(x); }

// Exact walk for a given slot k and fractional part x (the reference's stack algorithm, see above).
template <class Mask>
HB_HDNI int alias_walk_exact(int N, const double* wq, int stride, double scale, int k, double x) {
    Mask under = 0, over = 0;
    for (int q = 0; q < N; ++q) {
        const double u = wq[q * stride] * scale;
        if (u <= 1.0) under |= ((Mask)1 << q); else over |= ((Mask)1 << q);
    }
    double Uk = wq[k * stride] * scale;    // final aliasU(k) unless k is demoted from the overfull stack
    int Kk = k;                            // aliasK(k) - 1
    if (over != 0) {
        int ov = mask_top<Mask>(over);
        double uov = wq[ov * stride] * scale;
        bool carry = false;
        int cq = 0;
        double cu = 0.0;
        for (;;) {
            int un;
            double uun;
            if (carry) { un = cq; uun = cu; }
            else {
                if (under == 0) break;
                un = mask_top<Mask>(under);
                under ^= ((Mask)1 << un);
                uun = wq[un * stride] * scale;
            }
            if (un == k) Kk = ov;
            uov = uov - (1 - uun);
            carry = uov < 1.0;
            if (carry) {
                if (ov == k) Uk = uov;
                cq = ov; cu = uov;
                over ^= ((Mask)1 << ov);
                if (over == 0) break;
                ov = mask_top<Mask>(over);
                uov = wq[ov * stride] * scale;
            }
        }
    }
    return (x < Uk) ? k + 1 : Kk + 1;
}
// Selection without the serial walk.  In exact arithmetic the walk is a merge of two running sums: with the underfull
// entries in pop order (descending index) having deficits d_t = 1 - u and the overfull ones (descending index) excesses
// e_t = u - 1, overfull entry n is demoted right after absorbing the first underfull entry m with D_m > E_n (D, E =
// running sums), the alias of underfull entry m is the first overfull entry n with E_n >= D_(m-1), a demoted overfull
// entry ends with aliasU = 1 + E_n - D_m and its alias is the next overfull entry.  Only slot k matters, so:
//   k underfull: x < u_k returns k at once; otherwise D = deficits of the underfull entries above k, then the overfull
//                entries are scanned from the top until their running excess exceeds D;
//   k overfull:  E = excesses of the overfull entries from the top down to k, then the underfull entries are scanned from
//                the top until their running deficit exceeds E.
// The reference evaluates the same quantities through a different sequence of roundings (|difference| < 1e-12 for
// N <= 64), so every comparison that decides the result is taken only when it is clear by `guard` = 1e-9; otherwise the
// exact walk above is run (about 1e-8 of the draws; exact ties, e.g. integer weights, always).  Same index as
// generate_alias_tables + select_weighted_value_precalc for every input (tests/test_core_vs_oracle.py: 24 M draws
// over seven weight families incl. ties and draws aimed at table boundaries).
template <class Mask>
HB_HD int alias_select_fast(int N, const double* wq, int stride, double scale, int k, double x) {
    const double guard = 1.e-9;
    // pass 1: classification only
    Mask under = 0;
#pragma unroll 4
    for (int q = 0; q < N; ++q) {
        const double u = wq[q * stride] * scale;
        under |= (u <= 1.0) ? ((Mask)1 << q) : (Mask)0;
    }
    const Mask all = (N >= (int)(8 * sizeof(Mask))) ? ~(Mask)0 : (((Mask)1 << N) - 1);
    const Mask over = all & ~under;
    const Mask kbit = (Mask)1 << k;
    const bool k_under = (under & kbit) != 0;
    const double uk = wq[k * stride] * scale;
    const Mask below = over & (kbit - 1);                   // overfull entries further down the stack than k
    int result = k + 1;
    bool done = k_under ? (x < uk) : (below == 0);          // the lowest overfull entry keeps (or aliases to) itself
    if (!done) {
        // pass 2a: the running sum that belongs to slot k (its own class, from the top of the stack down to k)
        Mask m = (k_under ? under : over) & ~(kbit | (kbit - 1));
        double target = k_under ? 0.0 : (uk - 1.0);
        while (m != 0) {
            const int q = mask_top<Mask>(m);
            m ^= (Mask)1 << q;
            target += fabs(wq[q * stride] * scale - 1.0);
        }
        // pass 2b: the other class from the top until its running sum passes `target`
        m = k_under ? over : under;
        double acc = 0.0;
        int hit = -1;
        bool unclear = false;
        while (m != 0) {
            const int q = mask_top<Mask>(m);
            m ^= (Mask)1 << q;
            acc += fabs(wq[q * stride] * scale - 1.0);
            const double d = acc - target;
            // k underfull: the first overfull entry with E >= D survives to alias k; k overfull: demoted once D > E
            if (d > guard) { hit = q; break; }
            if (d >= -guard) { unclear = true; break; }
        }
        if (unclear) {
            result = alias_walk_exact<Mask>(N, wq, stride, scale, k, x);
        } else if (k_under) {
            if (hit >= 0) result = hit + 1;                 // else the overfull entries ran out: aliasK(k) stays k
        } else if (hit >= 0) {                              // else never demoted: aliasU(k) >= 1 > x
            const double Uk = 1.0 + (target - acc);
            if (fabs(x - Uk) <= guard) result = alias_walk_exact<Mask>(N, wq, stride, scale, k, x);
            else if (!(x < Uk)) result = mask_top<Mask>(below) + 1;
        }
    }
    return result;
}
// The same selection with fixed trip counts (no per-lane loops over mask bits, whose length is the warp maximum): ONE
// top-down pass classifies every entry, forms slot k's own running sum (`target`, the entries of its class above it)
// and overwrites wq[q] with the running sum of the OTHER class down to q - non-decreasing along the pass, so the first
// entry whose running sum reaches `target` is found by a binary search instead of a second pass.  Sums are formed in
// the order of alias_select_fast (descending index within a class), hence the same values and the same decisions.
// CLOBBERS the staged weights; returns 0 where alias_select_fast would fall back to the exact walk (the caller stages the
// weights again and calls alias_walk_exact).
template <class Mask>
HB_HD int alias_select_scan(int N, double* wq, int stride, double scale, int k, double x) {
    const double guard = 1.e-9;
    const double uk = wq[k * stride] * scale;
    const bool k_under = uk <= 1.0;
    Mask under = 0;
    double target = k_under ? 0.0 : (uk - 1.0);
    double acc = 0.0;
#pragma unroll 4
    for (int q = N - 1; q >= 0; --q) {
        const double u = wq[q * stride] * scale;
        const bool un = u <= 1.0;
        under |= un ? ((Mask)1 << q) : (Mask)0;
        const double d = fabs(u - 1.0);
        const bool own = (un == k_under);
        if (own && q > k) target += d;
        if (!own) acc += d;
        wq[q * stride] = acc;
    }
    const Mask all = (N >= (int)(8 * sizeof(Mask))) ? ~(Mask)0 : (((Mask)1 << N) - 1);
    const Mask over = all & ~under;
    const Mask kbit = (Mask)1 << k;
    const Mask below = over & (kbit - 1);
    if (k_under ? (x < uk) : (below == 0)) return k + 1;
    const Mask other = k_under ? over : under;
    if (other == 0) return k + 1;
    // largest q with (running sum - target) >= -guard; the running sum at q = 0 is the class total
    if (!(wq[0] - target >= -guard)) return k + 1;          // the other class runs out first
    int lo = 0, hi = N - 1;                                  // invariant: predicate true at lo
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (wq[mid * stride] - target >= -guard) lo = mid; else hi = mid - 1;
    }
    // the first entry of the other class at or below lo (the running sum is constant across entries of k's own class)
    const Mask cand = other & ((lo >= (int)(8 * sizeof(Mask)) - 1) ? ~(Mask)0 : ((((Mask)1 << lo) << 1) - 1));
    const int hit = mask_top<Mask>(cand);
    const double acc_hit = wq[hit * stride];
    const double d = acc_hit - target;
    if (!(d > guard)) return 0;                              // within the guard band: exact walk
    if (k_under) return hit + 1;
    const double Uk = 1.0 + (target - acc_hit);
    if (fabs(x - Uk) <= guard) return 0;
    return (x < Uk) ? k + 1 : mask_top<Mask>(below) + 1;
}
template <class Mask, class R>
HB_HD int select_alias_staged_m(R& rng, int N, const double* wq, int stride, double totweight) {
    double x = rng.next() * N;
    const int k = (int)x;                  // floor: x >= 0
    x = x - k;
    return alias_select_fast<Mask>(N, wq, stride, N / totweight, k, x);
}
template <class R>
HB_HD int select_alias_staged(R& rng, int N, const double* wq, int stride, double totweight) {
    if (N <= 32) return select_alias_staged_m<uint32_t>(rng, N, wq, stride, totweight);
    return select_alias_staged_m<uint64_t>(rng, N, wq, stride, totweight);
}
// gather tab[occ[q]-1] into wq[q*stride] and return their sum in list order; loads issued four at a time
HB_HD double stage_occ(const double* __restrict__ tab, const occ_t* occ, int nel, double* wq, int stride) {
    double tot = 0.0;
    if (occ_words_ok(occ, nel)) {
        const uint32_t* ow = reinterpret_cast<const uint32_t*>(occ);
        const double* __restrict__ tab1 = tab - 1;
        for (int q0 = 0; q0 < nel; q0 += 4) {
            const uint32_t o4 = ow[q0 >> 2];
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = tab1[(o4 >> (8 * k)) & 0xffu];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                wq[(q0 + k) * stride] = v[k];
                tot = tot + v[k];
            }
        }
        return tot;
    }
    for (int q0 = 0; q0 < nel; q0 += 4) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (q0 + k < nel) ? tab[occ[q0 + k] - 1] : 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (q0 + k < nel) wq[(q0 + k) * stride] = v[k];
            tot = tot + v[k];
        }
    }
    return tot;
}

// gen_excit_mol_heat_bath (src/excit_gen_heat_bath_mol.F90:258-548; src/excit_gen_utils.f90:9-66,142-160),
// split into phases so that the kernel can regroup the work of a tile between them (the per-attempt control flow
// diverges strongly: ~20 % null excitations, spin-conditional slater_condon1 evaluations, a division-heavy pgen
// loop for singles).  The monolithic gen_excit_heat_bath below chains the phases for one attempt and is what the
// CPU parity harness and the probe kernel run; k_spawn_death calls the same phase functions with block-level
// queues in between.  The per-determinant weight lists (i_d_occ%weights, ij_weights_occ, ji_weights_occ) are columns
// of hb_i_w / hb_ij_w gathered at the occupied orbitals and are re-read where needed.
// select_weighted_value over the staged weights wq[q] = tab[occ[q]-1] with the fixed-trip-count scan; the staged weights
// are consumed (clobbered), and staged again for the exact walk in the rare guard-band case
template <class R>
HB_HD int select_alias_restage(R& rng, int N, const double* __restrict__ tab, const occ_t* occ, double* wq, int stride,
                               double totweight) {
    double x = rng.next() * N;
    const int k = (int)x;                  // floor: x >= 0
    x = x - k;
    const double scale = N / totweight;
    int r = (N <= 32) ? alias_select_scan<uint32_t>(N, wq, stride, scale, k, x) : alias_select_scan<uint64_t>(N, wq, stride, scale, k, x);
    if (r == 0) {
        stage_occ(tab, occ, N, wq, stride);
        r = (N <= 32) ? alias_walk_exact<uint32_t>(N, wq, stride, scale, k, x) : alias_walk_exact<uint64_t>(N, wq, stride, scale, k, x);
    }
    return r;
}
#define HB_I2(j, i) ((int64_t)((j) - 1) + nb * ((i) - 1))
#define HB_I3(a, j, i) ((int64_t)((a) - 1) + nb * (((j) - 1) + nb * ((i) - 1)))
#define HB_I4(b, a, j, i) ((int64_t)((b) - 1) + nb * (((a) - 1) + nb * (((j) - 1) + nb * ((i) - 1))))

struct HbState {
    int i, j, a, b;
    double i_tot, ij_tot, ji_tot;
    double psingle, hmod_ia, h_ia;   // h_ia = signed <D|H|D_i^a> (with permutation) when need_ia
    bool allowed;                    // still a candidate excitation
    bool need_ia;                    // i -> a is spin/symmetry allowed: slater_condon1(i,a) required
    bool dbl;                        // double excitation chosen
    bool perm_ia;
    unsigned need_k;                 // bit k set: ordering k (0: i->b, 1: j->a, 2: j->b) needs slater_condon1
};

// single -> (from,to,other) of ordering k of the pgen sum
HB_HD void hb_ordering(const HbState& st, int k, int& fr, int& to, int& ot) {
    fr = (k == 0) ? st.i : st.j;
    to = (k == 1) ? st.a : st.b;
    ot = (k == 0) ? st.j : st.i;
}
HB_HD bool hb_single_allowed(const Sys& s, int fr, int to) {
    const int isyma = cross_product(s, s.bf_sym[fr], s.gamma_sym);
    return s.bf_sym[to] == isyma && ms_of(to) == ms_of(fr);
}

// Phase A: select i, j (on-the-fly alias tables) and a (precomputed alias table); 2-3 random numbers.
// iw = hb_i_w (or a shared-memory copy); scr/stride = staging area of nel doubles for the weight list being selected
// from (first S_i at the occupied orbitals, then column i of hb_ij_w at the occupied orbitals).
template <int W, class R>
HB_HDN void hb_phase_a(R& rng, const Sys& s, const uint64_t* f, const occ_t* occ, HbState& st,
                       const double* __restrict__ iw, double* scr, int stride) {
    const int nel = s.nel;
    const int64_t nb = s.nbasis;
    st.ij_tot = 0.0; st.ji_tot = 0.0;
    st.j = 0; st.a = 0; st.b = 0;
    st.psingle = 0.0; st.hmod_ia = 0.0; st.h_ia = 0.0; st.perm_ia = false;
    st.dbl = true; st.need_ia = false; st.need_k = 0; st.allowed = false;
    st.i_tot = stage_occ(iw, occ, nel, scr, stride);
    st.i = occ[select_alias_restage(rng, nel, iw, occ, scr, stride, st.i_tot) - 1];
    st.ij_tot = stage_occ(s.hb_ij_w + nb * (st.i - 1), occ, nel, scr, stride);
    if (st.ij_tot > 0.0) {
        st.j = occ[select_alias_restage(rng, nel, s.hb_ij_w + nb * (st.i - 1), occ, scr, stride, st.ij_tot) - 1];
        st.allowed = fabs(s.hb_ija_tot[HB_I2(st.j, st.i)]) > 0.0;
    }
    if (st.allowed) {
        st.a = select_precalc(rng, (int)nb, s.hb_ija_U + HB_I3(1, st.j, st.i), s.hb_ija_K + HB_I3(1, st.j, st.i));
        if (det_test(f, st.a)) st.allowed = false;
        else st.need_ia = hb_single_allowed(s, st.i, st.a);
    }
}

// slater_condon1 request of the tile queues: |<D|H|D_fr^to>| with its sign
template <int W>
HB_HD double hb_sc1(const Sys& s, const uint64_t* f, const occ_t* occ, int fr, int to, bool& perm) {
    perm = excit_perm1<W>(f, fr, to);
    if (s.sc1T) {
        // slater_condon1_mol_excit through the branch-free rows (Sys::sc1T): entry j = {<ij|aj>, <ij|ja> or 0}; the entry
        // of the excited orbital itself and entry 0 (list padding) are {0, 0}, so every partial sum equals the
        // reference's, in occ_list order.  Loads issued four at a time.
        const unsigned ta = s.uhf ? (unsigned)(to - 1) : ((unsigned)(to - 1) >> 1);
        const D2* __restrict__ row = s.sc1T + ((size_t)((unsigned)(fr - 1) * (unsigned)s.sc1A + ta)) * (unsigned)(s.nbasis + 1);
        double h = one_body(s, fr, to);
        const int nel = s.nel;
        if (occ_words_ok(occ, nel)) {
            const uint32_t* ow = reinterpret_cast<const uint32_t*>(occ);
            for (int q0 = 0; q0 < nel; q0 += 4) {
                const uint32_t o4 = ow[q0 >> 2];
                D2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = row[(o4 >> (8 * k)) & 0xffu];
#pragma unroll
                for (int k = 0; k < 4; ++k) { h = h + v[k].x; h = h - v[k].y; }
            }
            return perm ? -h : h;
        }
        for (int q0 = 0; q0 < nel; q0 += 4) {
            D2 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = row[(q0 + k < nel) ? occ[q0 + k] : 0];
#pragma unroll
            for (int k = 0; k < 4; ++k) { h = h + v[k].x; h = h - v[k].y; }
        }
        return perm ? -h : h;
    }
    return slater_condon1_excit(s, occ, fr, to, perm);
}

// Phase C: single/double coin (1 random number iff need_ia), selection of b (1 random number iff double).
template <int W, class R>
HB_HDN void hb_phase_c(R& rng, const Sys& s, const uint64_t* f, HbState& st) {
    if (!st.allowed) return;
    const int64_t nb = s.nbasis;
    if (st.need_ia) {
        st.hmod_ia = fabs(st.h_ia);
        const double x = rng.next();
        const double wt = s.hb_ijab_tot[HB_I3(st.a, st.j, st.i)];
        if (st.hmod_ia < wt) st.psingle = st.hmod_ia / (wt + st.hmod_ia); else st.psingle = 0.5;
        st.dbl = !(x < st.psingle);
    } else {
        st.dbl = true; st.psingle = 0.0;
    }
    if (st.dbl) {
        if (s.hb_ijab_rec) {      // {aliasU, aliasK} of the drawn slot from one 32-byte record instead of two arrays
            double x = rng.next() * (int)nb;
            const int K = (int)floor(x);
            x = x - K;
            const HbRec* rec = s.hb_ijab_rec + HB_I4(1, st.a, st.j, st.i) + K;
            const double U = HB_LDCS(&rec->U);
            const int alias = HB_LDCS(&rec->K);
            st.b = (x < U) ? K + 1 : alias;
        } else {
            st.b = select_precalc<true>(rng, (int)nb, s.hb_ijab_U + HB_I4(1, st.a, st.j, st.i), s.hb_ijab_K + HB_I4(1, st.a, st.j, st.i));
        }
        if (det_test(f, st.b)) { st.allowed = false; return; }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int fr, to, ot;
            hb_ordering(st, k, fr, to, ot);
            if (hb_single_allowed(s, fr, to)) st.need_k |= (1u << k);
        }
    }
}

// One term of the generation probability of a single excitation (the q-th occupied orbital as spectator j).
// Returns false when the term is skipped (oq == i or oq == a).
HB_HD bool hb_single_term(const Sys& s, int i, int a, double hmod_ia, double ij_tot, int oq, double& term) {
    if (i == oq || a == oq) return false;
    const int64_t nb = s.nbasis;
    // hb_ijab%weights_tot(a,oq,i) and hb_ija%weights(a,oq,i) are the same number (the same sum): one load
    if (s.hb_ija_rec) {
        // packed record: {w = hb_ija%weights(a,oq,i), p = w / hb_ija%weights_tot(oq,i)} - the same division, done once
        // when the tables were built
        const HbRec* rec = s.hb_ija_rec + HB_I3(a, oq, i);
        const double wt = rec->w, pa = rec->p;
        double psq;
        if (hmod_ia < wt) psq = hmod_ia / (wt + hmod_ia); else psq = 0.5;
        term = (psq * (s.hb_ij_w[HB_I2(oq, i)] / ij_tot) * pa);
        return true;
    }
    const double wt = s.hb_ija_w[HB_I3(a, oq, i)];
    double psq;
    if (hmod_ia < wt) psq = hmod_ia / (wt + hmod_ia); else psq = 0.5;
    term = (psq * (s.hb_ij_w[HB_I2(oq, i)] / ij_tot) * (wt / s.hb_ija_tot[HB_I2(oq, i)]));
    return true;
}

// Phase F: generation probability, matrix element and excitation.  hm[k] = |slater_condon1| of ordering k (only
// read where need_k has the bit); pgen_single_sum = sum of hb_single_term over the occupied orbitals (singles).
// The tables are exactly symmetric - ij_weights(j,i) = ij_weights(i,j), hb_ija%weights_tot(j,i) = ...(i,j),
// hb_ija%weights(a,j,i) = hb_ijab%weights_tot(a,j,i) = ...(a,i,j), hb_ijab%weights invariant under i<->j and a<->b: the
// same sums of the same numbers (checked bitwise in tests/test_core_vs_oracle.py) - so the four-ordering formula below
// reads 7 numbers where it names 20, with every operation of the reference kept in its order.  ji_weights_occ_tot is
// only needed here and is summed for the double excitations that got this far.
template <int W>
HB_HDN void hb_phase_f(const Sys& s, const uint64_t* f, const occ_t* occ, const HbState& st, const double* hm,
                       double pgen_single_sum, const double* __restrict__ iw, Gen& g) {
    const int64_t nb = s.nbasis;
    g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false; g.nexcit = 2;
    if (!st.allowed) {
        g.allowed = false; g.hmatel = 0.0; g.pgen = 1.0;
        return;
    }
    const int i = st.i, j = st.j, a = st.a, b = st.b;
    if (st.dbl) {
        const double ji_tot = sum_occ(s.hb_ij_w + nb * (j - 1), occ, s.nel);
        const double wij = s.hb_ij_w[HB_I2(j, i)];
        const double Tij = s.hb_ija_tot[HB_I2(j, i)];
        const double Ta = s.hb_ija_w[HB_I3(a, j, i)], Tb = s.hb_ija_w[HB_I3(b, j, i)];
        const double wab = HB_LDCS(s.hb_ijab_w + HB_I4(b, a, j, i));
        const double pi_ = iw[i - 1] / st.i_tot;
        const double pj_ = iw[j - 1] / st.i_tot;
        const double pij = wij / st.ij_tot;      // ij_weights_occ(j_ind)/ij_weights_occ_tot
        const double pji = wij / ji_tot;         // ji_weights_occ(i_ind)/ji_weights_occ_tot
        double ps[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (st.need_k & (1u << k)) {
                const double wt = (k == 1) ? Ta : Tb;    // weights_tot(to, other, from) of ordering k
                if (hm[k] < wt) ps[k] = hm[k] / (wt + hm[k]); else ps[k] = 0.5;
            } else {
                ps[k] = 0.0;
            }
        }
        const double ra = Ta / Tij, rb = Tb / Tij, wa = wab / Ta, wb = wab / Tb;
        double pgen_ija = ((pi_) * (pij)) * ra * (1.0 - st.psingle) * wa;
        double pgen_ijb = ((pi_) * (pij)) * rb * (1.0 - ps[0]) * wb;
        double pgen_jia = ((pj_) * (pji)) * ra * (1.0 - ps[1]) * wa;
        double pgen_jib = ((pj_) * (pji)) * rb * (1.0 - ps[2]) * wb;
        g.pgen = pgen_ija + pgen_ijb + pgen_jia + pgen_jib;
        g.from1 = (i < j) ? i : j; g.from2 = (i < j) ? j : i;
        g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
        g.nexcit = 2;
        g.perm = excit_perm2<W>(f, g.from1, g.from2, g.to1, g.to2);
        g.hmatel = slater_condon2_excit(s, g.from1, g.from2, g.to1, g.to2, g.perm);
        g.allowed = true;
    } else {
        g.nexcit = 1; g.from1 = i; g.to1 = a; g.perm = st.perm_ia;
        g.hmatel = st.h_ia;
        g.pgen = pgen_single_sum * (iw[i - 1] / st.i_tot);
        g.allowed = true;
    }
}

template <int W, class R>
HB_HDN void gen_excit_heat_bath(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                Gen& g) {
    HbState st;
    double scr[HB_MAXNEL];
    hb_phase_a<W>(rng, s, f, occ, st, s.hb_i_w, scr, 1);
    if (st.allowed && st.need_ia) st.h_ia = hb_sc1<W>(s, f, occ, st.i, st.a, st.perm_ia);
    hb_phase_c<W>(rng, s, f, st);
    double hm[3] = {0.0, 0.0, 0.0};
    double psum = 0.0;
    if (st.allowed) {
        if (st.dbl) {
            for (int k = 0; k < 3; ++k)
                if (st.need_k & (1u << k)) {
                    int fr, to, ot;
                    hb_ordering(st, k, fr, to, ot);
                    bool pm;
                    hm[k] = fabs(hb_sc1<W>(s, f, occ, fr, to, pm));
                }
        } else {
            for (int q = 0; q < s.nel; ++q) {
                double t;
                if (hb_single_term(s, st.i, st.a, st.hmod_ia, st.ij_tot, occ[q], t)) psum = psum + t;
            }
        }
    }
    hb_phase_f<W>(s, f, occ, st, hm, psum, s.hb_i_w, g);
}

// k-th (1-based) unoccupied orbital, ascending (decode_det_occ_unocc, src/determinant_decoders.f90)
template <int W>
HB_HD int nth_unocc(const uint64_t* f, int k) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint64_t x = ~f[w];
        const int c = popc64(x);
        if (k <= c) {
            for (int q = 1; q < k; ++q) x &= x - 1;
            return w * 64 + ctz64(x) + 1;
        }
        k -= c;
    }
    return 0;
}
// select_weighted_value for lists of up to 2*HB_MAXLIST entries (explicit tables as in lib/local/alias.f90)
template <class R>
HB_HDN int select_weighted_value_big(R& rng, int N, const double* weights, double totweight) {
    double aliasU[2 * HB_MAXLIST];
    int aliasK[2 * HB_MAXLIST], under[2 * HB_MAXLIST], over[2 * HB_MAXLIST];
    generate_alias_tables(N, weights, totweight, aliasU, aliasK, under, over);
    return select_precalc(rng, N, aliasU, aliasK);
}
// gen_single_excit_heat_bath_exact (src/excit_gen_heat_bath_mol.F90:720-805) with find_ia_single_weights
// (src/excit_gen_utils.f90:162-218): weights |<D|H|D_i^a>| over all (occupied i, unoccupied a); i then a are drawn with
// on-the-fly alias tables.  O(N M) slater_condon1 evaluations per call - correctness-first (the reference caches the
// weights per determinant; here they are recomputed per single-excitation attempt).
template <int W, class R>
HB_HDN void gen_single_heat_bath_exact(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ, Gen& g) {
    const int nel = s.nel, nvirt = s.nbasis - s.nel;
    g.nexcit = 1; g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    double wi[HB_MAXNEL];
    double tot = 0.0;
    // slater_condon1_mol (src/hamiltonian_molecular.f90:141-197) is zero unless i and a share spin and symmetry; for the
    // allowed pairs it is the unchecked sum (subtracting the 0.0 of a forbidden exchange integral changes nothing)
    for (int q = 0; q < nel; ++q) {
        const int i = occ[q];
        double wsum = 0.0;
        for (int k = 1; k <= nvirt; ++k) {
            const int a = nth_unocc<W>(f, k);
            double w = 0.0;
            if (((a ^ i) & 1) == 0 && s.bf_sym[a] == s.bf_sym[i]) w = fabs(slater_condon1_excit(s, occ, i, a, false));
            wsum = wsum + w;
        }
        wi[q] = wsum;
        tot = tot + wsum;
    }
    if (tot < 1.e-12) {
        g.allowed = false; g.hmatel = 0.0; g.pgen = 1.0;
        return;
    }
    const int i_ind = select_weighted_value_list(rng, nel, wi, tot);
    const int i = occ[i_ind - 1];
    double wa[HB_MAXLIST * 2];
    for (int k = 1; k <= nvirt; ++k) {
        const int a = nth_unocc<W>(f, k);
        double w = 0.0;
        if (((a ^ i) & 1) == 0 && s.bf_sym[a] == s.bf_sym[i]) w = fabs(slater_condon1_excit(s, occ, i, a, false));
        wa[k - 1] = w;
    }
    const int a_ind = select_weighted_value_big(rng, nvirt, wa, wi[i_ind - 1]);
    const int a = nth_unocc<W>(f, a_ind);
    g.from1 = i; g.to1 = a;
    g.pgen = p.pattempt_single * (wi[i_ind - 1] / tot) * (wa[a_ind - 1] / wi[i_ind - 1]);
    g.perm = excit_perm1<W>(f, i, a);
    g.hmatel = slater_condon1_excit(s, occ, i, a, g.perm);
    g.allowed = true;
}

// One weight of find_ia_single_weights: |<D|H|D_i^a>| for an allowed (same spin, same symmetry) pair, else 0
HB_HD double hb_single_weight(const Sys& s, const occ_t* occ, int i, int a) {
    if (((a ^ i) & 1) == 0 && s.bf_sym[a] == s.bf_sym[i]) {
        if (s.sc1T) {     // the branch-free rows (see hb_sc1): the same partial sums in occ_list order
            const unsigned ta = s.uhf ? (unsigned)(a - 1) : ((unsigned)(a - 1) >> 1);
            const D2* __restrict__ row = s.sc1T + ((size_t)((unsigned)(i - 1) * (unsigned)s.sc1A + ta)) * (unsigned)(s.nbasis + 1);
            double h = one_body(s, i, a);
            const int nel = s.nel;
            for (int q0 = 0; q0 < nel; q0 += 4) {
                D2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = row[(q0 + k < nel) ? occ[q0 + k] : 0];
#pragma unroll
                for (int k = 0; k < 4; ++k) { h = h + v[k].x; h = h - v[k].y; }
            }
            return fabs(h);
        }
        return fabs(slater_condon1_excit(s, occ, i, a, false));
    }
    return 0.0;
}
// gen_single_excit_heat_bath_exact once the weights are known: wi[q] = sum over a of w(q, a) (a ascending), wall =
// the nel x nvirt weights, unocc = the unoccupied orbitals ascending.  Draws i then a (two random numbers).
template <int W, class R, class U>
HB_HDN void gen_single_heat_bath_select(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                        const double* wi, const double* wall, const U* unocc, Gen& g) {
    const int nel = s.nel, nvirt = s.nbasis - s.nel;
    g.nexcit = 1; g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    double tot = 0.0;
    for (int q = 0; q < nel; ++q) tot = tot + wi[q];
    if (tot < 1.e-12) {
        g.allowed = false; g.hmatel = 0.0; g.pgen = 1.0;
        return;
    }
    const int i_ind = select_weighted_value_list(rng, nel, wi, tot);
    const int i = occ[i_ind - 1];
    const double* wa = wall + (size_t)(i_ind - 1) * nvirt;
    const int a_ind = select_weighted_value_big(rng, nvirt, wa, wi[i_ind - 1]);
    const int a = unocc[a_ind - 1];
    g.from1 = i; g.to1 = a;
    g.pgen = p.pattempt_single * (wi[i_ind - 1] / tot) * (wa[a_ind - 1] / wi[i_ind - 1]);
    g.perm = excit_perm1<W>(f, i, a);
    g.hmatel = slater_condon1_excit(s, occ, i, a, g.perm);
    g.allowed = true;
}

// gen_excit_mol_heat_bath_uniform (src/excit_gen_heat_bath_mol.F90:550-718), excit_gen = heat_bath_uniform: single
// excitations from the renormalised uniform generator with probability pattempt_single, double excitations from the
// heat-bath tables (i, j, a, b in turn); the generation probability of a double needs no matrix elements.
// EXACT_SINGLE: excit_gen = heat_bath_single (src/excit_gen_heat_bath_mol.F90:720-805) - the same doubles, singles
// from gen_single_heat_bath_exact.
template <int W, class R>
HB_HDN void gen_double_heat_bath_uniform(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                         const double* __restrict__ iw, double* scr, int stride, Gen& g);
template <int W, bool EXACT_SINGLE = false, class R>
HB_HDN void gen_excit_heat_bath_uniform(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                        const uint8_t* su, const double* __restrict__ iw, double* scr, int stride, Gen& g) {
    const int nel = s.nel;
    const int64_t nb = s.nbasis;
    g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    if (rng.next() < p.pattempt_single) {
        if (EXACT_SINGLE) gen_single_heat_bath_exact<W>(rng, s, p, f, occ, g);
        else gen_single_renorm<W>(rng, s, p, f, occ, su, g);
        return;
    }
    gen_double_heat_bath_uniform<W>(rng, s, p, f, occ, iw, scr, stride, g);
}
// the double excitation of heat_bath_uniform / heat_bath_single (after the single/double coin)
template <int W, class R>
HB_HDN void gen_double_heat_bath_uniform(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                         const double* __restrict__ iw, double* scr, int stride, Gen& g) {
    const int nel = s.nel;
    const int64_t nb = s.nbasis;
    g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    g.nexcit = 2;
    const double i_tot = stage_occ(iw, occ, nel, scr, stride);
    const int iq = select_alias_staged(rng, nel, scr, stride, i_tot);
    int i = occ[iq - 1];
    const double wi = scr[(iq - 1) * stride];
    const double ij_tot = stage_occ(s.hb_ij_w + nb * (i - 1), occ, nel, scr, stride);
    bool allowed = false;
    int j = 0;
    double pgen = 0.0;
    if (ij_tot > 0.0) {
        const int jq = select_alias_staged(rng, nel, scr, stride, ij_tot);
        j = occ[jq - 1];
        const double wij = scr[(jq - 1) * stride];
        const double ji_tot = sum_occ(s.hb_ij_w + nb * (j - 1), occ, nel);
        allowed = fabs(s.hb_ija_tot[HB_I2(j, i)]) > 0.0;
        if (allowed)
            pgen = ((wi / i_tot) * (wij / ij_tot)) + ((iw[j - 1] / i_tot) * (s.hb_ij_w[HB_I2(i, j)] / ji_tot));
    }
    if (j < i) { const int t = i; i = j; j = t; }
    int a = 0, b = 0;
    if (allowed) {
        a = select_precalc(rng, (int)nb, s.hb_ija_U + HB_I3(1, j, i), s.hb_ija_K + HB_I3(1, j, i));
        if (fabs(s.hb_ijab_tot[HB_I3(a, j, i)]) > 0.0 && !det_test(f, a)) {
            b = select_precalc<true>(rng, (int)nb, s.hb_ijab_U + HB_I4(1, a, j, i), s.hb_ijab_K + HB_I4(1, a, j, i));
            allowed = !det_test(f, b);
        } else {
            allowed = false;
        }
    }
    g.allowed = allowed;
    if (allowed) {
        g.from1 = i; g.from2 = j;
        g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
        g.perm = excit_perm2<W>(f, i, j, g.to1, g.to2);
        g.hmatel = slater_condon2_excit(s, i, j, g.to1, g.to2, g.perm);
        g.pgen = (1.0 - p.pattempt_single) * pgen *
                 (((s.hb_ija_w[HB_I3(a, j, i)] / s.hb_ija_tot[HB_I2(j, i)]) *
                   (HB_LDCS(s.hb_ijab_w + HB_I4(b, a, j, i)) / s.hb_ijab_tot[HB_I3(a, j, i)])) +
                  ((s.hb_ija_w[HB_I3(b, j, i)] / s.hb_ija_tot[HB_I2(j, i)]) *
                   (HB_LDCS(s.hb_ijab_w + HB_I4(a, b, j, i)) / s.hb_ijab_tot[HB_I3(b, j, i)])));
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Power-Pitzer / Cauchy-Schwarz "occ" generators with uniformly selected ij
// (gen_excit_mol_power_pitzer_occ, src/excit_gen_power_pitzer_mol.F90:1260-1549; weights
// create_weighted_excitation_list_mol, src/hamiltonian_molecular.f90:348-390): O(M) weights sqrt|<ia|ai>| (or
// sqrt|<ia|ia>|) are formed on the fly over the unoccupied orbitals of i's spin and over the (spin, symmetry) class of b,
// each followed by an on-the-fly alias selection.  Lists of up to HB_MAXLIST entries live in thread-local memory.
// ------------------------------------------------------------------------------------------------
template <class R>
HB_HDN int select_weighted_value_list(R& rng, int N, const double* weights, double totweight) {
    if (N <= 64) return select_alias_staged(rng, N, weights, 1, totweight);   // table-free walk (same index)
    double aliasU[HB_MAXLIST];
    int aliasK[HB_MAXLIST], under[HB_MAXLIST], over[HB_MAXLIST];
    generate_alias_tables(N, weights, totweight, aliasU, aliasK, under, over);
    return select_precalc(rng, N, aliasU, aliasK);
}
HB_HD double pp_weight(const Sys& s, bool cauchy_schwarz, int i, int a) {
    // the weights depend on (i, a) only: once the engine has tabulated them (k_build_ppw: the same expression, evaluated
    // once per orbital pair) a weight is one load from an nbasis^2 table instead of an integral look-up and a square root
    if (s.ppw) return s.ppw[(cauchy_schwarz ? (size_t)s.nbasis * s.nbasis : 0) + (size_t)(i - 1) * s.nbasis + (a - 1)];
    return sqrt(fabs(cauchy_schwarz ? two_body(s, i, a, i, a) : two_body(s, i, a, a, i)));
}
// k-th (1-based) unoccupied orbital of the given spin parity (1 = alpha/odd orbitals, 0 = beta/even), ascending
template <int W>
HB_HD int nth_unocc_of_spin(const uint64_t* f, int nbasis, int alpha, int k) {
    const uint64_t par = alpha ? 0x5555555555555555ull : 0xAAAAAAAAAAAAAAAAull;   // orbital o <-> bit o-1
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint64_t x = ~f[w] & par;
        const int top = nbasis - 64 * w;
        if (top < 64) x &= (top <= 0) ? 0ull : ((1ull << top) - 1ull);
        const int c = popc64(x);
        if (k <= c) {
            for (int q = 1; q < k; ++q) x &= x - 1;
            return w * 64 + ctz64(x) + 1;
        }
        k -= c;
    }
    return 0;
}
// position (1-based) of unoccupied orbital b within the ascending list of unoccupied orbitals of its spin
template <int W>
HB_HD int unocc_rank_of(const uint64_t* f, int b) {
    const uint64_t par = (b & 1) ? 0x5555555555555555ull : 0xAAAAAAAAAAAAAAAAull;
    const int bw = (b - 1) >> 6, bb = (b - 1) & 63;
    int n = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint64_t x = ~f[w] & par;
        if (w > bw) x = 0;
        else if (w == bw) x &= (bb == 0) ? 0ull : ((1ull << bb) - 1ull);
        n += popc64(x);
    }
    return n + 1;
}
// iw/scr/stride: hb_i_w (or its shared-memory copy) and the staging area of nel doubles for the _occ_ij variants, whose
// (i, j) are drawn from ppm_i_d_weights / ppm_ij_d_weights (init_excit_mol_power_pitzer_orderM_ij, :585-647) - the same
// sums as hb_i_w / hb_ij_w of the heat-bath tables, which are used for them.
template <int W, class R>
HB_HDN void gen_excit_power_pitzer_occ(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                       const uint8_t* su, const double* __restrict__ iw, double* scr, int stride, Gen& g) {
    const bool cs = p.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC || p.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ;
    const bool weighted_ij = p.excit_gen == EXCIT_GEN_POWER_PITZER_OCC_IJ || p.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ;
    g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    if (rng.next() < p.pattempt_single) {
        gen_single_renorm<W>(rng, s, p, f, occ, su, g);
        return;
    }
    g.nexcit = 2;
    int i, j, ij_sym, ij_spin;
    double pgen_ij;
    if (weighted_ij) {
        // select_ij_heat_bath (src/excit_gen_utils.f90:9-66)
        const int nel = s.nel;
        const int64_t nb = s.nbasis;
        const double i_tot = stage_occ(iw, occ, nel, scr, stride);
        const int iq = select_alias_staged(rng, nel, scr, stride, i_tot);
        i = occ[iq - 1];
        const double wi = scr[(iq - 1) * stride];
        const double ij_tot = stage_occ(s.hb_ij_w + nb * (i - 1), occ, nel, scr, stride);
        if (!(ij_tot > 0.0)) {
            g.allowed = false; g.hmatel = 0.0; g.pgen = 1.0;
            return;
        }
        const int jq = select_alias_staged(rng, nel, scr, stride, ij_tot);
        j = occ[jq - 1];
        const double wij = scr[(jq - 1) * stride];
        const double ji_tot = sum_occ(s.hb_ij_w + nb * (j - 1), occ, nel);
        ij_spin = ms_of(i) + ms_of(j);
        ij_sym = sym_conj(s, cross_product(s, s.bf_sym[i], s.bf_sym[j]));
        pgen_ij = ((wi / i_tot) * (wij / ij_tot)) + ((iw[j - 1] / i_tot) * (s.hb_ij_w[HB_I2(i, j)] / ji_tot));
        if (j < i) { const int t = i; i = j; j = t; }
    } else {
        choose_ij(rng, s, occ, i, j, ij_sym, ij_spin);
        pgen_ij = 2.0 / (s.nel * (s.nel - 1));
    }
    const int ialpha = i & 1;
    // number of unoccupied orbitals of i's spin
    int ni = 0;
    {
        const uint64_t par = ialpha ? 0x5555555555555555ull : 0xAAAAAAAAAAAAAAAAull;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint64_t x = ~f[w] & par;
            const int top = s.nbasis - 64 * w;
            if (top < 64) x &= (top <= 0) ? 0ull : ((1ull << top) - 1ull);
            ni += popc64(x);
        }
    }
    double ia_w[HB_MAXLIST], jb_w[HB_MAXLIST];
    double ia_tot = 0.0, jb_tot = 0.0;
    bool a_found = false;
    int a = 0, b = 0, a_ind = 0, b_ind = 0;
    if (ni > 0) {
        // walk the unoccupied orbitals of i's spin in ascending order
        int k = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint64_t x = ~f[w] & (ialpha ? 0x5555555555555555ull : 0xAAAAAAAAAAAAAAAAull);
            const int top = s.nbasis - 64 * w;
            if (top < 64) x &= (top <= 0) ? 0ull : ((1ull << top) - 1ull);
            while (x) {
                const int orb = w * 64 + ctz64(x) + 1;
                x &= x - 1;
                ia_w[k] = pp_weight(s, cs, i, orb);
                ia_tot = ia_tot + ia_w[k];
                ++k;
            }
        }
        if (ia_tot > 0.0) {
            a_ind = select_weighted_value_list(rng, ni, ia_w, ia_tot);
            a = nth_unocc_of_spin<W>(f, s.nbasis, ialpha, a_ind);
            a_found = true;
        }
    }
    int isymb = 0, imsb = 0, nb_list = 0;
    if (a_found) {
        isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
        imsb = (ij_spin - ms_of(a) + 3) / 2;
        nb_list = nbss(s, imsb, isymb);
        if (nb_list > 0) {
            for (int k = 0; k < nb_list; ++k) {
                const int orb = ssbf(s, k + 1, imsb, isymb);
                if (orb != a) { jb_w[k] = pp_weight(s, cs, j, orb); jb_tot = jb_tot + jb_w[k]; }
                else jb_w[k] = 0.0;
            }
        }
    }
    g.allowed = false;
    if (a_found && jb_tot > 0.0) {
        b_ind = select_weighted_value_list(rng, nb_list, jb_w, jb_tot);
        b = ssbf(s, b_ind, imsb, isymb);
        if (!det_test(f, b)) {
            double pgen;
            if (ij_spin == 0) {
                pgen = ia_w[a_ind - 1] / ia_tot * jb_w[b_ind - 1] / jb_tot;
            } else {
                const int b_ind_rev = unocc_rank_of<W>(f, b);
                const int isyma = sym_conj(s, cross_product(s, ij_sym, isymb));
                const int na_list = nbss(s, imsb, isyma);
                double ja_tot = 0.0, ja_a = 0.0;
                for (int k = 0; k < na_list; ++k) {
                    const int orb = ssbf(s, k + 1, imsb, isyma);
                    double wk = 0.0;
                    if (orb != b) { wk = pp_weight(s, cs, j, orb); ja_tot = ja_tot + wk; }
                    if (orb == a) ja_a = wk;
                }
                pgen = (ia_w[a_ind - 1] * jb_w[b_ind - 1]) / (ia_tot * jb_tot) + (ia_w[b_ind_rev - 1] * ja_a) / (ia_tot * ja_tot);
            }
            g.pgen = p.pattempt_double * pgen * pgen_ij;
            g.from1 = (i < j) ? i : j; g.from2 = (i < j) ? j : i;
            g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
            g.allowed = true;
        }
    }
    if (g.allowed) {
        g.perm = excit_perm2<W>(f, g.from1, g.from2, g.to1, g.to2);
        g.hmatel = slater_condon2_excit(s, g.from1, g.from2, g.to1, g.to2, g.perm);
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
}

// ------------------------------------------------------------------------------------------------
// power_pitzer_orderN ('heat_bath_power_pitzer_ref'): every choice is one look-up in a precomputed alias table indexed
// through the mapping reference orbital -> orbital of this determinant.
// ------------------------------------------------------------------------------------------------
// get_excitation_locations (src/excitations.F90) + the spin pairing of find_diff_ref_cdet (src/excit_gen_utils.f90:220-269)
// from the bit strings: the reference's orbitals missing in the determinant ("holes": position in the reference's
// list and orbital, ascending) and the determinant's orbitals missing in the reference ("particles", ascending), then
// the reference's forward swaps so that every hole is paired with a particle of its spin.  Same pairs as the
// reference's merge walk over the two occupied lists (tests/test_core_vs_oracle.py compares them with the oracle's
// literal restatement).
template <int W>
HB_HDN int ref_det_diff(const uint64_t* f0, const uint64_t* f, uint8_t* hole_idx, uint8_t* hole_orb, uint8_t* part_orb) {
    int nh = 0, np = 0, base = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint64_t h = f0[w] & ~f[w];
        while (h) {
            const int bit = ctz64(h);
            h &= h - 1;
            hole_orb[nh] = (uint8_t)(w * 64 + bit + 1);
            hole_idx[nh] = (uint8_t)(base + popc64(f0[w] & ((1ull << bit) - 1ull)) + 1);
            nh++;
        }
        uint64_t q = f[w] & ~f0[w];
        while (q) {
            const int bit = ctz64(q);
            q &= q - 1;
            part_orb[np++] = (uint8_t)(w * 64 + bit + 1);
        }
        base += popc64(f0[w]);
    }
    for (int ii = 0; ii < nh; ++ii) {
        if (ms_of(hole_orb[ii]) != ms_of(part_orb[ii])) {
            int jj = ii + 1;
            while (ms_of(hole_orb[ii]) != ms_of(part_orb[jj])) jj++;
            const uint8_t t = part_orb[ii]; part_orb[ii] = part_orb[jj]; part_orb[jj] = t;
        }
    }
    return nh;
}
// det_info_t%ref_cdet_occ_list: the determinant's occupied orbitals in the order of the reference's
template <int W>
HB_HDN void find_diff_ref_cdet(const Sys& s, const uint64_t* f0, const uint64_t* f, uint8_t* ref_cdet) {
    uint8_t hole_idx[HB_MAXNEL], hole_orb[HB_MAXNEL], part_orb[HB_MAXNEL];
    const int nex = ref_det_diff<W>(f0, f, hole_idx, hole_orb, part_orb);
    for (int k = 0; k < s.nel; ++k) ref_cdet[k] = (uint8_t)s.ppn_occ[k];
    for (int ii = 0; ii < nex; ++ii) ref_cdet[hole_idx[ii] - 1] = part_orb[ii];
}
// gen_excit_mol_power_pitzer_occ_ref (src/excit_gen_power_pitzer_mol.F90:650-939), excit_gen = power_pitzer: ij uniform
// among the reference's occupied orbitals, a and b from the reference's alias tables, mapped onto this determinant
template <int W, class R>
HB_HDN void gen_excit_power_pitzer_ref(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ, Gen& g) {
    const int nel = s.nel, mv = s.max_nbss, nsym = s.nsym_tot;
    g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    if (rng.next() < p.pattempt_single) {     // gen_single_excit_mol_no_renorm (src/excit_gen_mol.f90:450-517)
        g.nexcit = 1;
        const int i = occ[(int)(rng.next() * nel)];
        const int imsa = ims_of(i);
        const int isyma = cross_product(s, s.bf_sym[i], s.gamma_sym);
        const int n = nbss(s, imsa, isyma);
        const int ind = (int)(n * rng.next()) + 1;
        g.from1 = i;
        int a = 0;
        if (n == 0) g.allowed = false;
        else { a = ssbf(s, ind, imsa, isyma); g.allowed = !det_test(f, a); }
        if (g.allowed) {
            g.to1 = a;
            g.pgen = p.pattempt_single * (1.0 / (nel * nbss(s, ims_of(a), s.bf_sym[a])));
            g.perm = excit_perm1<W>(f, i, a);
            g.hmatel = slater_condon1_excit(s, occ, i, a, g.perm);
        } else { g.hmatel = 0.0; g.pgen = 1.0; }
        return;
    }
    g.nexcit = 2;
    const int ind = (int)(rng.next() * nel * (nel - 1) / 2) + 1;
    const int j_ind = (int)(1.50 + sqrt(2 * ind - 1.750));
    const int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
    const double pgen_ij = 2.0 / (nel * (nel - 1));
    const int i_ref = s.ppn_occ[i_ind - 1], j_ref = s.ppn_occ[j_ind - 1];
    const int ij_spin = ms_of(i_ref) + ms_of(j_ref);
    const int sp = (ms_of(i_ref) < 0) ? 0 : 1;
    const int nv = s.pp_nvirt[sp];
    const int* virt = s.pp_virt[sp];
    const size_t sia = (size_t)s.pp_sia;
    bool a_found = false;
    int a_ind = 0, a_ref = 0;
    if (nv > 0) {
        a_ind = select_precalc(rng, nv, s.pp_ia.U + sia * (i_ind - 1), s.pp_ia.K + sia * (i_ind - 1));
        a_ref = virt[a_ind - 1];
        a_found = true;
    }
    uint8_t hole_idx[HB_MAXNEL], hole_orb[HB_MAXNEL], part_orb[HB_MAXNEL];
    int nex = 0, i = i_ref, j = j_ref, a = a_ref, ij_sym = 0, isymb = 0, imsb = 1;
    if (a_found) {
        nex = ref_det_diff<W>(p.f0, f, hole_idx, hole_orb, part_orb);
        for (int ii = 0; ii < nex; ++ii) {
            if (hole_idx[ii] == i_ind) i = part_orb[ii];
            else if (hole_idx[ii] == j_ind) j = part_orb[ii];
            if ((int)part_orb[ii] == a_ref) a = hole_orb[ii];
        }
        ij_sym = sym_conj(s, cross_product(s, s.bf_sym[i], s.bf_sym[j]));
        isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
        imsb = ims_of(j_ref);
    }
    g.allowed = false;
    int b = 0;
    if (a_found && nbss(s, imsb, isymb) > 0) {
        const size_t colb = (size_t)isymb + (size_t)nsym * (j_ind - 1);
        const int b_ind = select_precalc(rng, nbss(s, imsb, isymb), s.pp_jb.U + (size_t)mv * colb, s.pp_jb.K + (size_t)mv * colb);
        b = ssbf(s, b_ind, imsb, isymb);
        if (a != b && !det_test(f, b)) {
            double pgen;
            const double pa = s.pp_ia.w[sia * (i_ind - 1) + a_ind - 1] / s.pp_ia.tot[i_ind - 1];
            if (ij_spin == 0) {
                pgen = pa * s.pp_jb.w[(size_t)mv * colb + b_ind - 1] / s.pp_jb.tot[colb];
            } else {
                int b_ref = b;
                for (int ii = 0; ii < nex; ++ii)
                    if ((int)hole_orb[ii] == b) { b_ref = part_orb[ii]; break; }
                int b_rev = 0;
                for (int lo = 1, hi = nv; lo <= hi;) {      // binary_search in virt_list_{alpha,beta}
                    const int mid = (lo + hi) / 2;
                    if (virt[mid - 1] == b_ref) { b_rev = mid; break; }
                    if (virt[mid - 1] < b_ref) lo = mid + 1; else hi = mid - 1;
                }
                const int isyma = sym_conj(s, cross_product(s, ij_sym, isymb));
                int a_rev = 0;
                const int na = nbss(s, imsb, isyma);
                for (int k = 1; k <= na; ++k)
                    if (ssbf(s, k, imsb, isyma) == a) { a_rev = k; break; }
                const size_t cola = (size_t)isyma + (size_t)nsym * (j_ind - 1);
                pgen = pa * s.pp_jb.w[(size_t)mv * colb + b_ind - 1] / s.pp_jb.tot[colb] +
                       s.pp_ia.w[sia * (i_ind - 1) + b_rev - 1] / s.pp_ia.tot[i_ind - 1] *
                           s.pp_jb.w[(size_t)mv * cola + a_rev - 1] / s.pp_jb.tot[cola];
            }
            g.pgen = p.pattempt_double * pgen * pgen_ij;
            g.allowed = true;
        }
    }
    if (g.allowed) {
        g.from1 = (i < j) ? i : j; g.from2 = (i < j) ? j : i;
        g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
        g.perm = excit_perm2<W>(f, g.from1, g.from2, g.to1, g.to2);
        g.hmatel = slater_condon2_excit(s, g.from1, g.from2, g.to1, g.to2, g.perm);
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
}
// gen_excit_mol_power_pitzer_orderN (src/excit_gen_power_pitzer_mol.F90:941-1258)
template <int W, class R>
HB_HDN void gen_excit_power_pitzer_orderN(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ,
                                          const uint8_t* ref_cdet, Gen& g) {
    const int nel = s.nel, mv = s.max_nbss, nsym = s.nsym_tot, nall = s.nbasis / 2;
    g.from1 = 0; g.from2 = 0; g.to1 = 0; g.to2 = 0; g.perm = false;
    if (rng.next() < p.pattempt_single) {
        g.nexcit = 1;
        const int i_ind = select_precalc(rng, nel, s.ppn[PPN_IS].U, s.ppn[PPN_IS].K);
        const int i = ref_cdet[i_ind - 1];
        const int imsa = ims_of(i);
        const int isyma = cross_product(s, s.bf_sym[i], s.gamma_sym);
        const int n = nbss(s, imsa, isyma);
        int a_ind = 0, a = 0;
        if (n > 0) {
            a_ind = select_precalc(rng, n, s.ppn[PPN_IAS].U + (size_t)mv * i, s.ppn[PPN_IAS].K + (size_t)mv * i);
            a = ssbf(s, a_ind, imsa, isyma);
            g.allowed = !det_test(f, a);
        } else {
            g.allowed = false;
        }
        if (g.allowed) {
            g.pgen = (s.ppn[PPN_IS].w[i_ind - 1] / s.ppn[PPN_IS].tot[0]) *
                     (s.ppn[PPN_IAS].w[(size_t)mv * i + a_ind - 1] / s.ppn[PPN_IAS].tot[i]);
            g.pgen = p.pattempt_single * g.pgen;
            g.from1 = i; g.to1 = a;
            g.perm = excit_perm1<W>(f, i, a);
            g.hmatel = slater_condon1_excit(s, occ, i, a, g.perm);
        } else {
            g.hmatel = 0.0; g.pgen = 1.0;
        }
        return;
    }
    g.nexcit = 2;
    const int i_ind = select_precalc(rng, nel, s.ppn[PPN_ID].U, s.ppn[PPN_ID].K);
    int i = ref_cdet[i_ind - 1];
    const int j_ind = select_precalc(rng, nel, s.ppn[PPN_IJD].U + (size_t)nel * i, s.ppn[PPN_IJD].K + (size_t)nel * i);
    int j = ref_cdet[j_ind - 1];
    double pgen = 1.0;
    int ij_spin = 0;
    if (j != i) {
        g.allowed = true;
        pgen = (s.ppn[PPN_ID].w[i_ind - 1] / s.ppn[PPN_ID].tot[0]) *
               (s.ppn[PPN_IJD].w[(size_t)nel * i + j_ind - 1] / s.ppn[PPN_IJD].tot[i]);
        ij_spin = ms_of(i) + ms_of(j);
        pgen = pgen + ((s.ppn[PPN_ID].w[j_ind - 1] / s.ppn[PPN_ID].tot[0]) *
                       (s.ppn[PPN_IJD].w[(size_t)nel * j + i_ind - 1] / s.ppn[PPN_IJD].tot[j]));
        if (j < i) { const int t = i; i = j; j = t; }
    } else {
        g.allowed = false;
    }
    int a_ind = 0, a = 0, b_ind = 0, b = 0, ij_sym = 0, isymb = 0, imsb = 1;
    if (g.allowed) {
        a_ind = select_precalc(rng, nall, s.ppn[PPN_IAD].U + (size_t)nall * i, s.ppn[PPN_IAD].K + (size_t)nall * i);
        a = (ms_of(i) < 0) ? 2 * a_ind : 2 * a_ind - 1;      // all_list_beta / all_list_alpha
        if (det_test(f, a)) g.allowed = false;
    }
    if (g.allowed) {
        ij_sym = sym_conj(s, cross_product(s, s.bf_sym[i], s.bf_sym[j]));
        isymb = sym_conj(s, cross_product(s, ij_sym, s.bf_sym[a]));
        imsb = ims_of(j);
        if (nbss(s, imsb, isymb) == 0) g.allowed = false;
    }
    if (g.allowed) {
        const size_t colb = (size_t)isymb + (size_t)nsym * j;
        b_ind = select_precalc(rng, nbss(s, imsb, isymb), s.ppn[PPN_JBD].U + (size_t)mv * colb, s.ppn[PPN_JBD].K + (size_t)mv * colb);
        b = ssbf(s, b_ind, imsb, isymb);
        if (a != b && !det_test(f, b)) {
            const double pa = s.ppn[PPN_IAD].w[(size_t)nall * i + a_ind - 1] / s.ppn[PPN_IAD].tot[i];
            if (ij_spin == 0) {
                pgen = pgen * (pa * s.ppn[PPN_JBD].w[(size_t)mv * colb + b_ind - 1] / s.ppn[PPN_JBD].tot[colb]);
            } else {
                const int b_rev = (b + 1) >> 1;               // position of b in all_list_{alpha,beta}
                const int isyma = sym_conj(s, cross_product(s, ij_sym, isymb));
                int a_rev = 0;
                const int na = nbss(s, imsb, isyma);
                for (int k = 1; k <= na; ++k)
                    if (ssbf(s, k, imsb, isyma) == a) { a_rev = k; break; }
                const size_t cola = (size_t)isyma + (size_t)nsym * j;
                pgen = pgen * (pa * s.ppn[PPN_JBD].w[(size_t)mv * colb + b_ind - 1] / s.ppn[PPN_JBD].tot[colb] +
                               s.ppn[PPN_IAD].w[(size_t)nall * i + b_rev - 1] / s.ppn[PPN_IAD].tot[i] *
                                   s.ppn[PPN_JBD].w[(size_t)mv * cola + a_rev - 1] / s.ppn[PPN_JBD].tot[cola]);
            }
            pgen = p.pattempt_double * pgen;
        } else {
            g.allowed = false;
        }
    }
    if (g.allowed) {
        g.pgen = pgen;
        g.from1 = i; g.from2 = j;
        g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
        g.perm = excit_perm2<W>(f, g.from1, g.from2, g.to1, g.to2);
        g.hmatel = slater_condon2_excit(s, g.from1, g.from2, g.to1, g.to2, g.perm);
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Uniform electron gas (3D): analytic integrals, Slater-Condon rules and the no_renorm generator
// ------------------------------------------------------------------------------------------------
// coulomb_int_ueg_3d (src/ueg.f90:250-280): 1 / (pi L |k_i - k_a|^2)
HB_HD double ueg_coulomb(const Sys& s, int i, int a) {
    const K4 ki = s.ueg_k[i], ka = s.ueg_k[a];
    const int qx = ki.x - ka.x, qy = ki.y - ka.y, qz = ki.z - ka.z;
    return 1.0 / (s.ueg_piL * (qx * qx + qy * qy + qz * qz));
}
// get_two_e_int_ueg (src/ueg.f90:176-215): < ij || ab > with momentum and spin checks
HB_HD double ueg_two_e_int(const Sys& s, int i, int j, int a, int b) {
    const K4 ki = s.ueg_k[i], kj = s.ueg_k[j], ka = s.ueg_k[a], kb = s.ueg_k[b];
    double v = 0.0;
    if (ki.x + kj.x - ka.x - kb.x == 0 && ki.y + kj.y - ka.y - kb.y == 0 && ki.z + kj.z - ka.z - kb.z == 0) {
        if (ms_of(i) == ms_of(a) && ms_of(j) == ms_of(b)) v = v + ueg_coulomb(s, i, a);
        if (ms_of(i) == ms_of(b) && ms_of(j) == ms_of(a)) v = v - ueg_coulomb(s, i, b);
    }
    return v;
}
// slater_condon0_ueg (src/hamiltonian_ueg.f90:71-99,127-156; src/determinants.f90:403-425)
HB_HDN double slater_condon0_ueg(const Sys& s, const occ_t* occ) {
    double spe = 0.0;
    for (int i = 0; i < s.nel; ++i) spe = spe + s.sp_eigv[occ[i]];
    double ex = 0.0;
    for (int i = 0; i < s.nel; ++i)
        for (int j = i + 1; j < s.nel; ++j)
            if (((occ[i] ^ occ[j]) & 1) == 0) ex = ex - ueg_coulomb(s, occ[i], occ[j]);
    return spe + ex;
}
// slater_condon2_ueg_excit (src/hamiltonian_ueg.f90:186-223)
HB_HD double slater_condon2_ueg_excit(const Sys& s, int i, int a, int b, bool perm) {
    double h = 0.0;
    if (ms_of(i) == ms_of(a)) h = ueg_coulomb(s, i, a);
    if (ms_of(i) == ms_of(b)) h = h - ueg_coulomb(s, i, b);
    return perm ? -h : h;
}
// ueg_basis_index (src/ueg.f90:142-172)
HB_HD int ueg_basis_index(const Sys& s, int kx, int ky, int kz, int spin) {
    const int mn = kx < ky ? (kx < kz ? kx : kz) : (ky < kz ? ky : kz);
    const int mx = kx > ky ? (kx > kz ? kx : kz) : (ky > kz ? ky : kz);
    if (mn < -s.ueg_kmax || mx > s.ueg_kmax) return -1;
    int indx = s.ueg_lookup[kx * s.ueg_oi[0] + ky * s.ueg_oi[1] + kz * s.ueg_oi[2] + s.ueg_offset];
    if (spin < 0) indx = indx + 1;
    return indx;
}
// gen_excit_ueg_no_renorm (src/excit_gen_ueg.f90:27-101): choose_ij_k (:105-190), find_ab_ueg (:194-306),
// calc_pgen_ueg_no_renorm (:310-360).  Two random numbers (the second only if an a exists).
template <int W, class R>
HB_HDN void gen_excit_ueg_no_renorm(R& rng, const Sys& s, const uint64_t* f, const occ_t* occ, Gen& g) {
    const int nel = s.nel;
    g.nexcit = 2; g.perm = false; g.to1 = 0; g.to2 = 0;
    const int ind = (int)(rng.next() * nel * (nel - 1) / 2) + 1;
    const int j_ind = (int)(1.50 + sqrt(2 * ind - 1.750));
    const int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
    const int i = occ[i_ind - 1], j = occ[j_ind - 1];
    g.from1 = i; g.from2 = j;
    const int ij_spin = ms_of(i) + ms_of(j);
    const K4 ki = s.ueg_k[i], kj = s.ueg_k[j];
    const int kx = ki.x + kj.x, ky = ki.y + kj.y, kz = ki.z + kj.z;
    const uint64_t* __restrict__ t =
        s.ueg_tern + (size_t)(W + 1) * ((kx + s.ueg_tK) + (size_t)s.ueg_tD * ((ky + s.ueg_tK) + (size_t)s.ueg_tD * (kz + s.ueg_tK)));
    uint64_t poss[W];
    int max_na = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint64_t tc = (ij_spin == -2) ? (t[1 + w] << 1) : t[1 + w];
        poss[w] = ~f[w] & tc;
        max_na += popc64(poss[w]);
    }
    g.allowed = false;
    if (max_na > 0) {
        int a = (int)(max_na * rng.next()) + 1;
        int n = 0;
        bool found = false;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int c = popc64(poss[w]);
            if (!found) {
                if (n + c >= a) {
                    uint64_t x = poss[w];
                    for (int q = 1; q < a - n; ++q) x &= x - 1;
                    a = w * 64 + ctz64(x) + 1;
                    found = true;
                } else {
                    n += c;
                }
            }
        }
        const K4 ka = s.ueg_k[a];
        int b = ueg_basis_index(s, kx - ka.x, ky - ka.y, kz - ka.z, ij_spin == 2 ? 1 : -1);
        g.allowed = !det_test(f, b);
        if (a > b) { const int tmp = a; a = b; b = tmp; }
        g.to1 = a; g.to2 = b;
    }
    if (g.allowed) {
        g.pgen = 2.0 / (nel * (nel - 1) * max_na);
        if (ij_spin != 0) g.pgen = g.pgen * 2;
        g.perm = excit_perm2<W>(f, i, j, g.to1, g.to2);
        g.hmatel = slater_condon2_ueg_excit(s, i, g.to1, g.to2, g.perm);
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
}

// gen_excit_ueg_power_pitzer (src/excit_gen_ueg.f90:410-566), excit_gen = power_pitzer on the UEG: ij uniform, a from
// the alias table of i over the orbitals of its spin (weights |<ia|ai>|, s.pp_ia), b from momentum conservation
template <int W, class R>
HB_HDN void gen_excit_ueg_power_pitzer(R& rng, const Sys& s, const uint64_t* f, const occ_t* occ, Gen& g) {
    const int nel = s.nel, maxv = s.nbasis / 2;
    g.nexcit = 2; g.perm = false; g.to1 = 0; g.to2 = 0;
    const int ind = (int)(rng.next() * nel * (nel - 1) / 2) + 1;
    const int j_ind = (int)(1.50 + sqrt(2 * ind - 1.750));
    const int i_ind = ind - ((j_ind - 1) * (j_ind - 2)) / 2;
    const int i = occ[i_ind - 1], j = occ[j_ind - 1];
    g.from1 = i; g.from2 = j;
    const int ij_spin = ms_of(i) + ms_of(j);
    const K4 ki = s.ueg_k[i], kj = s.ueg_k[j];
    const int kx = ki.x + kj.x, ky = ki.y + kj.y, kz = ki.z + kj.z;
    const int a_ind = select_precalc(rng, maxv, s.pp_ia.U + (size_t)maxv * i, s.pp_ia.K + (size_t)maxv * i);
    const int a = 2 * a_ind - (i & 1);
    int b = 0, b_ind = 0;
    g.allowed = !det_test(f, a);
    if (g.allowed) {
        const K4 ka = s.ueg_k[a];
        const int spin_b = (ij_spin == 2) ? 1 : ((ij_spin == 0) ? -ms_of(a) : -1);
        b = ueg_basis_index(s, kx - ka.x, ky - ka.y, kz - ka.z, spin_b);
        if (b <= 0) g.allowed = false;
        else { b_ind = (b + 1) >> 1; g.allowed = !det_test(f, b); }
    }
    if (g.allowed) {
        const double* w = s.pp_ia.w + (size_t)maxv * i;
        if (ij_spin == 0) g.pgen = w[a_ind - 1] / s.pp_ia.tot[i];
        else g.pgen = (w[a_ind - 1] + w[b_ind - 1]) / s.pp_ia.tot[i];
        g.pgen = g.pgen * 2.0 / (nel * (nel - 1));
        g.allowed = (a != b);
    }
    if (g.allowed) {
        g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
        g.perm = excit_perm2<W>(f, i, j, g.to1, g.to2);
        g.hmatel = slater_condon2_ueg_excit(s, i, g.to1, g.to2, g.perm);
    } else {
        g.hmatel = 0.0; g.pgen = 1.0;
    }
}
template <int W, class R>
HB_HD void gen_excit(R& rng, const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ, const uint8_t* su,
                     Gen& g) {
    if (s.kind == SYS_UEG) {
        if (p.excit_gen == EXCIT_GEN_POWER_PITZER) gen_excit_ueg_power_pitzer<W>(rng, s, f, occ, g);
        else gen_excit_ueg_no_renorm<W>(rng, s, f, occ, g);
    }
    // the wide layout (W > 4, more than 254 spin-orbitals) is built for the UEG generators only: the molecular ones keep
    // byte orbital lists and nbasis^3 / nbasis^4 tables
    else if (W > 4) { g.allowed = false; g.hmatel = 0.0; g.pgen = 1.0; g.nexcit = 0; }
    else if (p.excit_gen == EXCIT_GEN_RENORM) gen_excit_renorm<W>(rng, s, p, f, occ, su, g);
    else if (p.excit_gen == EXCIT_GEN_NO_RENORM) gen_excit_no_renorm<W>(rng, s, p, f, occ, g);
    else if (p.excit_gen == EXCIT_GEN_POWER_PITZER) gen_excit_power_pitzer_ref<W>(rng, s, p, f, occ, g);
    else if (p.excit_gen == EXCIT_GEN_POWER_PITZER_ORDERN) {
        uint8_t ref_cdet[HB_MAXNEL];
        find_diff_ref_cdet<W>(s, p.f0, f, ref_cdet);
        gen_excit_power_pitzer_orderN<W>(rng, s, p, f, occ, ref_cdet, g);
    }
    else if (p.excit_gen == EXCIT_GEN_RENORM_SPIN) gen_excit_renorm<W, true>(rng, s, p, f, occ, su, g);
    else if (p.excit_gen == EXCIT_GEN_NO_RENORM_SPIN) gen_excit_no_renorm<W, true>(rng, s, p, f, occ, g);
    else if (p.excit_gen == EXCIT_GEN_POWER_PITZER_OCC || p.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC ||
             p.excit_gen == EXCIT_GEN_POWER_PITZER_OCC_IJ || p.excit_gen == EXCIT_GEN_CAUCHY_SCHWARZ_OCC_IJ) {
        double scr[HB_MAXNEL];
        gen_excit_power_pitzer_occ<W>(rng, s, p, f, occ, su, s.hb_i_w, scr, 1, g);
    }
    else if (p.excit_gen == EXCIT_GEN_HEAT_BATH_SINGLE) {
        double scr[HB_MAXNEL];
        gen_excit_heat_bath_uniform<W, true>(rng, s, p, f, occ, su, s.hb_i_w, scr, 1, g);
    } else if (p.excit_gen == EXCIT_GEN_HEAT_BATH_UNIFORM) {
        double scr[HB_MAXNEL];
        gen_excit_heat_bath_uniform<W>(rng, s, p, f, occ, su, s.hb_i_w, scr, 1, g);
    }
    else gen_excit_heat_bath<W>(rng, s, p, f, occ, g);
}

// ------------------------------------------------------------------------------------------------
// Spawning / death arithmetic
// ------------------------------------------------------------------------------------------------
// quasi-Newton weights: cdet%fock_sum (src/fciqmc.f90:321-322), calc_qn_spawned_weighting and calc_qn_weighting
// (src/spawning.F90:2063-2137)
HB_HD double qn_fock_sum(const Sys& s, const Params& p, const occ_t* occ) {
    double fs = 0.0;
    for (int k = 0; k < s.nel; ++k) fs = fs + p.sp_fock[occ[k]];
    return fs - p.ref_fock_sum;
}
HB_HD double qn_spawned_weighting(const Params& p, double spawner_dfock, const Gen& g) {
    double diagel = spawner_dfock;
    diagel = diagel + p.sp_fock[g.to1] - p.sp_fock[g.from1];
    if (g.nexcit == 2) diagel = diagel + p.sp_fock[g.to2] - p.sp_fock[g.from2];
    if (diagel < p.qn_threshold) diagel = p.qn_value;
    return 1.0 / diagel;
}
HB_HD double qn_weighting(const Params& p, double dfock) {
    return (dfock < p.qn_threshold) ? 1.0 / p.qn_value : 1.0 / dfock;
}
// attempt_to_spawn + stochastic_round_spawned_particle (src/spawning.F90:711-766,
// src/stoch_utils.f90:87-136).  Always draws exactly one random number.
template <class R>
HB_HD int64_t attempt_to_spawn(R& rng, const Params& p, double hmatel, double pgen, int64_t parent_pop) {
    double pspawn = p.tau * fabs(hmatel) / pgen;
    pspawn = pspawn * (double)p.real_factor;
    int64_t nspawn;
    if (pspawn < (double)p.spawn_cutoff) {
        nspawn = (pspawn > rng.next() * (double)p.spawn_cutoff) ? p.spawn_cutoff : 0;
    } else {
        nspawn = (int64_t)pspawn;
        double padd = pspawn - (double)nspawn;
        if (padd > rng.next()) nspawn++;
    }
    if (nspawn > 0) {
        int64_t sg = (parent_pop >= 0) ? nspawn : -nspawn;
        nspawn = (hmatel > 0.0) ? -sg : sg;
    }
    return nspawn;
}
// decide_nattempts (src/qmc_common.F90:379-406)
template <class R>
HB_HD int decide_nattempts(R& rng, double population) {
    int n = (int)population;
    if (n < 0) n = -n;
    double pextra = fabs(population) - n;
    if (fabs(pextra) > 1.e-12) {
        if (pextra > rng.next()) n++;
    }
    return n;
}
// stochastic_death (src/death.f90:11-130), no quasi-Newton / Chebyshev weights.  Returns the new
// population; kill_abs receives |kill| (ndeath contribution).
template <class R>
HB_HD int64_t stochastic_death(R& rng, const Params& p, double Kii, int64_t population, int64_t& kill_abs, double weight = 1.0) {
    const double pop_control = p.qn ? p.qn_pop_control : 1.0;
    double pd = p.tau * ((Kii - p.proj_energy_old) * weight + (p.proj_energy_old - p.shift) * pop_control) * 1.0;
    pd = pd * p.cheby_weight;
    int64_t apop = population < 0 ? -population : population;
    pd = pd * (double)apop;
    int64_t kill = (int64_t)pd;
    pd = pd - (double)kill;
    double r = rng.next();
    if (fabs(pd) > r) {
        if (pd > 0.0) kill++; else kill--;
    }
    kill_abs = kill < 0 ? -kill : kill;
    return (population < 0) ? population + kill : population - kill;
}
// stochastic_round (src/stoch_utils.f90:51-85)
template <class R>
HB_HD int64_t stochastic_round(R& rng, int64_t pop, int64_t cutoff) {
    int64_t ap = pop < 0 ? -pop : pop;
    if (ap < cutoff && pop != 0) {
        double r = rng.next() * (double)cutoff;
        if ((double)ap > r) return (pop < 0) ? -cutoff : cutoff;
        return 0;
    }
    return pop;
}

// Projected-energy contribution of one determinant (get_excitation src/excitations.F90:75-200 +
// update_proj_energy_mol src/energy_evaluation.F90:906-986).  Returns H_0j (incl. sign) or 0; sets
// is_ref when f == f0.
// hmatel_pair: the same between f (whose occupied list is occ) and any determinant f0 (get_hmatel, off-diagonal part).
template <int W>
HB_HD double hmatel_pair(const Sys& s, const uint64_t* f, const occ_t* occ, const uint64_t* f0, bool& is_ref);
template <int W>
HB_HDN double proj_energy_hmatel(const Sys& s, const Params& p, const uint64_t* f, const occ_t* occ, bool& is_ref) {
    return hmatel_pair<W>(s, f, occ, p.f0, is_ref);
}
template <int W>
HB_HD double hmatel_pair(const Sys& s, const uint64_t* f, const occ_t* occ, const uint64_t* f0, bool& is_ref) {
    int nx = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) nx += popc64(f[k] ^ f0[k]);
    is_ref = (nx == 0);
    if (nx == 0 || nx > 4) return 0.0;
    const int nexcit = nx / 2;
    const int shift = s.nel - nexcit;
    int from[2] = {0, 0}, to[2] = {0, 0};
    int iexcit1 = 0, iexcit2 = 0, iel1 = 0, iel2 = 0, perm = 0;
    for (int i = 0; i < W; ++i) {
        uint64_t f1 = f[i], f2 = f0[i];
        if (f1 == f2) {
            if (((iexcit1 - iexcit2) & 1) != 0) {
                int n = popc64(f1);
                iel1 += n; iel2 += n;
            }
            continue;
        }
        for (int jb = 0; jb < 64; ++jb) {
            bool t1 = (f1 >> jb) & 1ull, t2 = (f2 >> jb) & 1ull;
            if (t2) iel2++;
            if (t1) {
                iel1++;
                if (!t2) { iexcit1++; from[iexcit1 - 1] = i * 64 + jb + 1; perm += (shift - iel1 + iexcit1); }
            } else if (t2) {
                iexcit2++; to[iexcit2 - 1] = i * 64 + jb + 1; perm += (shift - iel2 + iexcit2);
            }
        }
    }
    bool pm = (((perm % 2) + 2) % 2) == 1;
    if (s.kind == SYS_UEG) {
        // update_proj_energy_ueg (src/energy_evaluation.F90:1072-1127): slater_condon2_ueg, doubles only
        if (nexcit != 2) return 0.0;
        const double h = ueg_two_e_int(s, from[0], from[1], to[0], to[1]);
        return pm ? -h : h;
    }
    if (nexcit == 1) {
        if (ms_of(from[0]) == ms_of(to[0]) && s.bf_sym[from[0]] == s.bf_sym[to[0]])
            return slater_condon1_excit(s, occ, from[0], to[0], pm);
        return 0.0;
    }
    if (ms_of(from[0]) + ms_of(from[1]) == ms_of(to[0]) + ms_of(to[1])) {
        int ij = cross_product(s, s.bf_sym[from[0]], s.bf_sym[from[1]]);
        int ab = cross_product(s, s.bf_sym[to[0]], s.bf_sym[to[1]]);
        if (ij == ab) return slater_condon2_excit(s, from[0], from[1], to[0], to[1], pm);
    }
    return 0.0;
}

// Excitation level of f relative to f0 (get_excitation_level)
template <int W>
HB_HD int excit_level(const uint64_t* f, const uint64_t* f0) {
    int nx = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) nx += popc64(f[k] ^ f0[k]);
    return nx / 2;
}

// check_if_determ (src/semi_stoch.F90:795-824): is f one of the deterministic states?  The reference walks a hash
// bucket; membership is all that matters, so the sorted copy of determ%dets is bisected instead.
template <int W>
HB_HDNI bool ss_check_if_determ(const uint64_t* __restrict__ sorted, int n, const uint64_t* f) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (det_less<W>(sorted + (size_t)mid * W, f)) lo = mid + 1; else hi = mid;
    }
    return lo < n && det_eq<W>(sorted + (size_t)lo * W, f);
}

// ------------------------------------------------------------------------------------------------
// CCMC: stochastic cluster selection and cluster algebra (src/ccmc_selection.f90, src/ccmc_utils.F90)
// ------------------------------------------------------------------------------------------------
struct CcmcArgs {
    long long nstates;          // psip_list%nstates
    long long nattempts;        // cluster-kernel attempts: nstochastic + nD0_select
    long long nstochastic;      // selection_data%nstochastic_clusters (enters pselect)
    long long D0_pos;           // 1-based position of the reference in the main list
    double D0_normalisation;    // population on the reference / real_factor
    double tot_abs_real_pop;    // cumulative population of all excitors but the reference
    int ex_level;               // qs%ref%ex_level (CC truncation level)
    int max_cluster_size;       // min(nel, ex_level+2, nstates-1)
    int nprocs;
    // set_cluster_selections (src/ccmc_selection.f90:874-948): full_nc = 0: all nattempts clusters are stochastic with
    // sizes from 0; full_nc = 1: composite clusters (size >= 2) are stochastic, every excitor is a non-composite cluster
    // of its own and the reference is selected nD0_select times
    int full_nc, min_cluster_size;
    long long nD0_select;
};
struct Cluster {
    int nexcitors, excitation_level, sign;   // excitation_level < 0: cluster not allowed (huge(0) in the reference)
    double pselect, amplitude;
    long long first_pos;                     // 1-based position of the first excitor (cdet%data)
};

// cumulative_excip_pop(pos): cumulative |population| / real_factor of excitors 1..pos with the reference skipped
// (cumulative_population, src/ccmc_utils.F90:427-563).  The engine keeps the exact integer prefix sums of the
// encoded populations; the quotient equals the reference's running double sum while the total stays below 2^53
// encoded units (2^22 walkers with real amplitudes) - see DESIGN.md.
HB_HD double ccmc_cum(const long long* __restrict__ cum_enc, long long pos, double real_factor) {
    return (double)cum_enc[pos - 1] / real_factor;
}
// binary_search_real_p (src/search.f90:378-481) followed by the step back over zero-population entries
// (src/ccmc_selection.f90:253-262)
HB_HDN long long ccmc_find_excitor(const long long* __restrict__ cum_enc, double item, long long istart, long long iend,
                                   double real_factor) {
    long long pos = istart;
    if (istart <= iend) {
        long long lo = istart, hi = iend;
        bool hit = false;
        while (hi != lo) {
            pos = (hi + lo) / 2;
            const double compare = item - ccmc_cum(cum_enc, pos, real_factor);
            if (fabs(compare) < 1.e-12) { hit = true; break; }
            else if (compare > 0.0) lo = pos + 1;
            else hi = pos;
        }
        if (!hit) {
            const double compare = item - ccmc_cum(cum_enc, hi, real_factor);
            if (fabs(compare) < 1.e-12) pos = hi;
            else if (compare > 0.0) pos = hi + 1;
            else pos = hi;
        }
    }
    for (;;) {
        if (pos == 1) break;
        if (fabs(ccmc_cum(cum_enc, pos, real_factor) - ccmc_cum(cum_enc, pos - 1, real_factor)) > 1.e-12) break;
        pos = pos - 1;
    }
    return pos;
}
template <int W>
HB_HD void excit_mask_of(int orb, uint64_t* m) {
    const int iw = (orb - 1) >> 6, ib = (orb - 1) & 63;
#pragma unroll
    for (int k = 0; k < W; ++k) m[k] = (k < iw) ? 0ull : ((k > iw) ? ~0ull : ((ib == 63) ? 0ull : (~0ull << (ib + 1))));
}
// collapse_cluster + collapse_excitor_onto_cluster (src/ccmc_utils.F90:132-295)
template <int W>
HB_HDN bool ccmc_collapse(const uint64_t* f0, const uint64_t* excitor, double excitor_population, uint64_t* cluster_excitor,
                          double& cluster_population) {
    uint64_t ee[W], ca[W], cc[W];
    bool clash = false;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        ee[k] = f0[k] ^ excitor[k];
        const uint64_t ce = f0[k] ^ cluster_excitor[k];
        const uint64_t ea = ee[k] & f0[k], ec = ee[k] & excitor[k];
        ca[k] = ce & f0[k];
        cc[k] = ce & cluster_excitor[k];
        if ((ec & cc[k]) != 0 || (ea & ca[k]) != 0) clash = true;
    }
    cluster_population = cluster_population * excitor_population;
    if (clash) return false;
#pragma unroll
    for (int ib = 0; ib < W; ++ib) {
        uint64_t x = ee[ib];
        while (x) {
            const int bit = ctz64(x);
            x &= x - 1;
            const int orb = ib * 64 + bit + 1;
            uint64_t mask[W];
            excit_mask_of<W>(orb, mask);
            int n = 0;
            if ((f0[ib] >> bit) & 1ull) {
                cluster_excitor[ib] &= ~(1ull << bit);
#pragma unroll
                for (int k = 0; k < W; ++k) n += popc64((mask[k] & ca[k]) | cc[k]);
            } else {
                cluster_excitor[ib] |= (1ull << bit);
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    uint64_t pm = (~mask[k]) & cc[k];
                    if (k == ib) pm &= ~(1ull << bit);
                    n += popc64(pm);
                }
            }
            if (n & 1) cluster_population = -cluster_population;
        }
    }
    return true;
}
// convert_excitor_to_determinant (src/ccmc_utils.F90:296-412): sign of the excitor applied to the reference
template <int W>
HB_HDN int ccmc_excitor_sign(const uint64_t* f0, const uint64_t* excitor, int excitor_level) {
    int nann = excitor_level, ncre = excitor_level, sign = 1;
#pragma unroll
    for (int ib = 0; ib < W; ++ib) {
        const uint64_t ex = f0[ib] ^ excitor[ib];
        uint64_t x = f0[ib] | ex;        // only these bit positions change state in the reference's loop
        while (x) {
            const int bit = ctz64(x);
            x &= x - 1;
            if ((f0[ib] >> bit) & 1ull) {
                if ((ex >> bit) & 1ull) nann--;
                else if ((nann + ncre) & 1) sign = -sign;
            } else {
                ncre--;
            }
        }
    }
    return sign;
}
// select_cluster (src/ccmc_selection.f90:93-392; linked = false, no discard threshold) with create_null_cluster
// (:394-460).  cf receives the collapsed excitor (cdet%f).
template <int W, class R>
HB_HDN void ccmc_select_cluster(R& rng, const Params& p, const CcmcArgs& a, const uint64_t* __restrict__ states,
                                const int64_t* __restrict__ pops, const long long* __restrict__ cum_enc, uint64_t* cf,
                                Cluster& cl) {
    const int min_size = a.min_cluster_size, max_size = a.max_cluster_size;
    const double rf = (double)p.real_factor;
    cl.pselect = (double)(a.nstochastic * a.nprocs);
    const double rand = rng.next();
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
    cl.first_pos = 0;
    if (cl.nexcitors == 0) {
#pragma unroll
        for (int k = 0; k < W; ++k) cf[k] = p.f0[k];
        cl.excitation_level = 0;
        cl.amplitude = a.D0_normalisation;
        cl.sign = 1;
        return;
    }
    double pop[HB_MAX_CLUSTER];
    for (int i = 0; i < cl.nexcitors; ++i) pop[i] = rng.next() * a.tot_abs_real_pop;
    for (int i = 1; i < cl.nexcitors; ++i) {      // insert_sort_real_p (src/sort.f90:827-852)
        int j = i - 1;
        const double tmp = pop[i];
        while (j >= 0 && !(pop[j] <= tmp)) { pop[j + 1] = pop[j]; j--; }
        pop[j + 1] = tmp;
    }
    long long prev_pos = 1;
    double cluster_population = 0.0;
    bool allowed = min_size <= max_size;
    for (int i = 1; i <= cl.nexcitors; ++i) {
        const long long pos = ccmc_find_excitor(cum_enc, pop[i - 1], prev_pos, a.nstates, rf);
        const double excitor_pop = (double)pops[pos - 1] / rf;
        uint64_t ex[W];
#pragma unroll
        for (int k = 0; k < W; ++k) ex[k] = states[(pos - 1) * W + k];
        if (i == 1) {
#pragma unroll
            for (int k = 0; k < W; ++k) cf[k] = ex[k];
            cl.first_pos = pos;
            cluster_population = excitor_pop;
            cl.pselect = cl.pselect / a.nprocs;
        } else {
            const bool ok = ccmc_collapse<W>(p.f0, ex, excitor_pop, cf, cluster_population);
            if (!ok) { allowed = false; break; }
            if (pos != prev_pos) cl.pselect = cl.pselect / a.nprocs;
        }
        cl.pselect = (cl.pselect * fabs(excitor_pop)) / a.tot_abs_real_pop;
        prev_pos = pos;
    }
    if (allowed) {
        cl.excitation_level = excit_level<W>(p.f0, cf);
        allowed = cl.excitation_level <= a.ex_level + 2;
    }
    if (allowed) {
        double fact = 1.0;
        for (int k = 2; k <= cl.nexcitors; ++k) fact = fact * k;
        cl.pselect = cl.pselect * fact;
        cl.sign = ccmc_excitor_sign<W>(p.f0, cf, cl.excitation_level);
        double norm_pow = 1.0;
        for (int k = 0; k < cl.nexcitors - 1; ++k) norm_pow = norm_pow * a.D0_normalisation;
        cl.amplitude = cluster_population / norm_pow;
    } else {
        cl.excitation_level = -1;
    }
}

}  // namespace hb
