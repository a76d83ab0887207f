// hande_b200: the spawning kernel of the original heat-bath generator (excit_gen = heat_bath, the bench headline) as a
// warp-synchronous wavefront.
//
// Why not "one warp per determinant": the per-attempt work is a chain of ~3,000 dependent scalar operations (two alias
// selections over the occupied orbitals, up to four slater_condon1 sums in occ_list order, a dozen fp64 divisions); the
// kernel is instruction-bound, and a warp that cooperates on one attempt executes the serial parts 32 times over.  So a
// LANE owns an attempt, but nothing is done under divergence: every warp owns 64 attempt slots and runs the generator
// (gen_excit_mol_heat_bath, src/excit_gen_heat_bath_mol.F90:258-548) as phases over dense, ballot-compacted queues of
// those slots -
//   A  i, j (alias selections over the occupied orbitals), a (precomputed alias row)          per slot
//   B  slater_condon1(i -> a) for the slots where that single excitation is allowed           queue
//   C  single/double coin, b (2 GB hb_ijab row: one packed 32-byte record)                    per slot
//   D  |slater_condon1| of the other orderings i->b, j->a, j->b that enter pgen               queue
//   E  the nel pgen terms of every single excitation, one term per lane                       queue
//   F  pgen, H_ij, attempt_to_spawn, child, owner, warp-aggregated append                     queue (doubles, then singles)
// - with no block barrier anywhere (warps drift apart and hide each other's L2/DRAM round trips), slot state in the
// warp's own shared memory, and the queues rebuilt with __ballot_sync.  Excitation choice, pgen, H_ij and nspawn are
// bit-identical to the oracle: identical operations in identical order; the table look-ups use the exact symmetries of
// the reference's tables (ij_weights(j,i) = ij_weights(i,j), hb_ija%weights(a,j,i) = hb_ijab%weights_tot(a,j,i) =
// ...(a,i,j), hb_ijab%weights invariant under i<->j and a<->b - they are the same sums of the same numbers, checked
// bitwise in tests/test_core_vs_oracle.py) to read 7 numbers where the four-ordering pgen formula names 20.
#pragma once
#include "hb_common.cuh"

namespace hbw {

constexpr int SLOTS = 64;      // attempt slots (and states per tile) of a warp: two per lane
constexpr int NWARP = 4;       // warps per block
constexpr int HEAVY = 4096;    // a state with more attempts than this is deferred to k_spawn_hb_heavy

enum { FL_ALLOWED = 1, FL_NEED_IA = 2, FL_DBL = 4, FL_PERM_IA = 8, FL_NEEDK = 0x70, FL_WKNOWN = 0x80 };

// byte offsets inside a warp's private shared-memory region (same arithmetic on host and device)
struct WarpSmem {
    int sf, shash, sdf, sitot, sijtot, sr3, sps, swab, spab, stage, sh1, shm, spsum, sterm, sscan, sij, socc, sbits, slo, sfl, qB,
        qD, qE, qF, total, term_chunk, noccw;
    __host__ __device__ WarpSmem(int W, int nel, bool qn) {
        int o = 0;
        sf = o;      o += SLOTS * W * 8;
        shash = o;   o += SLOTS * 8;
        sdf = o;     o += qn ? SLOTS * 8 : 0;        // quasi-Newton: fock_sum of each state
        sitot = o;   o += SLOTS * 8;                 // per slot: sum of S_i over the occupied orbitals
        sijtot = o;  o += SLOTS * 8;                 // per slot: sum of ij_weights(:, i) over the occupied orbitals
        sr3 = o;     o += SLOTS * 8;                 // per slot: the stream's next uniform (draw 3, later the spawn draw)
        sps = o;     o += SLOTS * 8;                 // per slot: psingle
        swab = o;    o += SLOTS * 8;                 // per slot: hb_ijab%weights(b,a,j,i) when the selection record held it
        spab = o;    o += SLOTS * 8;                 //           and its weights(b,a,j,i) / weights_tot(a,j,i)
        // phase A staging area ([q][lane] weights of the list being selected from), reused by the later phases
        stage = o;
        const int later = SLOTS * 8 * 5 + 16 * nel * 8;   // h1, hm[3], psum, pgen terms of 16 singles
        const int stage_bytes = (nel * 32 * 8 > later) ? nel * 32 * 8 : later;
        sh1 = stage; shm = stage + SLOTS * 8; spsum = stage + SLOTS * 8 * 4; sterm = stage + SLOTS * 8 * 5;
        term_chunk = (stage_bytes - SLOTS * 8 * 5) / (nel * 8);
        o += stage_bytes;
        sscan = o;   o += (SLOTS + 1) * 4; o = (o + 7) & ~7;
        sij = o;     o += SLOTS * 4;                 // i, j, a, b
        noccw = (nel + 3) >> 2;                      // occupied lists: four orbitals to a word, zero padded
        socc = o;    o += SLOTS * noccw * 4;
        sbits = o;   o += SLOTS;                     // per state: bit 0 sign of the population, bit 1 set_parent_flag
        slo = o;     o += SLOTS;                     // per slot: state of the tile it belongs to
        sfl = o;     o += SLOTS;
        qB = o;      o += SLOTS;
        qD = o;      o += 3 * SLOTS;
        qE = o;      o += SLOTS;
        qF = o;      o += SLOTS;
        total = (o + 15) & ~15;
    }
};

struct HeavyItem { long long state; long long pop; int natt, pad; };   // pop: the population before death (sign, initiator flag)
struct HeavyQueue { HeavyItem* items; unsigned* count; unsigned cap; };

struct OneDraw { double v; __device__ __forceinline__ double next() { return v; } };

__device__ __forceinline__ double u01(uint32_t lo, uint32_t hi) {
    const uint64_t u = ((uint64_t)hi << 32) | lo;
    return (double)(u >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ bool smem_det_test(const uint64_t* f, int orb) { return (f[(orb - 1) >> 6] >> ((orb - 1) & 63)) & 1ull; }

// The kernel's warps drift through different phases, so its code has to stay resident in the instruction cache: the
// big leaf routines (Philox block, alias selection, the occupied-list sums) are real functions, not inlined copies.
static __device__ __noinline__ uint4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    uint32_t r[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, r);
    return make_uint4(r[0], r[1], r[2], r[3]);
}
// The occupied list of a determinant is kept four orbitals to a 32-bit word (zero padded) so that a pass over it costs
// one shared-memory load per four orbitals; the table loads of up to 20 orbitals are issued together (one round trip).
constexpr int OCC_CHUNK = 2;    // words per batch (8 independent loads in flight)
// gather tab1[occ[q]] (tab1 = table pointer minus one element: orbitals are 1-based) into wq[q*32] and return the sum in
// list order (find_i_d_weights / ij_weights_occ, src/excit_gen_utils.f90:142-160).  STAGE = false: the sum only.
template <bool STAGE>
__device__ __forceinline__ double gather_occ(const double* __restrict__ tab1, const uint32_t* occw, int nel, double* wq) {
    double tot = 0.0;
    const int nfull = nel >> 2;
#pragma unroll 1
    for (int w0 = 0; w0 < nfull; w0 += OCC_CHUNK) {
        double v[OCC_CHUNK][4];
#pragma unroll
        for (int c = 0; c < OCC_CHUNK; ++c)
            if (w0 + c < nfull) {
                const uint32_t o4 = occw[w0 + c];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[c][k] = tab1[(o4 >> (8 * k)) & 0xffu];
            }
#pragma unroll
        for (int c = 0; c < OCC_CHUNK; ++c)
            if (w0 + c < nfull) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (STAGE) wq[(4 * (w0 + c) + k) * 32] = v[c][k];
                    tot = tot + v[c][k];
                }
            }
    }
    if (nel & 3) {
        const uint32_t o4 = occw[nfull];
        for (int k = 0; k < (nel & 3); ++k) {
            const double v = tab1[(o4 >> (8 * k)) & 0xffu];
            if (STAGE) wq[(4 * nfull + k) * 32] = v;
            tot = tot + v;
        }
    }
    return tot;
}
// slater_condon1_mol_excit (src/hamiltonian_molecular.f90:199-259) for the excitation i -> a: h = <i|h|a> plus the sum over
// the occupied list in occ_list order of the row entries {<ij|aj>, <ij|ja> or 0} (Sys::sc1T; the entry of i itself and the
// list padding are {0, 0}: adding them leaves every partial sum unchanged).  No branches, eight loads in flight.
static __device__ __forceinline__ double sc1_row(const D2* __restrict__ row, double h, const uint32_t* occw, int noccw) {
#pragma unroll 1
    for (int w0 = 0; w0 < noccw; w0 += 2) {
        D2 v[2][4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t o4 = (w0 + c < noccw) ? occw[w0 + c] : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[c][k] = row[(o4 >> (8 * k)) & 0xffu];
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) { h = h + v[c][k].x; h = h - v[c][k].y; }
    }
    return h;
}
__device__ __forceinline__ double sc1_lean(const Sys& s, const uint32_t* occw, int noccw, int i, int a) {
    const unsigned ta = s.uhf ? (unsigned)(a - 1) : ((unsigned)(a - 1) >> 1);
    const D2* row = s.sc1T + ((size_t)((unsigned)(i - 1) * (unsigned)s.sc1A + ta)) * (unsigned)(s.nbasis + 1);
    return sc1_row(row, one_body(s, i, a), occw, noccw);
}

template <int W>
struct WarpCtx {
    uint64_t* sf; uint64_t* shash; double* sdf; double* sitot; double* sijtot; double* sr3; double* sps; double* swab;
    double* spab; double* stage; double* sh1; double* shm; double* spsum; double* sterm; int* sscan; uchar4* sij; uint32_t* socc;
    uint8_t* sbits; uint8_t* slo; uint8_t* sfl; uint8_t* qB; uint8_t* qD; uint8_t* qE; uint8_t* qF;
    const double* siw1;      // shared-memory copy of hb_i_w, 1-based: siw1[o] = S_o
    int term_chunk, noccw;
    __device__ WarpCtx(unsigned char* base, const WarpSmem& L, const double* siw_) {
        sf = (uint64_t*)(base + L.sf); shash = (uint64_t*)(base + L.shash); sdf = (double*)(base + L.sdf);
        sitot = (double*)(base + L.sitot); sijtot = (double*)(base + L.sijtot); sr3 = (double*)(base + L.sr3);
        sps = (double*)(base + L.sps); swab = (double*)(base + L.swab); spab = (double*)(base + L.spab);
        stage = (double*)(base + L.stage);
        sh1 = (double*)(base + L.sh1); shm = (double*)(base + L.shm); spsum = (double*)(base + L.spsum);
        sterm = (double*)(base + L.sterm); sscan = (int*)(base + L.sscan); sij = (uchar4*)(base + L.sij);
        socc = (uint32_t*)(base + L.socc); sbits = base + L.sbits; slo = base + L.slo; sfl = base + L.sfl; qB = base + L.qB;
        qD = base + L.qD; qE = base + L.qE; qF = base + L.qF;
        siw1 = siw_ - 1; term_chunk = L.term_chunk; noccw = L.noccw;
    }
};

// Phase boundary.  The warps of a block need no data from each other, but they are kept in the same phase: the kernel's
// code (~150 KB) is several times the instruction cache, and warps drifting through different phases thrash it
// (measured: instruction-cache hit rate 69 % free-running against 95 % in step).
__device__ __forceinline__ void phase_sync() { __syncthreads(); }

// All spawning attempts [0, T) of the tile whose states (sf, shash, socc, sbits, sdf) and exclusive attempt scan
// (sscan[0..nst]) sit in the warp's shared memory; att_off is added to the attempt index that keys the random stream
// (0, or the first attempt of a chunk of a heavy state).  nrounds is the same for every warp of the block (>= the rounds
// this warp needs): all of them pass the phase boundaries together.
template <int W, class Mask>
__device__ void warp_spawn_rounds(const Sys& s, const Params& p, WarpCtx<W>& c, int nst, int T, int nrounds, uint32_t att_off,
                                  int64_t* __restrict__ spawn, unsigned long long* __restrict__ head, long long block_size,
                                  const int* __restrict__ proc_map, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nel = s.nel;
    const int64_t nb = s.nbasis;
    constexpr int E = W + 2;
    for (int rnd = 0; rnd < nrounds; ++rnd) {
        const int base = rnd * SLOTS;
        // ================= phase A: i, j, a =================
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const int slot = sub * 32 + lane;
            const int a_idx = base + slot;
            uint8_t fl = 0;
            if (a_idx < T) {
                int lo = 0, hi = nst;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (c.sscan[mid] <= a_idx) lo = mid; else hi = mid;
                }
                const uint32_t att = (uint32_t)(a_idx - c.sscan[lo]) + att_off;
                const uint32_t* occw = c.socc + lo * c.noccw;
                const uint8_t* occ = reinterpret_cast<const uint8_t*>(occw);
                const uint64_t* f = c.sf + lo * W;
                const uint64_t hsh = c.shash[lo];
                const uint4 ra_ = philox_block(((uint32_t)RNG_SPAWN << 24) | 0u, att, (uint32_t)hsh, (uint32_t)(hsh >> 32), p.seed, p.cycle);
                const uint4 rb_ = philox_block(((uint32_t)RNG_SPAWN << 24) | 1u, att, (uint32_t)hsh, (uint32_t)(hsh >> 32), p.seed, p.cycle);
                const uint32_t r[8] = {ra_.x, ra_.y, ra_.z, ra_.w, rb_.x, rb_.y, rb_.z, rb_.w};
                double* wq = c.stage + lane;
                // i from S_i over the occupied orbitals (select_ij_heat_bath, src/excit_gen_utils.f90:9-66)
                const double i_tot = gather_occ<true>(c.siw1, occw, nel, wq);
                double x = u01(r[0], r[1]) * nel;
                int k = (int)x;
                x = x - k;
                const int i = occ[alias_select_fast<Mask>(nel, wq, 32, nel / i_tot, k, x) - 1];
                // j from ij_weights(:, i) over the occupied orbitals
                const double ij_tot = gather_occ<true>(s.hb_ij_w + nb * (i - 1) - 1, occw, nel, wq);
                int j = 0, a = 0;
                bool allowed = false;
                if (ij_tot > 0.0) {
                    x = u01(r[2], r[3]) * nel;
                    k = (int)x;
                    x = x - k;
                    j = occ[alias_select_fast<Mask>(nel, wq, 32, nel / ij_tot, k, x) - 1];
                    allowed = fabs(s.hb_ija_tot[HB_I2(j, i)]) > 0.0;
                }
                bool need_ia = false;
                if (allowed) {
                    // a from the precomputed alias row hb_ija(:, j, i)
                    x = u01(r[4], r[5]) * (int)nb;
                    const int K = (int)floor(x);
                    x = x - K;
                    const HbRec* rec = s.hb_ija_rec + HB_I3(1, j, i) + K;
                    const double U = rec->U;
                    const int alias = rec->K;
                    a = (x < U) ? K + 1 : alias;
                    if (smem_det_test(f, a)) allowed = false;
                    else need_ia = hb_single_allowed(s, i, a);
                }
                fl = (allowed ? FL_ALLOWED : 0) | (need_ia ? FL_NEED_IA : 0);
                c.slo[slot] = (uint8_t)lo;
                c.sij[slot] = make_uchar4((unsigned char)i, (unsigned char)j, (unsigned char)a, 0);
                c.sitot[slot] = i_tot;
                c.sijtot[slot] = ij_tot;
                c.sr3[slot] = u01(r[6], r[7]);
            }
            c.sfl[slot] = fl;
        }
        phase_sync();
        // ================= queue B: slots whose i -> a single excitation is allowed =================
        int nB = 0;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
            const int slot = sub * 32 + lane;
            const bool want = (c.sfl[slot] & FL_NEED_IA) != 0;
            const unsigned m = __ballot_sync(0xffffffffu, want);
            if (want) c.qB[nB + __popc(m & lt)] = (uint8_t)slot;
            nB += __popc(m);
        }
        phase_sync();
        // ================= phase B: slater_condon1(i, a) =================
        for (int r0 = lane; r0 < nB; r0 += 32) {
            const int slot = c.qB[r0];
            const int lo = c.slo[slot];
            const uchar4 q = c.sij[slot];
            uint64_t ff[W];
#pragma unroll
            for (int w = 0; w < W; ++w) ff[w] = c.sf[lo * W + w];
            const bool pm = excit_perm1<W>(ff, q.x, q.z);
            const double h = sc1_lean(s, c.socc + lo * c.noccw, c.noccw, q.x, q.z);
            c.sh1[slot] = pm ? -h : h;
            if (pm) c.sfl[slot] |= FL_PERM_IA;
        }
        phase_sync();
        // ================= phase C: single/double coin, b =================
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const int slot = sub * 32 + lane;
            uint8_t fl = c.sfl[slot];
            if (fl & FL_ALLOWED) {
                const int lo = c.slo[slot];
                const uchar4 q = c.sij[slot];
                const int i = q.x, j = q.y, a = q.z;
                const uint64_t hsh = c.shash[lo];
                const uint32_t att = (uint32_t)(base + slot - c.sscan[lo]) + att_off;
                const uint4 rc_ = philox_block(((uint32_t)RNG_SPAWN << 24) | 2u, att, (uint32_t)hsh, (uint32_t)(hsh >> 32), p.seed, p.cycle);
                const double r3 = c.sr3[slot], r4 = u01(rc_.x, rc_.y), r5 = u01(rc_.z, rc_.w);
                bool dbl = true;
                double psingle = 0.0, rb = r3, rs = r4;      // draws: b, attempt_to_spawn
                if (fl & FL_NEED_IA) {
                    const double hmod = fabs(c.sh1[slot]);
                    const double wt = s.hb_ija_rec[HB_I3(a, j, i)].w;          // = hb_ijab%weights_tot(a,j,i)
                    if (hmod < wt) psingle = hmod / (wt + hmod); else psingle = 0.5;
                    dbl = !(r3 < psingle);
                    rb = r4; rs = dbl ? r5 : r4;
                }
                c.sps[slot] = psingle;
                if (dbl) {
                    fl |= FL_DBL;
                    double x = rb * (int)nb;
                    const int K = (int)floor(x);
                    x = x - K;
                    const HbRec* rec = s.hb_ijab_rec + HB_I4(1, a, j, i) + K;
                    // the whole 32-byte record, streamed (no reuse): {aliasU, weight}, {aliasK, -, weight / weights_tot}
                    const double2 uw = __ldcs(reinterpret_cast<const double2*>(rec));
                    const int4 kp = __ldcs(reinterpret_cast<const int4*>(rec) + 1);
                    const int b = (x < uw.x) ? K + 1 : kp.x;
                    if (b == K + 1) {
                        fl |= FL_WKNOWN;
                        c.swab[slot] = uw.y;
                        c.spab[slot] = __hiloint2double(kp.w, kp.z);
                    }
                    if (smem_det_test(c.sf + lo * W, b)) {
                        fl = 0;
                    } else {
                        c.sij[slot] = make_uchar4((unsigned char)i, (unsigned char)j, (unsigned char)a, (unsigned char)b);
                        if (hb_single_allowed(s, i, b)) fl |= 0x10;
                        if (hb_single_allowed(s, j, a)) fl |= 0x20;
                        if (hb_single_allowed(s, j, b)) fl |= 0x40;
                    }
                }
                c.sr3[slot] = rs;
                c.sfl[slot] = fl;
            }
        }
        phase_sync();
        // ================= queues D (other orderings), E (singles), F (survivors: doubles first) =================
        int nD = 0, nE = 0, nFd = 0;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
            const int slot = sub * 32 + lane;
            const uint8_t fl = c.sfl[slot];
            const bool isd = (fl & FL_ALLOWED) && (fl & FL_DBL), iss = (fl & FL_ALLOWED) && !(fl & FL_DBL);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const bool want = isd && (fl & (0x10 << k));
                const unsigned m = __ballot_sync(0xffffffffu, want);
                if (want) c.qD[nD + __popc(m & lt)] = (uint8_t)(slot | (k << 6));
                nD += __popc(m);
            }
            const unsigned ms = __ballot_sync(0xffffffffu, iss);
            if (iss) c.qE[nE + __popc(ms & lt)] = (uint8_t)slot;
            nE += __popc(ms);
            const unsigned md = __ballot_sync(0xffffffffu, isd);
            if (isd) c.qF[nFd + __popc(md & lt)] = (uint8_t)slot;
            nFd += __popc(md);
        }
        __syncwarp();
        for (int r0 = lane; r0 < nE; r0 += 32) c.qF[nFd + r0] = c.qE[r0];
        const int nF = nFd + nE;
        phase_sync();
        // ================= phase D: |slater_condon1| of the orderings i->b, j->a, j->b =================
        for (int r0 = lane; r0 < nD; r0 += 32) {
            const int rq = c.qD[r0];
            const int slot = rq & 63, k = rq >> 6;
            const int lo = c.slo[slot];
            const uchar4 q = c.sij[slot];
            const int fr = (k == 0) ? q.x : q.y, to = (k == 1) ? q.z : q.w;
            c.shm[k * SLOTS + slot] = fabs(sc1_lean(s, c.socc + lo * c.noccw, c.noccw, fr, to));
        }
        // ================= phase E: generation probability of the singles (sum over spectator orbitals) =================
        for (int c0 = 0; c0 < nE; c0 += c.term_chunk) {
            const int nc = min(c.term_chunk, nE - c0);
            int rr = lane / nel, qq = lane - rr * nel;          // item w = rr * nel + qq, advanced by 32 per round
            for (int w = lane; w < nc * nel; w += 32) {
                const int slot = c.qE[c0 + rr];
                const uchar4 q = c.sij[slot];
                const int i = q.x, a = q.z;
                const int oq = reinterpret_cast<const uint8_t*>(c.socc + c.slo[slot] * c.noccw)[qq];
                double term = 0.0;
                if (i != oq && a != oq) {
                    const double hmod = fabs(c.sh1[slot]);
                    const HbRec* rec = s.hb_ija_rec + HB_I3(a, oq, i);
                    const double Taq = rec->w;      // hb_ija%weights(a,oq,i) = hb_ijab%weights_tot(a,oq,i)
                    const double paq = rec->p;      // hb_ija%weights(a,oq,i) / hb_ija%weights_tot(oq,i)
                    double psq;
                    if (hmod < Taq) psq = hmod / (Taq + hmod); else psq = 0.5;
                    term = (psq * (s.hb_ij_w[HB_I2(oq, i)] / c.sijtot[slot]) * paq);
                }
                c.sterm[w] = term;
                qq += 32;
                while (qq >= nel) { qq -= nel; rr++; }
            }
            __syncwarp();
            for (int r1 = lane; r1 < nc; r1 += 32) {
                double psum = 0.0;
                for (int q1 = 0; q1 < nel; ++q1) psum = psum + c.sterm[r1 * nel + q1];
                c.spsum[c.qE[c0 + r1]] = psum;
            }
            __syncwarp();
        }
        phase_sync();
        // ================= phase F: pgen, H_ij, spawn =================
        for (int r0 = 0; r0 < nF; r0 += 32) {
            const int rr = r0 + lane;
            int64_t nspawn = 0;
            uint64_t child[W];
            int dest = 0, pflag = 0;
            if (rr < nF) {
                const int slot = c.qF[rr];
                const int lo = c.slo[slot];
                const uchar4 q = c.sij[slot];
                const uint8_t fl = c.sfl[slot];
                const int i = q.x, j = q.y, a = q.z, b = q.w;
                uint64_t f[W];
#pragma unroll
                for (int w = 0; w < W; ++w) f[w] = c.sf[lo * W + w];
                const double i_tot = c.sitot[slot];
                Gen g;
                g.allowed = true;
                if (fl & FL_DBL) {
                    const double ij_tot = c.sijtot[slot];
                    const double ji_tot = gather_occ<false>(s.hb_ij_w + nb * (j - 1) - 1, c.socc + lo * c.noccw, nel, nullptr);
                    const double wij = s.hb_ij_w[HB_I2(j, i)];                  // = ij_weights(i,j)
                    // hb_ija%weights(a,j,i) = hb_ijab%weights_tot(a,j,i) = ...(a,i,j), and the same divided by
                    // hb_ija%weights_tot(j,i) = ...(i,j)
                    const HbRec* ra_ = s.hb_ija_rec + HB_I3(a, j, i);
                    const HbRec* rb_ = s.hb_ija_rec + HB_I3(b, j, i);
                    const double Ta = ra_->w, ra = ra_->p, Tb = rb_->w, rb = rb_->p;
                    double wab, wa;                                             // hb_ijab%weights(b,a,j,i) and it / weights_tot(a,j,i)
                    if (fl & FL_WKNOWN) { wab = c.swab[slot]; wa = c.spab[slot]; }
                    else {
                        const HbRec* rw = s.hb_ijab_rec + HB_I4(b, a, j, i);
                        wab = __ldcs(&rw->w); wa = __ldcs(&rw->p);
                    }
                    const double pi_ = c.siw1[i] / i_tot;
                    const double pj_ = c.siw1[j] / i_tot;
                    const double pij = wij / ij_tot;
                    const double pji = wij / ji_tot;
                    double ps[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        ps[k] = 0.0;
                        if (fl & (0x10 << k)) {
                            const double hm = c.shm[k * SLOTS + slot];
                            const double wt = (k == 1) ? Ta : Tb;
                            if (hm < wt) ps[k] = hm / (wt + hm); else ps[k] = 0.5;
                        }
                    }
                    const double wb = wab / Tb;
                    const double pgen_ija = ((pi_) * (pij)) * ra * (1.0 - c.sps[slot]) * wa;
                    const double pgen_ijb = ((pi_) * (pij)) * rb * (1.0 - ps[0]) * wb;
                    const double pgen_jia = ((pj_) * (pji)) * ra * (1.0 - ps[1]) * wa;
                    const double pgen_jib = ((pj_) * (pji)) * rb * (1.0 - ps[2]) * wb;
                    g.pgen = pgen_ija + pgen_ijb + pgen_jia + pgen_jib;
                    g.from1 = (i < j) ? i : j; g.from2 = (i < j) ? j : i;
                    g.to1 = (a < b) ? a : b; g.to2 = (a < b) ? b : a;
                    g.nexcit = 2;
                    g.perm = excit_perm2<W>(f, g.from1, g.from2, g.to1, g.to2);
                    g.hmatel = slater_condon2_excit(s, g.from1, g.from2, g.to1, g.to2, g.perm);
                } else {
                    g.nexcit = 1; g.from1 = i; g.to1 = a; g.from2 = 0; g.to2 = 0;
                    g.perm = (fl & FL_PERM_IA) != 0;
                    g.hmatel = c.sh1[slot];
                    g.pgen = c.spsum[slot] * (c.siw1[i] / i_tot);
                }
                double hmq = g.hmatel;
                if (p.qn) hmq = hmq * qn_spawned_weighting(p, c.sdf[lo], g);   // spawn_standard (src/spawning.F90:101-103)
                hmq = hmq * p.cheby_weight;
                OneDraw rng{c.sr3[slot]};
                const uint8_t bits = c.sbits[lo];
                nspawn = attempt_to_spawn(rng, p, hmq, g.pgen, (bits & 1) ? (int64_t)-1 : (int64_t)1);
                if (nspawn != 0) {
                    make_child<W>(f, g, child);
                    // create_spawned_particle[_initiator]_truncated (src/spawning.F90:1186-1319)
                    if (p.trunc_level >= 0 && excit_level<W>(child, p.f0) > p.trunc_level) {
                        nspawn = 0;
                    } else {
                        // assign_particle_processor (src/spawning.F90:770-838)
                        dest = (p.nprocs > 1) ? proc_map[owner_slot(child, s.nbasis, p.hash_seed, p.nprocs, p.nslots)] : 0;
                        pflag = p.initiator ? ((bits >> 1) & 1) : 0;
                    }
                }
            }
            const unsigned has = __ballot_sync(0xffffffffu, nspawn != 0);
            if (nspawn != 0) {
                // add_[flagged_]spawned_particle (src/spawning.F90:907-1018): warp-aggregated pointer bump per destination
                const unsigned peers = (p.nprocs > 1) ? __match_any_sync(has, dest) : has;
                const int leader = __ffs(peers) - 1;
                const int rank = __popc(peers & lt);
                unsigned long long slot0 = 0;
                if (lane == leader) slot0 = atomicAdd(&head[dest], (unsigned long long)__popc(peers));
                slot0 = __shfl_sync(peers, slot0, leader);
                const long long sl = (long long)slot0 + rank;
                if (sl < block_size) {
                    int64_t* dst = spawn + ((long long)dest * block_size + sl) * E;
                    if (W == 2) {
                        reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(child[0], child[1]);
                        reinterpret_cast<longlong2*>(dst)[1] = make_longlong2((long long)nspawn, (long long)pflag);
                    } else {
#pragma unroll
                        for (int w = 0; w < W; ++w) dst[w] = (int64_t)child[w];
                        dst[W] = nspawn;
                        dst[W + 1] = pflag;
                    }
                } else {
                    atomicOr(err, 1);  // spawn%error: no space left in the spawning array
                }
            }
        }
        phase_sync();
    }
}

// Per-state part of the idet loop (src/fciqmc.f90:315-371): decode, set_parent_flag, update_proj_energy_mol,
// decide_nattempts, stochastic_death.  Fills the tile entry `st` of the warp; returns the number of spawning attempts.
template <int W>
__device__ __forceinline__ int state_prologue(const Sys& s, const Params& p, WarpCtx<W>& c, int st, const uint64_t* f, int64_t pop,
                                              double Kii, bool do_death, int64_t* __restrict__ pops, long long idx, double& pe,
                                              double& d0, long long& ndeath, long long& npart) {
    const int nel = s.nel;
#pragma unroll
    for (int k = 0; k < W; ++k) c.sf[st * W + k] = f[k];
    uint8_t* occ = reinterpret_cast<uint8_t*>(c.socc + st * c.noccw);
    c.socc[st * c.noccw + c.noccw - 1] = 0u;       // zero padding of the last word
    decode_det<W>(f, occ);
    const uint64_t h = det_hash64<W>(f);
    c.shash[st] = h;
    const double real_pop = (double)pop / (double)p.real_factor;
    // set_parent_flag (src/ifciqmc.f90:13-57)
    c.sbits[st] = (uint8_t)((pop < 0 ? 1 : 0) | ((fabs(real_pop) > p.initiator_pop) ? 0 : 2));
    double dfock = 0.0;
    if (p.qn) { dfock = qn_fock_sum(s, p, occ); c.sdf[st] = dfock; }
    // decide_nattempts (src/qmc_common.F90:379-406): the stream of purpose RNG_NATTEMPTS is drawn from only for a
    // fractional population
    int natt = (int)real_pop;
    if (natt < 0) natt = -natt;
    const double pextra = fabs(real_pop) - natt;
    if (fabs(pextra) > 1.e-12) {
        const uint4 r = philox_block(((uint32_t)RNG_NATTEMPTS << 24) | 0u, 0u, (uint32_t)h, (uint32_t)(h >> 32), p.seed, p.cycle);
        if (pextra > u01(r.x, r.y)) natt++;
    }
    if (do_death) {
        // update_proj_energy_mol (src/energy_evaluation.F90:906-986)
        bool is_ref;
        const double hm = proj_energy_hmatel<W>(s, p, f, occ, is_ref);
        if (is_ref) d0 += real_pop; else pe += hm * real_pop;
        const uint4 r = philox_block(((uint32_t)RNG_DEATH << 24) | 0u, 0u, (uint32_t)h, (uint32_t)(h >> 32), p.seed, p.cycle);
        OneDraw rng{u01(r.x, r.y)};
        int64_t kill_abs;
        const double death_weight = p.qn ? qn_weighting(p, dfock) : 1.0;
        const int64_t newpop = stochastic_death(rng, p, Kii, pop, kill_abs, death_weight);
        pops[idx] = newpop;
        ndeath += kill_abs;
        npart += newpop < 0 ? -newpop : newpop;
    }
    return natt;
}

template <int W, class Mask>
__global__ void __launch_bounds__(NWARP * 32, 4)
k_spawn_hb(Sys s, Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ pops, const double* __restrict__ dat,
           long long nstates, int64_t* __restrict__ spawn, unsigned long long* __restrict__ head, long long block_size,
           const int* __restrict__ proc_map, SpawnPartials* __restrict__ partials, int* __restrict__ err, HeavyQueue hq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WarpSmem L(W, s.nel, p.qn != 0);
    double* siw = reinterpret_cast<double*>(smem_raw);
    const int siw_bytes = (s.nbasis * 8 + 15) & ~15;
    for (int k = tid; k < s.nbasis; k += NWARP * 32) siw[k] = s.hb_i_w[k];
    __syncthreads();
    WarpCtx<W> c(smem_raw + siw_bytes + warp * L.total, L, siw);
    double pe = 0.0, d0 = 0.0;
    long long ndeath = 0, npart = 0, nattempts = 0;
    const long long ntile = (nstates + SLOTS - 1) / SLOTS;
    const long long gwarp = (long long)blockIdx.x * NWARP + warp, nwarps = (long long)gridDim.x * NWARP;
    __shared__ int s_rounds[NWARP];
    const long long niter = (ntile + nwarps - 1) / nwarps;
    for (long long it = 0; it < niter; ++it) {
        const long long t = gwarp + it * nwarps;      // t >= ntile: an empty tile (the warp only keeps in step)
        int natt[2];
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const long long idx = t * SLOTS + sub * 32 + lane;
            natt[sub] = 0;
            if (idx < nstates) {
                uint64_t f[W];
                load_det<W>(states + idx * W, f);
                const int64_t pop = __ldcs(pops + idx);
                const double Kii = __ldcs(dat + idx);
                int n = state_prologue<W>(s, p, c, sub * 32 + lane, f, pop, Kii, true, pops, idx, pe, d0, ndeath, npart);
                nattempts += n;
                if (n > HEAVY) {     // a determinant with a huge population: its attempts are spread over the whole grid later
                    const unsigned k = atomicAdd(hq.count, 1u);
                    if (k < hq.cap) { HeavyItem hi; hi.state = idx; hi.pop = pop; hi.natt = n; hi.pad = 0; hq.items[k] = hi; n = 0; }
                }
                natt[sub] = n;
            }
        }
        const int i0 = warp_incl_scan(natt[0]);
        const int t0 = __shfl_sync(0xffffffffu, i0, 31);
        const int i1 = warp_incl_scan(natt[1]) + t0;
        const int T = __shfl_sync(0xffffffffu, i1, 31);
        c.sscan[lane] = i0 - natt[0];
        c.sscan[32 + lane] = i1 - natt[1];
        if (lane == 0) c.sscan[SLOTS] = T;
        __syncwarp();
        const int nst = (int)max(1LL, min((long long)SLOTS, nstates - t * SLOTS));
        if (lane == 0) s_rounds[warp] = (T + SLOTS - 1) / SLOTS;
        __syncthreads();
        int nrounds = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) nrounds = max(nrounds, s_rounds[w]);
        __syncthreads();       // s_rounds is rewritten by the next tile
        warp_spawn_rounds<W, Mask>(s, p, c, nst, T, nrounds, 0u, spawn, head, block_size, proc_map, err);
    }
    // deterministic block reduction of the estimators
    __shared__ double sred[2][NWARP];
    __shared__ long long lred[3][NWARP];
    const double r0 = warp_sum_d(pe), r1 = warp_sum_d(d0);
    const long long r2 = warp_sum_ll(ndeath), r3 = warp_sum_ll(npart), r4 = warp_sum_ll(nattempts);
    if (lane == 0) { sred[0][warp] = r0; sred[1][warp] = r1; lred[0][warp] = r2; lred[1][warp] = r3; lred[2][warp] = r4; }
    __syncthreads();
    if (tid == 0) {
        SpawnPartials out;
        out.pe = 0.0; out.d0 = 0.0; out.ndeath = 0; out.npart = 0; out.nattempts = 0;
        for (int w = 0; w < NWARP; ++w) {
            out.pe += sred[0][w]; out.d0 += sred[1][w];
            out.ndeath += lred[0][w]; out.npart += lred[1][w]; out.nattempts += lred[2][w];
        }
        partials[blockIdx.x] = out;
    }
}

// The attempts of the deferred (heavy) determinants: chunks of 64 attempts of one state per warp, over the whole grid.
template <int W, class Mask>
__global__ void __launch_bounds__(NWARP * 32, 4)
k_spawn_hb_heavy(Sys s, Params p, const uint64_t* __restrict__ states, int64_t* __restrict__ spawn, unsigned long long* __restrict__ head, long long block_size,
                 const int* __restrict__ proc_map, int* __restrict__ err, HeavyQueue hq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nitems = min(*hq.count, hq.cap);
    if (nitems == 0) return;
    const WarpSmem L(W, s.nel, p.qn != 0);
    double* siw = reinterpret_cast<double*>(smem_raw);
    const int siw_bytes = (s.nbasis * 8 + 15) & ~15;
    for (int k = tid; k < s.nbasis; k += NWARP * 32) siw[k] = s.hb_i_w[k];
    __syncthreads();
    WarpCtx<W> c(smem_raw + siw_bytes + warp * L.total, L, siw);
    const long long gwarp = (long long)blockIdx.x * NWARP + warp, nwarps = (long long)gridDim.x * NWARP;
    for (unsigned it = 0; it < nitems; ++it) {
        const HeavyItem item = hq.items[it];
        const long long nchunk = ((long long)item.natt + SLOTS - 1) / SLOTS;
        __syncthreads();
        if (lane == 0) {
            uint64_t f[W];
#pragma unroll
            for (int k = 0; k < W; ++k) f[k] = states[item.state * W + k];
            double pe = 0.0, d0 = 0.0;
            long long nd = 0, np = 0;
            // sign and initiator flag come from the population BEFORE death, which the spawn kernel saved in the item
            state_prologue<W>(s, p, c, 0, f, item.pop, 0.0, false, nullptr, 0, pe, d0, nd, np);
        }
        __syncwarp();
        const long long niter = (nchunk + nwarps - 1) / nwarps;       // the same for every warp (phase boundaries)
        for (long long k2 = 0; k2 < niter; ++k2) {
            const long long ch = gwarp + k2 * nwarps;
            const int n = (ch < nchunk) ? (int)min((long long)SLOTS, (long long)item.natt - ch * SLOTS) : 0;
            if (lane == 0) { c.sscan[0] = 0; c.sscan[1] = n; }
            __syncwarp();
            warp_spawn_rounds<W, Mask>(s, p, c, 1, n, 1, (uint32_t)(ch * SLOTS), spawn, head, block_size, proc_map, err);
        }
    }
}

}  // namespace hbw
