// hande_b200: semi-stochastic projection on the device (src/semi_stoch.F90, separate annihilation - the reference's
// default projection mode - and deterministic_annihilation of src/annihilation.f90:488-535).
//
// Data (SemiStoch in hb_common.cuh): every deterministic determinant of every rank (determ%dets, rank by rank, each rank's
// part sorted like its main list), a globally sorted copy for check_if_determ, this rank's slice of the deterministic
// Hamiltonian stored BY COLUMN (one column per deterministic state of this rank, rows = every deterministic state, in
// row order - the order csrpgemv(transpose) accumulates in, lib/local/csr.f90:188-194), the position of each of this
// rank's deterministic states in the main list (determ%indices) and one bit per main-list state (determ%flags).
#pragma once
#include "hb_common.cuh"

// determ%indices: where this rank's deterministic states sit in the (sorted) main list.  ntot2 != nullptr: the list
// length is ntot2[0] + ntot2[1] (new + surviving states of the merge that has just run, still on the device).
template <int W>
__global__ void __launch_bounds__(256)
k_ss_locate(const uint64_t* __restrict__ states, long long nstates, const int* __restrict__ ntot2,
            const uint64_t* __restrict__ local, int nloc, long long* __restrict__ idx, uint32_t* __restrict__ bits,
            int* __restrict__ miss) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nloc) return;
    if (ntot2) nstates = (long long)ntot2[0] + (long long)ntot2[1];
    uint64_t f[W];
    load_det<W>(local + (size_t)j * W, f);
    long long lo = 0, hi = nstates;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        uint64_t g[W];
        load_det<W>(states + mid * W, g);
        if (det_less<W>(g, f)) lo = mid + 1; else hi = mid;
    }
    bool hit = false;
    if (lo < nstates) {
        uint64_t g[W];
        load_det<W>(states + lo * W, g);
        hit = det_eq<W>(g, f);
    }
    idx[j] = lo;
    if (hit) atomicOr(&bits[lo >> 5], 1u << (lo & 31));
    else atomicAdd(miss, 1);
}

// <D_i|H|D_j> - H00 delta_ij of create_determ_hamil (src/semi_stoch.F90:533-685): get_hmatel(dets(:,i), dets_this_proc(:,j))
template <int W>
HB_HDN double ss_hmatel(const Sys& s, const Params& p, const uint64_t* f1, const uint64_t* f2) {
    int nx = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) nx += popc64(f1[k] ^ f2[k]);
    if (nx > 4) return 0.0;
    occ_t occ[HB_MAXNEL];
    decode_det<W>(f1, occ);
    if (nx == 0) return ((s.kind == SYS_UEG) ? slater_condon0_ueg(s, occ) : slater_condon0(s, occ)) - p.H00;
    bool same;
    return hmatel_pair<W>(s, f1, occ, f2, same);
}

// One warp per column (deterministic state j of this rank): pass 0 counts the elements above depsilon, pass 1 stores
// them in row order.  weight = the quasi-Newton weight of the spawnee j (transp = .true.), rho = pop_control - weight.
template <int W>
__global__ void __launch_bounds__(256)
k_ss_hamil(Sys s, Params p, const uint64_t* __restrict__ all, int tot, const int* __restrict__ pad,
           const uint64_t* __restrict__ local, int nloc, int displ, int pass, long long* __restrict__ colptr,
           int* __restrict__ row, double* __restrict__ val, double* __restrict__ rho) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= nloc) return;
    uint64_t f2[W];
    load_det<W>(local + (size_t)j * W, f2);
    double weight = 1.0;
    if (p.qn) {
        occ_t occ[HB_MAXNEL];
        decode_det<W>(f2, occ);
        weight = qn_weighting(p, qn_fock_sum(s, p, occ));
    }
    long long base = (pass == 1) ? colptr[j] : 0;
    long long cnt = 0;
    for (int i0 = 0; i0 < tot; i0 += 32) {
        const int i = i0 + lane;
        double h = 0.0;
        if (i < tot) {
            uint64_t f1[W];
            load_det<W>(all + (size_t)i * W, f1);
            h = ss_hmatel<W>(s, p, f1, f2);
            // the diagonal is the pair i == j + displs(iproc) only (the determinants of the space are distinct)
            h = weight * h;
        }
        const bool keep = fabs(h) > 1.e-12;      // depsilon (lib/local/const.F90:91)
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (pass == 1 && keep) {
            const long long o = base + __popc(m & ((1u << lane) - 1u));
            row[o] = pad[i];
            val[o] = h;
        }
        base += __popc(m);
        cnt += __popc(m);
    }
    if (lane == 0 && pass == 0) {
        colptr[j] = cnt;
        const double pc = p.qn ? p.qn_pop_control : 1.0;
        rho[j] = pc - weight;
    }
    (void)displ;
}

// determ_proj_separate_annihil (src/semi_stoch.F90:1009-1094) + deterministic_annihilation (src/annihilation.f90:488-535):
//   vector_j <- -tau (E_proj rho_j - S pop_control) v_j, then += (-tau H_ij) v_i for the rows i of column j in row order,
// then the result is stochastically rounded to the amplitude resolution and added to the state's population.
// One warp per column: 32 elements are loaded (coalesced) and multiplied at a time, then every lane adds the 32 products
// in row order (shuffles), so the sum is the reference's sequential sum bit for bit while the loads run in parallel.
template <int W>
__global__ void __launch_bounds__(256)
k_ss_project(Params p, const uint64_t* __restrict__ local, int nloc, const double* __restrict__ full, int self0,
             const long long* __restrict__ colptr, const int* __restrict__ row, const double* __restrict__ val,
             const double* __restrict__ rho, double* __restrict__ vec, const long long* __restrict__ idx,
             int64_t* __restrict__ pops) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= nloc) return;
    const double pc = p.qn ? p.qn_pop_control : 1.0;
    double y = (-p.tau * (p.proj_energy_old * rho[j] - p.shift * pc)) * full[self0 + j];
    const double a = -1.0 * p.tau;
    const long long z1 = colptr[j + 1];
    long long z0 = colptr[j];
    // 128 rows per step: the four loads of a lane are in flight together, and the products of the next step are fetched
    // while the current 128 are added (the additions are one dependent chain, the loads are not)
    constexpr int C = 4;
    double tn[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const long long z = z0 + 32 * c + lane;
        tn[c] = (z < z1) ? a * __ldcs(val + z) * full[__ldcs(row + z)] : 0.0;
    }
    for (; z0 < z1; z0 += 32 * C) {
        double t[C];
#pragma unroll
        for (int c = 0; c < C; ++c) t[c] = tn[c];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const long long z = z0 + 32 * (C + c) + lane;
            tn[c] = (z < z1) ? a * __ldcs(val + z) * full[__ldcs(row + z)] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const long long left = z1 - (z0 + 32 * c);
            if (left >= 32) {
                double u[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) u[k] = __shfl_sync(0xffffffffu, t[c], k);
#pragma unroll
                for (int k = 0; k < 32; ++k) y = y + u[k];
            } else if (left > 0) {
                for (int k = 0; k < (int)left; ++k) y = y + __shfl_sync(0xffffffffu, t[c], k);
            }
        }
    }
    if (lane != 0) return;
    vec[j] = y;
    double scaled = y * (double)p.real_factor;
    const int64_t sign = (scaled < 0.0) ? -1 : 1;
    scaled = fabs(scaled);
    int64_t nspawn = (int64_t)scaled;
    scaled = scaled - (double)nspawn;
    uint64_t f[W];
    load_det<W>(local + (size_t)j * W, f);
    PhiloxStream rng;
    rng.begin(p.seed, p.cycle, RNG_DETERM, det_hash64<W>(f, HB_NW(p)), 0);
    if (scaled > rng.next()) nspawn++;
    const long long k = idx[j];
    pops[k] = pops[k] + sign * nspawn;
}

template <int W>
static int ss_locate(hb200_engine* e, int buf, const int* ntot2) {
    SemiStoch& S = e->ss;
    const size_t nwords = ((size_t)e->cfg.walker_length + 31) / 32 + 1;
    CK(cudaMemsetAsync(S.d_bits, 0, nwords * sizeof(uint32_t), e->stream));
    CK(cudaMemsetAsync(S.d_miss, 0, sizeof(int), e->stream));
    if (S.nloc > 0) {
        k_ss_locate<W><<<(S.nloc + 255) / 256, 256, 0, e->stream>>>(e->d_states[buf], e->nstates, ntot2, S.d_local, S.nloc,
                                                                    S.d_idx, S.d_bits, S.d_miss);
        CK(cudaGetLastError());
    }
    return 0;
}
template <int W>
static int ss_hamil(hb200_engine* e, int pass) {
    SemiStoch& S = e->ss;
    if (S.nloc == 0) return 0;
    const unsigned nb = (unsigned)(((long long)S.nloc * 32 + 255) / 256);
    k_ss_hamil<W><<<nb, 256, 0, e->stream>>>(e->sys, e->par, S.d_all, S.tot, S.d_pad, S.d_local, S.nloc, S.displ, pass, S.d_colptr,
                                             S.d_row, S.d_val, S.d_rho);
    CK(cudaGetLastError());
    return 0;
}
template <int W>
static int ss_project(hb200_engine* e, const Params& p) {
    SemiStoch& S = e->ss;
    if (S.nloc == 0) return 0;
    k_ss_project<W><<<(unsigned)(((long long)S.nloc * 32 + 255) / 256), 256, 0, e->stream>>>(p, S.d_local, S.nloc, S.d_full, p.iproc * S.maxsz, S.d_colptr,
                                                                 S.d_row, S.d_val, S.d_rho, S.d_vec, S.d_idx, e->d_pops[e->cur]);
    CK(cudaGetLastError());
    return 0;
}
