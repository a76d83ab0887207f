"""Synthetic inputs for benchmarks and tests.

* `synthetic_fcidump(norb, nelec)`: an 8-fold-symmetric, C1-symmetry RHF FCIDUMP in the text format the
  reference reads (`read (ir,*) x, i, a, j, b`, src/read_in.F90:535), so one file feeds the oracle, this engine
  and (elsewhere) a real HANDE binary.  S50 = (50, 20), S40 = (40, 16).
* `random_walkers(...)`: distinct random determinants with unit (distribution A) or 1+Exp(1) (distribution B)
  populations, sorted in the reference's list order.
"""
from __future__ import annotations

import io

import numpy as np

SEED = 20261017


def synthetic_integrals(norb, seed=SEED):
    """Return (eps[norb], h[norb,norb], eri dict arrays) for the canonical 8-fold unique (pq|rs), chemists' notation."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    p = np.arange(1, norb + 1)
    eps = -2.0 + 0.1 * p
    g1 = rng.standard_normal((norb, norb))
    g1 = np.tril(g1) + np.tril(g1, -1).T
    h = 0.02 * g1 * np.exp(-np.abs(p[:, None] - p[None, :]) / 5.0)
    h[np.diag_indices(norb)] = 0.0
    h = h + np.diag(eps)
    # canonical pairs pq (p>=q), ordered by tri index
    pp, qq = np.tril_indices(norb)
    pp, qq = pp + 1, qq + 1
    npair = len(pp)
    a, b = np.tril_indices(npair)          # pair index pq >= rs
    P, Q, R, S = pp[a], qq[a], pp[b], qq[b]
    g = rng.standard_normal(len(a))
    val = 0.5 * np.exp(-(np.abs(P - Q) + np.abs(R - S)) / 4.0) * np.exp(-np.abs((P + Q) / 2.0 - (R + S) / 2.0) / 10.0) \
        * (1.0 + 0.1 * g)
    return eps, h, (P, Q, R, S, val)


def synthetic_fcidump(norb, nelec, ms2=0, seed=SEED, path=None):
    eps, h, (P, Q, R, S, val) = synthetic_integrals(norb, seed)
    out = io.StringIO()
    out.write(f" &FCI NORB={norb},NELEC={nelec},MS2={ms2},\n  ORBSYM=" + ",".join(["1"] * norb) + ",\n  ISYM=1 UHF=.FALSE.\n &END\n")
    body = np.column_stack([val, P, Q, R, S])
    np.savetxt(out, body, fmt="%23.16e %3d %3d %3d %3d")
    pp, qq = np.tril_indices(norb)
    np.savetxt(out, np.column_stack([h[pp, qq], pp + 1, qq + 1, 0 * pp, 0 * pp]), fmt="%23.16e %3d %3d %3d %3d")
    np.savetxt(out, np.column_stack([eps, np.arange(1, norb + 1), 0 * eps, 0 * eps, 0 * eps]),
               fmt="%23.16e %3d %3d %3d %3d")
    out.write("%23.16e %3d %3d %3d %3d\n" % (0.0, 0, 0, 0, 0))
    text = out.getvalue()
    if path is not None:
        with open(path, "w") as f:
            f.write(text)
    return text


def synthetic_fcidump_uhf(norb, nelec, ms2=0, seed=SEED, path=None):
    """UHF variant (UHF=.TRUE.: NORB = 2*norb spin-orbitals, odd = alpha, even = beta; src/read_in.F90:300-420): the
    spatial integrals of `synthetic_integrals` scaled differently per spin channel - (aa|aa) x 1, (bb|bb) x 0.9,
    (aa|bb) = (bb|aa) x 0.95; h_beta = 0.97 h_alpha off the diagonal, eps_beta = eps_alpha + 0.013 - so that the four
    two-body channels and the spin-dependent one-body terms of the UHF code paths all differ."""
    eps, h, (P, Q, R, S, val) = synthetic_integrals(norb, seed)
    out = io.StringIO()
    out.write(f" &FCI NORB={2 * norb},NELEC={nelec},MS2={ms2},\n  ORBSYM=" + ",".join(["1"] * (2 * norb)) +
              ",\n  ISYM=1 UHF=.TRUE.\n &END\n")
    so = lambda p, spin: 2 * p - 1 + spin          # spatial p (1-based), spin 0 = alpha / 1 = beta
    for s1, s2, c in ((0, 0, 1.0), (1, 1, 0.9), (0, 1, 0.95), (1, 0, 0.95)):
        np.savetxt(out, np.column_stack([c * val, so(P, s1), so(Q, s1), so(R, s2), so(S, s2)]), fmt="%23.16e %3d %3d %3d %3d")
    pp, qq = np.tril_indices(norb)
    for spin, c, de in ((0, 1.0, 0.0), (1, 0.97, 0.013)):
        hv = np.where(pp == qq, h[pp, qq] + de, c * h[pp, qq])
        np.savetxt(out, np.column_stack([hv, so(pp + 1, spin), so(qq + 1, spin), 0 * pp, 0 * pp]), fmt="%23.16e %3d %3d %3d %3d")
    for spin, de in ((0, 0.0), (1, 0.013)):
        np.savetxt(out, np.column_stack([eps + de, so(np.arange(1, norb + 1), spin), 0 * eps, 0 * eps, 0 * eps]),
                   fmt="%23.16e %3d %3d %3d %3d")
    out.write("%23.16e %3d %3d %3d %3d\n" % (0.0, 0, 0, 0, 0))
    text = out.getvalue()
    if path is not None:
        with open(path, "w") as f:
            f.write(text)
    return text


def random_dets(n, nbasis, nalpha, nbeta, seed=1):
    """n distinct random determinants (alpha = odd orbitals, beta = even), sorted ascending in the reference order
    (unsigned compare, last word most significant).  Returns uint64 array (n, W)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    norb = nbasis // 2
    W = (nbasis + 63) // 64
    need = n
    chunks = []
    total = None
    while True:
        m = int(need * 1.1) + 16
        f = np.zeros((m, W), dtype=np.uint64)
        for nocc, off in ((nalpha, 0), (nbeta, 1)):
            # choose nocc of norb spatial orbitals per row: argsort of random keys
            keys = rng.random((m, norb))
            sel = np.argpartition(keys, nocc - 1, axis=1)[:, :nocc]
            orb0 = 2 * sel + off                      # 0-based spin-orbital = bit position
            for k in range(nocc):
                w = orb0[:, k] // 64
                bit = (orb0[:, k] % 64).astype(np.uint64)
                for iw in range(W):
                    msk = w == iw
                    f[msk, iw] |= (np.uint64(1) << bit[msk])
        chunks.append(f)
        total = np.unique(np.concatenate(chunks), axis=0)
        if len(total) >= n:
            break
        need = n - len(total)
    total = total[:n] if len(total) > n else total
    return sort_dets(total)


def sort_dets(f):
    """Sort rows in the reference's list order (bit_str_cmp, src/bit_utils.F90:452-479)."""
    W = f.shape[1]
    order = np.lexsort(tuple(f[:, k] for k in range(W)))   # last key (highest word) is primary
    return np.ascontiguousarray(f[order])


def random_walkers(n, nbasis, nalpha, nbeta, real_factor=1, dist="A", seed=1):
    """(states, pops) for walker distribution A (|pop| = 1) or B (|pop| = 1 + Exp(1)), sign Bernoulli(1/2)."""
    f = random_dets(n, nbasis, nalpha, nbeta, seed)
    rng = np.random.Generator(np.random.Philox(key=seed + 7919))
    sign = np.where(rng.random(len(f)) < 0.5, -1, 1).astype(np.int64)
    if dist == "A":
        mag = np.full(len(f), real_factor, dtype=np.int64)
    else:
        mag = np.floor((1.0 + rng.exponential(1.0, len(f))) * real_factor).astype(np.int64)
        if real_factor == 1:
            mag = np.maximum(mag, 1)
    return f, sign * mag


def murmur_owner_torch(w, nbasis, nprocs, nslots=1, seed=7):
    """Owner rank of determinants (rows of int64 tensor w[n, W]) by HANDE's rule, evaluated with torch ops:
    MurmurHash2 over ceil(nbasis/32) 32-bit words, seed 7, Fortran modulo (src/spawning.F90:770-838)."""
    import torch
    M32 = 0xFFFFFFFF
    m = 0x5BD1E995
    nw = (nbasis + 31) // 32
    h = torch.full((w.shape[0],), (seed ^ (nw * 4)) & M32, dtype=torch.int64, device=w.device)
    for i in range(nw):
        k = (w[:, i // 2] >> (32 * (i % 2))) & M32
        k = (k * m) & M32
        k = k ^ (k >> 24)
        k = (k * m) & M32
        h = (h * m) & M32
        h = h ^ k
    h = h ^ (h >> 13)
    h = (h * m) & M32
    h = h ^ (h >> 15)
    h = torch.where(h >= 2**31, h - 2**32, h)
    return torch.remainder(h, nprocs * nslots) % nprocs


def random_walkers_torch(n, nbasis, nalpha, nbeta, real_factor, device, seed=1, nprocs=1, iproc=0, chunk=4_000_000):
    """n distinct random determinants owned by rank `iproc`, sorted in the reference's list order, |pop| = 1
    (distribution A of SURVEY.md 8d), generated on the GPU with torch (data plumbing only).
    Returns numpy (states[n, W] uint64, pops[n] int64)."""
    import torch
    norb = nbasis // 2
    W = (nbasis + 63) // 64
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + iproc)
    parts = []
    have = 0
    if norb > 128:
        chunk = min(chunk, max(1 << 16, (1 << 29) // norb))     # the (chunk, norb) random matrix stays below 2 GB
    while have < n:
        m = chunk
        w = torch.zeros((m, max(W, 2)), dtype=torch.int64, device=device)
        for nocc, off in ((nalpha, 0), (nbeta, 1)):
            sel = torch.rand((m, norb), device=device, generator=g).topk(nocc, dim=1).indices  # nocc of norb
            bit = 2 * sel + off
            w.scatter_add_(1, bit // 64, torch.ones_like(bit) << (bit % 64))    # distinct bits: add == or
        if nprocs > 1:
            w = w[murmur_owner_torch(w, nbasis, nprocs) == iproc]
        parts.append(w)
        have += w.shape[0]
    w = torch.cat(parts)[: int(n * 1.02) + 16]
    del parts
    # sort: unsigned compare, last word most significant (bit_str_cmp): two stable sorts, low word first
    flip = torch.tensor(-2**63, dtype=torch.int64, device=device)
    for wi in range(W):
        key = w[:, wi] ^ flip if wi == 0 or True else w[:, wi]
        idx = torch.sort(key, stable=True).indices
        w = w[idx]
    keep = torch.ones(w.shape[0], dtype=torch.bool, device=device)
    keep[1:] = (w[1:] != w[:-1]).any(dim=1)
    w = w[keep][:n]
    sign = torch.where(torch.rand(w.shape[0], device=device, generator=g) < 0.5, -1, 1).to(torch.int64)
    pops = sign * int(real_factor)
    states = w[:, :W].contiguous().cpu().numpy().view(np.uint64)
    return states, pops.cpu().numpy()


def random_excitors(n, occ0, nbasis, max_level, seed=1):
    """n distinct random excitors of the reference `occ0` (1-based spin-orbitals) with excitation level 1..max_level
    (spin-conserving replacements), sorted in the reference's list order; the reference itself is not included.
    Synthetic CCMC excip lists for throughput measurements (BASELINE config 5)."""
    rng = np.random.Generator(np.random.Philox(key=seed + 31))
    W = (nbasis + 63) // 64
    occ0 = np.asarray(sorted(occ0))
    occ_s = [occ0[occ0 % 2 == 1], occ0[occ0 % 2 == 0]]                       # alpha (odd), beta (even)
    allo = np.arange(1, nbasis + 1)
    virt = np.setdiff1d(allo, occ0)
    virt_s = [virt[virt % 2 == 1], virt[virt % 2 == 0]]
    f0 = np.zeros(W, dtype=np.uint64)
    for o in occ0:
        f0[(o - 1) // 64] |= np.uint64(1) << np.uint64((o - 1) % 64)
    out = None
    need = n
    while True:
        m = int(need * 1.3) + 64
        f = np.tile(f0, (m, 1))
        level = rng.integers(1, max_level + 1, size=m)
        for k in range(max_level):
            act = level > k
            spin = rng.integers(0, 2, size=m)
            for sp in (0, 1):
                sel = act & (spin == sp)
                cnt = int(sel.sum())
                if cnt == 0:
                    continue
                i = occ_s[sp][rng.integers(0, len(occ_s[sp]), size=cnt)]
                a = virt_s[sp][rng.integers(0, len(virt_s[sp]), size=cnt)]
                rows = np.nonzero(sel)[0]
                for orb, setbit in ((i, False), (a, True)):
                    w = (orb - 1) // 64
                    bit = np.uint64(1) << ((orb - 1) % 64).astype(np.uint64)
                    for iw in range(W):
                        msk = w == iw
                        if setbit:
                            f[rows[msk], iw] |= bit[msk]
                        else:
                            f[rows[msk], iw] &= ~bit[msk]
        # keep rows with the right electron count (an orbital hit twice changes it) and drop the reference
        nel = np.zeros(m, dtype=np.int64)
        for iw in range(W):
            nel += np.array([bin(int(x)).count("1") for x in f[:, iw]]) if m < 4096 else _popcount64(f[:, iw])
        keep = (nel == len(occ0)) & ~(f == f0).all(axis=1)
        f = f[keep]
        out = f if out is None else np.concatenate([out, f])
        out = sort_dets(out)
        out = out[np.concatenate([[True], (out[1:] != out[:-1]).any(axis=1)])]
        if len(out) >= n:
            break
        need = n - len(out)
    out = out[np.sort(rng.permutation(len(out))[:n])]
    return np.ascontiguousarray(out)


def _popcount64(x):
    x = x.astype(np.uint64)
    x = x - ((x >> np.uint64(1)) & np.uint64(0x5555555555555555))
    x = (x & np.uint64(0x3333333333333333)) + ((x >> np.uint64(2)) & np.uint64(0x3333333333333333))
    x = (x + (x >> np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    return ((x * np.uint64(0x0101010101010101)) >> np.uint64(56)).astype(np.int64)
