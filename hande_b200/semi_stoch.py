"""Host side of the semi-stochastic projection (semi_stoch = {...}; src/semi_stoch.F90).

The engine keeps the deterministic space on the device (hb200_set_determ_space: membership table, the rank's slice of the
deterministic Hamiltonian, the projection and its annihilation every cycle); choosing the space is host logic in the
reference too and is mirrored here:

  * create_high_pop_space (src/semi_stoch.F90:1198-1316): every rank offers its min(target, nstates) most populated
    determinants (find_most_populated_dets, :1318-1385), the target_size most populated offers are kept
    (find_indices_of_most_populated_dets, :1387-1448).  Both walk their input once, replacing the current minimum
    (first slot among equal minima) by any later entry that is strictly larger - restated with a heap keyed (|pop|, slot),
    which evicts exactly that entry.
  * init_semi_stoch_t's bookkeeping (:134-377): sizes all-gathered, each rank's determinants sorted in list order, the
    ranks' lists concatenated.
"""
import heapq

import numpy as np

from . import synthetic


def _allgatherv(comm, a):
    """MPI_Allgatherv of one numpy array per rank over the driver's fixed-record allgather_bytes: [size] arrays"""
    a = np.ascontiguousarray(a)
    raw = np.frombuffer(a.tobytes(), dtype=np.uint8)
    lens = comm.allgather_bytes(np.frombuffer(np.int64(len(raw)).tobytes(), dtype=np.uint8))
    lens = [int(np.frombuffer(np.ascontiguousarray(x).tobytes(), dtype=np.int64)[0]) for x in lens]
    m = max(max(lens), 1)
    buf = np.zeros(m, dtype=np.uint8)
    buf[:len(raw)] = raw
    rows = comm.allgather_bytes(buf)
    return [np.frombuffer(np.ascontiguousarray(rows[r][:lens[r]]).tobytes(), dtype=a.dtype) for r in range(len(lens))]


def _most_populated(abs_pops, nout, chunk=1 << 20):
    """Slots -> input index after one pass of the reference's replace-the-minimum selection (nout <= len(abs_pops))."""
    n = len(abs_pops)
    slots = list(range(nout))
    if nout == 0:
        return np.zeros(0, dtype=np.int64)
    heap = [(int(abs_pops[i]), i) for i in range(nout)]      # (|pop|, slot)
    heapq.heapify(heap)
    for c0 in range(nout, n, chunk):
        a = abs_pops[c0:c0 + chunk]
        cand = np.nonzero(a > heap[0][0])[0]                  # the minimum only grows: a superset of the real candidates
        for k in cand:
            v = int(a[k])
            if v > heap[0][0]:
                _, slot = heap[0]
                heapq.heapreplace(heap, (v, slot))
                slots[slot] = c0 + int(k)
    return np.asarray(slots, dtype=np.int64)


def find_most_populated_dets(states, pops, ndets_out):
    """find_most_populated_dets (src/semi_stoch.F90:1318-1385): (dets_out, pops_out) in slot order"""
    ap = np.abs(np.asarray(pops, dtype=np.int64))
    idx = _most_populated(ap, int(ndets_out))
    return np.asarray(states)[idx], ap[idx]


def find_indices_of_most_populated_dets(pops, nind_out):
    """find_indices_of_most_populated_dets (src/semi_stoch.F90:1387-1448): 0-based indices, -1 for unused slots"""
    ap = np.abs(np.asarray(pops, dtype=np.int64))
    n = len(ap)
    out = np.full(int(nind_out), -1, dtype=np.int64)
    k = min(int(nind_out), n)
    out[:k] = _most_populated(ap, k)
    return out


def create_high_pop_space(comm, states, pops, target_size):
    """create_high_pop_space (src/semi_stoch.F90:1198-1316) for this rank: its deterministic determinants, unsorted.
    comm: the driver's communicator (fciqmc.SerialComm / TorchDist) - allgather_bytes is all that is used."""
    states = np.ascontiguousarray(states, dtype=np.uint64)
    W = states.shape[1] if states.ndim == 2 else 1
    nstates = len(states)
    ndets = min(int(target_size), nstates)
    determ_dets, determ_pops = find_most_populated_dets(states, pops, ndets)
    all_pops = _allgatherv(comm, np.ascontiguousarray(determ_pops, dtype=np.int64))
    all_ndets = [len(a) for a in all_pops]
    displs = np.concatenate([[0], np.cumsum(all_ndets)])
    ndets_tot = int(displs[-1])
    determ_size = min(int(target_size), ndets_tot)
    indices = find_indices_of_most_populated_dets(np.concatenate(all_pops) if ndets_tot else np.zeros(0, np.int64), determ_size)
    me = comm.rank
    mine = [int(i - displs[me]) for i in indices if displs[me] <= i < displs[me + 1]]
    return determ_dets[mine].reshape(-1, W)


def create_ci_determ_space(sys, occ0, ex_level, owner=None):
    """create_ci_determ_space (src/semi_stoch.F90:1763-1823): every determinant within ex_level excitations of the
    reference that has its spin polarisation and its symmetry (enumerate_determinants with ref_sym = sys%symmetry) - the
    point-group product of the orbitals for read_in systems, the total momentum for the UEG.  owner(f) -> bool keeps
    the determinants of this rank (add_det_to_determ_space with check_proc); None keeps all."""
    from itertools import combinations
    occ0 = [int(o) for o in occ0]
    nb = int(sys.nbasis)
    virt = [o for o in range(1, nb + 1) if o not in occ0]
    ueg = hasattr(sys, "kvec")

    def symmetry(occ):
        if ueg:
            return tuple(int(x) for x in np.asarray(sys.kvec)[list(occ)].sum(axis=0))
        return int(sys.symmetry_orb_list(list(occ)))
    ref_sym = symmetry(occ0)
    out = []
    for level in range(0, min(int(ex_level), len(occ0), len(virt)) + 1):
        for holes in combinations(occ0, level):
            nalpha_out = sum(o % 2 for o in holes)
            kept = [o for o in occ0 if o not in holes]
            for parts in combinations(virt, level):
                if sum(o % 2 for o in parts) != nalpha_out:        # odd orbitals are alpha: ms conserved
                    continue
                occ = sorted(kept + list(parts))
                if symmetry(occ) != ref_sym:
                    continue
                f = np.asarray(sys.encode(occ), dtype=np.uint64).reshape(-1)
                if owner is None or owner(f):
                    out.append(f)
    W = (nb + 63) // 64
    return np.asarray(out, dtype=np.uint64).reshape(-1, W)


def gather_determ_space(comm, dets_this_proc):
    """init_semi_stoch_t (src/semi_stoch.F90:236-340): sort this rank's determinants in list order, all-gather sizes and
    determinants.  Returns (determ%dets [tot x W], determ%sizes)."""
    d = np.ascontiguousarray(dets_this_proc, dtype=np.uint64)
    W = d.shape[1]
    d = synthetic.sort_dets(d) if len(d) else d
    parts = [x.reshape(-1, W) for x in _allgatherv(comm, d)]
    sizes = np.asarray([len(x) for x in parts], dtype=np.int32)
    return (np.concatenate(parts) if sizes.sum() else np.zeros((0, W), dtype=np.uint64)), sizes


def read_determ_from_file(path, owner=None):
    """read_determ_from_file (src/semi_stoch.F90:1450-1645) for this rank: the determinants of a stored space that the
    rank owns (add_det_to_determ_space with check_proc).  The reference keeps the space in SEMI.STOCH.<id>.H5
    (dataset `dets`); without HDF5 here the same array is a .npy file (uint64, one row per determinant)."""
    dets = np.ascontiguousarray(np.load(path), dtype=np.uint64)
    if dets.ndim == 1:
        dets = dets.reshape(-1, 1)
    if owner is None:
        return dets
    keep = [bool(owner(f)) for f in dets]
    return dets[keep].reshape(-1, dets.shape[1])


def write_determ_to_file(path, dets):
    """write_determ_to_file (src/semi_stoch.F90:1647-1723): determ%dets, all ranks' determinants, written by the parent"""
    np.save(path, np.ascontiguousarray(dets, dtype=np.uint64))


def init_semi_stoch(eng, comm, target_size, space="high", sys=None, occ0=None, ci_ex_level=-1, owner=None, path=None):
    """semi_stoch = { space = "high", size = target_size }, { space = "ci", ci_space = { ex_level = ... } } or
    { space = "read" } at the current iteration: choose the space (from the engine's list, or by enumeration) and hand it to
    hb200_set_determ_space.  Returns (dets, sizes)."""
    if space == "ci":
        mine = create_ci_determ_space(sys, occ0, ci_ex_level, owner)
    elif space == "read":
        mine = read_determ_from_file(path, owner)
    else:
        f, p, _ = eng.download_psips()
        mine = create_high_pop_space(comm, f, p, target_size)
    dets, sizes = gather_determ_space(comm, mine)
    eng.set_determ_space(dets, sizes)
    return dets, sizes
