"""Host side of the semi-stochastic projection on the CPU: the choice of the deterministic space
(hande_b200.semi_stoch, mirror of create_high_pop_space) against the oracle's, on one and on several ranks, and the
driver's start-iteration logic (do_fciqmc with the oracle-backed test engine) against the oracle's own run."""
import threading

import numpy as np

from hande_b200 import read_in as R
from hande_b200 import semi_stoch as SS
from hande_b200.fciqmc import QmcIn, do_fciqmc, _SingleProcess
from oracle.pyoracle import HUGE, Oracle


class ThreadComm:
    """allgather_bytes between `size` threads of this process (the collective create_high_pop_space uses)"""

    def __init__(self, rank, size, shared):
        self.rank, self.size, self.sh = rank, size, shared

    def allgather_bytes(self, arr):
        self.sh["slot"][self.rank] = np.asarray(arr, dtype=np.uint8).copy()
        self.sh["bar"].wait()
        out = np.stack(self.sh["slot"])
        self.sh["bar"].wait()
        return out


def _grown_oracle(fcidump_path, nprocs):
    o = Oracle()
    o.read_fcidump(fcidump_path("ne"), nel=10, ms=0, sym=0, cas=(8, 22))
    o.set_qmc(tau=0.002, seed=18, D0_population=10, ncycles=10, nreport=40, target_particles=50000, walker_length=50000,
              spawned_walker_length=5000, ex_level=5, nprocs=nprocs, real_amplitudes=1, spawn_cutoff=0.01)
    o.init()
    o.run()
    return o


def test_high_pop_space_matches_oracle_on_four_ranks(fcidump_path):
    """create_high_pop_space + the sort / all-gather of init_semi_stoch_t (src/semi_stoch.F90:134-377, 1198-1448): the
    host mirror picks, rank by rank, exactly the determinants the oracle picks (whose semi-stochastic runs reproduce the
    reference's tables), including the ties among equal populations"""
    nprocs, size = 4, 120
    o = _grown_oracle(fcidump_path, nprocs)
    lists = [o.get_psips(r) for r in range(nprocs)]
    assert sum(len(x[0]) for x in lists) > 3 * size
    o.set_semi_stoch(space="high", size=size)
    o.init_semi_stoch()
    dets_o, sizes_o = o.determ_space()
    shared = {"slot": [None] * nprocs, "bar": threading.Barrier(nprocs)}
    res = [None] * nprocs

    def work(r):
        comm = ThreadComm(r, nprocs, shared)
        mine = SS.create_high_pop_space(comm, lists[r][0], lists[r][1], size)
        res[r] = SS.gather_determ_space(comm, mine)

    th = [threading.Thread(target=work, args=(r,)) for r in range(nprocs)]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in range(nprocs):
        dets, sizes = res[r]
        assert (sizes == sizes_o).all() and (dets == dets_o).all()
    assert sizes_o.sum() == size
    # the oracle now holds the space: zero-population deterministic states stay in the lists over further cycles
    o.iterate(5, 1000, 0.002, -0.1, -0.1)
    for r in range(nprocs):
        v, flags = o.determ_vector(r)
        assert int((flags == 0).sum()) == sizes_o[r] and len(v) == sizes_o[r]


def test_selection_with_ties_and_short_lists():
    """find_most_populated_dets / find_indices_of_most_populated_dets (src/semi_stoch.F90:1318-1448) restated literally
    and compared with the heap form on inputs full of equal populations; a list shorter than the target"""
    def literal(pops, nout):
        pops = [abs(int(x)) for x in pops]
        out = list(range(nout))
        val = pops[:nout]
        mi = min(range(nout), key=lambda i: (val[i], i))
        for i in range(nout, len(pops)):
            if pops[i] > val[mi]:
                out[mi], val[mi] = i, pops[i]
                mi = min(range(nout), key=lambda k: (val[k], k))
        return out
    rng = np.random.default_rng(5)
    for trial in range(20):
        n, k = int(rng.integers(5, 300)), int(rng.integers(1, 40))
        pops = rng.integers(-6, 7, n)
        k = min(k, n)
        assert list(SS._most_populated(np.abs(pops), k, chunk=17)) == literal(pops, k)
    idx = SS.find_indices_of_most_populated_dets(np.array([3, -9, 4]), 5)
    assert list(idx) == [0, 1, 2, -1, -1]
    st = np.arange(3, dtype=np.uint64).reshape(-1, 1)
    d = SS.create_high_pop_space(_SingleProcess(), st, np.array([3, -9, 4]), 10)
    assert sorted(d[:, 0].tolist()) == [0, 1, 2]


def test_driver_starts_the_projection_at_the_requested_iteration(fcidump_path):
    """do_fciqmc with semi_stoch = { space = "high", size = 40, start_iteration = 33 } (mid report loop) against the
    oracle's own run with the same options: every row of the report table, reference dSFMT stream, one rank"""
    from tests.oracle_engine import make_engine_cls
    path = fcidump_path("he2_avdz")
    kw = dict(nel=4, ms=0, sym=HUGE, cas=(-1, -1))
    o = Oracle()
    o.read_fcidump(path, **kw)
    q = dict(tau=0.01, seed=7, D0_population=200, ncycles=10, nreport=12, target_particles=400, real_amplitudes=1,
             spawn_cutoff=0.01, walker_length=4000, spawned_walker_length=2000)
    o.set_qmc(**q)
    o.set_semi_stoch(space="high", size=12, start_iteration=33)
    o.init()
    rows_o = o.run()
    dets_o, sizes_o = o.determ_space()
    assert sizes_o.sum() == 12
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.01, rng_seed=7, init_pop=200, mc_cycles=10, nreports=12, target_population=400, real_amplitudes=True,
                spawn_cutoff=0.01, state_size=4000, spawned_state_size=2000, semi_stoch_space="high", semi_stoch_size=12,
                semi_stoch_start_iteration=33)
    res = do_fciqmc(s, qmc, engine_cls=make_engine_cls(path, kw, rng_kind=0))
    dets, sizes = res.determ_space
    assert (dets == dets_o).all() and (sizes == sizes_o).all()
    rows = np.array(res.rows)
    assert len(rows) == len(rows_o) == 13
    for a, b in zip(rows, rows_o):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6]
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-11 * max(1.0, abs(b[k])), (a, b)


def test_ci_space_matches_oracle_and_reference_printout(fcidump_path):
    """create_ci_determ_space (src/semi_stoch.F90:1763-1823): the host enumeration gives, rank by rank, the space the
    oracle builds for the reference's ueg_n7_rs1_e1_SS_cisdtq run (358 determinants, 83..98 per rank as the reference
    prints), and the oracle's space for a molecular point group (He2, doubles)"""
    from hande_b200.fciqmc import owner_of
    from hande_b200.ueg import UegSystem
    from tests.conftest import load_golden
    g = load_golden("ueg_ss_np4")
    o = Oracle()
    o.init_ueg(**g["ueg"])
    o.set_ref_det(g["ref_det"])
    o.set_qmc(**dict(g["qmc"], nreport=1))
    o.set_semi_stoch(**g["semi_stoch"])
    o.init()
    o.run()                                   # start_iteration = 1: the space is built in the first cycle
    dets_o, sizes_o = o.determ_space()
    s = UegSystem(g["ueg"]["nel"], g["ueg"]["ms"], g["ueg"]["rs"], g["ueg"]["cutoff"])
    parts = []
    for r in range(4):
        mine = SS.create_ci_determ_space(s, g["ref_det"], 4, owner=lambda f: owner_of(f, s.nbasis, 4, 1) == r)
        parts.append(SS.gather_determ_space(_SingleProcess(), mine)[0])
    assert [len(p) for p in parts] == list(sizes_o) and sum(len(p) for p in parts) == g["kat"]["determ_size"]
    assert min(len(p) for p in parts) == g["kat"]["determ_min"] and max(len(p) for p in parts) == g["kat"]["determ_max"]
    assert (np.concatenate(parts) == dets_o).all()
    # molecular: D2h He2, all symmetry-allowed singles and doubles of the reference
    path = fcidump_path("he2_avdz")
    kw = dict(nel=4, ms=0, sym=HUGE, cas=(-1, -1))
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=0.01, seed=7, D0_population=100, ncycles=2, nreport=1, target_particles=1e6, real_amplitudes=1)
    o.set_semi_stoch(space="ci", ci_ex_level=2)
    o.init()
    o.run()
    dets_o, sizes_o = o.determ_space()
    sm = R.read_in(path, **kw)
    mine = SS.create_ci_determ_space(sm, o.reference()["occ"], 2)
    dets, sizes = SS.gather_determ_space(_SingleProcess(), mine)
    assert len(dets) == len(dets_o) > 20 and (dets == dets_o).all()


def test_driver_with_a_ci_space_on_the_polarised_ueg():
    """do_fciqmc with semi_stoch = { space = "ci", ci_space = { ex_level = 4 } } on the reference's fully polarised
    7-electron UEG (one rank, reference dSFMT stream) against the oracle's own run: every row"""
    from hande_b200.ueg import UegSystem
    from tests.conftest import load_golden
    from tests.oracle_engine import make_engine_cls
    g = load_golden("ueg_ss_np4")
    u = g["ueg"]
    o = Oracle()
    o.init_ueg(**u)
    o.set_ref_det(g["ref_det"])
    q = dict(g["qmc"], nreport=15, nprocs=1)
    o.set_qmc(**q)
    o.set_semi_stoch(space="ci", ci_ex_level=4, start_iteration=21)
    o.init()
    rows_o = o.run()
    s = UegSystem(u["nel"], u["ms"], u["rs"], u["cutoff"])
    qmc = QmcIn(tau=q["tau"], rng_seed=q["seed"], init_pop=q["D0_population"], mc_cycles=q["ncycles"], nreports=15,
                target_population=q["target_particles"], real_amplitudes=True, spawn_cutoff=0.01, excit_gen="no_renorm",
                state_size=q["walker_length"], spawned_state_size=q["spawned_walker_length"], semi_stoch_space="ci",
                semi_stoch_ci_ex_level=4, semi_stoch_start_iteration=21, reference_det=g["ref_det"])
    res = do_fciqmc(s, qmc, engine_cls=make_engine_cls(None, None, rng_kind=0, ueg=(u["nel"], u["ms"], u["rs"], u["cutoff"]),
                                                       ref_det=g["ref_det"]))
    assert int(res.determ_space[1].sum()) == 358
    rows = np.array(res.rows)
    assert len(rows) == len(rows_o) == 16
    for a, b in zip(rows, rows_o):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-11 * max(1.0, abs(b[k])), (a, b)


def test_driver_writes_and_reads_a_determ_space(fcidump_path, tmp_path):
    """write_determ_to_file / read_determ_from_file (src/semi_stoch.F90:1450-1723) through do_fciqmc: a run stores the
    space it picked; a second run that reads it from the first iteration on propagates exactly like the oracle handed
    the same determinants (init_semi_stoch_t with read_determ_space)"""
    from tests.oracle_engine import make_engine_cls
    path = fcidump_path("he2_avdz")
    kw = dict(nel=4, ms=0, sym=HUGE, cas=(-1, -1))
    s = R.read_in(path, **kw)
    store = str(tmp_path / "SEMI.STOCH.0.npy")
    base = dict(tau=0.01, rng_seed=7, init_pop=200, mc_cycles=10, nreports=6, target_population=400, real_amplitudes=True,
                spawn_cutoff=0.01, state_size=4000, spawned_state_size=2000)
    res = do_fciqmc(s, QmcIn(semi_stoch_space="high", semi_stoch_size=15, semi_stoch_start_iteration=31,
                             semi_stoch_write_file=store, **base), engine_cls=make_engine_cls(path, kw, rng_kind=0))
    dets = res.determ_space[0]
    assert len(dets) == 15 and (np.load(store) == dets).all()
    res2 = do_fciqmc(s, QmcIn(semi_stoch_space="read", semi_stoch_read_file=store, **base),
                     engine_cls=make_engine_cls(path, kw, rng_kind=0))
    assert (res2.determ_space[0] == dets).all()
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=0.01, seed=7, D0_population=200, ncycles=10, nreport=6, target_particles=400, real_amplitudes=1,
              spawn_cutoff=0.01, walker_length=4000, spawned_walker_length=2000)
    o.init()
    o.init_semi_stoch(dets, [15])        # before the first cycle, as start_iteration = 1 does (the oracle's iteration-0
    rows_o = o.run()                     # row therefore already counts the 15 states; the driver's counts the reference)
    assert len(res2.rows) == len(rows_o) == 7
    for a, b in zip(np.array(res2.rows), rows_o):
        assert a[0] == b[0] and (a[5] == b[5] or a[0] == 0) and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-11 * max(1.0, abs(b[k])), (a, b)


def test_driver_starts_the_projection_relative_to_the_shift(fcidump_path):
    """semi_stoch = { shift_start_iteration = 12 }: start_iteration becomes huge(0) until the report loop in which the
    shift starts to vary fixes it to that iteration + 12 + 1 (src/lua_hande_calc.f90:1712-1715,
    src/qmc_common.F90:1200-1204); driver against the oracle's own run, every row"""
    from tests.oracle_engine import make_engine_cls
    path = fcidump_path("he2_avdz")
    kw = dict(nel=4, ms=0, sym=HUGE, cas=(-1, -1))
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=0.01, seed=7, D0_population=200, ncycles=10, nreport=14, target_particles=260, real_amplitudes=1,
              spawn_cutoff=0.01, walker_length=4000, spawned_walker_length=2000)
    o.set_semi_stoch(space="high", size=12, shift_start_iteration=12)
    o.init()
    rows_o = o.run()
    started = next(i for i, r in enumerate(rows_o) if r[1] != 0.0)        # first row with a varying shift
    assert 1 < started < 10 and o.determ_space()[1].sum() == 12
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.01, rng_seed=7, init_pop=200, mc_cycles=10, nreports=14, target_population=260, real_amplitudes=True,
                spawn_cutoff=0.01, state_size=4000, spawned_state_size=2000, semi_stoch_space="high", semi_stoch_size=12,
                semi_stoch_shift_start_iteration=12)
    res = do_fciqmc(s, qmc, engine_cls=make_engine_cls(path, kw, rng_kind=0))
    assert (res.determ_space[0] == o.determ_space()[0]).all()
    for a, b in zip(np.array(res.rows), rows_o):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-11 * max(1.0, abs(b[k])), (a, b)
