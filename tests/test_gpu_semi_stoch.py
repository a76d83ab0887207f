"""Semi-stochastic projection on the device against the oracle (-m gpu): hb200_set_determ_space + the projection inside
hb200_iterate, and two ranks on one device through the staged hb200_determ_vector / hb200_determ_project calls
(src/semi_stoch.F90, src/annihilation.f90:488-535).  The oracle's semi-stochastic runs reproduce the reference's own
tables (tests/test_oracle_golden.py::test_semi_stochastic_projection)."""
import numpy as np
import pytest

from hande_b200 import semi_stoch as SS
from hande_b200.fciqmc import _SingleProcess
from tests.common import make_pair
from tests.test_gpu_two_ranks_one_device import _exchange_and_annihilate, _setup

pytestmark = pytest.mark.gpu


def _columns(rp, ci, mat, ncol):
    """oracle CSR (rows = all deterministic states, columns = the rank's) -> per-column (rows, values) in row order"""
    cols = [([], []) for _ in range(ncol)]
    for i in range(len(rp) - 1):
        for z in range(rp[i], rp[i + 1]):
            cols[ci[z]][0].append(i)
            cols[ci[z]][1].append(mat[z])
    return cols


def _compare_hamil(o, eng, rank, nloc):
    rp, ci, mat = o.determ_hamil(rank)
    cp, row, val = eng.determ_hamil()
    assert cp[-1] == len(mat) == len(val)
    cols = _columns(rp, ci, mat, nloc)
    for j in range(nloc):
        assert list(row[cp[j]:cp[j + 1]]) == cols[j][0], j
        assert (val[cp[j]:cp[j + 1]] == np.array(cols[j][1])).all(), j       # bit-exact: same arithmetic, same order
    return len(val)


@pytest.mark.parametrize("name,gen,real,init,tau,size,qn", [("h2o", "renorm", True, False, 0.003, 40, None),
                                                            ("s12", "heat_bath", True, True, 0.002, 60, None),
                                                            ("ueg14", "no_renorm", True, True, 0.002, 50, None),
                                                            ("h2o", "renorm", True, True, 0.003, 12, dict()),
                                                            ("ueg358", "no_renorm", True, False, 0.001, 25, None)])
def test_semi_stochastic_iterate_matches_oracle(name, gen, real, init, tau, size, qn):
    """Grow a population from the reference, pick the `size` most populated determinants (create_high_pop_space on the
    host == the oracle's own choice), build the space on the device, then 30 more cycles through hb200_iterate with the
    projection on: Hamiltonian slice, lists (zero-population deterministic states kept), projection vector and
    estimators all equal the oracle's."""
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real, initiator=init, quasi_newton=qn)
    rf = 2**31
    f0 = ref["f0"].reshape(1, -1)
    o.set_psips(f0, [80 * rf], [0.0])
    eng.upload_psips(f0, [80 * rf], [0.0])
    o.iterate(30, 1, tau, 0.0, 0.0)
    eng.iterate(30, tau, 0.0, 0.0, 1)
    fg, pg, dg = eng.download_psips()
    assert len(fg) > size
    # the host's choice of the space == the oracle's (create_high_pop_space)
    o.set_semi_stoch(space="high", size=size)
    o.init_semi_stoch()
    dets_o, sizes_o = o.determ_space()
    dets, sizes = SS.gather_determ_space(_SingleProcess(), SS.create_high_pop_space(_SingleProcess(), fg, pg, size))
    assert (sizes == sizes_o).all() and (dets == dets_o).all() and len(dets) == size
    eng.set_determ_space(dets, sizes)
    nnz = _compare_hamil(o, eng, 0, size)
    assert nnz > size
    cyc = 31
    for block in range(3):
        ro = o.iterate(10, cyc, tau, -0.03, -0.02 * block)
        rg = eng.iterate(10, tau, -0.03, -0.02 * block, cyc)
        cyc += 10
        assert rg["spawn_error"] == 0 and rg["psip_error"] == 0 and ro["error"] == 0
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo)
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all()
        vo, flags = o.determ_vector(0)
        assert (eng.determ_vector(1) == vo).all()
        assert int((flags == 0).sum()) == size
        assert rg["nspawn_events"] == ro["nspawn_events"] and rg["ndeath"] == ro["ndeath"]
        for key in ("proj_energy", "D0_population", "rspawn", "nparticles"):
            assert abs(rg[key] - ro[key]) <= 1e-12 * max(1.0, abs(ro[key])), key
    # every deterministic state is still in the list, whatever its population
    keys = {tuple(x) for x in fg.tolist()}
    assert all(tuple(x) in keys for x in dets.tolist())
    # switching the projection off again: all sizes zero
    eng.set_determ_space(np.zeros((0, s.W), dtype=np.uint64), np.zeros(1, dtype=np.int32))
    eng.iterate(1, tau, -0.03, -0.05, cyc)
    eng.close()


def test_semi_stochastic_space_with_unoccupied_determinants():
    """add_determ_dets_to_psip_list (src/semi_stoch.F90:728-793): deterministic states that are not in the main list are
    added with zero population and their diagonal element; a spawn from a deterministic state onto one is cancelled."""
    name, gen, tau = "h2o", "renorm", 0.003
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=True)
    rf = 2**31
    f0 = ref["f0"].reshape(1, -1)
    o.set_psips(f0, [80 * rf], [0.0])
    eng.upload_psips(f0, [80 * rf], [0.0])
    o.iterate(25, 1, tau, 0.0, 0.0)
    eng.iterate(25, tau, 0.0, 0.0, 1)
    f_grown, p_grown, _ = eng.download_psips()
    # restart both from the reference alone, with the grown list's most populated determinants as the space
    dets, sizes = SS.gather_determ_space(_SingleProcess(), SS.create_high_pop_space(_SingleProcess(), f_grown, p_grown, 30))
    o.set_psips(f0, [80 * rf], [0.0])
    eng.upload_psips(f0, [80 * rf], [0.0])
    o.init_semi_stoch(dets, sizes)
    eng.set_determ_space(dets, sizes)
    fo, po, do_ = o.get_psips()
    fg, pg, dg = eng.download_psips()
    assert len(fg) == len(fo) == 30 and (fg == fo).all() and (pg == po).all() and (dg == do_).all()
    assert int((pg == 0).sum()) == 29
    ro = o.iterate(20, 1, tau, 0.0, -0.05)
    rg = eng.iterate(20, tau, 0.0, -0.05, 1)
    fo, po, do_ = o.get_psips()
    fg, pg, dg = eng.download_psips()
    assert len(fg) == len(fo) and (fg == fo).all() and (pg == po).all() and (dg == do_).all()
    assert rg["nspawn_events"] == ro["nspawn_events"] and rg["ndeath"] == ro["ndeath"]
    eng.close()


@pytest.mark.parametrize("name,gen,real,init,tau,world", [("h2o", "renorm", True, False, 0.003, 2),
                                                          ("s12", "heat_bath", True, True, 0.004, 3)])
def test_semi_stochastic_ranks_on_one_device(name, gen, real, init, tau, world):
    """The projection across ranks (determ_proj_separate_annihil's mpi_allgatherv, src/semi_stoch.F90:1056-1066): every
    rank's hb200_determ_vector is gathered by the host and handed to hb200_determ_project, as a Fortran host without NCCL
    would; each rank's list is compared with the oracle's emulated rank after every cycle."""
    s, o, engs = _setup(name, gen, real, init, tau, world, 4000)
    shift, pe_old = -0.05, -0.1
    o.set_semi_stoch(space="high", size=90)
    o.init_semi_stoch()
    dets, sizes = o.determ_space()
    assert sizes.sum() == 90 and (sizes > 0).all()
    for r, e in enumerate(engs):
        e.set_determ_space(dets, sizes)
        _compare_hamil(o, e, r, int(sizes[r]))
    for cycle in range(1, 6):
        ro = o.iterate(1, cycle, tau, shift, pe_old)
        stats = [e.spawn_death(tau, shift, pe_old, cycle) for e in engs]
        full = np.concatenate([e.determ_vector(0) for e in engs])
        for e in engs:
            e.determ_project(tau, shift, pe_old, cycle, full)
        outs = _exchange_and_annihilate(s, o, engs, cycle)
        for d, e in enumerate(engs):
            fo, po, do_ = o.get_psips(d)
            fg, pg, dg = e.download_psips()
            assert len(fg) == len(fo) == outs[d]["nstates"], (cycle, d)
            assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), (cycle, d)
            assert (e.determ_vector(1) == o.determ_vector(d)[0]).all()
        assert sum(st["nspawn_events"] for st in stats) == ro["nspawn_events"]
        assert sum(st["ndeath"] for st in stats) == ro["ndeath"]
    for e in engs:
        e.close()


def test_determ_space_error_paths():
    """hb200_set_determ_space refuses a rank's determinants out of list order or repeated; hb200_upload_psips refuses a
    new list that lacks a deterministic state; load-balancing redistribution refuses to run with a space set"""
    from hande_b200.engine import EngineError
    s, o, eng, ref = make_pair("h2o", excit_gen="renorm", tau=0.003, real=True)
    rf = 2**31
    f0 = ref["f0"].reshape(1, -1)
    eng.upload_psips(f0, [80 * rf], [0.0])
    eng.iterate(20, 0.003, 0.0, 0.0, 1)
    f, p, d = eng.download_psips()
    dets, sizes = SS.gather_determ_space(_SingleProcess(), SS.create_high_pop_space(_SingleProcess(), f, p, 10))
    with pytest.raises(EngineError):
        eng.set_determ_space(dets[::-1].copy(), sizes)                      # descending
    with pytest.raises(EngineError):
        eng.set_determ_space(np.concatenate([dets[:9], dets[8:9]]), sizes)  # repeated
    eng.set_determ_space(dets, sizes)
    with pytest.raises(EngineError):
        eng.upload_psips(f0, [80 * rf], [0.0])                              # nine deterministic states missing
    with pytest.raises(EngineError):
        eng.redistribute_particles()
    # the engine is still usable: a complete list is accepted and propagates
    eng.upload_psips(f, p, d)
    out = eng.iterate(2, 0.003, 0.0, 0.0, 50)
    assert out["spawn_error"] == 0 and out["nstates"] >= 10
    # the asynchronous upload relocates the deterministic states too (here: the list shifted by one new state in front)
    f2, p2, d2 = eng.download_psips()
    import torch
    hs = torch.from_numpy(f2.view(np.int64).copy()).pin_memory()
    hp = torch.from_numpy(p2.copy()).pin_memory()
    hd = torch.from_numpy(d2.copy()).pin_memory()
    eng.upload_psips_begin_ptr(hs.data_ptr(), hp.data_ptr(), hd.data_ptr(), len(p2))
    eng.upload_psips_commit()
    v0 = eng.determ_vector(0)
    keys = {tuple(x): k for k, x in enumerate(f2.tolist())}
    assert (v0 == np.array([p2[keys[tuple(x)]] / rf for x in dets.tolist()])).all()
    eng.close()


def test_load_balancing_with_a_deterministic_space():
    """redistribute_load_balancing_dets + redistribute_semi_stoch_t (src/qmc_common.F90:597-650,1332-1390): with the
    projection on, slots change owner - the moved determinants are annihilated without the deterministic flags, then the
    space is rebuilt from determ%dets under the new proc_map (recreate_determ_space) - and the projection carries on;
    three ranks on one device against the oracle's emulated ranks."""
    from hande_b200.fciqmc import owner_of
    name, gen, real, init, tau, world, nslots = "h2o", "renorm", True, False, 0.003, 3, 20
    s, o, engs = _setup(name, gen, real, init, tau, world, 4000, nslots=nslots)
    shift, pe_old = -0.05, -0.1

    def cycle_all(cycle):
        o.iterate(1, cycle, tau, shift, pe_old)
        for e in engs:
            e.spawn_death(tau, shift, pe_old, cycle)
        full = np.concatenate([e.determ_vector(0) for e in engs])
        for e in engs:
            e.determ_project(tau, shift, pe_old, cycle, full)
        _exchange_and_annihilate(s, o, engs, cycle)

    def compare(tag):
        for d, e in enumerate(engs):
            fo, po, do_ = o.get_psips(d)
            fg, pg, dg = e.download_psips()
            assert len(fg) == len(fo), (tag, d, len(fg), len(fo))
            assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), (tag, d)

    o.set_semi_stoch(space="high", size=120)
    o.init_semi_stoch()
    dets, sizes = o.determ_space()
    for e in engs:
        e.set_determ_space(dets, sizes)
    for cycle in (1, 2):
        cycle_all(cycle)
    compare("before")
    # a new proc_map: a third of the slots change owner
    pmap = np.arange(world * nslots) % world
    rng = np.random.default_rng(3)
    moved = rng.choice(world * nslots, size=world * nslots // 3, replace=False)
    pmap[moved] = (pmap[moved] + 1) % world
    o.set_proc_map(pmap)
    lb_cycle = 0x80000000 | 3
    o.redistribute(lb_cycle)
    for e in engs:
        e.set_determ_space(dets[:0], np.zeros(world, dtype=np.int32))
        e.set_proc_map(pmap)
        e.redistribute_particles()
    _exchange_and_annihilate(s, o, engs, lb_cycle)
    own = np.array([owner_of(f, s.nbasis, world, nslots, proc_map=pmap) for f in dets])
    from hande_b200.synthetic import sort_dets
    new_dets = np.concatenate([sort_dets(dets[own == r]) for r in range(world)])   # each rank's part in list order
    new_sizes = np.array([(own == r).sum() for r in range(world)], dtype=np.int32)
    dets_o, sizes_o = o.determ_space()
    assert (new_sizes == sizes_o).all() and (new_dets == dets_o).all() and (new_sizes != sizes).any()
    for e in engs:
        e.set_determ_space(new_dets, new_sizes)
    compare("redistributed")
    for cycle in (3, 4, 5):
        cycle_all(cycle)
    compare("after")
    for d, e in enumerate(engs):
        assert (e.determ_vector(1) == o.determ_vector(d)[0]).all()
        e.close()
