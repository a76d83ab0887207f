"""world_size-2 gloo test (CPU) of the multi-rank host path: the product driver hande_b200.fciqmc.do_fciqmc with
TorchDist collectives, one process per rank, must reproduce the reference's np2 golden trajectory
(test_suite/fciqmc/np2/Ne-aug-cc-pVDZ-ci6qmc) when each rank's propagation is done by the oracle."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NROWS = 60


def _worker(rank, world, port, fcidump, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hande_b200 import read_in as R
    from hande_b200.fciqmc import QmcIn, TorchDist, do_fciqmc
    from tests.oracle_engine import make_engine_cls
    if fcidump in ("ueg_np2", "ueg_qn_real64_np2"):
        from hande_b200.ueg import UegSystem
        g = load_golden(fcidump)
        u = g["ueg"]
        ueg = (u["nel"], u["ms"], u["rs"], u["cutoff"])
        s = UegSystem(*ueg)
        gq = g["qmc"]
        qmc = QmcIn(tau=gq["tau"], rng_seed=gq["seed"], init_pop=gq["D0_population"], mc_cycles=gq["ncycles"],
                    nreports=NROWS, target_population=gq["target_particles"], state_size=gq["walker_length"],
                    spawned_state_size=gq["spawned_walker_length"], reference_det=g["ref_det"],
                    real_amplitudes=bool(gq.get("real_amplitudes", 0)), spawn_cutoff=gq.get("spawn_cutoff", 0.01))
        qn = g.get("quasi_newton")
        if qn is not None:     # the host's init_propagator is checked against the oracle's inside the stand-in engine
            qmc.quasi_newton, qmc.quasi_newton_threshold = True, qn["threshold"]
        res = do_fciqmc(s, qmc, comm=TorchDist(),
                        engine_cls=make_engine_cls(None, None, rng_kind=0, ueg=ueg, ref_det=g["ref_det"], quasi_newton=qn))
        if rank == 0:
            q.put(res.rows)
        dist.barrier()
        dist.destroy_process_group()
        return
    g = load_golden("ne_ci6_np2")
    kw = dict(nel=g["sys"]["nel"], ms=g["sys"]["ms"], sym=g["sys"]["sym"], cas=tuple(g["sys"]["cas"]))
    s = R.read_in(fcidump, **kw)
    gq = g["qmc"]
    qmc = QmcIn(tau=gq["tau"], rng_seed=gq["seed"], init_pop=gq["D0_population"], mc_cycles=gq["ncycles"],
                nreports=NROWS, target_population=gq["target_particles"], state_size=gq["walker_length"],
                spawned_state_size=gq["spawned_walker_length"], ex_level=gq["ex_level"])
    res = do_fciqmc(s, qmc, comm=TorchDist(), engine_cls=make_engine_cls(fcidump, kw, rng_kind=0))
    if rank == 0:
        q.put(res.rows)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["ne_ci6_np2", "ueg_np2", "ueg_qn_real64_np2"])
def test_np2_driver_reproduces_golden(fcidump_path, case):
    from oracle import pyoracle
    if not pyoracle.have_ref_lib():
        pytest.skip("oracle/_ref not built")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    path = fcidump_path("ne") if case == "ne_ci6_np2" else case
    port = {"ne_ci6_np2": 29577, "ueg_np2": 29579}.get(case, 29581)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gold = np.array(load_golden(case)["rows"])

    def pr(x):
        return float("%.10E" % x)
    assert len(rows) == NROWS + 1
    for i, r in enumerate(rows):
        gr = gold[i]
        assert gr[0] == r[0]
        for k in (1, 2, 3, 4):
            assert gr[k] == pr(r[k]), (i, k, gr[k], r[k])
        assert gr[5] == r[5] and gr[6] == r[6]
        assert abs(gr[7] - r[7]) < 0.6e-4


def _worker_semi_stoch(rank, world, port, fcidump, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hande_b200 import read_in as R
    from hande_b200.fciqmc import QmcIn, TorchDist, do_fciqmc
    from oracle.pyoracle import HUGE
    from tests.oracle_engine import make_engine_cls
    kw = dict(nel=4, ms=0, sym=HUGE, cas=(-1, -1))
    s = R.read_in(fcidump, **kw)
    qmc = QmcIn(tau=0.01, rng_seed=7, init_pop=200, mc_cycles=10, nreports=12, target_population=400, real_amplitudes=True,
                spawn_cutoff=0.01, state_size=4000, spawned_state_size=2000, semi_stoch_space="high", semi_stoch_size=30,
                semi_stoch_start_iteration=27)
    res = do_fciqmc(s, qmc, comm=TorchDist(), engine_cls=make_engine_cls(fcidump, kw, rng_kind=0))
    if rank == 0:
        q.put((res.rows, res.determ_space[0].tolist(), res.determ_space[1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_np2_driver_semi_stochastic(fcidump_path):
    """The multi-rank host path of the semi-stochastic projection over real collectives (gloo, one process per rank):
    do_fciqmc picks the most populated determinants across the ranks (create_high_pop_space over allgather), hands the
    space to every rank's engine mid report loop, and the per-cycle all-gather of the deterministic amplitudes feeds each
    rank's projection - against the oracle's own two-rank run (reference dSFMT stream), every report row."""
    from oracle import pyoracle
    if not pyoracle.have_ref_lib():
        pytest.skip("oracle/_ref not built")
    from oracle.pyoracle import HUGE, Oracle
    path = fcidump_path("he2_avdz")
    o = Oracle()
    o.read_fcidump(path, nel=4, ms=0, sym=HUGE, cas=(-1, -1))
    o.set_qmc(tau=0.01, seed=7, D0_population=200, ncycles=10, nreport=12, target_particles=400, real_amplitudes=1,
              spawn_cutoff=0.01, walker_length=4000, spawned_walker_length=2000, nprocs=2)
    o.set_semi_stoch(space="high", size=30, start_iteration=27)
    o.init()
    rows_o = o.run()
    dets_o, sizes_o = o.determ_space()
    assert sizes_o.sum() == 30 and (sizes_o > 0).all()
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_semi_stoch, args=(r, 2, 29587, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, dets, sizes = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sizes == sizes_o.tolist() and dets == dets_o.tolist()
    assert len(rows) == len(rows_o) == 13
    for a, b in zip(rows, rows_o):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-11 * max(1.0, abs(b[k])), (k, a, b)
