"""CPU test of the product's host-side CCMC driver (hande_b200.ccmc.do_ccmc: report loop, estimator averages, shift
update, nattempts column): with the propagation done by the oracle (reference dSFMT stream) it must reproduce the
reference's CCSD golden table test_suite/ccmc/np1/Ne-RHF-cc-pVDZ_ccmc."""
import numpy as np
import pytest

from hande_b200 import read_in as R
from hande_b200.ccmc import do_ccmc
from hande_b200.fciqmc import QmcIn
from oracle import pyoracle
from tests.conftest import load_golden
from tests.oracle_engine import make_engine_cls

NROWS = 150


def test_ccmc_driver_reproduces_golden(fcidump_path):
    if not pyoracle.have_ref_lib():
        pytest.skip("oracle/_ref not built")
    g = load_golden("ccmc_ne")
    kw = dict(nel=g["sys"]["nel"], ms=g["sys"]["ms"], sym=g["sys"]["sym"])
    path = fcidump_path(g["fcidump"])
    s = R.read_in(path, **kw)
    gq = g["qmc"]
    qmc = QmcIn(tau=gq["tau"], rng_seed=gq["seed"], init_pop=gq["D0_population"], mc_cycles=gq["ncycles"],
                nreports=NROWS, target_population=gq["target_particles"], state_size=gq["walker_length"],
                spawned_state_size=gq["spawned_walker_length"], ex_level=gq["ex_level"])
    res = do_ccmc(s, qmc, engine_cls=make_engine_cls(path, kw, rng_kind=0))
    gold = np.array(g["rows"])

    def pr(x):
        return float("%.10E" % x)
    assert len(res.rows) == NROWS + 1
    for i, r in enumerate(res.rows):
        gr = gold[i]
        assert gr[0] == r[0]
        for k in (1, 2, 3, 4):
            assert gr[k] == pr(r[k]), (i, k, gr[k], r[k])
        assert gr[5] == r[5] and gr[6] == r[6] and gr[8] == r[8], (i, gr, r)
        assert abs(gr[7] - r[7]) < 0.6e-4


def test_ccmc_driver_pattempt_update(fcidump_path):
    """qmc = { pattempt_update = true }: the host half (hande_b200.fciqmc.PattemptUpdate: rep_accum -> total,
    update_pattempt_single, stop when the shift varies) against the oracle's own report loop, whose np4 runs reproduce
    the reference's NH3 golden tables with pattempt_update (tests/test_oracle_golden.py)."""
    if not pyoracle.have_ref_lib():
        pytest.skip("oracle/_ref not built")
    g = load_golden("ccmc_nh3_renorm")
    kw = dict(nel=g["sys"]["nel"], ms=g["sys"]["ms"], sym=g["sys"]["sym"])
    path = fcidump_path(g["fcidump"])
    s = R.read_in(path, **kw)
    gq = g["qmc"]
    nrows = 160
    o = pyoracle.Oracle()
    o.read_fcidump(path, **kw)
    q = dict(gq, nprocs=1, nreport=nrows, target_particles=2500.0)
    o.set_qmc(**q)
    o.init()
    o.set_pattempt_update(True)
    rows, na = o.run_ccmc()
    log = o.pattempt_log()
    assert len(log) >= 3 and rows[-1][1] != 0.0       # several updates, then the shift varies and the updates stop
    qmc = QmcIn(tau=gq["tau"], rng_seed=gq["seed"], init_pop=gq["D0_population"], mc_cycles=gq["ncycles"],
                nreports=nrows, target_population=2500.0, state_size=gq["walker_length"],
                spawned_state_size=gq["spawned_walker_length"], ex_level=gq["ex_level"], real_amplitudes=True,
                excit_gen="renorm", pattempt_update=True)
    res = do_ccmc(s, qmc, engine_cls=make_engine_cls(path, kw, rng_kind=0))
    assert list(res.pattempt_log) == list(log)
    assert len(res.rows) == len(rows)
    for r, ro, n in zip(res.rows, rows, na):
        assert list(r[:8]) == list(ro[:8]) and r[8] == n


def test_ccmc_driver_quasi_newton_reproduces_golden(fcidump_path):
    """do_ccmc with qmc = { quasi_newton = true, ... } (host init_propagator -> hb200_set_quasi_newton) on the oracle
    stand-in: the reference's CCSDT quasi-Newton table test_suite/ccmc/np1/H2O-cc-pVDZ_ccsdtmc_qn, first 60 rows."""
    if not pyoracle.have_ref_lib():
        pytest.skip("oracle/_ref not built")
    g = load_golden("ccmc_h2o_ccsdt_qn")
    kw = dict(nel=g["sys"]["nel"], ms=g["sys"]["ms"], sym=g["sys"]["sym"])
    path = fcidump_path(g["fcidump"])
    s = R.read_in(path, **kw)
    gq, qn = g["qmc"], g["quasi_newton"]
    nrows = 60
    qmc = QmcIn(tau=gq["tau"], rng_seed=gq["seed"], init_pop=gq["D0_population"], mc_cycles=gq["ncycles"],
                nreports=nrows, target_population=gq["target_particles"], state_size=gq["walker_length"],
                spawned_state_size=gq["spawned_walker_length"], ex_level=gq["ex_level"], quasi_newton=True,
                quasi_newton_threshold=qn["threshold"], quasi_newton_value=qn["value"],
                quasi_newton_pop_control=qn["pop_control"])
    res = do_ccmc(s, qmc, engine_cls=make_engine_cls(path, kw, rng_kind=0, quasi_newton=qn))
    gold = np.array(g["rows"])

    def pr(x):
        return float("%.10E" % x)
    assert len(res.rows) == nrows + 1
    for i, r in enumerate(res.rows):
        gr = gold[i]
        assert gr[0] == r[0]
        for k in (1, 2, 3, 4):
            assert gr[k] == pr(r[k]), (i, k, gr[k], r[k])
        assert gr[5] == r[5] and gr[6] == r[6] and gr[8] == r[8], (i, gr, r)
