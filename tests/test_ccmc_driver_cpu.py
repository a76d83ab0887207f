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
