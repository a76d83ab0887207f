"""Host side of the wall-Chebyshev propagator (hande_b200/propagators.py + the sub-cycle loop of do_fciqmc) on the CPU:
the product driver, with each cycle propagated by the oracle stand-in engine on the reference's dSFMT stream, must
reproduce the reference's own table (test_suite/fciqmc_real_64/np1/H4-STO-3g_cheby) row for row - spectral range from
the Gershgorin bound, the five weights, their update with the shift, harmonic forcing of the shift."""
import numpy as np
import pytest

from tests.conftest import load_golden


def test_driver_reproduces_chebyshev_golden(fcidump_path):
    from oracle import pyoracle
    if not pyoracle.have_ref_lib():
        pytest.skip("oracle/_ref not built")
    from hande_b200 import read_in as R
    from hande_b200.fciqmc import QmcIn, do_fciqmc
    from tests.oracle_engine import make_engine_cls
    g = load_golden("h4_cheby")
    path = fcidump_path(g["fcidump"])
    kw = dict(sym=g["sys"]["sym"])
    s = R.read_in(path, **kw)
    gq = g["qmc"]
    qmc = QmcIn(tau=0.001, rng_seed=gq["seed"], init_pop=gq["D0_population"], mc_cycles=gq["ncycles"], nreports=gq["nreport"],
                target_population=gq["target_particles"], state_size=gq["walker_length"],
                spawned_state_size=gq["spawned_walker_length"], real_amplitudes=True, spawn_cutoff=gq["spawn_cutoff"],
                vary_shift_from_proje=True, shift_damping=gq["shift_damping"], chebyshev=True,
                chebyshev_order=g["chebyshev"]["order"], shift_harmonic_crit_damp=True,
                shift_harmonic_forcing_two_stage=True)
    res = do_fciqmc(s, qmc, engine_cls=make_engine_cls(path, kw, rng_kind=0))
    k = g["kat"]
    c = res.chebyshev
    assert c.order == 5
    gold = np.array(g["rows"])

    def pr(x):
        return float("%.10E" % x)
    assert len(res.rows) == len(gold) == gq["nreport"] + 1
    for i, r in enumerate(res.rows):
        gr = gold[i]
        assert gr[0] == r[0]
        for kk in (1, 2, 3, 4):
            assert gr[kk] == pr(r[kk]), (i, kk, gr[kk], r[kk])
        assert gr[5] == r[5] and gr[6] == r[6]
        assert abs(gr[7] - r[7]) < 0.6e-4


def test_host_chebyshev_weights_match_the_reference_printout(fcidump_path):
    from hande_b200 import read_in as R
    from hande_b200.propagators import Chebyshev
    g = load_golden("h4_cheby")
    s = R.read_in(fcidump_path(g["fcidump"]), sym=0)
    occ0 = R.set_reference_det(s)
    c = Chebyshev(s, s.slater_condon0(occ0), order=5)
    k = g["kat"]
    assert float("%.8E" % c.spectral_range[1]) == k["spectral_range"]
    assert [float("%.8E" % x) for x in c.zeroes] == k["zeroes"]
    assert [float("%.8E" % x) for x in c.weights] == k["weights"]
    c2 = Chebyshev(s, s.slater_condon0(occ0), order=5, skip_gershgorin=True)
    assert 0.0 < c2.spectral_range[1] < c.spectral_range[1]
