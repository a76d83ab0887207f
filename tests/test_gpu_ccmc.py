"""GPU parity of the CCMC path (-m gpu): hb200_ccmc_spawn / hb200_ccmc_iterate against the oracle
(which reproduces the reference's CCSD golden table with the dSFMT stream) under the same Philox stream.  Bit-exact:
cluster choice, spawned / killed excips (spawn list), the annihilated excip list; 1e-12: block-reduced estimators."""
import numpy as np
import pytest

from tests.common import make_pair, sort_rows

pytestmark = pytest.mark.gpu


def _grow(o, tau, ncycles, pop0, rf):
    """Excip list after `ncycles` oracle cycles from the reference (realistic cluster amplitudes for the parity test)."""
    ref = o.reference()
    o.set_psips(ref["f0"].reshape(1, -1), [pop0 * rf], [0.0])
    pe_old = 0.0
    for c in range(1, ncycles + 1):
        st, _ = o.ccmc_stage_spawn(c, tau, 0.0, pe_old)
        o.stage_annihilate()
        if st["D0_population"] != 0:
            pe_old = st["proj_energy"] / st["D0_population"]
    return pe_old


CASES = [
    # system, generator, real amplitudes, ex_level, tau, warm-up cycles, full_non_composite[, quasi-Newton options]
    ("ne_vdz", "renorm", False, 3, 0.008, 100, False, dict(threshold=1e-5, value=1.0, pop_control=1.0)),   # quasi-Newton CCMC
    ("s12", "renorm", True, 3, 0.0005, 40, True, dict(threshold=0.3, value=1.0)),                          # ... full_nc
    ("ne_vdz", "renorm", False, 2, 0.01, 120, False),     # the reference's CCSD fixture
    ("ne_vdz", "no_renorm", True, 3, 0.005, 120, False),  # CCSDT, real amplitudes
    ("s12", "renorm", True, 3, 0.001, 40, False),
    ("s12", "heat_bath_uniform", True, 2, 0.001, 40, False),
    ("nh3", "renorm_spin", True, 3, 0.002, 60, False),    # the reference's per-generator CCSDT fixture system
    ("nh3", "heat_bath_single", True, 3, 0.003, 40, False),
    ("nh3", "power_pitzer_orderN", True, 3, 0.001, 80, False),
    ("s40", "renorm", False, 3, 0.0003, 30, False),         # two-word bit strings
    ("ueg6", "no_renorm", True, 2, 0.01, 80, False),      # CCMC on the UEG (doubles only)
    ("ne_vdz", "renorm", False, 3, 0.005, 100, True),     # full_nc: non-composite clusters + in-place death
    ("s12", "renorm", True, 2, 0.001, 40, True),
]


@pytest.mark.parametrize("case", CASES)
def test_ccmc_stage_and_cycle_parity(case):
    name, gen, real, exl, tau, warm, full_nc = case[:7]
    qn = case[7] if len(case) > 7 else None
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real, ex_level=exl, walker_length=1 << 18,
                               spawned_walker_length=1 << 17, quasi_newton=qn)
    o.ccmc_set_full_nc(full_nc)
    eng.ccmc_set_full_nc(full_nc)
    rf = 2**31 if real else 1
    pe_old = _grow(o, tau, warm, 200, rf)
    f, pops, dat = o.get_psips()
    assert len(f) > 30
    eng.upload_psips(f, pops, dat)
    shift = -0.02
    ps_on = name != "ueg6"
    if ps_on:      # qmc_in%pattempt_update statistics of the spawning attempts (cluster and non-composite kernels)
        o.set_pattempt_update(True)
        eng.set_pattempt(ref["pattempt_single"], ref["pattempt_double"], True)
        o.ps_stats(0, reset=True)
    for cycle in range(warm + 1, warm + 5):
        st_o, sd_o = o.ccmc_stage_spawn(cycle, tau, shift, pe_old)
        st_g = eng.ccmc_spawn(tau, shift, pe_old, cycle, exl)
        sd_g = eng.download_spawn()
        assert st_g["nattempts"] == st_o["nattempts"] and st_g["nattempts_spawn"] == st_o["nattempts_spawn"]
        assert st_g["D0_normalisation"] == st_o["D0_normalisation"]
        assert st_g["nspawn_events"] == st_o["nspawn_events"] == len(sd_o)
        assert (sort_rows(sd_g) == sort_rows(sd_o)).all()
        assert st_g["ndeath"] == st_o["ndeath"] and st_g["ndeath_nc"] == st_o["ndeath_nc"]
        for key in ("proj_energy", "D0_population"):
            assert abs(st_g[key] - st_o[key]) <= 1e-11 * max(1.0, abs(st_o[key])), key
        eng.annihilate_spawn()
        out = eng.annihilate_main(cycle)
        o.stage_annihilate()
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo) == out["nstates"]
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all()
        if st_o["D0_population"] != 0:
            pe_old = st_o["proj_energy"] / st_o["D0_population"]
    if ps_on:
        a, b = eng.get_ps_stats(reset=True), o.ps_stats(0)
        assert a[1] == b[1] and a[3] == b[3] and a[1] + a[3] > 100
        assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and abs(a[2] - b[2]) <= 1e-12 * abs(b[2])
    eng.close()


def test_ccmc_iterate_from_reference():
    """hb200_ccmc_iterate: 80 cycles from the reference in blocks of 10 against the oracle cycle by cycle."""
    name, gen, real, exl, tau = "ne_vdz", "renorm", False, 2, 0.01
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real, ex_level=exl, walker_length=1 << 18,
                               spawned_walker_length=1 << 17)
    f0 = ref["f0"].reshape(1, -1)
    o.set_psips(f0, [30], [0.0])
    eng.upload_psips(f0, [30], [0.0])
    cyc = 1
    for block in range(8):
        pe_sum = d0_sum = rsp = 0.0
        for c in range(cyc, cyc + 10):
            st, _ = o.ccmc_stage_spawn(c, tau, 0.0, -0.01 * block)
            o.stage_annihilate()
            pe_sum += st["proj_energy"]; d0_sum += st["D0_population"]
            if st["nattempts_spawn"] > 0:
                rsp += st["nspawn_events"] / st["nattempts_spawn"]
        rg = eng.ccmc_iterate(10, tau, 0.0, -0.01 * block, cyc, exl)
        cyc += 10
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo)
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all()
        assert abs(rg["proj_energy"] - pe_sum) <= 1e-11 * max(1.0, abs(pe_sum))
        assert abs(rg["D0_population"] - d0_sum) <= 1e-11 * max(1.0, abs(d0_sum))
        assert abs(rg["rspawn"] - rsp) <= 1e-12 * max(1.0, rsp)
        assert rg["nstates"] == len(fo)
    assert len(fg) > 10
    eng.close()
