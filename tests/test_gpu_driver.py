"""Whole-run tests of the host driver on the GPU engine (-m gpu): trajectory parity with the oracle under the same
Philox stream, and reblocked projected/shift energies against the oracle's run with the REFERENCE stream (dSFMT)."""
import numpy as np
import pytest

from hande_b200 import read_in as R
from hande_b200.fciqmc import QmcIn, do_fciqmc
from oracle.pyoracle import Oracle
from tests.blocking import optimal_error, ratio_with_error
from tests.common import system_path

pytestmark = pytest.mark.gpu


def _oracle_rows(path, kw, rng_kind, **q):
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(rng_kind=rng_kind, literal_event_int32=0 if rng_kind else 1, **q)
    o.init()
    return o.run()


def test_trajectory_matches_oracle_philox():
    path, kw = system_path("h2o")
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.003, rng_seed=7, init_pop=10, mc_cycles=20, nreports=40, target_population=1e7,
                state_size=-5, spawned_state_size=-1)
    res = do_fciqmc(s, qmc)
    rows = _oracle_rows(path, kw, 1, tau=0.003, seed=7, D0_population=10, ncycles=20, nreport=40,
                        target_particles=1e7, walker_length=178571, spawned_walker_length=31250)
    assert len(res.rows) == len(rows) == 41
    for g, o in zip(res.rows, rows):
        assert g[0] == o[0] and g[4] == o[4] and g[5] == o[5] and g[6] == o[6], (g, o)   # iterations, psips, states, events
        assert g[3] == o[3]                                                                 # N_0 (integer walkers)
        assert abs(g[2] - o[2]) <= 1e-10 * max(1.0, abs(o[2]))
        assert abs(g[7] - o[7]) <= 1e-12


def test_reblocked_energies_agree_with_reference_stream():
    """iFCIQMC with real amplitudes on H2O CAS(8,13) in the variable-shift regime: GPU (Philox) vs oracle with
    the reference's dSFMT stream.  Projected energy (ratio of means) and shift must agree within combined 2 sigma."""
    path, kw = system_path("h2o")
    s = R.read_in(path, **kw)
    common = dict(tau=0.004, ncycles=10, target=4000.0, nrep=700, skip=250)
    qmc = QmcIn(tau=common["tau"], rng_seed=3, init_pop=50, mc_cycles=common["ncycles"], nreports=common["nrep"],
                target_population=common["target"], state_size=200000, spawned_state_size=100000, initiator=True,
                real_amplitudes=True)
    g = np.array(do_fciqmc(s, qmc).rows)[1:]
    o = _oracle_rows(path, kw, 0, tau=common["tau"], seed=7, D0_population=50, ncycles=common["ncycles"],
                     nreport=common["nrep"], target_particles=common["target"], walker_length=200000,
                     spawned_walker_length=100000, initiator_approx=1, real_amplitudes=1)[1:]
    k = common["skip"]
    eg, sg = ratio_with_error(g[k:, 2], g[k:, 3])
    eo, so = ratio_with_error(o[k:, 2], o[k:, 3])
    assert abs(eg - eo) < 2.0 * np.hypot(sg, so), (eg, sg, eo, so)
    mg, esg = optimal_error(g[k:, 1])
    mo, eso = optimal_error(o[k:, 1])
    assert abs(mg - mo) < 2.0 * np.hypot(esg, eso), (mg, esg, mo, eso)
    # and both sit at the correlation energy of this active space (FCI ~ -0.1 Eh scale): sanity, not parity
    assert -0.5 < eg < 0.0


def test_heat_bath_reblocked_energies_agree_with_reference_stream():
    """Level-3 gate for the headline generator: real-amplitude iFCIQMC with excit_gen = heat_bath on NH3 6-31G (the
    system of the reference's per-generator golden tables) in the variable-shift regime - GPU engine (Philox) against
    the oracle run on the REFERENCE's dSFMT stream (which reproduces those golden tables).  Reblocked projected energy
    and shift must agree within combined 2 sigma."""
    path, kw = system_path("nh3")
    s = R.read_in(path, **kw)
    common = dict(tau=0.005, ncycles=10, target=2000.0, nrep=1500, skip=600)
    qmc = QmcIn(tau=common["tau"], rng_seed=5, init_pop=50, mc_cycles=common["ncycles"], nreports=common["nrep"],
                target_population=common["target"], state_size=200000, spawned_state_size=100000, initiator=True,
                real_amplitudes=True, excit_gen="heat_bath")
    g = np.array(do_fciqmc(s, qmc).rows)[1:]
    o = _oracle_rows(path, kw, 0, tau=common["tau"], seed=7, D0_population=50, ncycles=common["ncycles"],
                     nreport=common["nrep"], target_particles=common["target"], walker_length=200000,
                     spawned_walker_length=100000, initiator_approx=1, real_amplitudes=1, excit_gen="heat_bath")[1:]
    k = common["skip"]
    eg, sg = ratio_with_error(g[k:, 2], g[k:, 3])
    eo, so = ratio_with_error(o[k:, 2], o[k:, 3])
    assert abs(eg - eo) < 2.0 * np.hypot(sg, so), (eg, sg, eo, so)
    mg, esg = optimal_error(g[k:, 1])
    mo, eso = optimal_error(o[k:, 1])
    assert abs(mg - mo) < 2.0 * np.hypot(esg, eso), (mg, esg, mo, eso)
    assert -0.2 < eg < -0.05       # NH3 6-31G correlation energy is about -0.125 Eh (exact CCSDT -0.124920)


def test_ueg_trajectory_and_fci_energy():
    """UEG (the reference's ueg_n10_rs2_e4 fixture system: 6 electrons, 66 plane-wave spin-orbitals, rs = 2).
    (i) trajectory parity of the host driver + GPU engine with the oracle under the same Philox stream;
    (ii) the reblocked projected energy agrees with the exact correlation energy quoted in the reference's input
    file (test_suite/fciqmc/np2/ueg_n10_rs2_e4_fciqmc/ueg.fciqmc.in: "-0.176123766865", HANDE FCI)."""
    from hande_b200.ueg import UegSystem
    ueg = (6, 0, 2.0, 2.0)
    ref_det = [1, 2, 3, 10, 11, 14]
    s = UegSystem(*ueg)
    qmc = QmcIn(tau=0.005, rng_seed=122, init_pop=10, mc_cycles=10, nreports=60, target_population=90000,
                state_size=50000, spawned_state_size=5000, reference_det=ref_det)
    res = do_fciqmc(s, qmc)
    o = Oracle()
    o.init_ueg(*ueg)
    o.set_ref_det(ref_det)
    o.set_qmc(rng_kind=1, literal_event_int32=0, tau=0.005, seed=122, D0_population=10, ncycles=10, nreport=60,
              target_particles=90000, walker_length=50000, spawned_walker_length=5000)
    o.init()
    rows = o.run()
    assert res.H00 == o.reference()["H00"]
    assert len(res.rows) == len(rows) == 61
    for g, r in zip(res.rows, rows):
        assert g[0] == r[0] and g[4] == r[4] and g[5] == r[5] and g[6] == r[6] and g[3] == r[3], (g, r)
        assert abs(g[2] - r[2]) <= 1e-10 * max(1.0, abs(r[2]))
    # (ii) production-like run in the variable-shift regime, real amplitudes
    qmc = QmcIn(tau=0.02, rng_seed=5, init_pop=200, mc_cycles=10, nreports=1000, target_population=20000,
                state_size=400000, spawned_state_size=200000, real_amplitudes=True, reference_det=ref_det)
    g = np.array(do_fciqmc(s, qmc).rows)[1:]
    k = 400
    e, se = ratio_with_error(g[k:, 2], g[k:, 3])
    assert abs(e - (-0.176123766865)) < 4.0 * se + 2e-4, (e, se)
    m, sm = optimal_error(g[k:, 1])
    assert abs(m - (-0.176123766865)) < 4.0 * sm + 2e-3, (m, sm)


def test_ccmc_driver_trajectory_and_energy():
    """CCMC host driver on the GPU engine (the reference's CCSD fixture system, Ne cc-pVDZ):
    (i) report-loop trajectory identical to the oracle's under the same Philox stream;
    (ii) reblocked projected energy agrees within combined 2 sigma with the oracle run on the REFERENCE stream (dSFMT),
    which itself reproduces the reference's golden table."""
    from hande_b200.ccmc import do_ccmc
    path, kw = system_path("ne_vdz")
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.01, rng_seed=5691, init_pop=10, mc_cycles=10, nreports=60, target_population=20000,
                state_size=200000, spawned_state_size=100000, ex_level=2)
    res = do_ccmc(s, qmc)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(rng_kind=1, tau=0.01, seed=5691, D0_population=10, ncycles=10, nreport=60, target_particles=20000,
              walker_length=200000, spawned_walker_length=100000, ex_level=2)
    o.init()
    rows, na = o.run_ccmc()
    assert len(res.rows) == len(rows) == 61
    for g, r, a in zip(res.rows, rows, na):
        assert g[0] == r[0] and g[4] == r[4] and g[5] == r[5] and g[6] == r[6] and g[8] == a, (g, r, a)
        assert abs(g[2] - r[2]) <= 1e-10 * max(1.0, abs(r[2])) and abs(g[3] - r[3]) <= 1e-10 * max(1.0, abs(r[3]))
    # (ii) variable-shift regime, real amplitudes
    common = dict(tau=0.01, ncycles=10, target=3000.0, nrep=700, skip=300)
    qmc = QmcIn(tau=common["tau"], rng_seed=3, init_pop=100, mc_cycles=common["ncycles"], nreports=common["nrep"],
                target_population=common["target"], state_size=200000, spawned_state_size=100000, ex_level=2,
                real_amplitudes=True)
    g = np.array(do_ccmc(s, qmc).rows)[1:]
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(rng_kind=0, tau=common["tau"], seed=11, D0_population=100, ncycles=common["ncycles"], nreport=common["nrep"],
              target_particles=common["target"], walker_length=200000, spawned_walker_length=100000, ex_level=2,
              real_amplitudes=1)
    o.init()
    orow = o.run_ccmc()[0][1:]
    k = common["skip"]
    eg, sg = ratio_with_error(g[k:, 2], g[k:, 3])
    eo, so = ratio_with_error(orow[k:, 2], orow[k:, 3])
    assert abs(eg - eo) < 2.0 * np.hypot(sg, so) + 1e-4, (eg, sg, eo, so)
    assert -0.25 < eg < -0.15      # CCSD correlation energy of Ne/cc-pVDZ is about -0.19 Eh


@pytest.mark.parametrize("gen,tau", [("renorm", 0.0007), ("heat_bath_single", 0.003), ("power_pitzer_occ_ij", 0.0015)])
def test_ccmc_driver_pattempt_update(gen, tau):
    """qmc = { pattempt_update = true } through do_ccmc on the GPU engine, NH3 CCSDT (the system of the reference's
    per-generator golden tables): device-accumulated p_single_double sums -> host update_pattempt_single.  Same Philox
    stream as the oracle: identical trajectory, pattempt_single after each change to summation-order accuracy."""
    from hande_b200.ccmc import do_ccmc
    path, kw = system_path("nh3")
    s = R.read_in(path, **kw)
    nrep = 60
    qmc = QmcIn(tau=tau, rng_seed=30513, init_pop=200, mc_cycles=10, nreports=nrep, target_population=1200,
                state_size=200000, spawned_state_size=100000, ex_level=3, real_amplitudes=True, excit_gen=gen,
                pattempt_update=True)
    res = do_ccmc(s, qmc)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(rng_kind=1, tau=tau, seed=30513, D0_population=200, ncycles=10, nreport=nrep, target_particles=1200,
              walker_length=200000, spawned_walker_length=100000, ex_level=3, real_amplitudes=1, spawn_cutoff=0.01,
              excit_gen=gen)
    o.init()
    o.set_pattempt_update(True)
    rows, na = o.run_ccmc()
    log = o.pattempt_log()
    assert len(log) >= 2 and len(res.pattempt_log) == len(log)
    for a, b in zip(res.pattempt_log, log):
        assert abs(a - b) <= 1e-12 * b
    assert res.vary_shift and len(res.rows) == len(rows) == nrep + 1
    for g, r, a in zip(res.rows, rows, na):
        assert g[0] == r[0] and g[5] == r[5] and g[6] == r[6] and g[8] == a, (g, r, a)
        assert abs(g[1] - r[1]) <= 1e-9 * max(1.0, abs(r[1])) and abs(g[4] - r[4]) <= 1e-10 * max(1.0, abs(r[4]))
        assert abs(g[2] - r[2]) <= 1e-10 * max(1.0, abs(r[2])) and abs(g[3] - r[3]) <= 1e-10 * max(1.0, abs(r[3]))


def test_wall_chebyshev_trajectory_matches_oracle_philox():
    """qmc = { chebyshev = { chebyshev_order = 5 } } through do_fciqmc on the GPU engine (weights set per sub-cycle with
    hb200_set_propagator_weight) against the oracle's own Chebyshev run under the same Philox stream: every report row."""
    from tests.conftest import load_golden
    import gzip, os, tempfile
    g = load_golden("h4_cheby")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fcidump", g["fcidump"] + ".INTDUMP.gz")
    path = os.path.join(tempfile.gettempdir(), "hande_b200_h4.fcidump")
    with gzip.open(src, "rt") as fi, open(path, "w") as fo:
        fo.write(fi.read())
    s = R.read_in(path, sym=0)
    qmc = QmcIn(tau=0.001, rng_seed=11, init_pop=200, mc_cycles=2, nreports=25, target_population=2e4, state_size=1 << 16,
                spawned_state_size=1 << 16, real_amplitudes=True, vary_shift_from_proje=True, shift_damping=0.05,
                chebyshev=True, chebyshev_order=5, shift_harmonic_crit_damp=True, shift_harmonic_forcing_two_stage=True)
    res = do_fciqmc(s, qmc)
    o = Oracle()
    o.read_fcidump(path, sym=0)
    o.set_qmc(rng_kind=1, literal_event_int32=0, tau=1.0, seed=11, D0_population=200, ncycles=2, nreport=25,
              target_particles=2e4, real_amplitudes=1, spawn_cutoff=0.01, vary_shift_from_proje=1, shift_damping=0.05,
              walker_length=1 << 16, spawned_walker_length=1 << 16)
    o.init()
    o.init_chebyshev(order=5, harmonic_forcing=0.05 ** 2 / 4.0)
    rows = o.run()
    assert len(res.rows) == len(rows) == 26 and not res.error
    assert res.vary_shift and rows[-1][1] != 0.0          # the shift varied: update_chebyshev was exercised
    for a, b in zip(res.rows, rows):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-10 * max(1.0, abs(b[k])), (k, a, b)


def test_semi_stochastic_trajectory_matches_oracle_philox():
    """semi_stoch = { space = "high", size = 60, start_iteration = 45 } through do_fciqmc on the GPU engine (the host picks
    the space from the downloaded list, hb200_set_determ_space, projection inside hb200_iterate) against the oracle's
    own semi-stochastic run under the same Philox stream: every report row, with the shift varying.  The oracle's semi-stochastic runs reproduce the reference's tables with the dSFMT stream."""
    path, kw = system_path("h2o")
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.003, rng_seed=7, init_pop=300, mc_cycles=10, nreports=30, target_population=2000, real_amplitudes=True,
                state_size=1 << 17, spawned_state_size=1 << 16, semi_stoch_space="high", semi_stoch_size=60,
                semi_stoch_start_iteration=45)
    res = do_fciqmc(s, qmc)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(rng_kind=1, literal_event_int32=0, tau=0.003, seed=7, D0_population=300, ncycles=10, nreport=30,
              target_particles=2000, real_amplitudes=1, spawn_cutoff=0.01, walker_length=1 << 17,
              spawned_walker_length=1 << 16)
    o.set_semi_stoch(space="high", size=60, start_iteration=45)
    o.init()
    rows = o.run()
    dets_o, sizes_o = o.determ_space()
    assert len(res.rows) == len(rows) == 31 and not res.error
    assert (res.determ_space[0] == dets_o).all() and int(res.determ_space[1].sum()) == 60
    assert res.vary_shift and rows[-1][1] != 0.0
    for a, b in zip(res.rows, rows):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-10 * max(1.0, abs(b[k])), (k, a, b)


def test_semi_stochastic_ci_space_trajectory_matches_oracle_philox():
    """semi_stoch = { space = "ci", ci_space = { ex_level = 2 } } from the first iteration (the whole CISD space is added
    to the one-determinant list with zero population, add_determ_dets_to_psip_list) through do_fciqmc on the GPU engine
    against the oracle's own run under the same Philox stream, initiator approximation on."""
    path, kw = system_path("h2o")
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.003, rng_seed=7, init_pop=100, mc_cycles=10, nreports=15, target_population=1500, real_amplitudes=True,
                initiator=True, state_size=1 << 17, spawned_state_size=1 << 16, semi_stoch_space="ci", semi_stoch_ci_ex_level=2)
    res = do_fciqmc(s, qmc)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(rng_kind=1, literal_event_int32=0, tau=0.003, seed=7, D0_population=100, ncycles=10, nreport=15,
              target_particles=1500, real_amplitudes=1, spawn_cutoff=0.01, initiator_approx=1, walker_length=1 << 17,
              spawned_walker_length=1 << 16)
    o.set_semi_stoch(space="ci", ci_ex_level=2)
    o.init()
    rows = o.run()
    dets_o, sizes_o = o.determ_space()
    assert len(res.rows) == len(rows) == 16 and not res.error
    assert (res.determ_space[0] == dets_o).all() and len(dets_o) > 100
    assert rows[1][5] >= len(dets_o)                  # every deterministic state is in the list from the first cycle on
    for a, b in zip(res.rows, rows):
        assert a[0] == b[0] and a[5] == b[5] and a[6] == b[6], (a, b)
        for k in (1, 2, 3, 4):
            assert abs(a[k] - b[k]) <= 1e-10 * max(1.0, abs(b[k])), (k, a, b)
