"""The oracle is pinned against the reference's own golden trajectories (test_suite benchmark.out.*).

With the reference's dSFMT stream (compiled from the reference tree into oracle/_ref) the C++ restatement must
reproduce every printed column of the report table: shift, sum H0j Nj, N0, # H psips, # states, # spawn events,
R_spawn.  tools/golden_compare.py runs the complete tables (all rows verified when the fixtures were made);
here a leading slice of each table keeps the CPU suite to a few minutes.
"""
import numpy as np
import pytest

from tests.conftest import load_golden
from oracle import pyoracle
from oracle.pyoracle import Oracle, HUGE

pytestmark = pytest.mark.skipif(not pyoracle.have_ref_lib() and not __import__("os").path.isdir("/root/reference"),
                                reason="oracle/_ref (reference dSFMT) not built")


def _pr(x):
    return float("%.10E" % x)


def _run(case, fcidump_path, nrows, wide=False):
    g = load_golden(case)
    o = Oracle(wide=wide)
    if "ueg" in g:
        o.init_ueg(**g["ueg"])
        o.set_ref_det(g["ref_det"])
    else:
        s = g["sys"]
        o.read_fcidump(fcidump_path(g["fcidump"]), nel=s.get("nel", 0), ms=s.get("ms", HUGE), sym=s.get("sym", HUGE),
                       cas=tuple(s.get("cas", (-1, -1))))
    q = dict(g["qmc"])
    q["nreport"] = min(q["nreport"], nrows)
    o.set_qmc(**q)
    if g.get("pattempt_parallel", -1.0) >= 0:
        o.set_pattempt_parallel(g["pattempt_parallel"])
    if "quasi_newton" in g:
        o.set_quasi_newton(True, **g["quasi_newton"])
    if "semi_stoch" in g:
        o.set_semi_stoch(**g["semi_stoch"])
    elif "pop_real_bits" in g:
        o.set_semi_stoch(space="none", pop_real_bits=g["pop_real_bits"])
    o.init()
    if g.get("vary_shift"):
        o.set_vary_shift(True)
    if "harmonic_forcing" in g:
        o.set_harmonic_forcing(g["harmonic_forcing"])
    if "chebyshev" in g:
        hi, z, w = o.init_chebyshev(order=g["chebyshev"]["order"], harmonic_forcing=g["chebyshev"]["harmonic_forcing"])
        k = g["kat"]        # the spectral range, zeroes and weights the reference prints (9 significant digits)
        assert float("%.8E" % hi) == k["spectral_range"]
        assert [float("%.8E" % x) for x in z] == k["zeroes"] and [float("%.8E" % x) for x in w] == k["weights"]
    if g.get("ccmc"):
        o.ccmc_set_full_nc(bool(g.get("full_nc")))
        o.ccmc_set_pattempt_update(bool(g.get("pattempt_update")))
        rows, na = o.run_ccmc()
        rows = np.concatenate([rows, na.reshape(-1, 1).astype(float)], axis=1)
    else:
        rows = o.run()
    gold = np.array(g["rows"])
    n = min(len(gold), len(rows))
    assert n == q["nreport"] + 1
    for i in range(n):
        gr, r = gold[i], rows[i]
        assert gr[0] == r[0]
        for k in (1, 2, 3, 4):  # printed es17.10
            assert gr[k] == _pr(r[k]), (case, i, k, gr[k], r[k])
        assert gr[5] == r[5] and gr[6] == r[6], (case, i)
        assert abs(gr[7] - r[7]) < 0.6e-4
        if len(gr) > 8:
            assert gr[8] == r[8], (case, i)     # "# attempts" column of the CCMC table
    return o


def test_h2o_renorm_integer_np1(fcidump_path):
    o = _run("h2o", fcidump_path, 90)
    ref = o.reference()
    assert abs(ref["H00"] - (-76.02403856)) < 5e-9           # JSON block of the golden output
    assert abs(ref["pattempt_single"] - 0.04938272) < 5e-9
    assert abs(o.ecore - (-52.265550979288)) < 5e-13          # "E_core =" line
    assert list(ref["occ"]) == [1, 2, 3, 4, 5, 6, 7, 8]


def test_ne_initiator_np1(fcidump_path):
    # BASELINE config 1: integer-walker iFCIQMC, uniform (renorm) generator, ~1e5 walkers at the end
    _run("ne_init", fcidump_path, 200)


def test_ne_ci6_np2_hash_sharding(fcidump_path):
    _run("ne_ci6_np2", fcidump_path, 160)


def test_ne_ci6_np4_hash_sharding(fcidump_path):
    _run("ne_ci6_np4", fcidump_path, 160)


def test_ne_ci6_real_amplitudes_np2(fcidump_path):
    _run("ne_ci6_real64_np2", fcidump_path, 130)


def test_h4_wall_chebyshev_np1(fcidump_path):
    """the wall-Chebyshev propagator (order 5: init_chebyshev with the Gershgorin bound, five weighted
    sub-cycles per cycle, update_chebyshev) with harmonic forcing of the shift, real amplitudes - every row of the
    reference's table"""
    _run("h4_cheby", fcidump_path, 30)


def test_no_renorm_fciqmc_and_n2(fcidump_path):
    """Two more of the reference's FCIQMC tables: the no_renorm generator in an FCIQMC run (Ne CISDTQ, two ranks;
    complete 751-row table verified with tools/golden_compare.py) and N2 / STO-3G in D2h with nel and ms taken from the
    FCIDUMP header (two ranks, every one of its 1001 rows)"""
    _run("ne_cisdtq_no_renorm_np2", fcidump_path, 150)
    _run("n2_harmonic_np2", fcidump_path, 1000)


def test_real_amplitude_force_32(fcidump_path):
    """real_amplitude_force_32 (amplitudes stored times 2^11, src/particle_t_utils.f90): the truncated Ne runs of
    fciqmc_real_32/np{2,4}; the complete 1201-row tables were verified with tools/golden_compare.py
    (ne_ci6_real32_np2, ne_ci6_real32_np4)"""
    _run("ne_ci6_real32_np2", fcidump_path, 90)
    _run("ne_ci6_real32_np4", fcidump_path, 60)
    # and the UEG runs of fciqmc_real_32/np{2,4}/ueg_n10_rs2_e4_fciqmc_real_32 (complete 1001-row tables verified likewise)
    _run("ueg_real32_np2", fcidump_path, 150)
    _run("ueg_real32_np4", fcidump_path, 150)


def test_semi_stochastic_projection(fcidump_path):
    """semi-stochastic projection (src/semi_stoch.F90, deterministic_annihilation of
    src/annihilation.f90:488-535), every row of both tables of the reference that the container can run:
      * ueg_ss_np4: CI space of the quadruples (358 determinants over 4 ranks, 83..98 per rank as the reference prints),
        separate annihilation (the engine's mode): all-gathered vector, transposed CSR product, stochastic rounding of
        the deterministic amplitudes into the main list, deterministic states kept at zero population;
      * he2_ss: the 100 most populated determinants picked at iteration 1000 (create_high_pop_space), combined
        annihilation (the deterministic spawns travel through the spawned list)"""
    o = _run("ueg_ss_np4", fcidump_path, 200)
    k = load_golden("ueg_ss_np4")["kat"]
    dets, sizes = o.determ_space()
    assert len(dets) == k["determ_size"] and sizes.min() == k["determ_min"] and sizes.max() == k["determ_max"]
    assert abs(o.reference()["H00"] - k["H00"]) < 5e-9
    o = _run("he2_ss", fcidump_path, 200)
    k = load_golden("he2_ss")["kat"]
    dets, sizes = o.determ_space()
    assert len(dets) == k["determ_size"] and abs(o.reference()["H00"] - k["H00"]) < 5e-9
    rp, ci, mat = o.determ_hamil(0)
    assert rp[-1] == len(mat) and len(mat) > 100
    # the same He2 run in the reference's default mode, separate annihilation with the most populated determinants (the
    # engine's mode and the host driver's space, on one rank), and with the CISD space of the reference determinant (69
    # determinants after the point-group filter, as the reference prints)
    for case in ("he2_ss_sep", "he2_ss_cisd"):
        o = _run(case, fcidump_path, 200)
        assert len(o.determ_space()[0]) == load_golden(case)["kat"]["determ_size"]


def test_ueg_np2_np4(fcidump_path):
    # gen_excit_ueg_no_renorm, slater_condon0/2_ueg, update_proj_energy_ueg, plane-wave basis
    # ordering; the complete 1001-row tables were verified with tools/golden_compare.py (ueg_np2, ueg_np4)
    for case in ("ueg_np2", "ueg_np4"):
        o = _run(case, fcidump_path, 300)
        g = load_golden(case)
        assert abs(o.reference()["H00"] - g["kat"]["H00"]) < 5e-9     # "H00" of the golden JSON block
        t = o.ueg_tables()
        assert abs(t["L"] - g["kat"]["L"]) < 5e-9 and o.nbasis == g["kat"]["nbasis"]
        assert abs(o.basis()["sp_eigv"][2] - g["kat"]["sp_eigv_3"]) < 5e-10


def test_wide_oracle_build_reproduces_golden_tables(fcidump_path):
    """liboracle_wide.so (the same source with 32-word determinants, the checker of the wide device layout) on fixtures
    the 4-word build reproduces too: UEG np2 and the real-amplitude Ne CI6 np2 run"""
    _run("ueg_np2", fcidump_path, 150, wide=True)
    _run("ne_ci6_real64_np2", fcidump_path, 60, wide=True)


def test_ueg_quasi_newton_np2(fcidump_path):
    # quasi-Newton propagator (calc_qn_spawned_weighting, calc_qn_weighting, quasi_newton_pop_control,
    # sp_fock of the 3D UEG with exchange and Madelung terms); the complete 1001-row table was verified with
    # tools/golden_compare.py ueg_qn_real64_np2
    o = _run("ueg_qn_real64_np2", fcidump_path, 150)
    q = o.quasi_newton()
    assert q["threshold"] == 1.0 and q["value"] == 1.0 and q["pop_control"] == 1.0


def test_ccmc_ccsd_ne_np1(fcidump_path):
    # select_cluster / collapse_cluster / spawner_ccmc / stochastic_ccmc_death; the complete
    # 451-row table was verified with tools/golden_compare.py ccmc_ne
    o = _run("ccmc_ne", fcidump_path, 200)
    ref = o.reference()
    assert abs(ref["H00"] - (-128.48877555)) < 5e-9 and abs(ref["pattempt_single"] - 0.04511278) < 5e-9


def test_ccmc_ccsd_h2o_np2(fcidump_path):
    # two ranks: time-varying hash owner (hash_shift / move_freq), redistribute_particles, per-rank dSFMT streams;
    # the complete 176-row table was verified with tools/golden_compare.py ccmc_h2o_np2
    _run("ccmc_h2o_np2", fcidump_path, 45)


@pytest.mark.parametrize("case,nrows", [("ccmc_h2o_ccsdt_qn", 150), ("ccmc_h2o_ccsdt_qn_fullnc", 120)])
def test_ccmc_quasi_newton_np1(fcidump_path, case, nrows):
    """CCSDT with the quasi-Newton propagator (calc_qn_spawned_weighting in spawner_ccmc, calc_qn_weighting and
    quasi_newton_pop_control in stochastic_ccmc_death[_nc]), stochastic and full_non_composite selection"""
    _run(case, fcidump_path, nrows)


def test_ccmc_ccsdt_full_non_composite_np2(fcidump_path):
    # ccmc = { full_non_composite = true }, CCSDT in a CAS, two ranks: select_nc_cluster, do_nc_ccmc_propagation,
    # stochastic_ccmc_death_nc, deterministic reference selections (all 91 rows verified with tools/golden_compare.py)
    _run("ccmc_h2o_ccsdt_fullnc_np2", fcidump_path, 40)


@pytest.mark.parametrize("gen,nrows", [("hb", 150), ("hb_uni", 150), ("hb_single", 70), ("ppM", 120), ("ppMij", 120),
                                       ("csM", 120), ("csMij", 120), ("renorm", 250), ("no_renorm", 250),
                                       ("renorm_spin", 250), ("no_renorm_spin", 200), ("ppN", 150)])
def test_ccmc_ccsdt_nh3_np4_excitation_generators(fcidump_path, gen, nrows):
    """test_suite/ccmc_real_64/np4/NH3-6-31g_ccsdt_excit_gens: the reference's one golden trajectory per excitation
    generator (CCSDT, four ranks, real amplitudes) - the pin for heat_bath, heat_bath_uniform, heat_bath_single and the
    Power-Pitzer / Cauchy-Schwarz generators, and for qmc_in%pattempt_update (pattempt_single follows the spawn
    statistics; the printed '# pattempt_single changed to be:' values are compared too).  Comparable until the shift
    varies (blocking-on-the-fly with auto_shift_damping then changes the damping); tools/golden_compare.py verified
    every such row: hb 327, hb_uni 460, hb_single 179, ppM 691, ppMij 818, csM 701, csMij 817, renorm 1508,
    no_renorm 1424, renorm_spin 1476 (pattempt_parallel from find_parallel_spin_prob_mol: 0.22360108 in the golden
    JSON block), no_renorm_spin 373 (pattempt_parallel = 0.22 from the input), ppN ('heat_bath_power_pitzer_ref' =
    power_pitzer_orderN) 653."""
    case = "ccmc_nh3_" + gen
    g = load_golden(case)
    assert all(r[1] == 0.0 for r in g["rows"][:nrows + 1])
    o = _run(case, fcidump_path, nrows)
    last_it = g["rows"][nrows][0]
    got = o.ccmc_pattempt_log()
    # a change printed after row `it` happens at the end of the following report loop
    want = [v for it, v in g["pattempt_changes"] if it + g["qmc"]["ncycles"] <= last_it]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert float("%.10E" % a) == b
    if g["pattempt_update"] and nrows >= 60:
        assert len(want) >= 1
    if gen == "renorm_spin":
        assert abs(o.pattempt_parallel() - 0.22360108) < 5e-9


def test_dsfmt_and_murmur_known_answers():
    # seed 7 -> first close-open double 0.73384649635214716; MurmurHash2(0xFF, 4 bytes, seed 7) = -1594541972
    pyoracle.use_ref_lib()
    x = pyoracle.dsfmt_stream(7, 4)
    assert abs(x[0] - 0.73384649635214716) < 1e-17
    key = np.array([0xFF], dtype=np.uint32)
    L = pyoracle.lib()
    h_ref = np.int32(np.uint32(L.orc_ref_murmur2(key.ctypes.data, 4, 7)))
    h_own = np.int32(np.uint32(L.orc_murmur2(key.ctypes.data, 4, 7)))
    assert h_ref == -1594541972 and h_own == h_ref
    rng = np.random.default_rng(1)
    for n in (4, 8, 12, 16, 7, 13):
        buf = rng.integers(0, 256, size=n, dtype=np.uint8)
        assert L.orc_ref_murmur2(buf.ctypes.data, n, 7) == L.orc_murmur2(buf.ctypes.data, n, 7)
