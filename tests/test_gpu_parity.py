"""GPU parity tests (-m gpu): every call goes through the C ABI of libhande_b200.so and is compared with the
oracle fed the same Philox stream.  Bit-exact: excitation choice, nspawn, spawn lists, sort, annihilation, merged
main list (states, pops) and dat (same summation order, no FMA).  1e-12 relative: block-reduced estimators."""
import numpy as np
import pytest

from hande_b200.engine import EngineError
from tests.common import make_pair, random_population, sort_rows

pytestmark = pytest.mark.gpu


def _check_gen(name, gen, real, tau, n=200, nattempt=4, cycles=(1, 7)):
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real)
    f, pops, dat = random_population(s, o, n, real, seed=3)
    nd = len(f)
    for cyc in cycles:
        att = np.tile(np.arange(nattempt, dtype=np.uint32), nd)
        ff = np.repeat(f, nattempt, axis=0)
        pp = np.repeat(pops, nattempt)
        io, do, ns = eng.gen_excit_batch(ff, pp, att, cyc, tau)
        for k in range(len(pp)):
            io_o, do_o, ns_o = o.gen_excit_philox(ff[k], cyc, int(att[k]), int(pp[k]), tau)
            assert io_o[6] == io[k, 6], (k, io_o, io[k])
            if io_o[6]:
                assert list(io_o[:6]) == list(io[k, :6]), (k, io_o, io[k])
            assert do_o[0] == do[k, 0] and do_o[1] == do[k, 1], (k, do_o, do[k])
            assert ns_o == ns[k], (k, ns_o, ns[k])
    sc = eng.sc0_batch(f)
    for k in range(nd):
        assert sc[k] == o.sc0(f[k])
    eng.close()


def test_gen_excit_renorm_h2o():
    _check_gen("h2o", "renorm", False, 0.003)


def test_gen_excit_no_renorm_h2o():
    _check_gen("h2o", "no_renorm", False, 0.003)


def test_gen_excit_renorm_ne_real():
    _check_gen("ne", "renorm", True, 0.005, n=100)


def test_gen_excit_renorm_two_words():
    _check_gen("s40", "renorm", False, 0.01, n=100)


def test_gen_excit_heat_bath():
    _check_gen("s10", "heat_bath", True, 0.01, n=150, nattempt=6)


def test_gen_excit_heat_bath_uniform():
    _check_gen("s10", "heat_bath_uniform", True, 0.01, n=150, nattempt=6)


@pytest.mark.parametrize("gen", ["heat_bath", "heat_bath_uniform", "renorm"])
def test_gen_excit_s50_bench_system(gen):
    """The configuration bench.py times (S50: nbasis 100, W = 2, nel 20, the 2 GB hb_ijab tables): excitation choice,
    pgen, H_ij and nspawn of 500 determinants x 6 attempts x 2 cycles against the oracle."""
    _check_gen("s50", gen, True, 2e-5, n=500, nattempt=6)


@pytest.mark.parametrize("gen", ["renorm", "heat_bath", "heat_bath_uniform", "power_pitzer_orderN", "renorm_spin"])
def test_gen_excit_uhf(gen):
    """UHF integral channels on the device (four two-body channels, spin-resolved one-body and sc1 tables)"""
    _check_gen("s10u", gen, True, 0.01, n=150, nattempt=6)


def test_gen_excit_renorm_spin_and_no_renorm_spin():
    # choose_ij_spin_mol variants of the uniform generators
    _check_gen("h2o", "renorm_spin", False, 0.003, n=120)
    _check_gen("nh3", "no_renorm_spin", True, 0.002, n=120)
    _check_gen("s40", "renorm_spin", True, 0.01, n=80)


def test_gen_excit_power_pitzer_orderN():
    # 'heat_bath_power_pitzer_ref': tables built on the device (hb200_build_power_pitzer_orderN) - any differing bit in
    # a weight, a total or an alias entry would show up in pgen or in the excitation chosen
    _check_gen("nh3", "power_pitzer_orderN", True, 0.002, n=150, nattempt=6)
    _check_gen("h2o", "power_pitzer_orderN", False, 0.003, n=120)
    _check_gen("s12", "power_pitzer_orderN", True, 0.01, n=120)
    _check_gen("s40", "power_pitzer_orderN", False, 0.01, n=60)


def test_gen_excit_power_pitzer_reference_mapped():
    # excit_gen = power_pitzer (gen_excit_mol_power_pitzer_occ_ref): pp_ia_d / pp_jb_d built on the device
    _check_gen("nh3", "power_pitzer", True, 0.002, n=150, nattempt=6)
    _check_gen("h2o", "power_pitzer", False, 0.003, n=120)
    _check_gen("s40", "power_pitzer", False, 0.01, n=60)


def test_tables_need_the_reference():
    from hande_b200 import read_in as R
    from hande_b200.engine import Engine
    from tests.common import system_path
    path, kw = system_path("nh3")
    s = R.read_in(path, **kw)
    for gen in ("power_pitzer", "power_pitzer_orderN"):
        eng = Engine(s, excit_gen=gen, pattempt_single=0.1, pattempt_double=0.9, walker_length=1024, spawned_walker_length=512)
        eng.upload_psips(np.zeros((0, 1), dtype=np.uint64), [], [])
        with pytest.raises(EngineError):
            eng.iterate(1, 0.001, 0.0, 0.0, 1)         # tables are built by set_reference
        eng.close()


def test_pattempt_parallel_on_device():
    """find_parallel_spin_prob_mol on the device against the oracle (whose value reproduces the 0.22360108 printed in
    the reference's NH3 renorm_spin golden output)."""
    for name in ("nh3", "h2o"):
        s, o, eng, ref = make_pair(name, excit_gen="renorm_spin", tau=0.003)
        pp = eng.set_pattempt_parallel(-1.0)
        assert abs(pp - o.pattempt_parallel()) <= 1e-13
        if name == "nh3":
            assert abs(pp - 0.22360108) < 5e-9
        eng.close()


def test_gen_excit_heat_bath_single():
    _check_gen("s10", "heat_bath_single", True, 0.01, n=150, nattempt=6)
    _check_gen("h2o", "heat_bath_single", False, 0.003, n=100)


def test_gen_excit_power_pitzer_and_cauchy_schwarz_occ():
    # the O(M) on-the-fly variants with uniformly chosen ij
    _check_gen("h2o", "power_pitzer_occ", False, 0.003, n=120)
    _check_gen("s10", "cauchy_schwarz_occ", True, 0.01, n=120, nattempt=6)
    _check_gen("s10", "power_pitzer_occ_ij", True, 0.01, n=120, nattempt=6)
    _check_gen("s12", "cauchy_schwarz_occ_ij", False, 0.01, n=100, nattempt=4)


def test_gen_excit_ueg():
    # gen_excit_ueg_no_renorm + slater_condon0_ueg on the device, W = 2 and W = 3
    _check_gen("ueg6", "no_renorm", False, 0.005, n=150, nattempt=5)
    _check_gen("ueg14", "no_renorm", True, 0.002, n=150, nattempt=5)
    # gen_excit_ueg_power_pitzer, table built on the device (init_excit_ueg_power_pitzer)
    _check_gen("ueg6", "power_pitzer", False, 0.005, n=150, nattempt=5)
    _check_gen("ueg14", "power_pitzer", True, 0.002, n=150, nattempt=5)


@pytest.mark.parametrize("name,gen", [("ueg358", "no_renorm"), ("ueg2042", "no_renorm")])
def test_gen_excit_ueg_wide(name, gen):
    """Bit strings wider than 4 words (BASELINE configs[3], ~1000 plane waves): find_ab_ueg over 32-word strings,
    ternary_conserve rows re-strided to the device width, 16-bit occupied lists, slater_condon0_ueg"""
    _check_gen(name, gen, True, 0.002, n=150, nattempt=5)


@pytest.mark.parametrize("name", ["s10", "s50"])
def test_heat_bath_tables_match_oracle(name):
    """every entry of the device-built tables (S50: 1e8-entry ijab_w / ijab_U / ijab_K, the tables the bench reads)"""
    s, o, eng, ref = make_pair(name, excit_gen="heat_bath")
    hb = o.heat_bath_tables()
    nb = hb["nb"]
    names = ["i_weights", "ij_weights", "ija_w", "ija_U", "ija_tot", "ijab_w", "ijab_U", "ijab_tot", "ija_K", "ijab_K"]
    for which, nm in enumerate(names):
        ref_t = np.asarray(hb[nm])
        got = eng.heat_bath_table(which, len(ref_t))
        if nm in ("ija_U", "ija_K", "ijab_U", "ijab_K"):
            # rows with zero total weight are never initialised/used by the reference: compare the used rows only
            tot = np.asarray(hb["ija_tot"] if nm.startswith("ija_") else hb["ijab_tot"])
            used = np.repeat(np.abs(tot) > 0, nb)
            assert (got[used] == ref_t[used]).all(), nm
        else:
            assert (got == ref_t).all(), nm
    eng.close()


def test_radix_sort_spawn_list():
    s, o, eng, ref = make_pair("s40", spawned_walker_length=1 << 16)
    rng = np.random.default_rng(4)
    for n in (1, 2, 37, 255, 256, 257, 5000, 40000):
        keys = rng.integers(0, 2**63 - 1, size=(n, 2), dtype=np.int64)
        keys[:, 1] &= (1 << (s.nbasis - 64)) - 1
        keys[rng.integers(0, n, size=n // 3)] = keys[0]          # duplicates
        sd = np.concatenate([keys, rng.integers(-5, 6, size=(n, 1)), np.arange(n).reshape(-1, 1)], axis=1)
        eng.upload_spawn(sd)
        eng.annihilate_spawn()
        got = eng.download_spawn()
        # reference order: unsigned compare, word 1 most significant, stable
        order = np.lexsort((sd[:, 0].astype(np.uint64), sd[:, 1].astype(np.uint64)))
        assert (got == sd[order]).all(), n
    eng.close()


CASES = [
    # name, generator, real, initiator, tau, nwalkers, ex_level
    ("h2o", "renorm", False, False, 0.003, 3000, -1),
    ("h2o", "no_renorm", False, True, 0.003, 3000, -1),
    ("ne_cas", "renorm", True, False, 0.004, 4000, 5),
    ("ne", "renorm", True, True, 0.005, 5000, -1),
    ("s40", "renorm", False, True, 0.02, 3000, -1),
    ("s12", "heat_bath", True, True, 0.01, 2500, -1),
    ("s10u", "heat_bath", True, True, 0.01, 2500, -1),          # UHF channels
    ("s10u", "renorm", False, False, 0.01, 2500, -1),
    ("s50", "heat_bath", True, True, 2e-5, 6000, -1),          # the bench configuration
    ("s50", "heat_bath_uniform", True, True, 2e-5, 6000, -1),
    ("s50", "renorm", True, False, 2e-4, 6000, -1),
    ("s12", "heat_bath_uniform", True, True, 0.01, 2500, -1),
    ("s12", "heat_bath_single", True, True, 0.01, 2500, -1),
    ("nh3", "renorm_spin", True, True, 0.003, 2500, -1),
    ("nh3", "power_pitzer_orderN", True, True, 0.002, 2500, -1),
    ("nh3", "power_pitzer", True, False, 0.002, 2500, -1),
    ("s12", "power_pitzer_orderN", False, False, 0.004, 2500, 4),
    ("h2o", "no_renorm_spin", False, False, 0.003, 2500, -1),
    ("h2o", "power_pitzer_occ", False, True, 0.003, 2500, -1),
    ("s12", "cauchy_schwarz_occ", True, False, 0.004, 2500, -1),
    ("s12", "power_pitzer_occ_ij", True, True, 0.004, 2500, -1),
    ("ueg6", "no_renorm", False, False, 0.01, 3000, -1),
    ("ueg14", "no_renorm", True, True, 0.004, 4000, -1),
    ("ueg14", "power_pitzer", True, True, 0.004, 4000, -1),
    ("ueg358", "no_renorm", True, True, 0.004, 4000, -1),       # wide layout
    ("ueg2042", "no_renorm", True, False, 0.004, 4000, -1),     # BASELINE configs[3] size
    ("ueg2042", "no_renorm", False, True, 0.004, 3000, -1),
]


@pytest.mark.parametrize("name,gen,real,init,tau,n,exl", CASES)
def test_stage_and_cycle_parity(name, gen, real, init, tau, n, exl):
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real, initiator=init, ex_level=exl)
    f, pops, dat = random_population(s, o, n, real, seed=5)
    o.set_psips(f, pops, dat)
    eng.upload_psips(f, pops, dat)
    shift, pe_old = -0.05, -0.11
    rf = 2**31 if real else 1
    # qmc_in%pattempt_update statistics (refused for heat_bath and the UEG as in src/check_input.F90:192-197)
    ps_on = gen != "heat_bath" and not name.startswith("ueg")
    if ps_on:
        o.set_pattempt_update(True)
        eng.set_pattempt(ref["pattempt_single"], ref["pattempt_double"], True)
    else:
        with pytest.raises(EngineError):
            eng.set_pattempt(ref["pattempt_single"], ref["pattempt_double"], True)
    for cycle in (1, 2, 3):
        if ps_on and cycle == 3:      # a changed pattempt_single takes effect in the generators and in the sums
            o.set_pattempt(0.21, 0.79)
            eng.set_pattempt(0.21, 0.79, True)
        st_o, sd_o = o.stage_spawn(cycle, tau, shift, pe_old)
        st_g = eng.spawn_death(tau, shift, pe_old, cycle)
        sd_g = eng.download_spawn()
        assert st_g["nspawn_events"] == st_o["nspawn_events"] == len(sd_o)
        assert (sort_rows(sd_g) == sort_rows(sd_o)).all()
        assert st_g["ndeath"] == st_o["ndeath"]
        assert abs(st_g["proj_energy"] - st_o["proj_energy"]) <= 1e-12 * max(1.0, abs(st_o["proj_energy"]))
        assert abs(st_g["D0_population"] - st_o["D0_population"]) <= 1e-12 * max(1.0, abs(st_o["D0_population"]))
        eng.comm_spawn()
        eng.annihilate_spawn()
        srt = eng.download_spawn()
        k = srt[:, : s.W].astype(np.uint64)
        order = np.lexsort(tuple(k[:, w] for w in range(s.W)))
        assert (order == np.arange(len(srt))).all() or (k[order] == k).all()
        out = eng.annihilate_main(cycle)
        o.stage_annihilate()
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo) == out["nstates"]
        assert (fg == fo).all()
        assert (pg == po).all()
        assert (dg == do_).all()
        assert out["nparticles"] == pytest.approx(np.abs(po).sum() / rf, rel=1e-14)
    if ps_on:
        _check_ps_stats(eng, o)
    eng.close()


def _check_ps_stats(eng, o):
    """p_single_double_coll_t sums of the report loop: counts exact, sums to summation-order accuracy"""
    a, b = eng.get_ps_stats(reset=True), o.ps_stats(0)
    assert a[1] == b[1] and a[3] == b[3] and a[1] + a[3] > 100
    assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and abs(a[2] - b[2]) <= 1e-12 * abs(b[2])
    assert (eng.get_ps_stats() == 0).all()


@pytest.mark.parametrize("name,gen,real,init,tau", [("h2o", "renorm", False, False, 0.003),
                                                     ("ne", "renorm", True, True, 0.005),
                                                     ("s12", "heat_bath", True, True, 0.002),
                                                     ("s12", "heat_bath_uniform", True, False, 0.0004),
                                                     ("ueg6", "no_renorm", False, False, 0.005),
                                                     ("ueg14", "no_renorm", True, True, 0.002),
                                                     ("ueg6", "power_pitzer", True, False, 0.005),
                                                     ("ueg358", "no_renorm", True, True, 0.002),
                                                     ("ueg2042", "no_renorm", True, False, 0.0003)])
def test_iterate_from_single_determinant(name, gen, real, init, tau):
    """Population growth from the reference determinant: 60 cycles in blocks of 10 through hb200_iterate."""
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real, initiator=init)
    rf = 2**31 if real else 1
    f0 = ref["f0"].reshape(1, -1)
    o.set_psips(f0, [50 * rf], [0.0])
    eng.upload_psips(f0, [50 * rf], [0.0])
    cyc = 1
    for block in range(6):
        ro = o.iterate(10, cyc, tau, 0.0, -0.02 * block)
        rg = eng.iterate(10, tau, 0.0, -0.02 * block, cyc)
        cyc += 10
        assert rg["spawn_error"] == 0 and rg["psip_error"] == 0 and ro["error"] == 0
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo)
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all()
        assert rg["nspawn_events"] == ro["nspawn_events"] and rg["ndeath"] == ro["ndeath"]
        assert rg["nattempts"] == ro["nattempts"]
        for key in ("proj_energy", "D0_population", "rspawn", "nparticles"):
            assert abs(rg[key] - ro[key]) <= 1e-12 * max(1.0, abs(ro[key])), key
    assert len(fg) > 20
    eng.close()


@pytest.mark.parametrize("name,gen,real,init,tau,qn", [("ueg6", "no_renorm", True, False, 0.005, dict(threshold=1.0)),
                                                        ("h2o", "renorm", False, True, 0.003, dict()),
                                                        ("s12", "heat_bath", True, True, 0.002, dict(threshold=0.3, value=0.5))])
def test_quasi_newton_propagator(name, gen, real, init, tau, qn):
    """qmc = { quasi_newton = true }: spawn amplitudes scaled by calc_qn_spawned_weighting, death by
    calc_qn_weighting and quasi_newton_pop_control; 40 cycles from the reference determinant against the oracle, whose
    quasi-Newton runs reproduce the reference's UEG golden table."""
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real, initiator=init, quasi_newton=qn)
    rf = 2**31 if real else 1
    f0 = ref["f0"].reshape(1, -1)
    o.set_psips(f0, [50 * rf], [0.0])
    eng.upload_psips(f0, [50 * rf], [0.0])
    cyc = 1
    for block in range(4):
        ro = o.iterate(10, cyc, tau, 0.01, -0.02 * block)
        rg = eng.iterate(10, tau, 0.01, -0.02 * block, cyc)
        cyc += 10
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo)
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all()
        assert rg["nspawn_events"] == ro["nspawn_events"] and rg["ndeath"] == ro["ndeath"]
    assert len(fg) > 10
    eng.close()


def test_edge_cases_empty_and_overflow():
    s, o, eng, ref = make_pair("h2o", tau=0.003, walker_length=64, spawned_walker_length=32)
    # empty main list
    eng.upload_psips(np.zeros((0, 1), dtype=np.uint64), [], [])
    out = eng.iterate(2, 0.003, 0.0, 0.0, 1)
    assert out["nstates"] == 0 and out["nspawn_events"] == 0 and out["spawn_error"] == 0
    # a single huge population: spawn list (32) and main list (64) overflow -> flags, no abort
    eng.upload_psips(ref["f0"].reshape(1, -1), [2000000], [0.0])
    out = eng.iterate(1, 0.003, 0.0, 0.0, 5)
    assert out["spawn_error"] == 1
    assert out["nstates"] <= 64
    eng.close()


def test_async_upload_begin_commit():
    """hb200_upload_psips_begin/_commit: a list uploaded on the copy stream while another one propagates becomes the
    current list at commit and then evolves exactly like a synchronously uploaded one."""
    s, o, eng, ref = make_pair("h2o", tau=0.003)
    fa, pa, da = random_population(s, o, 2000, False, seed=5)
    fb, pb, db = random_population(s, o, 1500, False, seed=8)
    eng.upload_psips(fa, pa, da)
    keep = [np.ascontiguousarray(fb), np.ascontiguousarray(pb), np.ascontiguousarray(db)]
    eng.upload_psips_begin_ptr(keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, len(pb))
    eng.iterate(2, 0.003, 0.0, 0.0, 1)          # list A propagates while B is in flight
    eng.upload_psips_commit()
    f1, p1, d1 = eng.download_psips()
    assert (f1 == fb).all() and (p1 == pb).all() and (d1 == db).all()
    out_async = eng.iterate(3, 0.003, 0.0, -0.1, 11)
    g1 = eng.download_psips()
    eng.upload_psips(fb, pb, db)
    out_sync = eng.iterate(3, 0.003, 0.0, -0.1, 11)
    g2 = eng.download_psips()
    assert all((x == y).all() for x, y in zip(g1, g2))
    assert out_async["nparticles"] == out_sync["nparticles"] and out_async["proj_energy"] == out_sync["proj_energy"]
    eng.close()


@pytest.mark.parametrize("initiator", [False, True])
def test_annihilation_of_long_runs(initiator):
    """Runs of thousands of spawn events on one determinant (the reference near convergence): k_annihilate hands runs
    longer than 32 elements to the warp-per-run kernel (galloping search for the end of the run, integer sums,
    initiator flag algebra) - annihilate_spawn_t[_initiator] + annihilate_main_list + insert_new_walkers against the
    oracle on the same list."""
    import ctypes as C
    from hande_b200 import synthetic
    s, o, eng, ref = make_pair("s40", real=False, initiator=initiator, spawned_walker_length=1 << 17)
    f, pops, dat = random_population(s, o, 300, False, seed=7)
    o.set_psips(f, pops, dat)
    eng.upload_psips(f, pops, dat)
    rng = np.random.default_rng(5)
    have = {tuple(x) for x in f}
    new = np.array([x for x in synthetic.random_dets(400, s.nbasis, s.nalpha, s.nbeta, seed=99) if tuple(x) not in have])
    runs = [(f[10], 5000), (f[150], 40), (f[151], 33), (f[299], 32), (new[0], 700), (new[1], 32), (new[2], 33), (new[3], 1)]
    runs += [(f[k], 1) for k in rng.integers(0, len(f), 200)] + [(new[k], int(rng.integers(1, 4))) for k in range(4, 200)]
    rows = []
    for det, n in runs:
        for _ in range(n):
            pop = int(rng.integers(1, 4)) * (1 if rng.random() < 0.5 else -1)
            rows.append(list(det.view(np.int64)) + [pop, int(rng.random() < 0.5) if initiator else 0])
    sd = np.array(rows, dtype=np.int64)[rng.permutation(len(rows))]
    eng.upload_spawn(sd)
    eng.annihilate_spawn()
    out = eng.annihilate_main(1)
    res = np.zeros(4)
    o.L.orc_rank_annihilate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    assert o.L.orc_rank_annihilate(o.h, 0, np.ascontiguousarray(sd).ctypes.data_as(C.c_void_p), len(sd),
                                   res.ctypes.data_as(C.c_void_p)) == 0
    fo, po, do_ = o.get_psips()
    fg, pg, dg = eng.download_psips()
    assert len(fg) == len(fo) == out["nstates"] and len(fg) > 300
    assert (fg == fo).all() and (pg == po).all() and (dg == do_).all()
    eng.close()


def test_propagator_weight_in_spawn_and_death():
    """hb200_set_propagator_weight (wall-Chebyshev sub-cycle weights) in attempt_to_spawn and stochastic_death"""
    s, o, eng, ref = make_pair("s12", excit_gen="heat_bath", tau=0.002, real=True, initiator=True)
    f, pops, dat = random_population(s, o, 2500, True, seed=4)
    o.set_psips(f, pops, dat)
    eng.upload_psips(f, pops, dat)
    cyc = 1
    for w in (2.5, 0.9, 0.4, 1.0):
        o.set_propagator_weight(w)
        eng.set_propagator_weight(w)
        ro = o.iterate(2, cyc, 0.002, -0.05, -0.1)
        rg = eng.iterate(2, 0.002, -0.05, -0.1, cyc)
        cyc += 2
        assert rg["spawn_error"] == 0 and rg["psip_error"] == 0 and ro["error"] == 0
        fo, po, do_ = o.get_psips()
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo) and (fg == fo).all() and (pg == po).all() and (dg == do_).all(), w
        assert rg["nspawn_events"] == ro["nspawn_events"] and rg["ndeath"] == ro["ndeath"]
    eng.close()


def test_host_built_heat_bath_tables_are_accepted():
    """hb200_set_excit_tables: a host that has already run init_excit_mol_heat_bath hands its excit_gen_heat_bath_t over
    (here: the oracle's tables); the generator must behave exactly as with the tables built on the device."""
    from hande_b200 import read_in as R
    from hande_b200.engine import Engine
    from tests.common import system_path
    s, o, eng, ref = make_pair("s12", excit_gen="heat_bath", tau=0.01, real=True)
    hb = o.heat_bath_tables()
    tabs = dict(i_weights=hb["i_weights"], ij_weights=hb["ij_weights"], ija_weights=hb["ija_w"], ija_aliasU=hb["ija_U"],
                ija_aliasK=hb["ija_K"], ija_weights_tot=hb["ija_tot"], ijab_weights=hb["ijab_w"], ijab_aliasU=hb["ijab_U"],
                ijab_aliasK=hb["ijab_K"], ijab_weights_tot=hb["ijab_tot"])
    e2 = Engine(s, excit_gen="heat_bath", pattempt_single=ref["pattempt_single"], pattempt_double=ref["pattempt_double"],
                real_amplitudes=True, spawn_cutoff=0.01, walker_length=1 << 17, spawned_walker_length=1 << 16, seed=11,
                heat_bath_tables=tabs)
    e2.set_reference(ref["f0"], ref["H00"])
    f, pops, dat = random_population(s, o, 150, True, seed=3)
    att = np.tile(np.arange(6, dtype=np.uint32), len(f))
    ff, pp = np.repeat(f, 6, axis=0), np.repeat(pops, 6)
    a = eng.gen_excit_batch(ff, pp, att, 3, 0.01)
    b = e2.gen_excit_batch(ff, pp, att, 3, 0.01)
    for x, y in zip(a, b):
        assert (x == y).all()
    for k in range(0, len(pp), 7):
        io_o, do_o, ns_o = o.gen_excit_philox(ff[k], 3, int(att[k]), int(pp[k]), 0.01)
        assert io_o[6] == b[0][k, 6] and do_o[0] == b[1][k, 0] and do_o[1] == b[1][k, 1] and ns_o == b[2][k]
    eng.close(); e2.close()


@pytest.mark.parametrize("name,gen,real", [("h2o", "renorm", False), ("s12", "heat_bath", True), ("nh3", "power_pitzer_orderN", True),
                                           ("s10u", "heat_bath_uniform", True), ("ueg14", "no_renorm", True)])
def test_injected_random_numbers(name, gen, real):
    """Level 1 of the correctness contract: the generator and attempt_to_spawn fed an injected list of uniform numbers
    (hb200_gen_excit_batch_rn) against the oracle fed the same list (its list-driven generator is what reproduces the
    reference's golden tables through dSFMT): choice and nspawn exact, pgen and H_ij bit-identical, and the same number
    of random numbers consumed."""
    tau = 0.01
    s, o, eng, ref = make_pair(name, excit_gen=gen, tau=tau, real=real)
    f, pops, dat = random_population(s, o, 300, real, seed=9)
    nrn = 48      # rejection loops of the uniform generators can draw many; a list that is too short continues identically on both sides
    rn = np.random.default_rng(17).random((len(f), nrn))
    io, do, ns, nu = eng.gen_excit_batch_rn(f, pops, rn, tau)
    nallowed = 0
    for k in range(len(f)):
        io_o, do_o, ns_o, k_o = o.gen_excit_spawn_list(f[k], int(pops[k]), tau, rn[k])
        assert io_o[6] == io[k, 6], (k, io_o, io[k])
        if io_o[6]:
            assert list(io_o[:6]) == list(io[k, :6]), (k, io_o, io[k])
            nallowed += 1
        assert do_o[0] == do[k, 0] and do_o[1] == do[k, 1], (k, do_o, do[k])
        assert ns_o == ns[k] and k_o == nu[k], (k, ns_o, ns[k], k_o, nu[k])
    assert nallowed > 100
    eng.close()
