"""CPU parity of the engine's __host__ __device__ core (hb_core.cuh, built for the host by tests/hdcheck) against
the oracle: identical Philox stream -> bit-exact excitation choice, pgen, H_ij and nspawn; Slater-Condon rules,
owner hash and the heat-bath generator's normalisation (sum of pgen over all sampled excitations)."""
import numpy as np
import pytest

from hande_b200 import read_in as R
from hande_b200 import synthetic
from oracle.pyoracle import Oracle, philox_stream, EXCIT_GEN
from tests.hdcheck.harness import HdCheck


def _setup(path, kw, excit_gen, tau=0.01, real=False, seed=11):
    s = R.read_in(path, **kw)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=tau, seed=seed, excit_gen=excit_gen, rng_kind=1, real_amplitudes=int(real), spawn_cutoff=0.01)
    o.init()
    ref = o.reference()
    hb = o.heat_bath_tables() if (excit_gen.startswith("heat_bath") or excit_gen.endswith("_ij")) else None
    rf = 2**31 if real else 1
    cutoff = int(np.ceil(0.01 * rf)) if real else 0
    h = HdCheck(s, EXCIT_GEN[excit_gen], ref["pattempt_single"], ref["pattempt_double"], tau, 0.0, 0.0, rf, cutoff,
                seed, ref["f0"], ref["H00"], hb=hb)
    if excit_gen.endswith("_spin"):
        h.set_pattempt_parallel(o.pattempt_parallel())
    if excit_gen == "power_pitzer":
        h.set_power_pitzer(*o.power_pitzer_tables())
    if excit_gen == "power_pitzer_orderN":
        h.set_power_pitzer_orderN(*o.power_pitzer_orderN_tables())
    return s, o, h


def _compare_attempts(s, o, h, dets, pops, tau, ncycle=2, nattempt=6):
    nchecked = 0
    for f, pop in zip(dets, pops):
        for cyc in range(1, 1 + ncycle):
            for att in range(nattempt):
                io_o, do_o, ns_o = o.gen_excit_philox(f, cyc, att, int(pop), tau)
                io_h, do_h, ns_h = h.gen_excit_philox(f, cyc, att, int(pop))
                assert io_o[6] == io_h[6], (f, cyc, att, io_o, io_h)
                if io_o[6]:
                    assert list(io_o[:6]) == list(io_h[:6]), (f, cyc, att, io_o, io_h)
                assert do_o[0] == do_h[0] and do_o[1] == do_h[1], (f, cyc, att, do_o, do_h)   # bit-exact doubles
                assert ns_o == ns_h
                nchecked += 1
    return nchecked


@pytest.mark.parametrize("gen", ["renorm", "no_renorm", "renorm_spin", "no_renorm_spin"])
def test_uniform_generators_h2o(fcidump_path, gen):
    kw = dict(nel=10, ms=0, sym=0, cas=(8, 13))
    s, o, h = _setup(fcidump_path("h2o"), kw, gen, tau=0.003)
    dets = synthetic.random_dets(150, s.nbasis, s.nalpha, s.nbeta, seed=3)
    pops = np.where(np.arange(len(dets)) % 2 == 0, 3, -2)
    n = _compare_attempts(s, o, h, dets, pops, 0.003)
    assert n > 1000
    for f in dets[:40]:
        assert h.sc0(f) == o.sc0(f)
        assert h.murmur(f) == o.murmur_bit_string(f)
        hm, isref = h.proj_hmatel(f)
        assert isref == int((f == o.reference()["f0"]).all())


@pytest.mark.parametrize("gen", ["power_pitzer_occ", "cauchy_schwarz_occ", "power_pitzer_occ_ij", "cauchy_schwarz_occ_ij"])
def test_power_pitzer_occ_generators(fcidump_path, s10, gen):
    """SURVEY 8a row a10 (O(M) variants with uniform ij): bit-exact against the oracle on a D2h molecule (symmetry
    classes of different sizes) and on the C1 synthetic system."""
    kw = dict(nel=10, ms=0, sym=0, cas=(8, 13))
    s, o, h = _setup(fcidump_path("h2o"), kw, gen, tau=0.003)
    dets = synthetic.random_dets(100, s.nbasis, s.nalpha, s.nbeta, seed=3)
    pops = np.where(np.arange(len(dets)) % 2 == 0, 3, -2)
    assert _compare_attempts(s, o, h, dets, pops, 0.003) > 1000
    s, o, h = _setup(s10, {}, gen, tau=0.01, real=True)
    dets = synthetic.random_dets(80, s.nbasis, s.nalpha, s.nbeta, seed=9)
    pops = np.where(np.arange(len(dets)) % 3 == 0, -(2**31), 2**32 + 17)
    assert _compare_attempts(s, o, h, dets, pops, 0.01, ncycle=2, nattempt=6) > 900


@pytest.mark.parametrize("gen", ["power_pitzer_orderN", "power_pitzer"])
def test_power_pitzer_orderN_generator(fcidump_path, s10, gen):
    """SURVEY 8a row a10, the reference-mapped O(N) variant ('heat_bath_power_pitzer_ref'): alias-table look-ups through
    find_diff_ref_cdet; tables from the oracle (pinned on the NH3 ppN golden table)."""
    for path, kw, tau, real in ((fcidump_path("h2o"), dict(nel=10, ms=0, sym=0, cas=(8, 13)), 0.003, False),
                                (fcidump_path("nh3_631g"), dict(nel=10, ms=0, sym=0), 0.002, True), (s10, {}, 0.01, True)):
        s, o, h = _setup(path, kw, gen, tau=tau, real=real)
        dets = synthetic.random_dets(100, s.nbasis, s.nalpha, s.nbeta, seed=3)
        dets[0] = o.reference()["f0"]
        pops = np.where(np.arange(len(dets)) % 2 == 0, 3, -2) * (2**31 if real else 1)
        assert _compare_attempts(s, o, h, dets, pops, tau, nattempt=8) > 1200


def test_ne_large_basis_two_sym(fcidump_path):
    kw = dict(nel=10, ms=0, sym=0)
    s, o, h = _setup(fcidump_path("ne"), kw, "renorm", tau=0.005, real=True)
    dets = synthetic.random_dets(60, s.nbasis, s.nalpha, s.nbeta, seed=5)
    pops = np.full(len(dets), 2**31 + 12345)
    _compare_attempts(s, o, h, dets, pops, 0.005)


@pytest.fixture(scope="module")
def s10(tmp_path_factory):
    p = tmp_path_factory.mktemp("syn") / "s10.fcidump"
    synthetic.synthetic_fcidump(10, 8, path=str(p))
    return str(p)


@pytest.mark.parametrize("gen,real", [("renorm", False), ("no_renorm_spin", True), ("heat_bath", True), ("heat_bath_uniform", True),
                                      ("power_pitzer_orderN", True), ("power_pitzer_occ", False)])
def test_uhf_channels(tmp_path_factory, gen, real):
    """UHF integral store (four two-body channels, spin-resolved one-body terms; src/molecular_integrals.F90:736-845):
    the device core against the oracle on a synthetic UHF FCIDUMP.  The reference ships no UHF integral file in this
    checkout (CN-UHF-cc-pVDZ has none), so both sides are restatements here - parity unpinned for the UHF channels."""
    p = tmp_path_factory.mktemp("syn") / "s10u.fcidump"
    synthetic.synthetic_fcidump_uhf(10, 8, path=str(p))
    s, o, h = _setup(str(p), {}, gen, tau=0.01, real=real)
    assert s.uhf and len(s.v2) == 4
    dets = synthetic.random_dets(100, s.nbasis, s.nalpha, s.nbeta, seed=5)
    pops = np.where(np.arange(len(dets)) % 3 == 0, -(2**31), 2**32 + 17) if real else np.where(np.arange(len(dets)) % 2 == 0, 3, -2)
    n = _compare_attempts(s, o, h, dets, pops, 0.01, ncycle=2, nattempt=6)
    assert n > 1000
    for f in dets[:40]:
        assert h.sc0(f) == o.sc0(f)
        hm, isref = h.proj_hmatel(f)


def test_heat_bath_generator_synthetic(s10):
    s, o, h = _setup(s10, {}, "heat_bath", tau=0.01, real=True)
    dets = synthetic.random_dets(120, s.nbasis, s.nalpha, s.nbeta, seed=9)
    pops = np.where(np.arange(len(dets)) % 3 == 0, -(2**31), 2**32 + 17)
    n = _compare_attempts(s, o, h, dets, pops, 0.01, ncycle=2, nattempt=8)
    assert n > 1500


@pytest.mark.parametrize("nel,ms,rs,cutoff", [(6, 0, 2.0, 2.0), (14, 0, 1.0, 3.0), (5, 1, 0.5, 1.0)])
def test_ueg_generator_and_slater_condon(nel, ms, rs, cutoff):
    """host tables (hande_b200/ueg.py) == oracle tables; gen_excit_ueg_no_renorm, slater_condon0_ueg,
    update_proj_energy_ueg of the engine core bit-exact against the oracle under the same Philox stream."""
    from hande_b200.ueg import UegSystem
    s = UegSystem(nel, ms, rs, cutoff)
    o = Oracle()
    o.init_ueg(nel, ms, rs, cutoff)
    t = o.ueg_tables()
    assert s.nbasis == o.nbasis and (s.kvec[1:] == t["kvec"]).all() and (s.lookup == t["lookup"]).all()
    assert (s.ternary_conserve == t["ternary"]).all() and (s.sp_eigv[1:] == o.basis()["sp_eigv"]).all()
    tau = 0.01
    o.set_qmc(tau=tau, seed=11, rng_kind=1, excit_gen="no_renorm")
    o.init()
    ref = o.reference()
    assert s.slater_condon0(list(ref["occ"])) == ref["H00"]
    h = HdCheck(s, EXCIT_GEN["no_renorm"], 0.0, 1.0, tau, 0.0, 0.0, 1, 0, 11, ref["f0"], ref["H00"])
    dets = synthetic.random_dets(80, s.nbasis, s.nalpha, s.nbeta, seed=3)
    pops = np.where(np.arange(len(dets)) % 2 == 0, 3, -2)
    n = _compare_attempts(s, o, h, dets, pops, tau, ncycle=2, nattempt=6)
    assert n > 900
    for f in dets[:40]:
        assert h.sc0(f) == o.sc0(f)
        assert h.murmur(f) == o.murmur_bit_string(f)
    # projected-energy matrix elements: double excitations of the reference conserve momentum only sometimes
    nz = 0
    occ0 = list(ref["occ"])
    rng = np.random.default_rng(1)
    for _ in range(300):
        i, j = sorted(int(x) for x in rng.choice(occ0, 2, replace=False))
        virt = [v for v in range(1, s.nbasis + 1) if v not in occ0]
        a, b = sorted(int(x) for x in rng.choice(virt, 2, replace=False))
        f = ref["f0"].copy()
        for orb, on in ((i, 0), (j, 0), (a, 1), (b, 1)):
            w, bit = (orb - 1) // 64, np.uint64((orb - 1) % 64)
            f[w] = (f[w] | (np.uint64(1) << bit)) if on else (f[w] & ~(np.uint64(1) << bit))
        hm, isref = h.proj_hmatel(f)
        assert not isref
        assert hm == o.proj_hmatel(f)
        nz += hm != 0.0
    hm, isref = h.proj_hmatel(ref["f0"])
    assert isref and hm == 0.0


@pytest.mark.parametrize("nel,ms,rs,cutoff", [(6, 0, 2.0, 2.0), (7, 1, 1.0, 3.0)])
def test_ueg_power_pitzer_generator(nel, ms, rs, cutoff):
    """gen_excit_ueg_power_pitzer (src/excit_gen_ueg.f90:362-566): (i) the oracle's pgen is the true
    sampling probability - each excitation appears with the frequency it reports and sum pgen + P(null) = 1 - (the
    reference's only fixtures for it use even_selection CCMC); (ii) the engine core is bit-exact against the oracle."""
    from hande_b200.ueg import UegSystem
    s = UegSystem(nel, ms, rs, cutoff)
    o = Oracle()
    o.init_ueg(nel, ms, rs, cutoff)
    tau = 0.01
    o.set_qmc(tau=tau, seed=11, rng_kind=1, excit_gen="power_pitzer")
    o.init()
    ref = o.reference()
    tables, _, _, _ = o.power_pitzer_tables()
    h = HdCheck(s, EXCIT_GEN["power_pitzer"], 0.0, 1.0, tau, 0.0, 0.0, 1, 0, 11, ref["f0"], ref["H00"])
    h.set_ueg_power_pitzer(tables[0])
    dets = synthetic.random_dets(80, s.nbasis, s.nalpha, s.nbeta, seed=3)
    dets[0] = ref["f0"]
    pops = np.where(np.arange(len(dets)) % 2 == 0, 3, -2)
    assert _compare_attempts(s, o, h, dets, pops, tau, ncycle=2, nattempt=6) > 300
    rng = np.random.default_rng(5)
    n = 120000
    for f in [ref["f0"], dets[1]]:
        counts, pg = {}, {}
        nnull = 0
        for _ in range(n):
            io, do, k = o.gen_excit_list(f, rng.random(16))
            if not io[6]:
                nnull += 1
                continue
            key = tuple(io[:5])
            counts[key] = counts.get(key, 0) + 1
            pg[key] = do[0]
        total = sum(pg.values()) + nnull / n
        assert abs(total - 1.0) < 0.02, total
        for key, c in counts.items():
            if pg[key] * n > 400:
                assert abs(c / n - pg[key]) < 5.0 * np.sqrt(pg[key] / n), (key, c / n, pg[key])


def test_heat_bath_uniform_generator_synthetic(s10):
    s, o, h = _setup(s10, {}, "heat_bath_uniform", tau=0.01, real=True)
    dets = synthetic.random_dets(120, s.nbasis, s.nalpha, s.nbeta, seed=9)
    pops = np.where(np.arange(len(dets)) % 3 == 0, -(2**31), 2**32 + 17)
    n = _compare_attempts(s, o, h, dets, pops, 0.01, ncycle=2, nattempt=8)
    assert n > 1500


def test_heat_bath_single_generator(fcidump_path, s10):
    # exact single excitations (weights |<D|H|D_i^a>|), doubles as heat_bath_uniform; h2o has symmetry-forbidden pairs
    s, o, h = _setup(fcidump_path("h2o"), dict(nel=10, ms=0, sym=0, cas=(8, 13)), "heat_bath_single", tau=0.003)
    dets = synthetic.random_dets(60, s.nbasis, s.nalpha, s.nbeta, seed=3)
    assert _compare_attempts(s, o, h, dets, np.where(np.arange(len(dets)) % 2 == 0, 3, -2), 0.003, nattempt=8) > 800
    s, o, h = _setup(s10, {}, "heat_bath_single", tau=0.01, real=True)
    dets = synthetic.random_dets(120, s.nbasis, s.nalpha, s.nbeta, seed=9)
    pops = np.where(np.arange(len(dets)) % 3 == 0, -(2**31), 2**32 + 17)
    n = _compare_attempts(s, o, h, dets, pops, 0.01, ncycle=2, nattempt=8)
    assert n > 1500


def test_slater_condon_two_word_bitstrings(tmp_path):
    # 40 spatial orbitals -> 80 spin orbitals -> W = 2 words
    p = tmp_path / "s40.fcidump"
    synthetic.synthetic_fcidump(40, 10, path=str(p))
    s, o, h = _setup(str(p), {}, "renorm")
    assert s.W == 2
    dets = synthetic.random_dets(40, s.nbasis, s.nalpha, s.nbeta, seed=2)
    rng = np.random.default_rng(0)
    for f in dets:
        assert h.sc0(f) == o.sc0(f)
        occ = s.decode(f)
        virt = [v for v in range(1, s.nbasis + 1) if v not in occ]
        for _ in range(6):
            i = int(rng.choice(occ))
            a = int(rng.choice([v for v in virt if s.ms[v] == s.ms[i]]))
            assert h.sc1(f, i, a) == o.sc1(f, i, a)
            i, j = sorted(int(x) for x in rng.choice(occ, 2, replace=False))
            cand = [(a, b) for a in virt for b in virt if a < b and s.ms[a] + s.ms[b] == s.ms[i] + s.ms[j]]
            a, b = cand[int(rng.integers(len(cand)))]
            assert h.sc2(f, i, j, a, b) == o.sc2(f, i, j, a, b)
        assert h.murmur(f) == o.murmur_bit_string(f)
    _compare_attempts(s, o, h, dets[:20], np.ones(20, dtype=np.int64), 0.01)


def test_philox_stream_matches_oracle():
    f = np.array([0x123456789ABCDEF, 0xFEDCBA98], dtype=np.uint64)
    from tests.hdcheck import harness
    harness.build()
    import ctypes as C
    L = C.CDLL(harness.LIB)
    L.hd_philox_stream.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, C.c_int,
                                   C.c_void_p]
    for purpose in range(5):
        out = np.zeros(9)
        L.hd_philox_stream(77, 123, purpose, f.ctypes.data, 2, 5, 9, out.ctypes.data)
        ref = philox_stream(77, 123, purpose, f, 5, 9)
        assert (out == ref).all()
        assert ((out >= 0) & (out < 1)).all()


@pytest.mark.parametrize("gen", ["power_pitzer", "power_pitzer_orderN", "heat_bath", "heat_bath_uniform", "heat_bath_single", "power_pitzer_occ",
                                 "cauchy_schwarz_occ", "power_pitzer_occ_ij", "cauchy_schwarz_occ_ij"])
def test_heat_bath_pgen_normalisation(s10, gen):
    """SURVEY 8c gap-filler: heat-bath has no single-rank golden trajectory, so pin it statistically.  The
    generator reports pgen for the excitation it produced; over many samples each excitation must appear with
    that frequency, and the reported pgen of all distinct excitations plus the null fraction must sum to one."""
    s = R.read_in(s10)
    o = Oracle()
    o.read_fcidump(s10)
    o.set_qmc(tau=0.01, seed=11, excit_gen=gen, rng_kind=1)
    o.init()
    dets = synthetic.random_dets(3, s.nbasis, s.nalpha, s.nbeta, seed=4)
    rng = np.random.default_rng(5)
    n = 90000
    for f in [o.reference()["f0"], dets[1]]:
        counts, pg = {}, {}
        nnull = 0
        for _ in range(n):
            io, do, k = o.gen_excit_list(f, rng.random(96))
            if not io[6]:
                nnull += 1
                continue
            key = tuple(io[:5])
            counts[key] = counts.get(key, 0) + 1
            if key in pg:
                assert abs(pg[key] - do[0]) <= 4e-16 * do[0]   # function of the excitation only (sum order may differ)
            pg[key] = do[0]
        tot = sum(pg.values()) + nnull / n
        assert abs(tot - 1.0) < 0.02, tot
        for key, c in counts.items():
            e = n * pg[key]
            if e > 50:
                assert abs(c - e) < 5.5 * np.sqrt(e), (key, c, e)


def test_alias_prefix_sum_selection_equals_table_walk():
    """The engine's table-free alias selection (hb_core.cuh select_alias_staged: only the drawn slot's aliasU/aliasK are
    tracked through the stack walk) must return the index of the reference's generate_alias_tables +
    select_weighted_value_precalc (lib/local/alias.f90:68-184) for every weight vector and draw, including exact and
    near ties (7 weight families, draws aimed at table boundaries)."""
    import ctypes as C
    from tests.hdcheck import harness
    harness.build()
    L = C.CDLL(harness.LIB)
    L.hd_alias_selftest.restype = C.c_longlong
    L.hd_alias_selftest.argtypes = [C.c_longlong, C.c_ulonglong, C.c_void_p]
    out = np.zeros(2, dtype=np.int64)
    bad = L.hd_alias_selftest(400000, 7, out.ctypes.data)
    assert bad == 0


def _neighbours(s, f0_occ, rng, n):
    """random determinants within two excitations of random starting points (so that pairs are often connected)"""
    out = []
    occ0 = list(f0_occ)
    for _ in range(n):
        occ = list(occ0)
        for _k in range(int(rng.integers(0, 3))):
            spin = int(rng.integers(0, 2))
            mine = [o for o in occ if o % 2 == spin]
            free = [v for v in range(1, s.nbasis + 1) if v % 2 == spin and v not in occ]
            if mine and free:
                occ.remove(int(rng.choice(mine)))
                occ.append(int(rng.choice(free)))
        out.append(np.asarray(s.encode(sorted(occ)), dtype=np.uint64).reshape(-1))
    return out


@pytest.mark.parametrize("system", ["h2o", "ne", "ueg"])
def test_deterministic_hamiltonian_elements(fcidump_path, system):
    """The element function of the semi-stochastic projection's Hamiltonian (hb_semistoch.cuh ss_hmatel = hmatel_pair +
    the diagonal, compiled for the host) against the oracle's get_hmatel (src/hamiltonian_molecular.f90:12-71,
    src/hamiltonian_ueg.f90:12-69) on pairs of determinants zero to four excitations apart, bit for bit; and
    check_if_determ's bisection against set membership"""
    if system == "ueg":
        from hande_b200.ueg import UegSystem
        s = UegSystem(6, 0, 2.0, 2.0)
        o = Oracle()
        o.init_ueg(6, 0, 2.0, 2.0)
        o.set_qmc(tau=0.01, seed=11, excit_gen="no_renorm", rng_kind=1)
        o.init()
        ref = o.reference()
        h = HdCheck(s, EXCIT_GEN["no_renorm"], 0.0, 1.0, 0.01, 0.0, 0.0, 1, 0, 11, ref["f0"], ref["H00"])
    else:
        kw = dict(nel=10, ms=0, sym=0, cas=(8, 13)) if system == "h2o" else dict(nel=10, ms=0, sym=0, cas=(8, 22))
        s, o, h = _setup(fcidump_path(system), kw, "renorm")
        ref = o.reference()
    rng = np.random.default_rng(7)
    dets = _neighbours(s, ref["occ"], rng, 120)
    nz = diag = 0
    for f1 in dets[:60]:
        for f2 in dets[60:] + [f1]:
            a = h.hmatel_pair(f1, f2)
            b = o.get_hmatel(f1, f2)
            if (f1 == f2).all():
                b = b - ref["H00"]
                diag += 1
            assert a == b, (f1, f2, a, b)
            nz += a != 0.0
    assert nz > (40 if system == "ueg" else 200) and diag >= 60      # the UEG conserves momentum: most pairs vanish
    uniq = sorted({tuple(int(x) for x in f[::-1]) for f in dets})          # list order: last word most significant
    sd = np.array([t[::-1] for t in uniq], dtype=np.uint64)
    members = {tuple(int(x) for x in f) for f in sd}
    for f in dets + _neighbours(s, ref["occ"], rng, 60):
        assert h.check_if_determ(sd, f) == (tuple(int(x) for x in f) in members)
