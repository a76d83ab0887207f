"""Shared helpers for the parity tests (test infrastructure: may use oracle/)."""
import gzip
import os

import numpy as np

from hande_b200 import read_in as R
from hande_b200 import synthetic
from oracle.pyoracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_TMP = os.path.join("/tmp", "hande_b200_tests")

SYSTEMS = {
    "h2o": dict(fcidump="h2o", kw=dict(nel=10, ms=0, sym=0, cas=(8, 13))),
    "ne": dict(fcidump="ne", kw=dict(nel=10, ms=0, sym=0)),
    "ne_vdz": dict(fcidump="ne_vdz", kw=dict(nel=10, ms=0, sym=0)),   # the reference's CCMC fixture (28 spin-orbitals, D2h)
    "nh3": dict(fcidump="nh3_631g", kw=dict(nel=10, ms=0, sym=0)),    # the reference's per-generator CCSDT fixture (C3v in Cs)
    "ne_cas": dict(fcidump="ne", kw=dict(nel=10, ms=0, sym=0, cas=(8, 22))),
    "s10": dict(synthetic=(10, 8), kw={}),
    "s12": dict(synthetic=(12, 8), kw={}),
    "s10u": dict(synthetic_uhf=(10, 8), kw={}),   # UHF: four two-body channels, spin-dependent one-body terms
    "s40": dict(synthetic=(40, 10), kw={}),    # W = 2
    "s50": dict(synthetic=(50, 20), kw={}),    # the bench system (BASELINE configs[1]): 100 spin-orbitals, W = 2, 20 electrons
    # uniform electron gas: (electrons, ms, rs, cutoff)
    "ueg6": dict(ueg=(6, 0, 2.0, 2.0)),        # the reference's np2/np4 fixture system, 66 spin-orbitals (W = 2)
    "ueg14": dict(ueg=(14, 0, 1.0, 4.0)),      # 186 spin-orbitals (W = 3)
    # wide layout (more than 254 spin-orbitals: 32-word device lists, 16-bit occupied lists, compressed-key sort)
    "ueg358": dict(ueg=(14, 0, 1.0, 6.0), wide=True),     # 358 spin-orbitals (host W = 6)
    "ueg2042": dict(ueg=(14, 0, 1.0, 19.0), wide=True),   # BASELINE configs[3]: 14 electrons, 1021 plane waves (W = 32)
}


def system_path(name):
    os.makedirs(_TMP, exist_ok=True)
    spec = SYSTEMS[name]
    path = os.path.join(_TMP, name + ".fcidump")
    if not os.path.exists(path):
        if "fcidump" in spec:
            with gzip.open(os.path.join(GOLDEN, "fcidump", spec["fcidump"] + ".INTDUMP.gz"), "rt") as fi:
                txt = fi.read()
            with open(path, "w") as fo:
                fo.write(txt)
        elif "synthetic_uhf" in spec:
            synthetic.synthetic_fcidump_uhf(*spec["synthetic_uhf"], path=path)
        else:
            synthetic.synthetic_fcidump(*spec["synthetic"], path=path)
    return path, spec["kw"]


def make_pair(name, *, excit_gen="renorm", tau=0.01, seed=11, real=False, initiator=False, ex_level=-1,
              walker_length=1 << 17, spawned_walker_length=1 << 16, engine=True, device=0, quasi_newton=None):
    """Host system + oracle (Philox stream, symmetric initiator event rule) + GPU engine with identical options."""
    o = Oracle(wide=SYSTEMS[name].get("wide", False))
    if "ueg" in SYSTEMS[name]:
        from hande_b200.ueg import UegSystem
        s = UegSystem(*SYSTEMS[name]["ueg"])
        o.init_ueg(*SYSTEMS[name]["ueg"])
        excit_gen = "power_pitzer" if excit_gen == "power_pitzer" else "no_renorm"
    else:
        path, kw = system_path(name)
        s = R.read_in(path, **kw)
        o.read_fcidump(path, **kw)
    o.set_qmc(tau=tau, seed=seed, excit_gen=excit_gen, rng_kind=1, real_amplitudes=int(real), spawn_cutoff=0.01,
              initiator_approx=int(initiator), ex_level=ex_level, literal_event_int32=0, walker_length=walker_length,
              spawned_walker_length=spawned_walker_length)
    if quasi_newton is not None:
        o.set_quasi_newton(True, **quasi_newton)
    o.init()
    ref = o.reference()
    eng = None
    if engine:
        from hande_b200.engine import Engine
        eng = Engine(s, excit_gen=excit_gen, pattempt_single=ref["pattempt_single"],
                     pattempt_double=ref["pattempt_double"], real_amplitudes=real, spawn_cutoff=0.01,
                     initiator_approx=initiator, trunc_level=ex_level, walker_length=walker_length,
                     spawned_walker_length=spawned_walker_length, seed=seed, device=device,
                     pattempt_parallel=(o.pattempt_parallel() if excit_gen.endswith("_spin") else -1.0))
        eng.set_reference(ref["f0"], ref["H00"])
        if quasi_newton is not None:
            q = o.quasi_newton()
            eng.set_quasi_newton(q["sp_fock"], q["ref_fock_sum"], q["threshold"], q["value"], q["pop_control"])
    return s, o, eng, ref


def random_population(s, o, n, real, seed=1, dist="B", include_ref=True):
    """Random sorted walker list with true diagonal elements; optionally forces the reference into the list."""
    rf = 2**31 if real else 1
    f, pops = synthetic.random_walkers(n, s.nbasis, s.nalpha, s.nbeta, real_factor=rf, dist=dist, seed=seed)
    if include_ref:
        f0 = o.reference()["f0"]
        if not (f == f0).all(axis=1).any():
            f = synthetic.sort_dets(np.concatenate([f, f0.reshape(1, -1)]))
            rng = np.random.default_rng(seed)
            pops = np.concatenate([pops, [7 * rf]])
            pops = pops[rng.permutation(len(pops))]
    H00 = o.reference()["H00"]
    dat = np.array([o.sc0(x) - H00 for x in f])
    return f, pops.astype(np.int64), dat


def sort_rows(a):
    a = np.asarray(a)
    if len(a) == 0:
        return a
    order = np.lexsort(tuple(a[:, k] for k in range(a.shape[1] - 1, -1, -1)))
    return a[order]
