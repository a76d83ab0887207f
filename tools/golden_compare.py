"""Compare the oracle (dSFMT back-end) with a HANDE test_suite benchmark table.

Test-infrastructure helper: reads /root/reference (only available in the build container).
Usage: python tools/golden_compare.py <case> [max_rows]
"""
import re
import sys
import time
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import Oracle, HUGE  # noqa: E402

TS = "/root/reference/test_suite/"
CASES = {
    "h2o": dict(dir="fciqmc/np1/H2O-RHF-cc-pVTZ", bench="benchmark.out.9712b5a3.inp=h2o.in", int_file="INTDUMP",
                sys=dict(nel=10, ms=0, sym=0, cas=(8, 13)),
                qmc=dict(tau=0.003, seed=7, D0_population=10, ncycles=20, nreport=250, target_particles=1e7,
                         walker_length=178571, spawned_walker_length=31250)),
    "ne_init": dict(dir="ifciqmc/np1/Ne-RHF-aug-cc-pVDZ", bench="benchmark.out.9712b5a3.inp=hande_dump.in",
                    int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0),
                    qmc=dict(tau=0.005, seed=7, D0_population=10, ncycles=20, nreport=320, target_particles=80000,
                             initiator_approx=1, walker_length=5 * 10**6 // 32, spawned_walker_length=10**6 // 32)),
    "cn_uhf": dict(dir="fciqmc/np1/CN-UHF-cc-pVDZ", bench="benchmark.out.9712b5a3.inp=cn.in", int_file="INTDUMP",
                   sys=dict(nel=13, ms=1, cas=(9, 12)),
                   qmc=dict(tau=0.005, seed=7, D0_population=10, ncycles=10, nreport=200, target_particles=77000,
                            walker_length=500 * 10**6 // 32, spawned_walker_length=100 * 10**6 // 48)),
    "ne_ci6_np2": dict(dir="fciqmc/np2/Ne-aug-cc-pVDZ-ci6qmc", bench="benchmark.out.9712b5a3.inp=ne.ciqmc.in",
                       int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)),
                       qmc=dict(tau=0.002, seed=18, D0_population=10, ncycles=10, nreport=1200,
                                target_particles=50000, walker_length=50000, spawned_walker_length=5000,
                                ex_level=5, nprocs=2)),
    "ne_ci6_real64_np2": dict(dir="fciqmc_real_64/np2/Ne-aug-cc-pVDZ-ci6qmc_real_64",
                              bench="benchmark.out.9712b5a3.inp=ne.ciqmc.in", int_file="INTDUMP",
                              sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)),
                              qmc=dict(tau=0.002, seed=18, D0_population=10, ncycles=10, nreport=1200,
                                       target_particles=50000, walker_length=50000, spawned_walker_length=50000,
                                       ex_level=5, nprocs=2, real_amplitudes=1, spawn_cutoff=0.01)),
    # amplitudes stored times 2^11 (real_amplitude_force_32): the same run as ne_ci6_real64_np2 at the coarser resolution
    "ne_ci6_real32_np2": dict(dir="fciqmc_real_32/np2/Ne-aug-cc-pVDZ-ci6qmc_real_32",
                              bench="benchmark.out.9712b5a3.inp=ne.ciqmc.in", int_file="INTDUMP",
                              sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)), pop_real_bits=11,
                              qmc=dict(tau=0.002, seed=18, D0_population=10, ncycles=10, nreport=1200,
                                       target_particles=50000, walker_length=50000, spawned_walker_length=50000,
                                       ex_level=5, nprocs=2, real_amplitudes=1, spawn_cutoff=0.01)),
    "ne_ci6_real32_np4": dict(dir="fciqmc_real_32/np4/Ne-aug-cc-pVDZ-ci6qmc_real_32",
                              bench="benchmark.out.9712b5a3.inp=ne.ciqmc.in", int_file="INTDUMP",
                              sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)), pop_real_bits=11,
                              qmc=dict(tau=0.002, seed=18, D0_population=10, ncycles=10, nreport=1200,
                                       target_particles=50000, walker_length=50000, spawned_walker_length=50000,
                                       ex_level=5, nprocs=4, real_amplitudes=1, spawn_cutoff=0.01)),
    # the no_renorm excitation generator in an FCIQMC run (integer walkers, truncation at quadruples, two ranks)
    "ne_cisdtq_no_renorm_np2": dict(dir="fciqmc/np2/Ne-aug-cc-pVDZ-cisdtq_no_renorm",
                                    bench="benchmark.out.9712b5a3.inp=ne.no_renorm.in", int_file="INTDUMP",
                                    sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)),
                                    qmc=dict(tau=0.003, seed=14373, D0_population=100, ncycles=10, nreport=750,
                                             target_particles=50000, walker_length=3571428, spawned_walker_length=1562500,
                                             ex_level=4, nprocs=2, excit_gen="no_renorm")),
    # harmonic forcing of the shift on its own (qmc = { shift_harmonic_forcing = 0.0004, shift_damping = 0.04 }), two ranks
    "n2_harmonic_np2": dict(dir="fciqmc/np2/N2-RHF-STO-3G-harmonic_pop_control",
                            bench="benchmark.out.9712b5a3.inp=n2_sto3g_fciqmc_harmonic_pop_control.in",
                            int_file="n2_sto3g.fcidump", sys=dict(), harmonic_forcing=0.0004,
                            qmc=dict(tau=0.001, seed=21, D0_population=100, ncycles=10, nreport=1000, target_particles=5e6,
                                     shift_damping=0.04, walker_length=35714285, spawned_walker_length=31250000, nprocs=2)),
    "ne_ci6_np4": dict(dir="fciqmc/np4/Ne-aug-cc-pVDZ-ci6qmc", bench="benchmark.out.9712b5a3.inp=ne.ciqmc.in",
                       int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)),
                       qmc=dict(tau=0.002, seed=18, D0_population=10, ncycles=10, nreport=1200,
                                target_particles=50000, walker_length=50000, spawned_walker_length=5000,
                                ex_level=5, nprocs=4)),
    # wall-Chebyshev propagator, order 5, harmonic forcing of the shift (critical damping, two-stage)
    "h4_cheby": dict(dir="fciqmc_real_64/np1/H4-STO-3g_cheby", bench="benchmark.out.9712b5a3.inp=H4.in", int_file="INTDUMP",
                     sys=dict(sym=0), chebyshev=dict(order=5, harmonic_forcing=0.05 ** 2 / 4.0),
                     qmc=dict(tau=1.0, seed=2004313765, D0_population=200, ncycles=1, nreport=30, target_particles=1e5,
                              real_amplitudes=1, spawn_cutoff=0.01, vary_shift_from_proje=1, shift_damping=0.05,
                              walker_length=71428571 // 64, spawned_walker_length=62500000 // 64)),
    # semi-stochastic projection: CI space of the quadruples, separate annihilation, 4 ranks
    "ueg_ss_np4": dict(dir="fciqmc_real_32/np4/ueg_n7_rs1_e1_SS_cisdtq", bench="benchmark.out.9712b5a3.inp=ueg.in",
                       ueg=dict(nel=7, ms=7, rs=1.0, cutoff=1.0), ref_det=[1, 3, 5, 7, 9, 11, 13], vary_shift=True,
                       semi_stoch=dict(space="ci", ci_ex_level=4, pop_real_bits=11),
                       qmc=dict(tau=0.01, seed=7, D0_population=1000, ncycles=10, nreport=200, target_particles=1000,
                                real_amplitudes=1, spawn_cutoff=0.01, walker_length=178571, spawned_walker_length=31250,
                                nprocs=4)),
    # the same projection with the deterministic spawns sent through the spawned list (combined annihilation), the 100
    # most populated determinants, switched on at iteration 1000
    "he2_ss": dict(dir="fciqmc_real_32/np1/He2-aug-cc-pVDZ_real_32_SS", bench="benchmark.out.9712b5a3.inp=he2.in",
                   int_file="INTDUMP", sys=dict(nel=4, ms=0), vary_shift=True,
                   semi_stoch=dict(space="high", size=100, start_iteration=1000, separate_annihilation=False, pop_real_bits=11),
                   qmc=dict(tau=0.01, seed=7, D0_population=1000, ncycles=10, nreport=200, target_particles=1000,
                            real_amplitudes=1, spawn_cutoff=0.01, walker_length=178571, spawned_walker_length=31250)),
    # the same run with the reference's default projection mode (separate annihilation - what the engine implements)
    "he2_ss_sep": dict(dir="fciqmc_real_32/np1/He2-aug-cc-pVDZ_real_32_SS_Extra_MPI", bench="benchmark.out.9712b5a3.inp=he2.in",
                       int_file="INTDUMP", sys=dict(nel=4, ms=0), vary_shift=True,
                       semi_stoch=dict(space="high", size=100, start_iteration=1000, pop_real_bits=11),
                       qmc=dict(tau=0.01, seed=7, D0_population=1000, ncycles=10, nreport=200, target_particles=1000,
                                real_amplitudes=1, spawn_cutoff=0.01, walker_length=178571, spawned_walker_length=31250)),
    # and with the CISD space of the reference determinant (point-group symmetry filter of enumerate_determinants)
    "he2_ss_cisd": dict(dir="fciqmc_real_32/np1/He2-aug-cc-pVDZ_real_32_SS_cisd", bench="benchmark.out.9712b5a3.inp=he2.in",
                        int_file="INTDUMP", sys=dict(nel=4, ms=0), vary_shift=True,
                        semi_stoch=dict(space="ci", ci_ex_level=2, start_iteration=1000, pop_real_bits=11),
                        qmc=dict(tau=0.01, seed=7, D0_population=1000, ncycles=10, nreport=200, target_particles=1000,
                                 real_amplitudes=1, spawn_cutoff=0.01, walker_length=178571, spawned_walker_length=31250)),
    # uniform electron gas: sys = ueg{electrons=6, ms=0, dim=3, cutoff=2, rs=2}, explicit reference
    "ueg_np2": dict(dir="fciqmc/np2/ueg_n10_rs2_e4_fciqmc", bench="benchmark.out.9712b5a3.inp=ueg.fciqmc.in",
                    ueg=dict(nel=6, ms=0, rs=2.0, cutoff=2.0), ref_det=[1, 2, 3, 10, 11, 14],
                    qmc=dict(tau=0.005, seed=122, D0_population=10, ncycles=10, nreport=1000, target_particles=90000,
                             walker_length=50000, spawned_walker_length=5000, nprocs=2)),
    "ueg_np4": dict(dir="fciqmc/np4/ueg_n10_rs2_e4_fciqmc", bench="benchmark.out.9712b5a3.inp=ueg.fciqmc.in",
                    ueg=dict(nel=6, ms=0, rs=2.0, cutoff=2.0), ref_det=[1, 2, 3, 10, 11, 14],
                    qmc=dict(tau=0.005, seed=122, D0_population=10, ncycles=10, nreport=1000, target_particles=90000,
                             walker_length=50000, spawned_walker_length=5000, nprocs=4)),
    # the same UEG with real amplitudes stored times 2^11 (real_amplitude_force_32), two and four ranks
    "ueg_real32_np2": dict(dir="fciqmc_real_32/np2/ueg_n10_rs2_e4_fciqmc_real_32", bench="benchmark.out.9712b5a3.inp=ueg.fciqmc.in",
                           ueg=dict(nel=6, ms=0, rs=2.0, cutoff=2.0), ref_det=[1, 2, 3, 10, 11, 14], pop_real_bits=11,
                           qmc=dict(tau=0.005, seed=122, D0_population=10, ncycles=10, nreport=1000, target_particles=90000,
                                    walker_length=50000, spawned_walker_length=50000, nprocs=2, real_amplitudes=1,
                                    spawn_cutoff=0.01)),
    "ueg_real32_np4": dict(dir="fciqmc_real_32/np4/ueg_n10_rs2_e4_fciqmc_real_32", bench="benchmark.out.9712b5a3.inp=ueg.fciqmc.in",
                           ueg=dict(nel=6, ms=0, rs=2.0, cutoff=2.0), ref_det=[1, 2, 3, 10, 11, 14], pop_real_bits=11,
                           qmc=dict(tau=0.005, seed=122, D0_population=10, ncycles=10, nreport=1000, target_particles=90000,
                                    walker_length=50000, spawned_walker_length=50000, nprocs=4, real_amplitudes=1,
                                    spawn_cutoff=0.01)),
    # quasi-Newton propagator on the same UEG, real amplitudes
    "ueg_qn_real64_np2": dict(dir="fciqmc_real_64/np2/ueg_qn_n10_rs2_e4_fciqmc_real_64", bench="benchmark.out.9712b5a3.inp=ueg.fciqmc.in",
                              ueg=dict(nel=6, ms=0, rs=2.0, cutoff=2.0), ref_det=[1, 2, 3, 10, 11, 14], quasi_newton=dict(threshold=1.0),
                              qmc=dict(tau=0.005, seed=122, D0_population=10, ncycles=10, nreport=1000, target_particles=90000,
                                       walker_length=50000, spawned_walker_length=50000, nprocs=2, real_amplitudes=1,
                                       spawn_cutoff=0.01)),
    # CCMC: CCSD on Ne cc-pVDZ, stochastic cluster selection, integer walkers
    "ccmc_ne": dict(dir="ccmc/np1/Ne-RHF-cc-pVDZ_ccmc", bench="benchmark.out.9712b5a3.inp=ne.ccsdmc.in",
                    int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0), ccmc=True,
                    qmc=dict(tau=0.01, seed=5691, D0_population=10, ncycles=10, nreport=450, target_particles=20000,
                             walker_length=3571428, spawned_walker_length=1562500, ex_level=2)),
    "ccmc_h2o_np2": dict(dir="ccmc/np2/H2O-cc-pVDZ_ccsdmc", bench="benchmark.out.9712b5a3.inp=ccmc.in",
                         int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0), ccmc=True,
                         qmc=dict(tau=0.01, seed=1660032958, D0_population=500, ncycles=10, nreport=175,
                                  target_particles=50000, walker_length=3571428 // 1, spawned_walker_length=1562500,
                                  ex_level=2, nprocs=2)),
    # CCSDT with the quasi-Newton propagator, default and full_non_composite selection (np1, integer walkers)
    "ccmc_h2o_ccsdt_qn": dict(dir="ccmc/np1/H2O-cc-pVDZ_ccsdtmc_qn", bench="benchmark.out.9712b5a3.inp=ccmc.in",
                              int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0), ccmc=True,
                              quasi_newton=dict(threshold=1e-5, value=1.0, pop_control=1.0),
                              qmc=dict(tau=0.008, seed=1660032958, D0_population=500, ncycles=5, nreport=150,
                                       target_particles=70000, walker_length=100 * 10**6 // 28,
                                       spawned_walker_length=50 * 10**6 // 32, ex_level=3)),
    "ccmc_h2o_ccsdt_qn_fullnc": dict(dir="ccmc/np1/H2O-cc-pVDZ_ccsdtmc_qn", bench="benchmark.out.9712b5a3.inp=ccmc_nc.in",
                                     int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0), ccmc=True, full_nc=True,
                                     quasi_newton=dict(threshold=1e-5, value=1.0, pop_control=1.0),
                                     qmc=dict(tau=0.015, seed=1660032958, D0_population=500, ncycles=5, nreport=200,
                                              target_particles=70000, walker_length=100 * 10**6 // 28,
                                              spawned_walker_length=50 * 10**6 // 32, ex_level=3)),
    # CCSDT, full_non_composite = true (select_nc_cluster, stochastic_ccmc_death_nc, deterministic reference attempts)
    "ccmc_h2o_ccsdt_fullnc_np2": dict(dir="ccmc/np2/H2O-cc-pVDZ_ccsdt", bench="benchmark.out.9712b5a3.inp=ccsdt.in",
                                      int_file="INTDUMP", sys=dict(nel=10, ms=0, sym=0, cas=(8, 22)), ccmc=True, full_nc=True,
                                      qmc=dict(tau=0.002, seed=1660032958, D0_population=2000, ncycles=10, nreport=90,
                                               target_particles=62327.28366800, walker_length=3571428,
                                               spawned_walker_length=1562500, ex_level=3, nprocs=2)),
}
# CCSDT on NH3 6-31G*, np4, real amplitudes: one calculation per excitation generator.  Rows are
# comparable until the shift starts to vary (the reference then runs blocking-on-the-fly with auto_shift_damping,
# which is outside the hot path); the generators with pattempt_update are compared up to the first update.
_NH3 = dict(hb=("heat_bath", 0.002, 10000), hb_uni=("heat_bath_uniform", 0.002, 12000),
            hb_single=("heat_bath_single", 0.003, 12000), ppM=("power_pitzer_occ", 0.0015, 14000),
            ppMij=("power_pitzer_occ_ij", 0.0015, 14000), csM=("cauchy_schwarz_occ", 0.0015, 14000),
            csMij=("cauchy_schwarz_occ_ij", 0.0015, 14000), no_renorm=("no_renorm", 0.0007, 12000),
            renorm=("renorm", 0.0007, 12000), renorm_spin=("renorm_spin", 0.0007, 11000),
            no_renorm_spin=("no_renorm_spin", 0.0007, 9000), ppN=("power_pitzer_orderN", 0.001, 10000))
for _k, (_g, _tau, _tp) in _NH3.items():
    CASES["ccmc_nh3_" + _k] = dict(dir="ccmc_real_64/np4/NH3-6-31g_ccsdt_excit_gens",
                                   bench=f"benchmark.out.9712b5a3.inp=nh3.ccsdt.{_k}.in", int_file="INTDUMP",
                                   sys=dict(nel=10, ms=0, sym=0), ccmc=True, until_shift=True,
                                   pattempt_update=_k not in ("hb", "hb_uni"),
                                   qmc=dict(tau=_tau, seed=30513, D0_population=200, ncycles=10, nreport=400,
                                            target_particles=_tp, walker_length=1000000, spawned_walker_length=400000,
                                            ex_level=3, nprocs=4, real_amplitudes=1, spawn_cutoff=0.01, excit_gen=_g))
CASES["ccmc_nh3_no_renorm_spin"]["pattempt_parallel"] = 0.22     # set in the input; renorm_spin computes it (0.22360108)

ROW_CCMC = re.compile(r"^\s*#?\s+(\d+)\s+(-?\d\.\d+E[+-]\d+)\s+(-?\d\.\d+E[+-]\d+)\s+(-?\d\.\d+E[+-]\d+)\s+"
                      r"(-?\d\.\d+E[+-]\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d\.\d+)\s+(\d+\.\d+)\s*$")
ROW = re.compile(r"^\s*#?\s+(\d+)\s+(-?\d\.\d+E[+-]\d+)\s+(-?\d\.\d+E[+-]\d+)\s+(-?\d\.\d+E[+-]\d+)\s+"
                 r"(-?\d\.\d+E[+-]\d+)\s+(\d+)\s+(\d+)\s+(\d\.\d+)\s+(\d+\.\d+)\s*$")


def parse_table(path):
    """rows: iterations, shift, proj_energy, D0, nparticles, nstates, nspawn_events, rspawn[, nattempts (CCMC tables)]"""
    rows = []
    for line in open(path):
        m = ROW.match(line)
        if m:
            rows.append([float(x) for x in m.groups()[:8]])
            continue
        m = ROW_CCMC.match(line)
        if m:
            g = [float(x) for x in m.groups()]
            rows.append(g[:7] + [g[8], g[7]])
    return np.array(rows)


def row_matches(g, r):
    """testcode tolerance is (1e-10 abs, 1e-10 rel) on printed values; the table prints es17.10."""
    def pr(x):
        return float("%.10E" % x)
    ok = (g[0] == r[0] and g[1] == pr(r[1]) and g[2] == pr(r[2]) and g[3] == pr(r[3]) and g[4] == pr(r[4])
          and g[5] == r[5] and g[6] == r[6] and abs(g[7] - r[7]) < 0.6e-4)
    if len(g) > 8:
        ok = ok and g[8] == r[8]
    return ok


def run_case(name, max_rows=None, quiet=False):
    c = CASES[name]
    d = TS + c["dir"] + "/"
    gold = parse_table(d + c["bench"])
    o = Oracle()
    if "ueg" in c:
        o.init_ueg(**c["ueg"])
        o.set_ref_det(c["ref_det"])
    else:
        s = dict(c["sys"])
        o.read_fcidump(d + c["int_file"], nel=s.get("nel", 0), ms=s.get("ms", HUGE), sym=s.get("sym", HUGE),
                       cas=s.get("cas", (-1, -1)))
    q = dict(c["qmc"])
    if c.get("until_shift"):
        q["nreport"] = next((i for i in range(len(gold)) if gold[i][1] != 0.0), len(gold)) - 1
    if max_rows is not None:
        q["nreport"] = min(q["nreport"], max_rows)
    o.set_qmc(**q)
    if "pattempt_parallel" in c:
        o.set_pattempt_parallel(c["pattempt_parallel"])
    if "quasi_newton" in c:
        o.set_quasi_newton(True, **c["quasi_newton"])
    if "semi_stoch" in c:
        o.set_semi_stoch(**c["semi_stoch"])
    elif "pop_real_bits" in c:
        o.set_semi_stoch(space="none", pop_real_bits=c["pop_real_bits"])
    o.init()
    if c.get("vary_shift"):
        o.set_vary_shift(True)
    if "harmonic_forcing" in c:
        o.set_harmonic_forcing(c["harmonic_forcing"])
    t = time.time()
    if c.get("ccmc"):
        o.ccmc_set_full_nc(bool(c.get("full_nc")))
        o.ccmc_set_pattempt_update(bool(c.get("pattempt_update")))
        rows, na = o.run_ccmc()
        rows = np.concatenate([rows, na.reshape(-1, 1).astype(float)], axis=1)
    else:
        rows = o.run()
    dt = time.time() - t
    n = min(len(gold), len(rows))
    if c.get("until_shift"):
        n = min([n] + [i for i in range(n) if gold[i][1] != 0.0 or rows[i][1] != 0.0][:1])
    bad = [i for i in range(n) if not row_matches(gold[i], rows[i])]
    if not quiet:
        print(f"{name}: compared {n} rows, mismatches {len(bad)}, oracle time {dt:.1f}s, ref {o.reference()}")
        for i in bad[:3]:
            print("  gold", gold[i])
            print("  orcl", rows[i])
    return gold, rows, bad


if __name__ == "__main__":
    run_case(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
