"""Host-side CCMC driver: the stand-in for the part of `do_ccmc` (reference src/ccmc.f90:271-960) that stays on the
host.  Per report loop it calls the engine once (`hb200_ccmc_iterate`: ncycles cycles of cluster selection, spawning,
death and annihilation on the GPU) and then does `end_report_loop` exactly as the FCIQMC driver does: estimator
averages, shift update, one output row in HANDE's CCMC table format (with the "# attempts" column,
src/qmc_io.f90:412-508).

Scope of this version: stochastic cluster selection (the default) and
`full_non_composite` (not linked, even_selection or multi-reference), real orbitals; any number of ranks (the engine re-hashes and redistributes the
excips every cycle as the reference does, src/qmc_common.F90:505-595).  Option names follow the Lua `ccmc{ qmc = {...},
reference = { ex_level = ... } }` table.
"""
from __future__ import annotations

import math
import time

import numpy as np

from . import read_in as _ri
from .engine import Engine
from .fciqmc import init_propagator, FciqmcResult, PattemptUpdate, QmcIn, _SingleProcess, list_sizes, owner_of

HEADER = (" #     iterations   Shift                 \\sum H_0j N_j         N_0                   # H psips"
          "                  # states  # spawn_events            # attempts   R_spawn    time    ")


def format_row(it, shift, pe, d0, npart, nstates, nev, nattempts, rspawn, t, comment=False):
    lead = " # " if comment else "   "
    return (f"{lead}{it:14d}  {shift:17.10E}    {pe:18.10E}    {d0:18.10E}    {npart:18.10E}"
            f"  {nstates:17d}  {nev:14d}  {nattempts:20d}  {rspawn:8.4f}  {t:8.4f}  ")


def do_ccmc(sys, qmc: QmcIn, comm=None, device=0, io=None, engine_cls=Engine, keep_engine=False):
    """ccmc{sys=sys, qmc={...}, reference={ex_level=qmc.ex_level}} on the GPU engine.  rows: iterations, shift,
    proj_energy, D0_population, nparticles, nstates, nspawn_events, rspawn, nattempts."""
    comm = comm or _SingleProcess()
    nprocs, iproc = comm.size, comm.rank
    if qmc.ex_level < 0:
        raise ValueError("ccmc: reference ex_level (the CC truncation level) must be given")
    if min(sys.nel, qmc.ex_level + 2) > 8:      # HB_MAX_CLUSTER: size of the engine's cluster-selection buffers
        raise ValueError("ccmc: clusters of more than 8 excitors (ex_level + 2 > 8) are not supported by the engine")
    is_ueg = getattr(sys, "kind", "read_in") == "ueg"
    if qmc.reference_det:
        occ0 = sorted(int(x) for x in qmc.reference_det)
    else:
        occ0 = sys.aufbau_reference() if is_ueg else _ri.set_reference_det(sys)
    f0 = sys.encode(occ0)
    H00 = sys.slater_condon0(occ0)
    if is_ueg:
        ps, pd = 0.0, 1.0
    elif qmc.pattempt_single < 0 or qmc.pattempt_double < 0:
        ps, pd = _ri.find_single_double_prob(sys, occ0)
    else:
        ps = qmc.pattempt_single / (qmc.pattempt_single + qmc.pattempt_double)
        pd = 1.0 - qmc.pattempt_single
    wl, sl = list_sizes(qmc, sys.W, nprocs)
    eng = engine_cls(sys, excit_gen=(("power_pitzer" if qmc.excit_gen == "power_pitzer" else "no_renorm") if is_ueg else qmc.excit_gen), pattempt_single=ps, pattempt_double=pd,
                     real_amplitudes=qmc.real_amplitudes, spawn_cutoff=qmc.spawn_cutoff, initiator_approx=False,
                     initiator_pop=qmc.initiator_population, trunc_level=qmc.ex_level, walker_length=wl,
                     spawned_walker_length=sl, seed=qmc.rng_seed, nprocs=nprocs, iproc=iproc, nslots=qmc.nslots,
                     device=device, pattempt_parallel=qmc.pattempt_parallel)
    eng.set_reference(f0, H00)
    if qmc.quasi_newton:      # init_sp_fock / init_quasi_newton (src/qmc.F90:1064-1160), as in do_fciqmc
        eng.set_quasi_newton(*init_propagator(sys, occ0, qmc))
    if qmc.full_non_composite:
        eng.ccmc_set_full_nc(True)
    if nprocs > 1:
        eng.comm_setup(comm, p2p=False)     # CCMC cycles are host-stepped: NCCL send/recv exchange
    real_factor = (1 << 31) if qmc.real_amplitudes else 1
    # initial_distribution + initial_cc_projected_energy (src/qmc_common.F90:799-925): all excips on the reference
    n0 = int(round(qmc.init_pop))
    if owner_of(f0, sys.nbasis, nprocs, qmc.nslots) == iproc:       # hash_shift = 0 at the start: plain owner rule
        eng.upload_psips(f0.reshape(1, -1), np.array([n0 * real_factor], dtype=np.int64), np.zeros(1))
    else:
        eng.upload_psips(np.zeros((0, sys.W), dtype=np.uint64), np.zeros(0, dtype=np.int64), np.zeros(0))
    proj_energy, D0, ntot_old, tot_nstates = 0.0, float(n0), float(n0), 1
    shift, vary_shift = qmc.initial_shift, False
    res = FciqmcResult(H00=H00, occ0=occ0)
    pupd = PattemptUpdate(eng, comm, ps, pd, io=io) if qmc.pattempt_update else None
    res.rows.append([0, shift, proj_energy, D0, ntot_old, tot_nstates, 0, 0.0, n0])
    if io is not None and iproc == 0:
        io.write(HEADER + "\n")
        io.write(format_row(0, shift, proj_energy, D0, ntot_old, tot_nstates, 0, n0, 0.0, 0.0, comment=True) + "\n")
    for ireport in range(1, qmc.nreports + 1):
        t0 = time.time()
        pe_old = 0.0 if abs(D0) < np.finfo(np.float64).tiny else proj_energy / D0
        first_cycle = (ireport - 1) * qmc.mc_cycles + 1
        o = eng.ccmc_iterate(qmc.mc_cycles, qmc.tau, shift, pe_old, first_cycle, qmc.ex_level)
        # local_energy_estimators + MPI_Allreduce + communicated_energy_estimators (src/energy_evaluation.F90:126-655)
        tot = comm.allreduce_sum(np.array([o["proj_energy"], o["D0_population"], o["rspawn"], o["nparticles"],
                                           float(o["nstates"]), float(o["nspawn_events"]), float(o["nattempts"]),
                                           float(bool(o["spawn_error"] or o["psip_error"]))]))
        proj_energy = float(tot[0]) / qmc.mc_cycles
        D0 = float(tot[1]) / qmc.mc_cycles
        rspawn = float(tot[2]) / (qmc.mc_cycles * nprocs)
        ntot = float(tot[3])
        tot_nstates, tot_nev = int(round(tot[4])), int(round(tot[5]))
        o = dict(o, nattempts=int(round(tot[6])))
        error = tot[7] > 0
        if vary_shift:   # update_shift (src/energy_evaluation.F90:659-711)
            shift = shift - math.log(ntot / ntot_old) * qmc.shift_damping / (1.0 * qmc.tau * qmc.mc_cycles) \
                - math.log(ntot / qmc.target_population) * 0.0 / (1.0 * qmc.tau * qmc.mc_cycles)
        ntot_old = ntot
        if not vary_shift and ntot > qmc.target_population:
            vary_shift = True
            shift = proj_energy / D0 if qmc.vary_shift_from_proje else qmc.vary_shift_from
        if pupd is not None:
            pupd.end_report_loop(vary_shift)
        it = ireport * qmc.mc_cycles
        res.rows.append([it, shift, proj_energy, D0, ntot, tot_nstates, tot_nev, rspawn, int(o["nattempts"])])
        if io is not None and iproc == 0:
            io.write(format_row(it, shift, proj_energy, D0, ntot, tot_nstates, tot_nev, int(o["nattempts"]), rspawn,
                                (time.time() - t0) / qmc.mc_cycles) + "\n")
        if error:
            res.error = True
            break
    res.shift, res.vary_shift = shift, vary_shift
    res.pattempt_log = pupd.log if pupd is not None else []
    if keep_engine:
        res.engine = eng
    else:
        eng.close()
    return res
