"""Host-side FCIQMC driver: the stand-in for the part of `do_fciqmc` (reference src/fciqmc.f90:12-521) that
stays on the host.  Per report loop it calls the engine once (`hb200_iterate`, ncycles MC cycles on the GPU), then
does what `end_report_loop` does on the CPU: sum the rep_loop buffer over ranks
(`update_energy_estimators`, src/energy_evaluation.F90:126-201), average the estimators, update the shift
(`update_shift`, :659-711), switch on variable shift (:637-651) and print one row in HANDE's table format
(`write_qmc_report`, src/qmc_io.f90:412-508) so pyhande/pyblock parse the output unchanged.

Option names follow the Lua `fciqmc{ qmc = {...} }` table (src/lua_hande_calc.f90:1244-1503).
"""
from __future__ import annotations

import math
import sys as _sys
import time
from dataclasses import dataclass, field

import numpy as np

from . import read_in as _ri
from .engine import Engine
from . import load_balancing as _lb
from . import propagators as _prop
from . import semi_stoch as _ss

HUGE = _ri.HUGE


@dataclass
class QmcIn:
    """qmc_in_t (src/qmc_data.f90:121-330); names = Lua keys."""
    tau: float = 0.001
    rng_seed: int = 7
    init_pop: float = 10.0
    mc_cycles: int = 20
    nreports: int = 10
    target_population: float = 1.0e7
    initial_shift: float = 0.0
    shift_damping: float = 0.05
    vary_shift_from: float = 0.0
    vary_shift_from_proje: bool = False
    initiator: bool = False
    initiator_population: float = 3.0
    real_amplitudes: bool = False
    spawn_cutoff: float = 0.01
    excit_gen: str = "renorm"
    pattempt_single: float = -1.0
    pattempt_double: float = -1.0
    pattempt_parallel: float = -1.0   # renorm_spin / no_renorm_spin; < 0: find_parallel_spin_prob_mol
    quasi_newton: bool = False            # qmc = { quasi_newton = true, ... } (FCIQMC); negative scalars: defaults
    quasi_newton_threshold: float = -1.0
    quasi_newton_value: float = -1.0
    quasi_newton_pop_control: float = -1.0
    pattempt_update: bool = False   # qmc_in%pattempt_update: pattempt_single follows the spawn statistics until the shift varies
    state_size: int = -5            # <0: MB (src/particle_t_utils.f90), >0: elements
    spawned_state_size: int = -1
    ex_level: int = -1              # reference = {ex_level = ...}: truncation level, -1 = none
    nslots: int = 1
    # wall-Chebyshev propagator (qmc_in%chebyshev*, src/qmc_data.f90:244-251) and the harmonic forcing of the shift
    # (src/qmc_data.f90:182-186, src/qmc.F90:195-212)
    chebyshev: bool = False
    chebyshev_order: int = 5
    chebyshev_shift: float = 0.0
    chebyshev_scale: float = 1.1
    chebyshev_skip_gershgorin: bool = False
    shift_harmonic_forcing: float = 0.0
    shift_harmonic_crit_damp: bool = False
    shift_harmonic_forcing_two_stage: bool = False
    # load_bal_in_t (src/qmc_data.f90:491-512) and fciqmc_in%doing_load_balancing; nslots above = load_balancing_slots
    load_balancing: bool = False
    load_balancing_pop: float = 1000.0
    percent_imbal: float = 0.05
    max_load_attempts: int = 2
    # semi_stoch_in_t (src/qmc_data.f90:305-338): space = "high" (the `size` most populated determinants) with the
    # reference's default projection mode (separate annihilation); start at start_iteration, or shift_start_iteration
    # iterations after the shift starts to vary
    semi_stoch_space: str = "none"    # "high" | "ci" (ci_space = { ex_level = semi_stoch_ci_ex_level }) | "read"
    semi_stoch_read_file: str = None  # space = "read": the stored determ%dets (.npy; the reference's SEMI.STOCH.<id>.H5)
    semi_stoch_write_file: str = None # write_determ_space: rank 0 stores determ%dets once the space is built
    semi_stoch_size: int = 0
    semi_stoch_ci_ex_level: int = -1
    semi_stoch_start_iteration: int = 1
    semi_stoch_shift_start_iteration: int = -1
    full_non_composite: bool = False  # ccmc = {full_non_composite = true} (CCMC only)
    reference_det: list = None      # reference = {det = {...}}: explicit reference determinant (1-based orbitals)


def init_propagator(sys, occ0, qmc):
    """init_sp_fock + init_quasi_newton (src/qmc.F90:1064-1160) with calc_fock_values_3d_ueg
    (src/hamiltonian_ueg.f90:299-395).  Returns (sp_fock[0..nbasis], ref_fock_sum, threshold, value, pop_control)."""
    nb, nel = sys.nbasis, sys.nel
    sp = np.zeros(nb + 1)
    sp[1:] = np.asarray(sys.sp_eigv, dtype=np.float64)[1:nb + 1]
    if getattr(sys, "kind", "read_in") == "ueg":
        mad_c = float(np.float32(-2.837297))       # a default-kind (single precision) literal in the reference
        for i in range(1, nb + 1):
            ex = 0.0
            for o in occ0:
                if (o % 2) == (i % 2) and o != i:
                    ex = ex - sys.coulomb_int(o, i)
            mad = mad_c * (0.75 / (math.pi * (sys.rs ** 3) * float(nel))) ** (1.0 / 3.0) if i in occ0 else 0.0
            sp[i] = sp[i] + ex + 0.5 * mad
    ref_fock_sum = 0.0
    for o in occ0:
        ref_fock_sum = ref_fock_sum + sp[o]
    if qmc.quasi_newton_threshold < 0.0:
        thr = sp[nel + 1] - sp[nel]
        if getattr(sys, "kind", "read_in") == "ueg":
            thr = 2.0 * thr
    else:
        thr = qmc.quasi_newton_threshold
    val = thr if qmc.quasi_newton_value < 0.0 else qmc.quasi_newton_value
    pc = 1.0 / thr if qmc.quasi_newton_pop_control < 0.0 else qmc.quasi_newton_pop_control
    return sp, float(ref_fock_sum), float(thr), float(val), float(pc)


class PattemptUpdate:
    """Host half of qmc_in%pattempt_update (src/qmc.F90:1049-1060): the engine accumulates this rank's
    p_single_double_coll_t sums (hb200_set_pattempt / hb200_get_ps_stats); once per report loop this does
    end_report_loop's vary_psingles branch (src/qmc_common.F90:1206-1231): communicate_pattempt_single_data,
    add_rep_accum_to_total and update_pattempt_single (src/spawning.F90:2217-2372)."""
    every_attempts = 10000.0       # src/excit_gens.f90:36-37
    every_min_attempts = 10.0

    def __init__(self, eng, comm, ps, pd, io=None):
        self.eng, self.comm, self.ps, self.pd, self.io = eng, comm, ps, pd, io
        self.total = np.zeros(4)   # h_pgen_singles_sum, excit_gen_singles, h_pgen_doubles_sum, excit_gen_doubles
        self.counter = 1.0
        self.vary_psingles = True
        self.log = []
        eng.set_pattempt(ps, pd, True)

    def end_report_loop(self, vary_shift):
        if not self.vary_psingles:
            return
        if vary_shift:
            self.vary_psingles = False
            self.eng.set_pattempt(self.ps, self.pd, False)
            if self.io is not None and self.comm.rank == 0:
                self.io.write(" # pattempt_single chosen to be: %17.10E\n" % self.ps)
            return
        self.total = self.total + self.comm.allreduce_sum(self.eng.get_ps_stats(reset=True))
        hs, ns, hd, nd = (float(x) for x in self.total)
        if (ns + nd) > self.counter * self.every_attempts and ns > self.counter * self.every_min_attempts \
                and nd > self.counter * self.every_min_attempts:
            self.counter += 1.0
            ps = (hs / ns) / ((hd / nd) + (hs / ns))
            if ps < 1.0 / self.every_attempts:
                ps = 1.0 / self.every_attempts
            pd = 1.0 - ps
            if pd < 1.0 / self.every_attempts:
                pd = 1.0 / self.every_attempts
                ps = 1.0 - pd
            self.ps, self.pd = ps, pd
            self.log.append(ps)
            self.eng.set_pattempt(ps, pd, True)
            if self.io is not None and self.comm.rank == 0:
                self.io.write(" # pattempt_single changed to be: %17.10E\n" % ps)


def murmurhash2(data: bytes, seed: int) -> int:
    """MurmurHash2 (public-domain algorithm; lib/external/MurmurHash2.c) - host copy for the owner of f0."""
    m, r = 0x5BD1E995, 24
    n = len(data)
    h = (seed ^ n) & 0xFFFFFFFF
    i = 0
    while n >= 4:
        k = int.from_bytes(data[i:i + 4], "little")
        k = (k * m) & 0xFFFFFFFF
        k ^= k >> r
        k = (k * m) & 0xFFFFFFFF
        h = (h * m) & 0xFFFFFFFF
        h ^= k
        i += 4
        n -= 4
    if n == 3:
        h ^= data[i + 2] << 16
    if n >= 2:
        h ^= data[i + 1] << 8
    if n >= 1:
        h ^= data[i]
        h = (h * m) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * m) & 0xFFFFFFFF
    h ^= h >> 15
    return h


def owner_of(f, nbasis, nprocs, nslots, proc_map=None, seed=7, slot=False):
    """assign_particle_processor (src/spawning.F90:770-838), shift == 0.  slot=True: the load-balancing slot instead."""
    nbytes = ((nbasis + 31) // 32) * 4
    h = murmurhash2(np.ascontiguousarray(f, dtype=np.uint64).tobytes()[:nbytes], seed)
    if h >= 2**31:
        h -= 2**32
    islot = h % (nprocs * nslots)         # Python % == Fortran modulo for positive divisor
    if slot:
        return islot
    slot = islot
    return (slot % nprocs) if proc_map is None else int(proc_map[slot])


def list_sizes(qmc: QmcIn, W, nprocs):
    """init_particle_t / init_spawn_store sizing (src/particle_t_utils.f90, src/qmc.F90:1418-1505)."""
    size_main = W * 8 + 8 + 4 + 8
    wl = qmc.state_size
    if wl < 0:
        wl = int((-float(wl) * 10**6) / size_main)
    size_sp = (W + 1 + (1 if qmc.initiator else 0)) * 8
    sl = qmc.spawned_state_size
    if sl < 0:
        sl = int((-float(sl) * 10**6) / (2 * size_sp))
    if sl % nprocs != 0:
        sl = int(math.ceil(np.float32(sl) / nprocs)) * nprocs
    return wl, sl


class _SingleProcess:
    """Collective layer for one rank (the serial build of the reference)."""
    rank, size = 0, 1

    def allreduce_sum(self, buf):
        return buf

    def broadcast_bytes(self, arr, src=0):
        return arr

    def allgather_bytes(self, arr):
        return np.asarray(arr, dtype=np.uint8).reshape(1, -1)

    def barrier(self):
        pass


class TorchDist:
    """Collective layer over torch.distributed (gloo on CPU tests, nccl on GPUs) - replaces MPI_Allreduce of rep_loop."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def allreduce_sum(self, buf):
        t = self.torch.tensor(buf, dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def broadcast_bytes(self, arr, src=0):
        t = self.torch.tensor(np.asarray(arr, dtype=np.uint8), device=self.device)
        self.dist.broadcast(t, src=src)
        return t.cpu().numpy()

    def barrier(self):
        self.dist.barrier()

    def allgather_bytes(self, arr):
        """MPI_Allgather of a fixed-size byte record -> [size][len(arr)]"""
        t = self.torch.tensor(np.asarray(arr, dtype=np.uint8), device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(out, t)
        return np.stack([x.cpu().numpy() for x in out])


def _merge_out(o, oc):
    """accumulate the outputs of consecutive hb200_iterate calls of one report loop"""
    if o is None:
        return dict(oc)
    for key in ("proj_energy", "D0_population", "rspawn", "nattempts_spawn", "walker_iterations"):
        o[key] = o[key] + oc[key]
    for key in ("nparticles", "nstates", "nspawn_events", "ndeath", "nattempts"):
        o[key] = oc[key]
    o["spawn_error"] = o["spawn_error"] or oc["spawn_error"]
    o["psip_error"] = o["psip_error"] or oc["psip_error"]
    return o


HEADER = (" #     iterations   Shift                 \\sum H_0j N_j         N_0                   # H psips"
          "                  # states  # spawn_events   R_spawn    time    ")


def format_row(it, shift, pe, d0, npart, nstates, nev, rspawn, t, comment=False):
    """write_qmc_report (src/qmc_io.f90:412-508)."""
    lead = " # " if comment else "   "
    return (f"{lead}{it:14d}  {shift:17.10E}    {pe:18.10E}    {d0:18.10E}    {npart:18.10E}"
            f"  {nstates:17d}  {nev:14d}  {rspawn:8.4f}  {t:8.4f}  ")


@dataclass
class FciqmcResult:
    rows: list = field(default_factory=list)   # iterations, shift, proj_energy, D0, nparticles, nstates, nspawn_events, rspawn
    shift: float = 0.0
    vary_shift: bool = False
    H00: float = 0.0
    occ0: list = None
    engine: object = None
    error: bool = False
    timings: list = field(default_factory=list)
    pattempt_log: list = field(default_factory=list)   # pattempt_single after each pattempt_update change
    load_balancing_log: list = field(default_factory=list)   # (first cycle, proc_map) of every load-balancing step
    chebyshev: object = None                                  # propagators.Chebyshev of the run (None: linear projector)
    determ_space: object = None                               # (determ%dets, determ%sizes) once the semi-stochastic projection is on


def do_fciqmc(sys, qmc: QmcIn, comm=None, device=0, io=None, engine_cls=Engine, keep_engine=False, psips=None):
    """fciqmc{sys=sys, qmc={...}} on the GPU engine.  `psips` = (states, pops, dat) restarts from a given list
    (this rank's share), else the initial distribution puts init_pop on the reference determinant."""
    comm = comm or _SingleProcess()
    nprocs, iproc = comm.size, comm.rank
    out = io if io is not None else None
    # init_reference (src/qmc.F90:1162-1226)
    is_ueg = getattr(sys, "kind", "read_in") == "ueg"
    if qmc.reference_det:
        occ0 = sorted(int(x) for x in qmc.reference_det)
    else:
        occ0 = sys.aufbau_reference() if is_ueg else _ri.set_reference_det(sys)
    f0 = sys.encode(occ0)
    H00 = sys.slater_condon0(occ0)
    # init_excit_gen (src/qmc.F90:910-1010)
    if is_ueg:
        ps, pd = 0.0, 1.0               # find_single_double_prob, src/qmc_common.F90:178-182: doubles only
    elif qmc.pattempt_single < 0 or qmc.pattempt_double < 0:
        ps, pd = _ri.find_single_double_prob(sys, occ0)
    else:
        ps = qmc.pattempt_single / (qmc.pattempt_single + qmc.pattempt_double)
        pd = 1.0 - qmc.pattempt_single
    wl, sl = list_sizes(qmc, sys.W, nprocs)
    # UEG: "renormalised excitation generators not implemented" -> gen_excit_ueg_no_renorm (src/qmc.F90:466-476)
    eng = engine_cls(sys, excit_gen=(("power_pitzer" if qmc.excit_gen == "power_pitzer" else "no_renorm") if is_ueg else qmc.excit_gen), pattempt_single=ps, pattempt_double=pd,
                     real_amplitudes=qmc.real_amplitudes, spawn_cutoff=qmc.spawn_cutoff,
                     initiator_approx=qmc.initiator, initiator_pop=qmc.initiator_population,
                     trunc_level=qmc.ex_level, walker_length=wl, spawned_walker_length=sl, seed=qmc.rng_seed,
                     nprocs=nprocs, iproc=iproc, nslots=qmc.nslots, device=device,
                     pattempt_parallel=qmc.pattempt_parallel)
    eng.set_reference(f0, H00)
    if qmc.quasi_newton:
        eng.set_quasi_newton(*init_propagator(sys, occ0, qmc))
    if nprocs > 1:
        eng.comm_setup(comm)
    real_factor = (1 << 31) if qmc.real_amplitudes else 1
    # initial_distribution (src/qmc.F90:1507-1646)
    if psips is None:
        if owner_of(f0, sys.nbasis, nprocs, qmc.nslots) == iproc:
            eng.upload_psips(f0.reshape(1, -1), np.array([int(round(qmc.init_pop)) * real_factor], dtype=np.int64),
                             np.zeros(1))
            nparticles_loc, d0_loc = float(int(round(qmc.init_pop))), float(int(round(qmc.init_pop)))
        else:
            eng.upload_psips(np.zeros((0, sys.W), dtype=np.uint64), np.zeros(0, dtype=np.int64), np.zeros(0))
            nparticles_loc, d0_loc = 0.0, 0.0
        pe_loc = 0.0
    else:
        eng.upload_psips(*psips)
        nparticles_loc = float(np.abs(psips[1]).sum()) / real_factor
        hit = np.nonzero((np.asarray(psips[0]).reshape(-1, sys.W) == f0).all(axis=1))[0]
        d0_loc = float(psips[1][hit[0]]) / real_factor if len(hit) else 0.0
        pe_loc = 0.0   # initial_ci_projected_energy: first-loop proj_energy_old uses D0 only if pe unknown
    # initial_ci_projected_energy (src/qmc_common.F90:697-797): allreduce of the starting estimators
    buf = comm.allreduce_sum(np.array([pe_loc, d0_loc, nparticles_loc, float(eng.nstates)]))
    proj_energy, D0, ntot_old, tot_nstates = float(buf[0]), float(buf[1]), float(buf[2]), int(round(buf[3]))
    shift, vary_shift = qmc.initial_shift, False
    # init_qmc (src/qmc.F90:195-212): harmonic forcing of the shift
    harmonic = (qmc.shift_damping ** 2) / 4.0 if qmc.shift_harmonic_crit_damp else qmc.shift_harmonic_forcing
    if harmonic != 0.0 and not qmc.shift_harmonic_forcing_two_stage:
        vary_shift = True
    tau = qmc.tau
    cheb = None
    if qmc.chebyshev:
        # init_chebyshev (src/propagators.f90:11-165); the wall-Chebyshev projector is independent of tau: tau = 1
        cheb = _prop.Chebyshev(sys, H00, qmc.chebyshev_order, qmc.chebyshev_shift, qmc.chebyshev_scale,
                               qmc.chebyshev_skip_gershgorin)
        tau = 1.0
    order = cheb.order if cheb else 1
    res = FciqmcResult(H00=H00, occ0=occ0)
    res.chebyshev = cheb
    pupd = PattemptUpdate(eng, comm, ps, pd, io=out) if qmc.pattempt_update else None
    res.rows.append([0, shift, proj_energy, D0, ntot_old, tot_nstates, 0, 0.0])
    if out is not None and iproc == 0:
        out.write(HEADER + "\n")
        out.write(format_row(0, shift, proj_energy, D0, ntot_old, tot_nstates, 0, 0.0, 0.0, comment=True) + "\n")
    mc_cycles_done = 0
    ss_on = (qmc.semi_stoch_space == "high" and qmc.semi_stoch_size > 0) or \
        (qmc.semi_stoch_space == "ci" and qmc.semi_stoch_ci_ex_level >= 0) or \
        (qmc.semi_stoch_space == "read" and qmc.semi_stoch_read_file is not None)
    ss_done = False
    if qmc.semi_stoch_space not in ("none", "high", "ci", "read"):
        raise ValueError("semi_stoch: space must be 'high', 'ci' or 'read'")
    if ss_on and cheb is not None:
        raise ValueError("semi_stoch with the wall-Chebyshev propagator is not supported")
    # shift_start_iteration overrides start_iteration: start_iter = huge(0) until the shift comes on
    # (read_semi_stoch_in, src/lua_hande_calc.f90:1712-1715)
    ss_start = qmc.semi_stoch_start_iteration if qmc.semi_stoch_shift_start_iteration == -1 else 2**31 - 1
    semi_stoch_iter = max(ss_start, mc_cycles_done + 1)                            # src/fciqmc.f90:228
    res.determ_space = None
    lb_needed, lb_attempts = False, 0
    proc_map = [i % nprocs for i in range(nprocs * qmc.nslots)]      # src/load_balancing.F90:170
    for ireport in range(1, qmc.nreports + 1):
        t0 = time.time()
        # get_sanitized_projected_energy (src/energy_evaluation.F90:1433-1456)
        pe_old = 0.0 if abs(D0) < np.finfo(np.float64).tiny else proj_energy / D0
        first_cycle = mc_cycles_done + (ireport - 1) * qmc.mc_cycles + 1
        if lb_needed:
            # load_balancing_wrapper at the first cycle after the report loop that found the imbalance
            # (src/qmc_common.F90:1019-1078): do_load_balancing on the all-reduced slot populations, then
            # redistribute_load_balancing_dets
            slot_list = comm.allreduce_sum(eng.slot_populations())
            needed, new_map, _ = _lb.do_load_balancing(slot_list, proc_map, nprocs, qmc.percent_imbal)
            if needed:
                proc_map = new_map
                lb_attempts += 1
                eng.set_proc_map(proc_map)
                if ss_done:
                    # redistribute_load_balancing_dets (src/qmc_common.F90:1332-1390): the moved determinants are
                    # annihilated without the deterministic flags, then redistribute_semi_stoch_t (:597-650) rebuilds the
                    # space from determ%dets under the new proc_map (recreate_determ_space)
                    all_dets = res.determ_space[0]
                    eng.set_determ_space(all_dets[:0], np.zeros(nprocs, dtype=np.int32))
                eng.redistribute(0x80000000 | first_cycle)
                if ss_done:
                    mine = all_dets[[owner_of(f, sys.nbasis, nprocs, qmc.nslots, proc_map) == iproc for f in all_dets]]
                    res.determ_space = _ss.gather_determ_space(comm, mine.reshape(-1, all_dets.shape[1]))
                    eng.set_determ_space(*res.determ_space)
                res.load_balancing_log.append((first_cycle, list(proc_map)))
            lb_needed = False
        if cheb is None:
            ncyc, cyc0, o = qmc.mc_cycles, first_cycle, None
            if ss_on and not ss_done and cyc0 <= semi_stoch_iter < cyc0 + ncyc:
                # should the semi-stochastic projection start now? (src/fciqmc.f90:300-305): the cycles before it, then
                # init_semi_stoch_t on the list as it stands
                k = semi_stoch_iter - cyc0
                if k > 0:
                    o = eng.iterate(k, tau, shift, pe_old, cyc0)
                    cyc0, ncyc = cyc0 + k, ncyc - k
                res.determ_space = _ss.init_semi_stoch(
                    eng, comm, qmc.semi_stoch_size, space=qmc.semi_stoch_space, sys=sys, occ0=occ0,
                    ci_ex_level=qmc.semi_stoch_ci_ex_level, path=qmc.semi_stoch_read_file,
                    owner=lambda f: owner_of(f, sys.nbasis, nprocs, qmc.nslots, proc_map) == iproc)
                if qmc.semi_stoch_write_file and iproc == 0:
                    _ss.write_determ_to_file(qmc.semi_stoch_write_file, res.determ_space[0])
                ss_done = True
            o = _merge_out(o, eng.iterate(ncyc, tau, shift, pe_old, cyc0))
        else:
            # `order` linear projectors per MC cycle (src/fciqmc.f90:298-299), each a full spawn / annihilation cycle
            # with its own weight; the engine's random stream is keyed by the sub-cycle index
            o = None
            for icycle in range(qmc.mc_cycles):
                for icheb in range(1, order + 1):
                    eng.set_propagator_weight(cheb.weights[icheb - 1])
                    o = _merge_out(o, eng.iterate(1, tau, shift, pe_old, (first_cycle + icycle - 1) * order + icheb))
        res.timings.append(eng.last_timing())
        # local_energy_estimators + MPI_Allreduce + communicated_energy_estimators
        # (src/energy_evaluation.F90:126-201, 320-655)
        loc = np.array([o["proj_energy"], o["D0_population"], o["rspawn"], o["nparticles"], float(o["nstates"]),
                        float(o["nspawn_events"]), float(bool(o["spawn_error"] or o["psip_error"]))])
        tot = comm.allreduce_sum(loc)
        proj_energy = float(tot[0]) / (qmc.mc_cycles * order)
        D0 = float(tot[1]) / (qmc.mc_cycles * order)
        rspawn = float(tot[2]) / (qmc.mc_cycles * order * nprocs)
        ntot = float(tot[3])
        tot_nstates, tot_nev = int(round(tot[4])), int(round(tot[5]))
        error = tot[6] > 0
        if qmc.load_balancing and nprocs > 1 and ntot > qmc.load_balancing_pop and lb_attempts < qmc.max_load_attempts:
            # communicated_energy_estimators (src/energy_evaluation.F90:555-563): check_imbalance on the per-rank populations
            per_rank = comm.allreduce_sum(np.eye(nprocs)[iproc] * o["nparticles"])
            lb_needed = _lb.check_imbalance(per_rank, float(per_rank.sum()) / nprocs, qmc.percent_imbal)
        if vary_shift:
            # update_shift (src/energy_evaluation.F90:659-711)
            shift = shift - math.log(ntot / ntot_old) * qmc.shift_damping / (1.0 * tau * qmc.mc_cycles) \
                - math.log(ntot / qmc.target_population) * harmonic / (1.0 * tau * qmc.mc_cycles) \
                if qmc.target_population > 0 else \
                shift - math.log(ntot / ntot_old) * qmc.shift_damping / (1.0 * tau * qmc.mc_cycles)
        ntot_old = ntot
        if not vary_shift and ntot > qmc.target_population:
            vary_shift = True
            shift = proj_energy / D0 if qmc.vary_shift_from_proje else qmc.vary_shift_from
            if qmc.semi_stoch_shift_start_iteration != -1:       # src/qmc_common.F90:1200-1204
                semi_stoch_iter = qmc.semi_stoch_shift_start_iteration + (mc_cycles_done + ireport * qmc.mc_cycles) + 1
        if pupd is not None:
            pupd.end_report_loop(vary_shift)
        if cheb is not None:
            cheb.update(shift)          # src/fciqmc.f90:427
        it = mc_cycles_done + ireport * qmc.mc_cycles
        res.rows.append([it, shift, proj_energy, D0, ntot, tot_nstates, tot_nev, rspawn])
        if out is not None and iproc == 0:
            out.write(format_row(it, shift, proj_energy, D0, ntot, tot_nstates, tot_nev, rspawn,
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
