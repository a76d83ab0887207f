#!/usr/bin/env python
"""Benchmark of the FCIQMC walker-propagation hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): real-valued iFCIQMC on the synthetic 8-fold-symmetric 50-orbital / 20-electron
FCIDUMP (S50), heat-bath excitation generator, 1e8 unit-population walkers per GPU (distribution A).
A step = one full MC cycle (spawn + death + estimators + annihilation + merge) over the resident walker list.
value = walker-iterations/s with the list resident in HBM; e2e = the same through the C ABI with the walker list
in pinned HOST memory, uploaded every step (hb200_upload_psips + hb200_iterate + result read-back).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NORB, NELEC = 50, 20
T_HB = 76.0  # heat-bath table bytes per attempt that cannot be L2-resident


def s50_system(norb=NORB, nelec=NELEC):
    from hande_b200 import read_in as R
    from hande_b200 import synthetic
    path = os.path.join(tempfile.gettempdir(), f"hande_b200_S{norb}_{nelec}.fcidump")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}"
        synthetic.synthetic_fcidump(norb, nelec, path=tmp)
        os.replace(tmp, path)
    return R.read_in(path), path


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


UEG = dict(electrons=14, ms=0, rs=1.0, cutoff=2.5)   # --system ueg: 14 electrons, 57 plane waves (114 spin-orbitals)
UEG1000 = dict(electrons=14, ms=0, rs=1.0, cutoff=19.0)   # --system ueg1000: BASELINE configs[3], 1021 plane waves (2042
                                                           # spin-orbitals, 32-word bit strings: the wide list layout)


def oracle_cpu_run(path, n_sample, ncycles, tau, nthreads, seed=1, excit_gen="heat_bath", ueg=None):
    """Time the oracle (CPU restatement of the reference path) on a bounded sample of the workload."""
    from hande_b200 import synthetic
    from oracle.pyoracle import Oracle
    ueg = ueg or UEG
    o = Oracle(wide=(path is None and ueg["cutoff"] > 5.0))
    if path is None:
        o.init_ueg(ueg["electrons"], ueg["ms"], ueg["rs"], ueg["cutoff"])
        excit_gen = "no_renorm"
    else:
        o.read_fcidump(path)
    o.set_qmc(tau=tau, seed=7, excit_gen=excit_gen, rng_kind=1, real_amplitudes=1, initiator_approx=1,
              literal_event_int32=0, walker_length=4 * n_sample, spawned_walker_length=2 * n_sample)
    t0 = time.time()
    o.init()
    t_init = time.time() - t0
    rf = 2**31
    f, pops = synthetic.random_walkers(n_sample, o.nbasis, o.info["nalpha"], o.info["nbeta"], real_factor=rf,
                                       dist="A", seed=seed)
    H00 = o.reference()["H00"]
    dat = np.array([o.sc0(x) - H00 for x in f])
    o.set_psips(f, pops, dat)
    r = o.cpu_baseline(nthreads, ncycles, tau, 0.0, 0.0)
    return r, t_init


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Fortran reference cannot be built
    here (no gfortran/MPI/Lua/HDF5), so this times the oracle port - one emulated MPI rank per host core."""
    if rank != 0:
        return
    s, path = s50_system()
    cores = os.cpu_count() or 1
    n_sample = 200000     # ~2 s timed on 16 cores (tens of core-seconds of CPU work)
    tau = args.tau
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_cpu_run(path, 2000, 1, tau, cores)
    t0 = time.time()
    tot_wi, tot_t = 0.0, 0.0
    nsteps = max(1, min(args.steps, 5))
    r, t_init = oracle_cpu_run(path, n_sample, nsteps, tau, cores)
    tot_wi += r["walker_iters"]; tot_t += r["seconds"]
    v = tot_wi / tot_t
    line = {
        "impl": "reference", "metric": "walker-iterations/s (FCIQMC MC cycles x walkers)", "value": v,
        "unit": "walker-iterations/s", "n_gpus": args.gpus, "steps": nsteps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / nsteps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int64+f64", "data": "synthetic",
        "config": {"workload": "S50 (50 orb / 20 el) real-valued iFCIQMC, heat_bath, unit walkers", "tau": tau},
        "cpu_baseline": {"value": v, "unit": "walker-iterations/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} replicas x {n_sample} walkers x {nsteps} cycles (oracle restatement, "
                                   f"not the Fortran binary; table init {t_init:.1f}s untimed)"},
        "e2e": {"value": v, "unit": "walker-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--walkers", type=float, default=0.0,
                    help="walkers (= occupied determinants) per GPU; 0 = from --scaling")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="auto: 1 GPU = BASELINE configs[1] (1e8 walkers); N > 1 GPUs = configs[2], strong scaling: "
                         "--total-walkers sharded by hash owner over the N ranks.  weak: 1e8 walkers per GPU")
    ap.add_argument("--total-walkers", type=float, default=1e9, help="strong scaling: walkers over all GPUs")
    ap.add_argument("--tau", type=float, default=0.0, help="0 => calibrate for R_spawn ~ 0.05")
    ap.add_argument("--excit-gen", default="heat_bath", choices=["heat_bath", "heat_bath_uniform", "heat_bath_single", "power_pitzer_orderN", "power_pitzer", "renorm", "renorm_spin", "no_renorm_spin", "no_renorm", "power_pitzer_occ",
                                                              "cauchy_schwarz_occ", "power_pitzer_occ_ij", "cauchy_schwarz_occ_ij"],
                    help="excitation generator (headline = heat_bath, BASELINE.json configs[1]; others are side measurements)")
    ap.add_argument("--system", default="s50", choices=["s50", "ueg", "ueg1000"],
                    help="s50 = BASELINE configs[1] (headline); ueg = side measurement on the 3D UEG (14 electrons, "
                         "114 plane-wave spin-orbitals: the largest basis of BASELINE configs[3] this version's W <= 4 holds)")
    ap.add_argument("--calc", default="fciqmc", choices=["fciqmc", "ccmc"],
                    help="ccmc: side line for BASELINE configs[4] (CCSDT-truncated CCMC on the synthetic 40-orbital FCIDUMP): "
                         "cluster-selection attempts/s, see tools/bench_ccmc.py")
    ap.add_argument("--excips", type=float, default=4e6, help="--calc ccmc: excips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--semi-stoch", type=int, default=0,
                    help="side measurement: semi-stochastic projection on, deterministic space = the reference and its first "
                         "N-1 single and double excitations (a connected space, added to the list with zero population)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.calc == "ccmc":
        from tools import bench_ccmc
        sys.argv = [sys.argv[0], "--excips", str(args.excips), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        bench_ccmc.main()
        return
    if args.semi_stoch > 0:
        args.no_e2e = True        # the end-to-end legs re-upload the host's list, which lacks the added deterministic states
    if args.scaling == "auto":
        args.scaling = "strong" if (world > 1 and args.walkers == 0.0) else "weak"
    if args.walkers == 0.0:
        args.walkers = args.total_walkers / world if args.scaling == "strong" else 1e8
    if args.impl == "reference":
        if args.tau == 0.0:
            args.tau = 5.3e-7     # what the GPU arm's calibration (R_spawn ~ 0.05) gives for this workload
        run_reference(args, rank, world)
        return

    import torch
    from hande_b200 import read_in as R
    from hande_b200 import synthetic
    from hande_b200.engine import Engine
    from hande_b200.fciqmc import TorchDist, _SingleProcess

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        comm = TorchDist(device=dev)
    else:
        comm = _SingleProcess()

    t_setup = time.time()
    if args.system in ("ueg", "ueg1000"):
        from hande_b200.ueg import UegSystem
        s, path = UegSystem(**(UEG if args.system == "ueg" else UEG1000)), None
        args.excit_gen = "power_pitzer" if args.excit_gen == "power_pitzer" else "no_renorm"
        occ0 = s.aufbau_reference()
        ps, pd = 0.0, 1.0
    else:
        if rank == 0:
            s, path = s50_system()
        if world > 1:
            torch.distributed.barrier()
        if rank != 0:
            s, path = s50_system()
        occ0 = R.set_reference_det(s)
        ps, pd = R.find_single_double_prob(s, occ0)
    n = int(args.walkers)
    f0 = s.encode(occ0)
    H00 = s.slater_condon0(occ0)
    wl = int(n * 1.25) + 4096
    sl = max(int(n * 0.15), 1 << 16) * 1      # per-destination block: 3x the ~5 % of attempts that spawn
    eng = Engine(s, excit_gen=args.excit_gen, pattempt_single=ps, pattempt_double=pd, real_amplitudes=True,
                 spawn_cutoff=0.01, initiator_approx=True, walker_length=wl, spawned_walker_length=sl * world, seed=7,
                 nprocs=world, iproc=rank, device=local_rank)
    eng.set_reference(f0, H00)
    if world > 1:
        eng.comm_setup(comm)
    rf = 1 << 31
    states, pops = synthetic.random_walkers_torch(n, s.nbasis, s.nalpha, s.nbeta, rf, dev, seed=1, nprocs=world,
                                                  iproc=rank)
    n = len(pops)
    torch.cuda.empty_cache()
    # pinned host copy of the walker list (the host side's particle_t)
    h_states = torch.empty((n, s.W), dtype=torch.int64).pin_memory()
    h_pops = torch.empty(n, dtype=torch.int64).pin_memory()
    h_dat = torch.empty(n, dtype=torch.float64).pin_memory()
    h_states.numpy()[:] = states.view(np.int64)
    h_pops.numpy()[:] = pops
    CH = 10_000_000
    for a in range(0, n, CH):  # dat(1,:) = <D|H|D> - H00 evaluated by the engine (sc0_ptr)
        h_dat.numpy()[a:a + CH] = eng.sc0_batch(states[a:a + CH]) - H00
    del states, pops

    def upload():
        eng.upload_psips_ptr(h_states.data_ptr(), h_pops.data_ptr(), h_dat.data_ptr(), n)

    # tau: calibrate for R_spawn ~ 0.05 (spawn events per attempt) on the first cycle, then restore the list
    tau = args.tau
    if tau == 0.0:
        # R_spawn(tau) saturates (real amplitudes: every allowed attempt spawns once tau*|H|/pgen >= spawn_cutoff),
        # so walk tau down until the response is linear, then scale to the target.
        target, trial = 0.05, 1.0e-4
        for _ in range(8):
            upload()
            o = eng.iterate(1, trial, 0.0, 0.0, 1)
            fr = comm.allreduce_sum(np.array([float(o["nspawn_events"]), float(o["nattempts_spawn"])]))
            r = fr[0] / max(fr[1], 1.0)
            if r < 1e-6:
                trial *= 10.0
            elif r > 2.0 * target:
                trial *= 0.1
            else:
                trial *= target / r
                if abs(r - target) < 0.1 * target:
                    break
        tau = float(f"{trial:.3g}")
    upload()
    ss_info = None
    if args.semi_stoch > 0:
        # semi-stochastic projection (src/semi_stoch.F90): a CISD-like deterministic space shared out by the hash-owner rule
        from hande_b200 import semi_stoch as SS
        from hande_b200.fciqmc import owner_of
        t_ss = time.time()
        occ = list(occ0)
        virt = [o for o in range(1, s.nbasis + 1) if o not in occ]
        space = [np.array(f0, dtype=np.uint64).reshape(-1)]

        def excite(frm, to):
            g = space[0].copy()
            for o in frm:
                g[(o - 1) >> 6] &= ~np.uint64(1 << ((o - 1) & 63))
            for o in to:
                g[(o - 1) >> 6] |= np.uint64(1 << ((o - 1) & 63))
            return g
        for i in occ:
            for a in virt:
                if (i - a) % 2 == 0 and len(space) < args.semi_stoch:
                    space.append(excite((i,), (a,)))
        for ii, i in enumerate(occ):
            for j in occ[ii + 1:]:
                for ia, a in enumerate(virt):
                    for b in virt[ia + 1:]:
                        if len(space) >= args.semi_stoch:
                            break
                        if (i % 2) + (j % 2) == (a % 2) + (b % 2):
                            space.append(excite((i, j), (a, b)))
        space = np.array(space, dtype=np.uint64).reshape(-1, s.W)
        mine = space[[owner_of(x, s.nbasis, world, 1) == rank for x in space]] if world > 1 else space
        dets, sizes = SS.gather_determ_space(comm, mine)
        eng.set_determ_space(dets, sizes)
        cp, _, _ = eng.determ_hamil()
        ss_info = {"size": int(sizes.sum()), "sizes": [int(x) for x in sizes], "nnz_this_rank": int(cp[-1]),
                   "set_space_s": time.time() - t_ss,
                   "space": "reference + its first single and double excitations (spin-conserving), zero population"}
    setup_s = time.time() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    shift, pe_old = 0.0, 0.0
    cyc = 100
    # ---- device-resident measurement
    for _ in range(args.warmup):
        eng.iterate(1, tau, shift, pe_old, cyc); cyc += 1
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.last_timing()["launches"]
    barrier()
    t0 = time.perf_counter()
    out = eng.iterate(args.steps, tau, shift, pe_old, cyc); cyc += args.steps
    barrier()
    wall = time.perf_counter() - t0
    tm = eng.last_timing()
    clocks = sampler.stop() if rank == 0 else None
    dev_s = tm["total_ms"] / 1e3           # CUDA events on the engine stream around the K cycles
    agg = comm.allreduce_sum(np.array([out["walker_iterations"], float(out["nattempts_spawn"]), float(n)]))
    tmax = float(np.max(comm.allreduce_sum(np.eye(world)[rank] * dev_s))) if world > 1 else dev_s
    value = agg[0] / tmax
    attempts_per_s = agg[1] / tmax
    launches = tm["launches"] - l0
    S = out["nstates"]
    A = out["nattempts_spawn"] / args.steps
    P = out["nspawn_events"]
    # ---- roofline of the dominant kernel (k_spawn_death): algorithmic bytes per launch / its launch duration
    Em, Es = 8 * (s.W + 2), 8 * (s.W + 2)
    t_hb = {"heat_bath": T_HB, "heat_bath_uniform": 12.0 + 2 * 16.0}.get(args.excit_gen, 0.0)   # others: L2-resident tables
    alg_bytes = S * Em + S * 8 + A * t_hb + P * Es
    k_ms = tm["spawn_kernel_ms"] / args.steps
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None   # DRAM bytes of one k_spawn_death launch from the committed ncu --set full capture of this workload
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.excit_gen == "heat_bath" and args.system == "s50":
        tj = json.load(open(tp))
        if int(tj.get("walkers_per_gpu", 0)) == int(args.walkers):
            for k, v in tj.get("dram_bytes_per_launch", {}).items():
                if k.startswith("k_spawn_death"):
                    traffic = v
    roofline = {"bound": "hbm", "kernel": "k_spawn_death", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json (ncu --set full capture of this workload, committed; not measured "
                                  "in this run)" if traffic else None,
                "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "stage_ms_per_step": {k: tm[k] / args.steps for k in ("spawn_ms", "comm_ms", "sort_ms", "annihilate_ms")}}
    # whole-cycle algorithmic traffic (SURVEY.md 8d B_alg) for context
    U = P
    b_alg = S * (2 * Em + 8) + S * Em + A * t_hb + P * 6 * Es + U * 2 * Es
    roofline["cycle_alg_bytes"] = b_alg
    roofline["cycle_frac"] = b_alg / (dev_s / args.steps) / 1e9 / peak

    # ---- end-to-end through the C ABI with HOST buffers, three integrations:
    #   e2e            the walker list lives in pinned host memory and a full copy of it is uploaded every step; the
    #                  estimators (result structs) come back; the propagated list stays on the device
    #   e2e_roundtrip  as above and the propagated list is downloaded to pinned host memory every step as well
    #   e2e_resident   the intended integration: list uploaded once, one hb200_iterate(ncycles) per
    #                  report loop, hb200_iter_in in / hb200_iter_out back
    e2e = e2e_rt = e2e_res = None
    if not args.no_e2e:
        nst = max(3, min(args.steps, 5))
        ptrs = (h_states.data_ptr(), h_pops.data_ptr(), h_dat.data_ptr(), n)

        def timed(body, nsteps):
            barrier()
            t0 = time.perf_counter()
            wi = 0.0
            for _ in range(nsteps):
                wi += body()
            barrier()
            dt = time.perf_counter() - t0
            wi_all = comm.allreduce_sum(np.array([wi]))[0]
            dt_max = float(np.max(comm.allreduce_sum(np.eye(world)[rank] * dt))) if world > 1 else dt
            return wi_all / dt_max

        def step_upload():
            # the host's copy of the NEXT list streams to the device (second stream, third buffer) while the current
            # one propagates; every timed step contains one full list upload, one hb200_iterate and the result structs
            nonlocal cyc
            eng.upload_psips_begin_ptr(*ptrs)
            o2 = eng.iterate(1, tau, shift, pe_old, cyc); cyc += 1
            eng.upload_psips_commit()
            return o2["walker_iterations"]

        upload(); eng.iterate(1, tau, shift, pe_old, cyc); cyc += 1
        upload()
        v = timed(step_upload, nst)
        e2e = {"value": v, "unit": "walker-iterations/s", "h2d_bytes_per_step": int(n * Em + 32),
               "d2h_bytes_per_step": 96 + 8 * 12, "steps": nst,
               "note": "upload-only: per step hb200_upload_psips_begin (full walker list from pinned host memory, overlapped "
                       "with the propagation of the resident list) + hb200_iterate(1) + hb200_upload_psips_commit + result "
                       "structs; the propagated list stays on the device (see e2e_roundtrip, e2e_resident)"}
        # round trip: the propagated list also returns to pinned host memory (separate buffers: the input is re-used)
        cap = int(n * 1.1) + 4096
        r_states = torch.empty((cap, s.W), dtype=torch.int64).pin_memory()
        r_pops = torch.empty(cap, dtype=torch.int64).pin_memory()
        r_dat = torch.empty(cap, dtype=torch.float64).pin_memory()
        nback = [0]

        def step_roundtrip():
            nonlocal cyc
            eng.upload_psips_begin_ptr(*ptrs)
            o2 = eng.iterate(1, tau, shift, pe_old, cyc); cyc += 1
            nback[0] = eng.download_psips_ptr(r_states.data_ptr(), r_pops.data_ptr(), r_dat.data_ptr(), cap)
            eng.upload_psips_commit()
            return o2["walker_iterations"]

        upload()
        v = timed(step_roundtrip, nst)
        e2e_rt = {"value": v, "unit": "walker-iterations/s", "h2d_bytes_per_step": int(n * Em + 32),
                  "d2h_bytes_per_step": int(nback[0] * Em + 192), "steps": nst,
                  "note": "per step: full list upload (overlapped) + hb200_iterate(1) + hb200_download_psips of the "
                          "propagated list + result structs"}
        del r_states, r_pops, r_dat
        # resident: one upload, then report loops of ncycles cycles per hb200_iterate call
        ncyc = 5

        def step_resident():
            nonlocal cyc
            o2 = eng.iterate(ncyc, tau, shift, pe_old, cyc); cyc += ncyc
            return o2["walker_iterations"]

        upload()
        v = timed(step_resident, 2)
        e2e_res = {"value": v, "unit": "walker-iterations/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 96,
                   "steps": 2, "cycles_per_call": ncyc,
                   "note": "list resident on the device across report loops: per step one hb200_iterate(ncycles=5) call, "
                           "hb200_iter_in in, hb200_iter_out back (wall clock around the calls); the one-off upload is "
                           "outside the timed region"}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns_cpu = 200000
        r, t_init = oracle_cpu_run(path, ns_cpu, 5, tau, cores, excit_gen=args.excit_gen,
                                   ueg=UEG1000 if args.system == "ueg1000" else UEG)
        cpu = {"value": r["walker_iters"] / r["seconds"], "unit": "walker-iterations/s", "cores": cores, "kind": "port",
               "sample": f"{cores} replicas x {ns_cpu} walkers x 5 cycles of the same {args.system} {args.excit_gen} workload "
                         f"(oracle restatement; {r['seconds']:.1f}s timed, table init {t_init:.1f}s untimed)"}

    if rank == 0:
        line = {
            "metric": "walker-iterations/s (FCIQMC MC cycles x walkers)", "value": value, "unit": "walker-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tmax / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": {"workload": ("S50 synthetic FCIDUMP (50 orb / 20 el, C1, 8-fold)" if args.system == "s50" else
                                    f"3D UEG (14 electrons, rs = 1, {s.nbasis} plane-wave spin-orbitals)") +
                                   f", real-valued iFCIQMC, {args.excit_gen}, {n:.3g} unit walkers per GPU (distribution A)",
                       "tau": tau, "R_spawn": float(P) / max(A, 1.0), "walkers_per_gpu": n, "walkers_total": int(agg[2]),
                       "nstates": int(S),
                       "l2": f"inputs ({n * Em / 1e9:.2g} GB walker list per GPU) " +
                             ("larger than L2; no flush needed" if n * Em > 2.5e8 else "NOT larger than L2 (side measurement)"),
                       "sharding": ("hash-owner (MurmurHash2); spawn blocks pushed peer-to-peer over NVLink while the spawning "
                                    "step runs" + ("" if getattr(eng, "p2p", False) else " [disabled: NCCL send/recv]"))
                                   if world > 1 else "single rank"},
            "spawn_attempts_per_s": attempts_per_s,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_roundtrip": e2e_rt, "e2e_resident": e2e_res,
            "gpu_launches": int(launches), "host_syncs_per_step": 1,
            "clocks": clocks, "wall_s_timed": wall, "setup_s": setup_s,
            "errors": {"spawn_error": out["spawn_error"], "psip_error": out["psip_error"]},
        }
        if ss_info is not None:
            line["semi_stoch"] = ss_info
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
