#!/usr/bin/env python
"""Side measurement (not the bench.py headline): CCMC cluster-selection throughput on one B200.

BASELINE.json configs[4]: CCSDT-truncated CCMC on a synthetic 40-orbital / 16-electron FCIDUMP (S40).  A synthetic
excip list (random excitors of level <= 3, |amplitude| = 1 + Exp(1), N/4 excips on the reference) is made resident
and K full cycles (cluster selection + spawning + death + sort + annihilation + merge) are timed.
  python tools/bench_ccmc.py [--excips 4e6] [--steps 10] [--warmup 3] [--ex-level 3]
Prints one JSON line: cluster attempts/s, ms per cycle.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--excips", type=float, default=4e6)   # > ~8e6 nears the number of distinct S40 excitors: slow set-up
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ex-level", type=int, default=3)
    ap.add_argument("--norb", type=int, default=40)
    ap.add_argument("--nelec", type=int, default=16)
    ap.add_argument("--tau", type=float, default=1e-5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from hande_b200.fciqmc import TorchDist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchDist(device=torch.device("cuda", local_rank))
    from hande_b200 import read_in as R
    from hande_b200 import synthetic
    from hande_b200.engine import Engine
    path = os.path.join(tempfile.gettempdir(), f"hande_b200_S{args.norb}_{args.nelec}.fcidump")
    if not os.path.exists(path) and rank == 0:
        synthetic.synthetic_fcidump(args.norb, args.nelec, path=path + ".tmp")
        os.replace(path + ".tmp", path)
    if world > 1:
        dist.barrier()
    s = R.read_in(path)
    occ0 = R.set_reference_det(s)
    f0 = s.encode(occ0)
    H00 = s.slater_condon0(occ0)
    ps, pd = R.find_single_double_prob(s, occ0)
    n = int(args.excips)
    rf = 1 << 31
    t0 = time.time()
    # every rank draws its own excitors (--excips per GPU); the first cycle's redistribute_particles moves them to their
    # owners under the time-varying hash (duplicates drawn on two ranks merge there), later cycles move ~1/32 of them
    f = synthetic.random_excitors(n, occ0, s.nbasis, args.ex_level, seed=1 + rank)
    rng = np.random.Generator(np.random.Philox(key=5 + rank))
    pops = (np.floor((1.0 + rng.exponential(1.0, len(f))) * rf) * np.where(rng.random(len(f)) < 0.5, -1, 1)).astype(np.int64)
    from hande_b200.fciqmc import owner_of
    if owner_of(f0, s.nbasis, world, 1) == rank:          # hash_shift = 0 at the start: plain owner of the reference
        f = np.concatenate([f, f0.reshape(1, -1)])
        pops = np.concatenate([pops, [int(len(f) * world // 4) * rf]])
    order = np.lexsort(tuple(f[:, k] for k in range(s.W)))
    f, pops = np.ascontiguousarray(f[order]), np.ascontiguousarray(pops[order])
    eng = Engine(s, excit_gen="renorm", pattempt_single=ps, pattempt_double=pd, real_amplitudes=True, spawn_cutoff=0.01,
                 trunc_level=args.ex_level, walker_length=int(len(f) * (1.5 if world == 1 else 2.5)) + 4096,
                 spawned_walker_length=max(int(len(f) * (1.0 if world == 1 else 2.0)), 1 << 16) * world, seed=7,
                 nprocs=world, iproc=rank, device=local_rank)
    eng.set_reference(f0, H00)
    if world > 1:
        eng.comm_setup(comm, p2p=False)
    dat = np.zeros(len(f))
    CH = 5_000_000
    for a in range(0, len(f), CH):
        dat[a:a + CH] = eng.sc0_batch(f[a:a + CH]) - H00
    eng.upload_psips(f, pops, dat)
    setup_s = time.time() - t0
    cyc = 1
    for _ in range(args.warmup):
        o = eng.ccmc_iterate(1, args.tau, 0.0, -0.1, cyc, args.ex_level); cyc += 1
    if world > 1:
        import torch
        torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    o = eng.ccmc_iterate(args.steps, args.tau, 0.0, -0.1, cyc, args.ex_level)
    if world > 1:
        torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        tot = comm.allreduce_sum(np.array([float(o["nattempts"]), float(o["nstates"]), float(o["nattempts_spawn"]),
                                           float(o["nspawn_events"]), float(o["spawn_error"] + o["psip_error"])]))
        dt = float(np.max(comm.allreduce_sum(np.eye(world)[rank] * dt)))
        o = dict(o, nattempts=int(tot[0]), nstates=int(tot[1]), nattempts_spawn=int(tot[2]), nspawn_events=int(tot[3]),
                 spawn_error=int(tot[4]), psip_error=0)
    if rank != 0:
        eng.close()
        dist.destroy_process_group()
        return
    line = {"metric": "CCMC cluster-selection attempts/s (side measurement)", "value": o["nattempts"] * args.steps / dt,
            "unit": "attempts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps,
            "config": {"workload": f"S{args.norb} synthetic FCIDUMP ({args.norb} orb / {args.nelec} el), CCMC ex_level "
                                   f"{args.ex_level}, renorm, {n:.3g} excips per GPU", "tau": args.tau,
                       "nattempts_per_cycle": int(o["nattempts"]), "nattempts_spawn_total": int(o["nattempts_spawn"]),
                       "nstates": int(o["nstates"]), "nspawn_events_last": int(o["nspawn_events"])},
            "timing": "wall clock around hb200_ccmc_iterate (the call synchronises every cycle)", "setup_s": setup_s,
            "errors": {"spawn_error": o["spawn_error"], "psip_error": o["psip_error"]}}
    print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
