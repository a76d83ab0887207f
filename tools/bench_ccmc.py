#!/usr/bin/env python
"""Side measurement (not the bench.py headline): CCMC cluster-selection throughput on one B200.

BASELINE.json configs[4]: CCSDT-truncated CCMC on a synthetic 40-orbital / 16-electron FCIDUMP (S40).  A synthetic
excip list (random excitors of level <= 3, |amplitude| = 1 + Exp(1), N/4 excips on the reference) is made resident
and K full cycles (cluster selection + spawning + death + sort + annihilation + merge) are timed.
  python tools/bench_ccmc.py [--excips 1e7] [--steps 10] [--warmup 3] [--ex-level 3]
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
    ap.add_argument("--excips", type=float, default=1e7)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ex-level", type=int, default=3)
    ap.add_argument("--norb", type=int, default=40)
    ap.add_argument("--nelec", type=int, default=16)
    ap.add_argument("--tau", type=float, default=1e-5)
    args = ap.parse_args()
    from hande_b200 import read_in as R
    from hande_b200 import synthetic
    from hande_b200.engine import Engine
    path = os.path.join(tempfile.gettempdir(), f"hande_b200_S{args.norb}_{args.nelec}.fcidump")
    if not os.path.exists(path):
        synthetic.synthetic_fcidump(args.norb, args.nelec, path=path)
    s = R.read_in(path)
    occ0 = R.set_reference_det(s)
    f0 = s.encode(occ0)
    H00 = s.slater_condon0(occ0)
    ps, pd = R.find_single_double_prob(s, occ0)
    n = int(args.excips)
    rf = 1 << 31
    t0 = time.time()
    f = synthetic.random_excitors(n, occ0, s.nbasis, args.ex_level, seed=1)
    rng = np.random.Generator(np.random.Philox(key=5))
    pops = (np.floor((1.0 + rng.exponential(1.0, len(f))) * rf) * np.where(rng.random(len(f)) < 0.5, -1, 1)).astype(np.int64)
    f = np.concatenate([f, f0.reshape(1, -1)])
    pops = np.concatenate([pops, [int(len(f) // 4) * rf]])
    order = np.lexsort(tuple(f[:, k] for k in range(s.W)))
    f, pops = np.ascontiguousarray(f[order]), np.ascontiguousarray(pops[order])
    eng = Engine(s, excit_gen="renorm", pattempt_single=ps, pattempt_double=pd, real_amplitudes=True, spawn_cutoff=0.01,
                 trunc_level=args.ex_level, walker_length=int(len(f) * 1.5) + 4096,
                 spawned_walker_length=max(int(len(f) * 1.0), 1 << 16), seed=7)
    eng.set_reference(f0, H00)
    dat = np.zeros(len(f))
    CH = 5_000_000
    for a in range(0, len(f), CH):
        dat[a:a + CH] = eng.sc0_batch(f[a:a + CH]) - H00
    eng.upload_psips(f, pops, dat)
    setup_s = time.time() - t0
    cyc = 1
    for _ in range(args.warmup):
        o = eng.ccmc_iterate(1, args.tau, 0.0, -0.1, cyc, args.ex_level); cyc += 1
    t0 = time.perf_counter()
    o = eng.ccmc_iterate(args.steps, args.tau, 0.0, -0.1, cyc, args.ex_level)
    dt = time.perf_counter() - t0
    line = {"metric": "CCMC cluster-selection attempts/s (side measurement)", "value": o["nattempts"] * args.steps / dt,
            "unit": "attempts/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps,
            "config": {"workload": f"S{args.norb} synthetic FCIDUMP ({args.norb} orb / {args.nelec} el), CCMC ex_level "
                                   f"{args.ex_level}, renorm, {len(f):.3g} excips", "tau": args.tau,
                       "nattempts_per_cycle": int(o["nattempts"]), "nattempts_spawn_total": int(o["nattempts_spawn"]),
                       "nstates": int(o["nstates"]), "nspawn_events_last": int(o["nspawn_events"])},
            "timing": "wall clock around hb200_ccmc_iterate (the call synchronises every cycle)", "setup_s": setup_s,
            "errors": {"spawn_error": o["spawn_error"], "psip_error": o["psip_error"]}}
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
