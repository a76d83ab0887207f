"""Peer-to-peer spawn exchange on ONE device (-m gpu): two (or three) processes, each an engine rank on cuda:0, run
hb200_iterate with the overlapped exchange - the spawning step launched in chunks (HB200_P2P_MIN_TILES=1 forces the
chunking on these small lists), every chunk's per-destination blocks pushed into the owner rank's receive buffer
through its CUDA IPC mapping, counts kept on the device, one host synchronisation per cycle - and compare every rank's
main list with the oracle's emulated rank after each block of cycles.  NCCL refuses several ranks on one GPU, so the
exchange ends with the host barrier (hb200_set_host_barrier -> gloo); on boxes with >= 2 GPUs tests/test_gpu_multi.py
runs the same worker with the NCCL collective instead."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,gen,real,init,tau,world,n", [("h2o", "renorm", 0, 0, 0.003, 2, 4000),
                                                            ("s12", "heat_bath", 1, 1, 0.004, 3, 4000),
                                                            ("s50", "heat_bath", 1, 1, 2e-5, 2, 6000)])
def test_p2p_exchange_ranks_on_one_device(name, gen, real, init, tau, world, n):
    env = dict(os.environ, HB200_TEST_ONE_DEVICE="1", HB200_P2P_MIN_TILES="1", HB200_TEST_WALKERS=str(n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           name, gen, str(real), str(init), str(tau)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == world
