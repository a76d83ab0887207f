"""Multi-GPU parity (-m gpu, needs >= 2 GPUs): hash-owner sharding + NCCL all-to-all vs the oracle's emulated ranks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("name,gen,real,init,tau", [("h2o", "renorm", 0, 0, 0.003), ("s12", "heat_bath", 1, 1, 0.004),
                                                     ("nh3", "power_pitzer_orderN", 1, 1, 0.002),
                                                     ("nh3", "renorm_spin", 1, 0, 0.002)])
def test_two_rank_parity(name, gen, real, init, tau):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           name, gen, str(real), str(init), str(tau)]
    # HB200_P2P_MIN_TILES=1: the spawning step is launched in chunks (as on bench-sized lists), each chunk pushed
    # peer-to-peer while the next one spawns
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HB200_P2P_MIN_TILES="1"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == nproc


@pytest.mark.parametrize("name,gen,real,init,tau", [("h2o", "renorm", 1, 0, 0.003), ("s12", "heat_bath", 1, 1, 0.004)])
def test_multi_rank_semi_stochastic(name, gen, real, init, tau):
    """Semi-stochastic projection over NCCL (determ_proj_separate_annihil's all-gather, src/semi_stoch.F90:1056-1066):
    after the plain cycles the 80 most populated determinants become the deterministic space and every rank's list
    keeps matching the oracle's emulated rank."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29539", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           name, gen, str(real), str(init), str(tau)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, HB200_P2P_MIN_TILES="1", HB200_TEST_SEMI_STOCH="80"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == nproc


def test_driver_load_balancing():
    """do_fciqmc with load balancing over NCCL: imbalance check, policy, redistribute_particles, direct_annihilation."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           "lb", "h2o"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == nproc
    # the same run with the semi-stochastic projection on: redistribute_semi_stoch_t in the driver
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HB200_TEST_SEMI_STOCH="60"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == nproc


@pytest.mark.parametrize("name,gen,real,exl,tau,full_nc", [("ne_vdz", "renorm", 0, 2, 0.01, 0), ("s12", "renorm", 1, 3, 0.001, 0),
                                                           ("nh3", "heat_bath_uniform", 1, 3, 0.002, 1)])
def test_multi_rank_ccmc_parity(name, gen, real, exl, tau, full_nc):
    """CCMC on 2 (or 4) GPUs: time-varying hash owner, redistribute_particles and the reference broadcast against the
    oracle's emulated ranks (which reproduce the reference's np2 CCMC golden table)."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
           "ccmc", name, gen, str(real), str(exl), str(tau), str(full_nc)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == nproc
