"""Worker for the multi-GPU parity test: launched with torchrun (one rank per GPU).  Every rank owns the
determinants HANDE's hash-owner rule assigns to it, runs MC cycles through hb200_iterate (NCCL all-to-all of the
spawn blocks inside) and compares its list with the oracle's emulated rank of the same index (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main_ccmc():
    """argv: ccmc <system> <generator> <real> <ex_level> <tau>: multi-rank CCMC (time-varying hash owner,
    redistribute_particles, D0 broadcast) against the oracle's emulated ranks."""
    import torch
    import torch.distributed as dist
    from hande_b200 import read_in as R
    from hande_b200.engine import Engine
    from hande_b200.fciqmc import TorchDist
    from oracle.pyoracle import Oracle
    from tests.common import system_path

    name, gen, real, exl, tau = sys.argv[2], sys.argv[3], bool(int(sys.argv[4])), int(sys.argv[5]), float(sys.argv[6])
    full_nc = len(sys.argv) > 7 and bool(int(sys.argv[7]))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = TorchDist(device=dev)
    if rank == 0:
        system_path(name)
    dist.barrier()
    path, kw = system_path(name)
    s = R.read_in(path, **kw)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=tau, seed=11, excit_gen=gen, rng_kind=1, real_amplitudes=int(real), spawn_cutoff=0.01, ex_level=exl,
              D0_population=300, walker_length=1 << 17, spawned_walker_length=1 << 16, nprocs=world)
    o.init()
    ref = o.reference()
    eng = Engine(s, excit_gen=gen, pattempt_single=ref["pattempt_single"], pattempt_double=ref["pattempt_double"],
                 real_amplitudes=real, spawn_cutoff=0.01, trunc_level=exl, walker_length=1 << 17,
                 spawned_walker_length=1 << 16, seed=11, nprocs=world, iproc=rank, device=local)
    eng.set_reference(ref["f0"], ref["H00"])
    o.ccmc_set_full_nc(full_nc)
    eng.ccmc_set_full_nc(full_nc)
    eng.comm_setup(comm, p2p=False)
    # warm-up on the oracle (all ranks emulated) to get a spread-out excip list, then hand each GPU its rank's share
    pe_old = 0.0
    warm = 40
    for c in range(1, warm + 1):
        st, _ = o.ccmc_stage_spawn(c, tau, 0.0, pe_old)
        o.stage_annihilate()
    f, pops, dat = o.get_psips(rank)
    eng.upload_psips(f, pops, dat)
    eng.ccmc_set_hash_shift(o.ccmc_hash_shift(), 5)
    o.set_pattempt_update(True)           # pattempt_update statistics, per rank
    eng.set_pattempt(ref["pattempt_single"], ref["pattempt_double"], True)
    ntot = 0
    for c in range(warm + 1, warm + 9):
        o.ccmc_stage_spawn(c, tau, -0.01, -0.1)
        o.stage_annihilate()
        rg = eng.ccmc_iterate(1, tau, -0.01, -0.1, c, exl)
        assert rg["spawn_error"] == 0 and rg["psip_error"] == 0
        fo, po, do_ = o.get_psips(rank)
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo), (rank, c, len(fg), len(fo))
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), (rank, c)
        ntot = comm.allreduce_sum(np.array([float(len(fg))]))[0]
    assert ntot > 50
    a, b = eng.get_ps_stats(), o.ps_stats(rank)
    assert a[1] == b[1] and a[3] == b[3] and abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and abs(a[2] - b[2]) <= 1e-12 * abs(b[2])
    print(f"rank {rank}: OK {len(fg)} states of {int(ntot)}", flush=True)
    eng.close()
    dist.destroy_process_group()


def main():
    import torch
    import torch.distributed as dist
    from hande_b200 import read_in as R
    from hande_b200.engine import Engine
    from hande_b200.fciqmc import TorchDist, owner_of
    from oracle.pyoracle import Oracle
    from tests.common import system_path, random_population

    name, gen, real, init, tau = sys.argv[1], sys.argv[2], bool(int(sys.argv[3])), bool(int(sys.argv[4])), float(sys.argv[5])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # HB200_TEST_ONE_DEVICE=1: every rank on cuda:0 (NCCL refuses that): gloo for the host collectives, the spawn
    # exchange goes peer-to-peer through CUDA IPC and ends with the host's barrier
    one_device = os.environ.get("HB200_TEST_ONE_DEVICE", "0") == "1"
    if one_device:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_device:
        dist.init_process_group("gloo")
        comm = TorchDist(device=None)
    else:
        dist.init_process_group("nccl", device_id=dev)
        comm = TorchDist(device=dev)
    nwalkers = int(os.environ.get("HB200_TEST_WALKERS", "4000"))
    path, kw = system_path(name) if rank == 0 else (None, None)
    dist.barrier()
    path, kw = system_path(name)
    s = R.read_in(path, **kw)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=tau, seed=11, excit_gen=gen, rng_kind=1, real_amplitudes=int(real), spawn_cutoff=0.01,
              initiator_approx=int(init), literal_event_int32=0, walker_length=1 << 17, spawned_walker_length=1 << 16,
              nprocs=world)
    o.init()
    ref = o.reference()
    eng = Engine(s, excit_gen=gen, pattempt_single=ref["pattempt_single"], pattempt_double=ref["pattempt_double"],
                 real_amplitudes=real, spawn_cutoff=0.01, initiator_approx=init, walker_length=1 << 17,
                 spawned_walker_length=1 << 16, seed=11, nprocs=world, iproc=rank, device=local,
                 pattempt_parallel=(o.pattempt_parallel() if gen.endswith("_spin") else -1.0))
    eng.set_reference(ref["f0"], ref["H00"])
    eng.comm_setup(comm, nccl=not one_device)
    assert eng.p2p or os.environ.get("HB200_NO_P2P", "0") == "1"
    ps_on = gen != "heat_bath"
    if ps_on:
        o.set_pattempt_update(True)
        eng.set_pattempt(ref["pattempt_single"], ref["pattempt_double"], True)
    f, pops, dat = random_population(s, o, nwalkers, real, seed=5)
    own = np.array([owner_of(x, s.nbasis, world, 1) for x in f])
    for r in range(world):
        m = own == r
        o.set_psips(f[m], pops[m], dat[m], rank=r)
        assert all(o.owner(x) == r for x in f[m][:50])
    m = own == rank
    eng.upload_psips(f[m], pops[m], dat[m])
    cyc = 1
    for block in range(3):
        ro = o.iterate(4, cyc, tau, -0.05, -0.1)
        rg = eng.iterate(4, tau, -0.05, -0.1, cyc)
        cyc += 4
        fo, po, do_ = o.get_psips(rank)
        fg, pg, dg = eng.download_psips()
        assert len(fg) == len(fo), (rank, len(fg), len(fo))
        assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), rank
        tot = comm.allreduce_sum(np.array([rg["proj_energy"], rg["D0_population"], float(rg["nspawn_events"]),
                                           float(rg["ndeath"]), float(rg["nstates"])]))
        assert abs(tot[0] - ro["proj_energy"]) <= 1e-12 * max(1.0, abs(ro["proj_energy"]))
        assert abs(tot[1] - ro["D0_population"]) <= 1e-12 * max(1.0, abs(ro["D0_population"]))
        assert tot[2] == ro["nspawn_events"] and tot[3] == ro["ndeath"] and tot[4] == ro["nstates"]
    nss = int(os.environ.get("HB200_TEST_SEMI_STOCH", "0"))
    if nss:
        # semi-stochastic projection across the GPUs: the space is chosen by the host driver code over the real
        # communicator (create_high_pop_space), the deterministic amplitudes are all-gathered with NCCL inside hb200_iterate
        from hande_b200 import semi_stoch as SS
        o.set_semi_stoch(space="high", size=nss)
        o.init_semi_stoch()
        dets_o, sizes_o = o.determ_space()
        dets, sizes = SS.init_semi_stoch(eng, comm, nss)
        assert (sizes == sizes_o).all() and (dets == dets_o).all() and sizes.sum() == nss and (sizes > 0).all()
        for block in range(2):
            ro = o.iterate(4, cyc, tau, -0.05, -0.1)
            rg = eng.iterate(4, tau, -0.05, -0.1, cyc)
            cyc += 4
            fo, po, do_ = o.get_psips(rank)
            fg, pg, dg = eng.download_psips()
            assert len(fg) == len(fo), (rank, len(fg), len(fo))
            assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), rank
            assert (eng.determ_vector(1) == o.determ_vector(rank)[0]).all()
            tot = comm.allreduce_sum(np.array([float(rg["nspawn_events"]), float(rg["ndeath"]), float(rg["nstates"])]))
            assert tot[0] == ro["nspawn_events"] and tot[1] == ro["ndeath"] and tot[2] == ro["nstates"]
    if ps_on:
        a, b = eng.get_ps_stats(), o.ps_stats(rank)
        assert a[1] == b[1] and a[3] == b[3] and abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and abs(a[2] - b[2]) <= 1e-12 * abs(b[2])
    print(f"rank {rank}: OK {len(fg)} states", flush=True)
    eng.close()
    dist.destroy_process_group()


def main_lb():
    """argv: lb <system>: do_fciqmc with load balancing on every rank (NCCL exchange): the imbalance check fires, the
    policy moves slots, redistribute_particles sends the determinants; afterwards every determinant sits on the rank
    the new proc_map names and the total population matches the report row."""
    import torch
    import torch.distributed as dist
    from hande_b200 import read_in as R
    from hande_b200.fciqmc import QmcIn, TorchDist, do_fciqmc, owner_of
    from tests.common import system_path

    name = sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = TorchDist(device=dev)
    if rank == 0:
        system_path(name)
    dist.barrier()
    path, kw = system_path(name)
    s = R.read_in(path, **kw)
    qmc = QmcIn(tau=0.003, init_pop=200, mc_cycles=5, nreports=12, target_population=1e9, real_amplitudes=True,
                excit_gen="renorm", nslots=20, load_balancing=True, load_balancing_pop=500, percent_imbal=0.001,
                max_load_attempts=2, state_size=1 << 17, spawned_state_size=1 << 16, initial_shift=0.3)
    nss = int(os.environ.get("HB200_TEST_SEMI_STOCH", "0"))
    if nss:     # the projection is switched on before the first load-balancing step, which then has to move the space too
        qmc.semi_stoch_space, qmc.semi_stoch_size, qmc.semi_stoch_start_iteration = "high", nss, 8
    res = do_fciqmc(s, qmc, comm=comm, device=local, keep_engine=True)
    assert not res.error
    assert 1 <= len(res.load_balancing_log) <= 2, res.load_balancing_log
    pmap = res.load_balancing_log[-1][1]
    assert pmap != [i % world for i in range(world * 20)]
    f, pops, _ = res.engine.download_psips()
    assert all(owner_of(x, s.nbasis, world, 20, proc_map=pmap) == rank for x in f)
    tot = comm.allreduce_sum(np.array([float(np.abs(pops).sum()) / 2**31, float(len(f))]))
    assert abs(tot[0] - res.rows[-1][4]) < 1e-9 * tot[0] and int(tot[1]) == res.rows[-1][5]
    if nss:
        dets, sizes = res.determ_space
        assert 10 < int(sizes.sum()) <= nss                      # min(target, states in the lists at iteration 8)
        assert res.load_balancing_log[-1][0] > 8                 # a load-balancing step ran with the projection on
        off = int(sizes[:rank].sum())
        mine = dets[off:off + int(sizes[rank])]
        assert all(owner_of(x, s.nbasis, world, 20, proc_map=pmap) == rank for x in mine)   # the space followed the slots
        keys = {tuple(x) for x in f.tolist()}
        assert all(tuple(x) in keys for x in mine.tolist())                                  # and is in the list
        assert len(res.engine.determ_vector(1)) == int(sizes[rank])
    print(f"rank {rank}: OK {len(f)} states after {len(res.load_balancing_log)} load-balancing steps", flush=True)
    res.engine.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "ccmc":
        main_ccmc()
    elif sys.argv[1] == "lb":
        main_lb()
    else:
        main()
