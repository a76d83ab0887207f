"""Multi-rank parity on ONE device (-m gpu): two engines with nprocs = 2 (iproc 0 and 1) share cuda:0.  NCCL refuses two
ranks on one GPU, so the test stages comm_spawn_t's exchange itself through the C ABI - hb200_spawn_counts +
hb200_download_spawn on the sender (the per-destination blocks of spawn%sdata, src/spawn_data.F90:686-721),
hb200_upload_spawn on the receiver, receive blocks ordered by source rank as MPI_Alltoallv leaves them - and every other
stage (hash-owner rule in the spawn kernel, per-destination append, sort, annihilation, merge) runs exactly as in a
multi-GPU run.  Each rank's main list is compared with the oracle's emulated rank after every cycle.  The NCCL
exchange itself is covered by tests/test_gpu_multi.py on boxes with >= 2 GPUs."""
import numpy as np
import pytest

from hande_b200 import read_in as R
from hande_b200.engine import Engine
from hande_b200.fciqmc import owner_of
from oracle.pyoracle import Oracle
from tests.common import random_population, system_path

pytestmark = pytest.mark.gpu


def _setup(name, gen, real, init, tau, world, nwalkers, nslots=1, skew=None):
    path, kw = system_path(name)
    s = R.read_in(path, **kw)
    o = Oracle()
    o.read_fcidump(path, **kw)
    o.set_qmc(tau=tau, seed=11, excit_gen=gen, rng_kind=1, real_amplitudes=int(real), spawn_cutoff=0.01,
              initiator_approx=int(init), literal_event_int32=0, walker_length=1 << 17, spawned_walker_length=1 << 16,
              nprocs=world, nslots=nslots)
    o.init()
    ref = o.reference()
    engs = []
    for r in range(world):
        e = Engine(s, excit_gen=gen, pattempt_single=ref["pattempt_single"], pattempt_double=ref["pattempt_double"],
                   real_amplitudes=real, spawn_cutoff=0.01, initiator_approx=init, walker_length=1 << 17,
                   spawned_walker_length=1 << 16, seed=11, nprocs=world, iproc=r, nslots=nslots, device=0,
                   pattempt_parallel=(o.pattempt_parallel() if gen.endswith("_spin") else -1.0))
        e.set_reference(ref["f0"], ref["H00"])
        engs.append(e)
    f, pops, dat = random_population(s, o, nwalkers, real, seed=5)
    own = np.array([owner_of(x, s.nbasis, world, nslots) for x in f])
    if skew is not None:       # a population imbalance between the slots, for the load-balancing test
        slot = np.array([owner_of(x, s.nbasis, world, nslots, slot=True) for x in f])
        pops = np.where(np.isin(slot, skew), pops * 6, pops)
    for r in range(world):
        m = own == r
        assert m.sum() > nwalkers // (3 * world)
        o.set_psips(f[m], pops[m], dat[m], rank=r)
        engs[r].upload_psips(f[m], pops[m], dat[m])
    return s, o, engs


@pytest.mark.parametrize("name,gen,real,init,tau,world,n", [("h2o", "renorm", False, False, 0.003, 2, 4000),
                                                            ("s12", "heat_bath", True, True, 0.004, 2, 4000),
                                                            ("nh3", "renorm_spin", True, False, 0.002, 3, 4000),
                                                            ("s50", "heat_bath", True, True, 2e-5, 2, 6000)])
def test_ranks_on_one_device_match_oracle_ranks(name, gen, real, init, tau, world, n):
    s, o, engs = _setup(name, gen, real, init, tau, world, n)
    E = s.W + 2
    shift, pe_old = -0.05, -0.1
    for cycle in range(1, 7):
        ro = o.iterate(1, cycle, tau, shift, pe_old)
        blocks, stats = [], []
        for e in engs:
            stats.append(e.spawn_death(tau, shift, pe_old, cycle))
            cnt = e.spawn_counts()
            sd = e.download_spawn()
            assert cnt.sum() == len(sd) == stats[-1]["nspawn_events"]
            off = np.concatenate([[0], np.cumsum(cnt)])
            blocks.append([sd[off[d]:off[d + 1]] for d in range(world)])
            # every element sits in the block of the rank that owns it
            for d in range(world):
                assert all(owner_of(x, s.nbasis, world, 1) == d for x in blocks[-1][d][:200, :s.W].astype(np.uint64))
        tot = np.zeros(5)
        for d, e in enumerate(engs):
            recv = np.concatenate([blocks[src][d] for src in range(world)]).reshape(-1, E)
            e.upload_spawn(recv)
            e.annihilate_spawn()
            out = e.annihilate_main(cycle)
            fo, po, do_ = o.get_psips(d)
            fg, pg, dg = e.download_psips()
            assert len(fg) == len(fo) == out["nstates"], (cycle, d, len(fg), len(fo))
            assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), (cycle, d)
            tot += [stats[d]["proj_energy"], stats[d]["D0_population"], stats[d]["nspawn_events"], stats[d]["ndeath"],
                    out["nstates"]]
        assert abs(tot[0] - ro["proj_energy"]) <= 1e-12 * max(1.0, abs(ro["proj_energy"]))
        assert abs(tot[1] - ro["D0_population"]) <= 1e-12 * max(1.0, abs(ro["D0_population"]))
        assert tot[2] == ro["nspawn_events"] and tot[3] == ro["ndeath"] and tot[4] == ro["nstates"]
    assert sum(e.nstates for e in engs) > n // 2
    for e in engs:
        e.close()


def _exchange_and_annihilate(s, o, engs, cycle):
    """comm_spawn_t staged through the C ABI (see the module docstring), then annihilation + merge on every rank."""
    world, E = len(engs), s.W + 2
    blocks = []
    for e in engs:
        cnt = e.spawn_counts()
        sd = e.download_spawn()
        assert cnt.sum() == len(sd)
        off = np.concatenate([[0], np.cumsum(cnt)])
        blocks.append([sd[off[d]:off[d + 1]] for d in range(world)])
    outs = []
    for d, e in enumerate(engs):
        recv = np.concatenate([blocks[src][d] for src in range(world)]).reshape(-1, E)
        e.upload_spawn(recv)
        e.annihilate_spawn()
        outs.append(e.annihilate_main(cycle))
    return outs


@pytest.mark.parametrize("name,gen,real,init,tau,world", [("h2o", "renorm", True, False, 0.003, 3),
                                                          ("s12", "heat_bath", True, True, 0.004, 2)])
def test_load_balancing_redistribution_matches_oracle(name, gen, real, init, tau, world):
    """do_load_balancing + redistribute_load_balancing_dets (src/load_balancing.F90:209-323, src/qmc_common.F90:505-595,
    1332-1390): slot populations from the device, the policy on the host, the moved determinants through the spawn
    list - against the oracle's emulated ranks, before and after more MC cycles with the new proc_map."""
    from hande_b200 import load_balancing as LB
    nslots = 20
    rng = np.random.default_rng(1)
    skew = rng.choice(world * nslots, size=world * nslots // 4, replace=False)
    s, o, engs = _setup(name, gen, real, init, tau, world, 4000, nslots=nslots, skew=skew)
    shift, pe_old = -0.05, -0.1

    def cycle_all(cycle):
        o.iterate(1, cycle, tau, shift, pe_old)
        for e in engs:
            e.spawn_death(tau, shift, pe_old, cycle)
        _exchange_and_annihilate(s, o, engs, cycle)

    def compare(tag):
        for d, e in enumerate(engs):
            fo, po, do_ = o.get_psips(d)
            fg, pg, dg = e.download_psips()
            assert len(fg) == len(fo), (tag, d, len(fg), len(fo))
            assert (fg == fo).all() and (pg == po).all() and (dg == do_).all(), (tag, d)

    for cycle in (1, 2):
        cycle_all(cycle)
    compare("before")
    # initialise_slot_pop on the device == oracle; MPI_AllReduce = the sum over the engines
    slot_list = np.zeros(world * nslots)
    for d, e in enumerate(engs):
        sp = e.slot_populations()
        assert (sp == o.slot_pop(d)).all()
        slot_list += sp
    needed, pmap, _ = LB.do_load_balancing(slot_list, o.proc_map(), world, 0.05)
    assert needed and o.do_load_balancing(0.05)
    assert (np.asarray(pmap) == o.proc_map()).all() and (np.asarray(pmap) != np.arange(world * nslots) % world).any()
    lb_cycle = 0x80000000 | 3
    o.redistribute(lb_cycle)
    nsent = 0.0
    for e in engs:
        e.set_proc_map(pmap)
        nsent += e.redistribute_particles()
    assert nsent > 0
    _exchange_and_annihilate(s, o, engs, lb_cycle)
    compare("redistributed")
    for d, e in enumerate(engs):
        fg = e.download_psips()[0]
        assert all(owner_of(x, s.nbasis, world, nslots, proc_map=pmap) == d for x in fg[:200])
    for cycle in (3, 4, 5):
        cycle_all(cycle)
    compare("after")
    for e in engs:
        e.close()
