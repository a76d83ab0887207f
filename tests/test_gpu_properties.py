"""Size-independent properties of the propagation at (near) benchmark size (-m gpu).  The oracle cannot run 1e7-1e8
walkers in test time, so these check what must hold at any size: the merged list stays strictly sorted in the
reference's order and free of duplicates and of empty entries, the reported population equals the sum over the list,
the same seed gives the same list (run-to-run determinism), and - the property the counter-based stream was designed
for - a run sharded over two ranks produces exactly the list of the single-rank run."""
import numpy as np
import pytest

from hande_b200 import read_in as R
from hande_b200.engine import Engine
from hande_b200.fciqmc import owner_of
from tests.common import system_path

pytestmark = pytest.mark.gpu


def _s50_engine(n, **kw):
    path, skw = system_path("s50")
    s = R.read_in(path, **skw)
    occ0 = R.set_reference_det(s)
    ps, pd = R.find_single_double_prob(s, occ0)
    eng = Engine(s, excit_gen="heat_bath", pattempt_single=ps, pattempt_double=pd, real_amplitudes=True, spawn_cutoff=0.01,
                 initiator_approx=True, walker_length=int(n * 1.3) + 4096, spawned_walker_length=max(int(n * 0.3), 1 << 16),
                 seed=7, **kw)
    eng.set_reference(s.encode(occ0), s.slater_condon0(occ0))
    return s, eng, s.slater_condon0(occ0)


def _sorted_strictly(f):
    """bit_str_cmp order: unsigned compare, last word most significant"""
    a, b = f[:-1], f[1:]
    lt = np.zeros(len(a), dtype=bool)
    eq = np.ones(len(a), dtype=bool)
    for w in range(f.shape[1] - 1, -1, -1):
        lt |= eq & (a[:, w] < b[:, w])
        eq &= a[:, w] == b[:, w]
    return bool(lt.all())


def test_properties_at_ten_million_walkers():
    import torch
    from hande_b200 import synthetic
    n = 10_000_000
    s, eng, H00 = _s50_engine(n)
    states, pops = synthetic.random_walkers_torch(n, s.nbasis, s.nalpha, s.nbeta, 1 << 31, torch.device("cuda", 0), seed=1)
    dat = np.concatenate([eng.sc0_batch(states[a:a + 5_000_000]) - H00 for a in range(0, len(states), 5_000_000)])
    tau = 5.3e-7
    runs = []
    for rep in range(2):
        eng.upload_psips(states, pops, dat)
        out = eng.iterate(3, tau, 0.0, 0.0, 1)
        assert out["spawn_error"] == 0 and out["psip_error"] == 0
        f, p, d = eng.download_psips()
        runs.append((f, p, d, out))
    f, p, d, out = runs[0]
    assert len(f) == out["nstates"] and len(f) > 0.9 * n
    assert _sorted_strictly(f)                                   # sorted, hence no duplicates
    assert (p != 0).all()                                        # remove_unoccupied_dets
    assert (np.abs(p) >= (1 << 31)).all()                        # every survivor was rounded up to a whole walker
    assert abs(out["nparticles"] - np.abs(p).sum() / 2**31) <= 1e-9 * out["nparticles"]
    assert 0.03 < out["nspawn_events"] / out["nattempts"] * 2 < 0.08        # the bench's spawning regime (R_spawn ~ 0.05)
    # new determinants carry <D|H|D> - H00 evaluated on the device: spot-check against sc0_batch
    idx = np.random.default_rng(0).integers(0, len(f), 2000)
    assert (eng.sc0_batch(f[idx]) - H00 == d[idx]).all()
    # run-to-run determinism (atomic append order differs between runs; the result must not)
    f2, p2, d2, out2 = runs[1]
    assert len(f2) == len(f) and (f2 == f).all() and (p2 == p).all() and (d2 == d).all()
    assert out2["proj_energy"] == out["proj_energy"] and out2["ndeath"] == out["ndeath"]
    eng.close()


def test_two_ranks_give_the_single_rank_list():
    """hash-owner sharding does not change the result: the union of the two ranks' lists equals the one-rank list"""
    from hande_b200 import synthetic
    n, tau, ncyc = 200_000, 5.3e-7, 4
    s, e1, H00 = _s50_engine(n)
    states, pops = synthetic.random_walkers(n, s.nbasis, s.nalpha, s.nbeta, real_factor=1 << 31, dist="B", seed=3)
    dat = e1.sc0_batch(states) - H00
    e1.upload_psips(states, pops, dat)
    world = 2
    engs = []
    own = np.array([owner_of(x, s.nbasis, world, 1) for x in states])
    for r in range(world):
        _, e, _ = _s50_engine(n, nprocs=world, iproc=r)
        m = own == r
        e.upload_psips(states[m], pops[m], dat[m])
        engs.append(e)
    E = s.W + 2
    for cycle in range(1, ncyc + 1):
        e1.iterate(1, tau, -0.01, -0.02, cycle)
        blocks = []
        for e in engs:
            e.spawn_death(tau, -0.01, -0.02, cycle)
            cnt = e.spawn_counts()
            sd = e.download_spawn()
            off = np.concatenate([[0], np.cumsum(cnt)])
            blocks.append([sd[off[d]:off[d + 1]] for d in range(world)])
        for d, e in enumerate(engs):
            e.upload_spawn(np.concatenate([blocks[src][d] for src in range(world)]).reshape(-1, E))
            e.annihilate_spawn()
            e.annihilate_main(cycle)
    f1, p1, d1 = e1.download_psips()
    parts = [e.download_psips() for e in engs]
    fu = np.concatenate([x[0] for x in parts]); pu = np.concatenate([x[1] for x in parts]); du = np.concatenate([x[2] for x in parts])
    order = np.lexsort(tuple(fu[:, w] for w in range(s.W)))
    assert len(fu) == len(f1)
    assert (fu[order] == f1).all() and (pu[order] == p1).all() and (du[order] == d1).all()
    for e in engs + [e1]:
        e.close()
