import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.common import make_pair
s, o, eng, ref = make_pair("s40", spawned_walker_length=1 << 17)
rng = np.random.default_rng(4)
for n in (5000, 8192, 8193, 10000, 16384, 20000, 40000, 100000):
    for dup in (False, True):
        keys = rng.integers(0, 2**63 - 1, size=(n, 2), dtype=np.int64)
        keys[:, 1] &= (1 << (s.nbasis - 64)) - 1
        if dup:
            keys[rng.integers(0, n, size=n // 3)] = keys[0]
        sd = np.concatenate([keys, rng.integers(-5, 6, size=(n, 1)), np.arange(n).reshape(-1, 1)], axis=1)
        eng.upload_spawn(sd)
        eng.annihilate_spawn()
        got = eng.download_spawn()
        order = np.lexsort((sd[:, 0].astype(np.uint64), sd[:, 1].astype(np.uint64)))
        exp = sd[order]
        ok = (got == exp).all()
        perm_ok = (np.sort(got[:, 3]) == np.arange(n)).all()
        kg = got[:, :2].astype(np.uint64)
        sorted_ok = (np.lexsort((kg[:, 0], kg[:, 1])) == np.arange(n)).all() if not dup else None
        nbad = int((got != exp).any(axis=1).sum())
        first = int(np.argmax((got != exp).any(axis=1))) if nbad else -1
        print(f"n={n} dup={dup} ok={ok} is_permutation={perm_ok} keys_sorted={sorted_ok} nbad={nbad} first_bad={first}")
        if nbad and nbad < 50:
            idx = np.nonzero((got != exp).any(axis=1))[0]
            print("   bad rows", idx[:10], got[idx[:4]], exp[idx[:4]])
