"""Flyvbjerg-Petersen reblocking (restating tools/pyblock/pyblock/blocking.py reblock/find_optimal_block and the ratio
error of tools/pyblock/pyblock/error.py:46-51 over plain numpy)."""
import numpy as np


def reblock(x):
    x = np.asarray(x, dtype=float)
    out = []
    while len(x) >= 2:
        n = len(x)
        mean = x.mean()
        var = x.var(ddof=1)
        se = np.sqrt(var / n)
        out.append((n, mean, se, se / np.sqrt(2 * (n - 1))))
        if n % 2:
            x = x[:-1]
        x = 0.5 * (x[0::2] + x[1::2])
    return out


def optimal_error(x):
    """Standard error at the first block level satisfying B^3 > 2 N (std_err/std_err_0)^4 (pyblock's criterion)."""
    st = reblock(x)
    n0, se0 = st[0][0], st[0][2]
    for i, (n, mean, se, sese) in enumerate(st):
        B = 2 ** i
        if B ** 3 > 2 * n0 * (se / se0) ** 4:
            return st[0][1], se
    return st[0][1], max(s[2] for s in st)


def ratio_with_error(a, b):
    ma, ea = optimal_error(a)
    mb, eb = optimal_error(b)
    n = len(a)
    cov = np.cov(a, b)[0, 1]
    r = ma / mb
    err = abs(r) * np.sqrt(max((ea / ma) ** 2 + (eb / mb) ** 2 - 2 * cov / (n * ma * mb), 0.0))
    return r, err
