"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/hande_b200.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "hande_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(hb200_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from hande_b200 import engine
    lib = engine.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(engine.ABI_SYMBOLS) == names


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors == sizeof() of the C structs, checked by compiling the header with gcc."""
    import subprocess
    from hande_b200 import engine
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "hande_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", '
                   'sizeof(hb200_iter_in), sizeof(hb200_iter_out), sizeof(hb200_config), '
                   'sizeof(hb200_system_read_in));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(engine.IterIn), ctypes.sizeof(engine.IterOut), ctypes.sizeof(engine.Config),
                     ctypes.sizeof(engine.SystemReadIn)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hande_b200 import engine
    lib = engine.load_library()
    cfg = engine.Config(nbasis=10, nel=4)
    assert not lib.hb200_create(ctypes.byref(cfg))
    assert b"no CUDA device" in lib.hb200_last_error()
    from hande_b200 import read_in as R, synthetic
    s = R.read_in(synthetic.synthetic_fcidump(6, 4), is_text=True)
    with pytest.raises(engine.EngineError):
        engine.Engine(s, pattempt_single=0.1, pattempt_double=0.9)


def test_host_owner_hash_matches_oracle(fcidump_path):
    import numpy as np
    from hande_b200 import read_in as R, synthetic
    from hande_b200.fciqmc import owner_of, list_sizes, QmcIn, format_row
    from oracle.pyoracle import Oracle
    o = Oracle()
    o.read_fcidump(fcidump_path("ne"), nel=10, ms=0, sym=0)
    s = R.read_in(fcidump_path("ne"), nel=10, ms=0, sym=0)
    o.set_qmc(nprocs=4, rng_kind=1, spawned_walker_length=4096)
    o.init()
    for f in synthetic.random_dets(300, s.nbasis, s.nalpha, s.nbeta, seed=8):
        assert owner_of(f, s.nbasis, 4, 1) == o.owner(f)
    # list sizing and the report-row format of the golden H2O output
    assert list_sizes(QmcIn(state_size=-5, spawned_state_size=-1), 1, 1) == (178571, 31250)
    row = format_row(20, 0.0, -1.6562228407E-02, 10.0, 13.0, 4, 2, 0.0070, 0.0)
    assert row == ("               20   0.0000000000E+00     -1.6562228407E-02      1.0000000000E+01      "
                   "1.3000000000E+01                  4               2    0.0070    0.0000  ")
    assert format_row(0, 0.0, 0.0, 10.0, 10.0, 1, 0, 0.0, 0.0, comment=True) == (
        " #              0   0.0000000000E+00      0.0000000000E+00      1.0000000000E+01      1.0000000000E+01"
        "                  1               0    0.0000    0.0000  ")
