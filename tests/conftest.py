import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fcidump_path(tmp_path_factory):
    """Decompress a committed FCIDUMP fixture (tests/golden/fcidump/<name>.INTDUMP.gz) to a temp file."""
    cache = {}

    def get(name):
        if name not in cache:
            dst = tmp_path_factory.mktemp("fcidump") / (name + ".INTDUMP")
            with gzip.open(os.path.join(GOLDEN, "fcidump", name + ".INTDUMP.gz"), "rb") as fi:
                dst.write_bytes(fi.read())
            cache[name] = str(dst)
        return cache[name]

    return get


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))
