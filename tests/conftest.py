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


def pytest_collection_modifyitems(config, items):
    # a kernel that never returns must not hold the GPU box until the harness's own limit: every GPU test gets a time
    # limit (pytest-timeout, thread method: the process is ended, later tests are reported as not run)
    for item in items:
        if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(420, method="thread"))


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
