"""Build libhande_b200.so (CUDA engine + C ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhande_b200.so")
SOURCES = ["hb_engine.cu"]
HEADERS = ["hb_core.cuh", os.path.join("..", "..", "include", "hande_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # no FMA contraction: sums must follow the reference's operation order (bit-exact excitation choice / nspawn)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libhande_b200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print("built", LIB)
