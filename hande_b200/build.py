"""Build libhande_b200.so (CUDA engine + C ABI) in-tree for sm_100a with nvcc.

The library is split into translation units that compile in parallel: hb_engine.cu (C ABI, stage drivers, sort /
annihilation / merge kernels), hb_spawn_tu.cu once per (W, generator group) and hb_ccmc_tu.cu once per W.  Object files
live in hande_b200/build/ (git-ignored); only the objects whose sources changed are recompiled."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhande_b200.so")
HEADERS = ["hb_core.cuh", "hb_common.cuh", "hb_spawn.cuh", "hb_ccmc.cuh", "hb_list.cuh", "hb_semistoch.cuh", os.path.join("..", "..", "include", "hande_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # no FMA contraction: sums must follow the reference's operation order (bit-exact excitation choice / nspawn)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
MAXW, NGROUPS = 4, 5


def units():
    """(object name, source, extra flags, headers it depends on)"""
    u = [("hb_engine.o", "hb_engine.cu", [], ["hb_core.cuh", "hb_common.cuh", HEADERS[-1]])]
    for w in range(1, MAXW + 1):
        for g in range(NGROUPS):
            u.append((f"hb_spawn_w{w}_g{g}.o", "hb_spawn_tu.cu", [f"-DHB_TU_W={w}", f"-DHB_TU_GROUP={g}"],
                      ["hb_core.cuh", "hb_common.cuh", "hb_spawn.cuh", HEADERS[-1]]))
        u.append((f"hb_ccmc_w{w}.o", "hb_ccmc_tu.cu", [f"-DHB_TU_W={w}"],
                  ["hb_core.cuh", "hb_common.cuh", "hb_ccmc.cuh", HEADERS[-1]]))
        u.append((f"hb_list_w{w}.o", "hb_list_tu.cu", [f"-DHB_TU_W={w}"],
                  ["hb_core.cuh", "hb_common.cuh", "hb_list.cuh", "hb_semistoch.cuh", HEADERS[-1]]))
    # wide layout (bit strings of 5..32 words, stored as 32): 16-bit occupied lists; the generators without
    # nbasis^3 / nbasis^4 tables only (the UEG's, group 4)
    wide = ["-DHB_TU_W=32", "-DHB_OCC16"]
    u.append(("hb_spawn_w32_g4.o", "hb_spawn_tu.cu", wide + ["-DHB_TU_GROUP=4"], ["hb_core.cuh", "hb_common.cuh", "hb_spawn.cuh", HEADERS[-1]]))
    u.append(("hb_ccmc_w32.o", "hb_ccmc_tu.cu", wide, ["hb_core.cuh", "hb_common.cuh", "hb_ccmc.cuh", HEADERS[-1]]))
    u.append(("hb_list_w32.o", "hb_list_tu.cu", wide, ["hb_core.cuh", "hb_common.cuh", "hb_list.cuh", "hb_semistoch.cuh", HEADERS[-1]]))
    return u


def _stale(obj, src, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in [src] + deps) or os.path.getmtime(__file__) > t


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = {u[1] for u in units()} | set(HEADERS)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in srcs)


def _compile(nvcc, obj, src, flags):
    cmd = [nvcc] + NVCC_FLAGS + flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return cmd, res


def build(force=False, verbose=False, jobs=None):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    todo = [(os.path.join(OBJ, o), s, f) for o, s, f, d in units() if force or _stale(os.path.join(OBJ, o), s, d)]
    log = []
    jobs = jobs or min(len(todo), os.cpu_count() or 4) or 1
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = [ex.submit(_compile, nvcc, o, s, f) for o, s, f in todo]
        for fu in futs:
            cmd, res = fu.result()
            log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if res.returncode != 0:
                sys.stderr.write(log[-1])
                raise RuntimeError("nvcc failed building " + cmd[-1])
    objs = [os.path.join(OBJ, o) for o, _, _, _ in units()]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(os.path.join(OBJ, "build.log"), "w") as f:     # ptxas -v output (registers, spills) of the last build
        f.write("\n".join(log))
    if res.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("nvcc failed linking libhande_b200.so")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="-f" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
