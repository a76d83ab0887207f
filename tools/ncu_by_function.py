"""Aggregate an `ncu --page source --csv` listing of k_spawn_death by the source FUNCTION of each SASS instruction
(tools/ncu_by_line.py maps SASS to lines through `nvdisasm -g` of the same object file; the function is the last
definition that starts at or before the line).  Usage: python tools/ncu_by_function.py <src.csv> <object file> [top]"""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_by_line as N  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(path):
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"^(?:HB_HDNI|HB_HDN|HB_HD|static|__device__|inline|template <[^>]*>\s*HB_HD\w*)[^;(]*?\b(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            out.append((i, m.group(1)))
    return out


def main():
    src, obj = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rows = list(csv.reader(open(src)))[2:]
    sl = N.sass_lines(obj, "k_spawn_death")
    n = min(len(sl), len(rows))
    core = functions(os.path.join(ROOT, "hande_b200", "csrc", "hb_core.cuh"))

    def fn(line):
        name = "?"
        for a, b in core:
            if a <= line:
                name = b
            else:
                break
        return name
    inst, smp = collections.Counter(), collections.Counter()
    for k in range(n):
        key = sl[k][0]
        if key is None:
            g = "?"
        elif key[0] == "hb_core.cuh":
            g = "hb_core.cuh:" + fn(key[1])
        else:
            g = key[0]
        inst[g] += int(rows[k][5]); smp[g] += int(rows[k][4])
    ti, ts = sum(inst.values()), sum(smp.values())
    print(f"total warp-inst {ti}  samples {ts}")
    for g, c in inst.most_common(top):
        print(f"{g:45s} inst {c / ti * 100:5.1f}%  samples {smp[g] / ts * 100:5.1f}%")


if __name__ == "__main__":
    main()
