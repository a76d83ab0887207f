"""Attribute an `ncu --page source --csv` SASS listing of one kernel to source lines using `nvdisasm -g` of the
same cubin (instruction order is identical).  Usage:
  python tools/ncu_by_line.py <src.csv> <lib.so> <mangled-substring> [top]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os


def sass_lines(lib, func_sub):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    txt = ""
    for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):     # one cubin per translation unit
        t = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
        if any(ln.startswith(".text.") and func_sub in ln for ln in t.split("\n")):
            txt = t
            break
    out = []
    cur = None
    active = False
    for ln in txt.split("\n"):
        if ln.startswith(".text."):
            active = func_sub in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append((cur, ln.strip()))
    return out


def main():
    src, lib, sub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))[2:]
    sl = sass_lines(lib, sub)
    if len(sl) != len(rows):
        print(f"WARNING: {len(sl)} SASS lines in cubin vs {len(rows)} in profile - cubin differs from profiled build")
    n = min(len(sl), len(rows))
    inst = collections.Counter(); smp = collections.Counter(); thr = collections.Counter()
    for k in range(n):
        key = sl[k][0][:2] if sl[k][0] else ("?", 0)
        inst[key] += int(rows[k][5]); smp[key] += int(rows[k][4]); thr[key] += int(rows[k][6])
    ti, ts = sum(inst.values()), sum(smp.values())
    print(f"total warp-inst {ti}  samples {ts}")
    for key, c in inst.most_common(top):
        print(f"{key[0]:>14}:{key[1]:<5} inst={c / ti * 100:5.1f}% smp={smp[key] / ts * 100:5.1f}% lanes={thr[key] / max(c, 1):5.1f}")


if __name__ == "__main__":
    main()
