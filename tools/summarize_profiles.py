"""Build profiles/<tag>_summary.md (+ copies of the small raw files) from a gpurun_out/ collection:
  <tag>_bench.json, <tag>_bench_reference.json, <tag>_launches.csv (ncu gpu__time_duration launch list),
  <tag>_prof_1e8.ncu-rep (ncu --set full of the hot kernels at the bench size), lib_<tag>.so (the profiled build)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

out = [f"# Profile summary {tag} (one B200, bench.py default workload: S50 heat-bath, 1e8 walkers)\n"]
bench = json.load(open(os.path.join(G, f"{tag}_bench.json")))
shutil.copy(os.path.join(G, f"{tag}_bench.json"), os.path.join(P, f"{tag}_bench.json"))
r = bench["roofline"]
out.append("## bench.py line (committed as `%s_bench.json`)\n" % tag)
out.append(f"* value: **{bench['value']:.4g} walker-iterations/s** ({bench['ms_per_step']:.2f} ms per MC cycle, "
           f"{bench['steps']} timed steps, {bench['warmup']} warm-up), spawn attempts/s {bench['spawn_attempts_per_s']:.4g}")
out.append(f"* stage ms per step: {json.dumps({k: round(v, 3) for k, v in r['stage_ms_per_step'].items()})}")
out.append(f"* roofline (k_spawn_death): achieved {r['achieved']:.1f} GB/s of {r['peak']} GB/s measured = "
           f"{r['frac']:.3f}; whole-cycle B_alg/time = {r['cycle_frac']:.3f} of peak")
if bench.get("e2e"):
    out.append(f"* e2e (list uploaded from pinned host memory every step): {bench['e2e']['value']:.4g} walker-iterations/s, "
               f"{bench['e2e']['h2d_bytes_per_step'] / 1e9:.2f} GB H2D per step")
if bench.get("cpu_baseline"):
    c = bench["cpu_baseline"]
    out.append(f"* cpu_baseline: {c['value']:.4g} walker-iterations/s on {c['cores']} cores ({c['kind']}; {c['sample']})")
out.append(f"* clocks: {json.dumps(bench['clocks'])}; gpu_launches in timed region: {bench['gpu_launches']}\n")
refp = os.path.join(G, f"{tag}_bench_reference.json")
if os.path.exists(refp):
    shutil.copy(refp, os.path.join(P, f"{tag}_bench_reference.json"))
    ref = json.load(open(refp))
    out.append(f"* `--impl reference` arm: {ref['value']:.4g} walker-iterations/s ({ref['cpu_baseline']['sample']})\n")

# ---- launch list
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, f"{tag}_launches.csv"))
    rows = list(csv.reader(open(lp)))
    hi = [k for k, rr in enumerate(rows) if "Kernel Name" in rr][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for rr in data:
        if len(rr) <= mv or rr[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", rr[kn]).replace("void ", "")
        if not name.startswith("k_"):
            continue  # torch kernels of the input generator are not part of the engine
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(rr[mv].replace(",", ""))
    step_kernels = {k: v for k, v in agg.items() if not k.startswith(("k_hb_", "k_build", "k_sc0", "k_abs"))}
    tot = sum(v[1] for v in step_kernels.values())
    out.append("## ncu launch list (`%s_launches.csv`: bench.py --walkers 1e7 --steps 2 --warmup 1; cold-cache, "
               "serialised: compare shares)\n" % tag)
    out.append("| engine kernel (per-cycle stages) | launches | total ms | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(step_kernels.items(), key=lambda x: -x[1][1]):
        out.append(f"| {n} | {c} | {t / 1e6:.3f} | {t / tot * 100:.1f}% |")
    out.append("")
    setup = {k: v for k, v in agg.items() if k not in step_kernels}
    out.append("one-off set-up kernels: " + ", ".join(f"{k} {v[1] / 1e6:.2f} ms" for k, v in setup.items()) + "\n")

# ---- full captures
rep = os.path.join(G, f"{tag}_prof_1e8.ncu-rep")
rawcsv = os.path.join(G, f"{tag}_raw_1e8.csv")
if os.path.exists(rep) or os.path.exists(rawcsv):
    if os.path.exists(rawcsv):
        raw = open(rawcsv).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__grid_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
    keep = [kn] + [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(P, f"{tag}_ncu_full_1e8.csv"), "w", newline="") as f:
        wtr = csv.writer(f)
        for rr in rows:
            wtr.writerow([rr[i] for i in keep])
    out.append("## ncu --set full at the bench size (1e8 walkers), key counters (`%s_ncu_full_1e8.csv`)\n" % tag)
    out.append("| kernel | ms | DRAM read GB | DRAM write GB | DRAM % | L2 % | L1 % | L2 hit % | issue % | lanes/inst | regs |\n"
               "|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    for rr in rows[2:]:
        def g(name):
            try:
                return float(rr[hdr.index(name)])
            except Exception:
                return float("nan")

        def gb(name):
            v = g(name)
            return v / 1e3 if units[hdr.index(name)].lower().startswith("mbyte") else v
        nm = re.sub(r"\(.*", "", rr[kn]).replace("void ", "")
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        traffic[nm] = (rd + wr) * 1e9
        out.append(f"| {nm} | {g('gpu__time_duration.sum'):.2f} | {rd:.2f} | {wr:.2f} | "
                   f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{g('lts__t_sector_hit_rate.pct'):.1f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{g('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {g('launch__registers_per_thread'):.0f} |")
    out.append("")
    json.dump({"walkers_per_gpu": 100000000, "source": f"profiles/{tag}_ncu_full_1e8.csv",
               "dram_bytes_per_launch": {k: v for k, v in traffic.items() if v == v}},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
    lib = os.path.join(G, f"lib_{tag}.so")
    if not os.path.exists(lib):
        # the object file of the profiled build (the library holds one cubin per translation unit; the spawn kernel of
        # the bench lives in the W = 2 / heat-bath unit)
        lib = os.path.join(ROOT, "hande_b200", "build", "hb_spawn_w2_g0.o")
    if os.path.exists(lib):
        tmp = os.path.join(G, f"{tag}_src_spawn.csv")
        if not os.path.exists(tmp) or os.path.exists(rep):
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_spawn_death"],
                                 capture_output=True, text=True).stdout
            open(tmp, "w").write(src)
        bl = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), tmp, lib, "k_spawn_death", "40"],
                            capture_output=True, text=True).stdout
        out.append("## k_spawn_death: instruction / stall-sample share by source line (tools/ncu_by_line.py)\n\n```\n" + bl + "```\n")
        bf = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), tmp, lib, "24"],
                            capture_output=True, text=True).stdout
        out.append("## k_spawn_death: the same listing summed by source function (tools/ncu_by_function.py)\n\n```\n" + bf + "```\n")

# ---- CCMC cluster kernel (tools/gpu_prof_ccmc.sh): before / after grouping a block's attempts by cluster size
ccmc_rows = []
for lab, fn in (("one thread per attempt in index order", "raw_ccmc_a.csv"),
                ("attempts of a block grouped by cluster size (final)", "raw_ccmc_b.csv")):
    pth = os.path.join(G, fn)
    if not os.path.exists(pth):
        pth = os.path.join(P, f"{tag}_ncu_" + fn.replace("raw_", ""))
    if not os.path.exists(pth):
        continue
    rows = list(csv.reader(open(pth).read().splitlines()))
    hdr = rows[0]
    rr = rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
            "lts__t_sector_hit_rate.pct"]
    keep = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(P, f"{tag}_ncu_" + fn.replace("raw_", "")), "w", newline="") as f:
        wtr = csv.writer(f)
        for r3 in rows[:3]:
            wtr.writerow([r3[i] for i in keep])

    def g(name):
        try:
            return float(rr[hdr.index(name)])
        except Exception:
            return float("nan")
    natt = g("launch__grid_size") * 256
    ccmc_rows.append(f"| {lab} | {natt:.3g} | {g('gpu__time_duration.sum'):.2f} | {g('smsp__inst_executed.sum') / natt:.0f} | "
                     f"{g('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | "
                     f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {g('launch__registers_per_thread'):.0f} |")
if ccmc_rows:
    out.append("## k_ccmc_cluster (CCSDT, S40; ncu --set full of one launch, `%s_ncu_ccmc_*.csv`)\n" % tag)
    out.append("| version | attempts in the launch | ms (under ncu) | warp-inst per attempt | lanes/inst | warps active % | regs |\n|---|---|---|---|---|---|---|")
    out += ccmc_rows
    out.append("")

# ---- side measurements committed next to the headline line
side = sorted(f for f in os.listdir(P) if f.startswith(f"{tag}_bench_") and f.endswith(".json") and "reference" not in f)
if side:
    out.append("## other measured lines (`profiles/%s_bench_*.json`)\n" % tag)
    out.append("| file | GPUs | workload | value | ms per step |\n|---|---|---|---|---|")
    for f in side:
        try:
            d = json.loads(open(os.path.join(P, f)).read().strip().splitlines()[-1])
        except Exception:
            continue
        out.append(f"| {f} | {d.get('n_gpus')} | {d['config']['workload']} | {d['value']:.4g} {d['unit']} | {d['ms_per_step']:.2f} |")
    out.append("")
gens = sorted(f for f in os.listdir(G) if f.startswith("gen_") and f.endswith(".json"))
if gens:
    out.append("## spawn kernel per generator (bench.py --excit-gen G, 1e8 walkers, one B200; `gpurun_out/gen_*.json` of the final build)\n")
    out.append("| generator | k_spawn_death ms | cycle ms | walker-iterations/s |\n|---|---|---|---|")
    for f in gens:
        try:
            d = json.loads(open(os.path.join(G, f)).read().strip().splitlines()[-1])
        except Exception:
            continue
        out.append(f"| {f[4:-5]} | {d['roofline']['kernel_ms']:.2f} | {d['ms_per_step']:.2f} | {d['value']:.4g} |")
    out.append("")

open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
