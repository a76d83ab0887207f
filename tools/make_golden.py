"""Generate tests/golden/ fixtures from the reference's own test_suite (run in the build container).

  * tests/golden/fcidump/{h2o,ne}.INTDUMP.gz   - the FCIDUMP inputs of the reference fixtures (data, not source)
  * tests/golden/<case>.json                   - every report-loop row of the reference's benchmark.out.* table
                                                 plus the system/qmc options of the corresponding *.in file

The tables are the reference's golden vectors for the hot path; the oracle must
reproduce them with the reference dSFMT stream before it is trusted as the checker for the CUDA path.
"""
import gzip
import json
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.golden_compare import CASES, TS, parse_table  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FCIDUMP_NAME = {"h2o": "h2o", "ne_init": "ne", "ne_ci6_np2": "ne", "ne_ci6_np4": "ne", "ne_ci6_real64_np2": "ne",
                "ne_ci6_real32_np2": "ne", "ne_ci6_real32_np4": "ne",
                "ccmc_ne": "ne_vdz", "ccmc_h2o_np2": "h2o_vdz",
                "ccmc_h2o_ccsdt_fullnc_np2": "h2o_vdz"}
FCIDUMP_NAME.update({k: "nh3_631g" for k in CASES if k.startswith("ccmc_nh3_")})
FCIDUMP_NAME["h4_cheby"] = "h4_sto3g"
FCIDUMP_NAME["ne_cisdtq_no_renorm_np2"] = "ne"
FCIDUMP_NAME["n2_harmonic_np2"] = "n2_sto3g"
FCIDUMP_NAME["he2_ss"] = "he2_avdz"
FCIDUMP_NAME["he2_ss_sep"] = "he2_avdz"
FCIDUMP_NAME["he2_ss_cisd"] = "he2_avdz"
FCIDUMP_NAME["ccmc_h2o_ccsdt_qn"] = "h2o_vdz"
FCIDUMP_NAME["ccmc_h2o_ccsdt_qn_fullnc"] = "h2o_vdz"


def pattempt_changes(path):
    """(iteration of the preceding table row, value) of every '# pattempt_single changed to be:' line"""
    out, last = [], 0
    for line in open(path):
        if "pattempt_single changed to be:" in line:
            out.append([last, float(line.split(":")[1])])
        else:
            t = line.split()
            if len(t) > 8 and t[0].isdigit():
                last = int(t[0])
    return out

os.makedirs(os.path.join(OUT, "fcidump"), exist_ok=True)
for name, c in CASES.items():
    d = TS + c["dir"] + "/"
    if "ueg" in c:
        rows = parse_table(d + c["bench"]).tolist()
        json.dump({"source": "test_suite/" + c["dir"] + "/" + c["bench"], "ueg": c["ueg"], "ref_det": c["ref_det"],
                   "qmc": c["qmc"], **({"quasi_newton": c["quasi_newton"]} if "quasi_newton" in c else {}),
                   **({"semi_stoch": c["semi_stoch"], "vary_shift": True} if "semi_stoch" in c else {}),
                   **({"pop_real_bits": c["pop_real_bits"]} if "pop_real_bits" in c else {}),
                   "columns": ["iterations", "shift", "proj_energy", "D0_population", "nparticles", "nstates",
                               "nspawn_events", "rspawn"],
                   # the "H00", box length, basis size and third eigenvalue the reference prints for the system
                   "kat": ({"H00": 11.13923996, "nbasis": 38, "determ_size": 358, "determ_min": 83, "determ_max": 98}
                           if name == "ueg_ss_np4" else
                           {"H00": 2.02890441, "L": 5.85836755, "nbasis": 66, "sp_eigv_3": 5.75143889E-01}),
                   "rows": rows}, open(os.path.join(OUT, name + ".json"), "w"))
        print(name, len(rows), "rows")
        continue
    if name not in FCIDUMP_NAME:
        continue
    dst = os.path.join(OUT, "fcidump", FCIDUMP_NAME[name] + ".INTDUMP.gz")
    if not os.path.exists(dst):
        with open(d + c["int_file"], "rb") as fi, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    rows = parse_table(d + c["bench"]).tolist()
    extra = {}
    if c.get("until_shift"):
        # keep the rows before the shift varies (the comparable part, see golden_compare.py) and a few after
        nz = [i for i, r in enumerate(rows) if r[1] != 0.0]
        rows = rows[:min(len(rows), (nz[0] if nz else len(rows)) + 3)]
        extra = {"until_shift": True, "pattempt_update": bool(c.get("pattempt_update")),
                 "pattempt_parallel": c.get("pattempt_parallel", -1.0),
                 "pattempt_changes": [pc for pc in pattempt_changes(d + c["bench"]) if pc[0] <= rows[-1][0]]}
    cols = ["iterations", "shift", "proj_energy", "D0_population", "nparticles", "nstates", "nspawn_events", "rspawn"]
    json.dump({"source": "test_suite/" + c["dir"] + "/" + c["bench"], "fcidump": FCIDUMP_NAME[name], "sys": c["sys"],
               "qmc": c["qmc"], "ccmc": bool(c.get("ccmc")), "full_nc": bool(c.get("full_nc")),
               **({"quasi_newton": c["quasi_newton"]} if "quasi_newton" in c else {}),
               **({"pop_real_bits": c["pop_real_bits"]} if "pop_real_bits" in c else {}),
               **({"harmonic_forcing": c["harmonic_forcing"]} if "harmonic_forcing" in c else {}),
               **({"semi_stoch": c["semi_stoch"], "vary_shift": True, "kat": {"H00": -5.69708312, "determ_size": 69 if name == "he2_ss_cisd" else 100}}
                  if "semi_stoch" in c else {}),
               **({"chebyshev": c["chebyshev"],
                   "kat": {"spectral_range": 2.92929139E+00,     # the "Initial estimate of spectral range" and the
                           "zeroes": [2.32507329E-01, 8.56209883E-01, 1.67308651E+00, 2.42378465E+00, 2.86996294E+00],
                           "weights": [4.30093969E+00, 1.16793793E+00, 5.97697725E-01, 4.12577909E-01, 3.48436555E-01]}}
                  if "chebyshev" in c else {}),
               "columns": cols + (["nattempts"] if c.get("ccmc") else []),
               "rows": rows, **extra}, open(os.path.join(OUT, name + ".json"), "w"))
    print(name, len(rows), "rows")
