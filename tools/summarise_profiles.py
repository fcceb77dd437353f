"""Summarise gpurun_out/*.ncu-rep and the launch list into tracked files under profiles/.

    python tools/summarise_profiles.py r01
writes profiles/<tag>_kernels.csv (one row per profiled launch: duration, DRAM bytes, throughput %,
registers, occupancy, top stall), profiles/<tag>_launches.csv (the dae launches of one bench step with
their device time and share) and updates profiles/traffic.json (DRAM bytes per launch of the CTC pair).
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_ld_sectors",
}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "?").split("(")[0].replace("void dae::", "")}
        for k, name in WANT.items():
            if k in d:
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                unit = u.get(k, "")
                if name == "duration_us":
                    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
                if name.endswith("_MB"):
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
                    v *= scale
                rec[name] = round(v, 3)
        stalls = {k.split("issue_stalled_")[1].split("_per")[0]: float(v) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("per_issue_active.ratio")
                  and "not_issued" not in k and v not in ("", "n/a")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        rec["top_stalls"] = " ".join(f"{k}={v:.2f}" for k, v in top)
        res.append(rec)
    return res


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    all_rows = []
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"*_{tag}.ncu-rep"))):
        fam = os.path.basename(rep).replace(f"_{tag}.ncu-rep", "")
        for r in raw_rows(rep):
            r["family"] = fam
            all_rows.append(r)
    cols = ["family", "kernel"] + list(WANT.values()) + ["top_stalls"]
    with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for r in all_rows:
            w.writerow({c: r.get(c, "") for c in cols})
    # traffic of the CTC pair (per launch, DRAM read+write)
    traffic = {}
    # one CTC loss+grad = every ctc_* launch of one repetition (bands + scan + fused block gradient, or the
    # per-frame chain + row gradient when that path is active)
    ctc = [r for r in all_rows if r["family"] == "ctc" and "ctc_" in r["kernel"]]
    if ctc:
        seen, tb = set(), 0.0
        for r in ctc:
            if r["kernel"] in seen:
                continue
            seen.add(r["kernel"])
            tb += (r.get("dram_read_MB", 0) + r.get("dram_write_MB", 0)) * 1e6
        traffic["ctc_lattice+ctc_grad"] = tb
        traffic["ctc_kernels"] = sorted(seen)
    for r in all_rows:
        traffic.setdefault("per_kernel", {})[r["family"] + ":" + r["kernel"]] = (r.get("dram_read_MB", 0) + r.get("dram_write_MB", 0)) * 1e6
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    # launch list
    ll = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(ll):
        lines = [l for l in open(ll) if not l.startswith("==")]
        rows = list(csv.DictReader(io.StringIO("".join(lines))))
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            us = v / 1e3 if r.get("Metric Unit") in ("ns", "nsecond") else (v * 1e3 if r.get("Metric Unit") in ("ms", "msecond") else v)
            k = r["Kernel Name"].split("(")[0].replace("void dae::", "")
            agg[k][0] += 1
            agg[k][1] += us
        tot = sum(v[1] for v in agg.values()) or 1.0
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_dae_time"])
            for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, n, round(us, 2), round(us / n, 2), round(us / tot, 4)])
    print(open(os.path.join(ROOT, "profiles", f"{tag}_kernels.csv")).read())
    if os.path.exists(os.path.join(ROOT, "profiles", f"{tag}_launches.csv")):
        print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv")).read())


if __name__ == "__main__":
    main()
