#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the JSON bench.py reads for `roofline.traffic`.

  python scripts/ncu_summary.py gpurun_out/r02_full.ncu-rep --workload cfg2 --reads-per-launch 2000000 \
         -o profiles/r02_ncu_summary.json

One entry per profiled kernel launch (first launch of each kernel name kept): duration, DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum), executed warp instructions, issue-active %, threads per
instruction, L2 hit rate, registers, achieved occupancy, top stall reasons.  Needs `ncu` on PATH (reads the
report; no GPU)."""
import argparse
import csv
import io
import json
import subprocess

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--reads-per-launch", type=int, default=2_000_000)
    ap.add_argument("-o", "--out", default="profiles/r02_ncu_summary.json")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=True):
        i = col.get(name)
        if i is None or r[i] in ("", "no data", "n/a"):
            return None
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return None
        return v * UNITS.get(units[i], 1.0) if scale else v

    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out, seen = [], set()
    for r in data:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("nb200::", "")
        if short in seen:
            continue
        seen.add(short)
        inst = val(r, "smsp__inst_executed.sum", False)
        st = sorted(((val(r, s, False) or 0.0, s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)
        e = {"name": short, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
             "duration_ms": (val(r, "gpu__time_duration.sum") or 0.0) * 1e3,
             "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
             "warp_inst": inst, "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
             "threads_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio", False),
             "alu_pipe_pct": val(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", False),
             "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct", False),
             "registers": val(r, "launch__registers_per_thread", False),
             "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
             "stalls_per_issue": {k: round(v, 3) for v, k in st[:5]}}
        e["dram_bytes"] = (e["dram_read_bytes"] or 0.0) + (e["dram_write_bytes"] or 0.0)
        if short.startswith("probe_kernel") and inst:
            e["warp_inst_per_read"] = inst / a.reads_per_launch
        out.append(e)
    with open(a.out, "w") as f:
        json.dump({"report": a.report, "workload": a.workload, "reads_per_launch": a.reads_per_launch,
                   "how": "ncu --set full --clock-control none --import-source on (one launch per kernel; cold-cache, serialised)",
                   "kernels": out}, f, indent=1)
    for e in out:
        print("%-28s %8.3f ms  dram %7.1f MB  inst %s  issue %s%%  thr/inst %s" % (e["name"][:28], e["duration_ms"], e["dram_bytes"] / 1e6,
              "%.3g" % e["warp_inst"] if e["warp_inst"] else "-", e["issue_active_pct"], e["threads_per_inst"]))


if __name__ == "__main__":
    main()
