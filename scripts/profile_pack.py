#!/usr/bin/env python
"""Turn one GPU call's raw outputs (gpurun_out/<prefix>_*) into the tracked files under profiles/:

  <tag>_launches.csv            the ncu launch list (gpu__time_duration.sum per launch) as captured
  <tag>_step_table.md           one step of the workload: launches, ms and share per kernel (from the launch list)
  <tag>_ncu_full_metrics.txt    key raw-page metrics of every kernel in the `ncu --set full` report
  <tag>_ncu_summary.json        scripts/ncu_summary.py (bench.py reads roofline.traffic from the r02 one)
  <tag>_sass_<kernel>.txt       cuobjdump -sass of the hot kernels in the built .so (instruction mix + the DPX / REDUX /
                                256-bit load lines that prove what the inner loops issue)

    python scripts/profile_pack.py --prefix gpurun_out/r2j --tag r02 [--batches 5]
Needs ncu and cuobjdump on PATH (reads reports and the .so; no GPU)."""
import argparse
import collections
import csv
import io
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|smsp__inst_executed\.sum$|"
                  r"smsp__issue_active\.avg\.pct|smsp__thread_inst_executed_per_inst_executed\.ratio|sm__warps_active\.avg\.pct|"
                  r"launch__(registers_per_thread|grid_size|block_size|occupancy_limit)|lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|"
                  r"sm__pipe_alu_cycles_active\.avg\.pct|sm__inst_executed_pipe_(alu|lsu|fma|xu|uniform)|sm__throughput\.avg\.pct|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|lts__t_bytes\.sum$|l1tex__t_bytes\.sum$|"
                  r"smsp__warps_eligible\.avg\.per_cycle_active|sm__cycles_elapsed\.max)")


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("nb200::", "")
    return re.sub(r"cub::CUB_\w+::detail::", "cub::", n)


def step_table(launch_csv, batches, out):
    rows = list(csv.reader(open(launch_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    idx = [i for i, r in enumerate(data) if "probe_kernel" in r[ki]]
    # step 4 of the run (3 warm-up steps before it): launches from its first probe to the next step's first probe
    a, b = idx[3 * batches], idx[4 * batches]
    agg = collections.OrderedDict()
    for r in data[a:b]:
        t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[r[ui]]
        e = agg.setdefault(short(r[ki])[:80], [0, 0.0])
        e[0] += 1; e[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("One timed step of the headline workload (launch list, ncu gpu__time_duration.sum: cold-cache, serialised —\ncompare SHARES with the live CUDA-event stage times in the bench JSON, not absolutes)\n\n")
        f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, v[0], v[1], 100 * v[1] / tot))
        f.write("| total | %d | %.3f | |\n" % (b - a, tot))
    return tot


def full_metrics(rep, out, header):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(header + "\n")
        for r in data:
            f.write("\n== %s  grid %s block %s\n" % (short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for i, h in enumerate(hdr):
                if KEEP.match(h) and r[i] not in ("", "n/a"):
                    f.write("%-86s %s %s\n" % (h, r[i], units[i]))


def sass(so, kernel_regex, out, title):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)
    picked = [b for b in blocks[1:] if re.search(kernel_regex, b.split("\n", 1)[0])]
    with open(out, "w") as f:
        f.write("# %s\n# cuobjdump -sass %s, functions matching /%s/\n" % (title, os.path.relpath(so, ROOT), kernel_regex))
        for b in picked:
            name = b.split("\n", 1)[0]
            ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b)
            mix = collections.Counter(i.split(".")[0] for i in ins)
            f.write("\n== %s\n%d instructions; mix: %s\n" % (name, len(ins), ", ".join("%s %d" % kv for kv in mix.most_common(24))))
            f.write("-- lines with DPX / REDUX / 256-bit loads / votes:\n")
            for line in b.split("\n"):
                if re.search(r"VIADDMNMX|VIMNMX|REDUX|LDG\.E\.(256|ENL2\.256)|\.256|VOTE|MATCH|LDGSTS|UBLKCP", line):
                    f.write(line.rstrip()[:150] + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prefix", required=True)
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--batches", type=int, default=5, help="probe launches per step (10 M reads = 5 batches of 2 M)")
    ap.add_argument("--reads-per-launch", type=int, default=2_000_000)
    a = ap.parse_args()
    P = os.path.join(ROOT, "profiles")
    lc = a.prefix + "_launches.csv"
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(P, a.tag + "_launches.csv"))
        tot = step_table(lc, a.batches, os.path.join(P, a.tag + "_step_table.md"))
        print("step table: %.2f ms per step under ncu" % tot)
    rep = a.prefix + "_full.ncu-rep"
    if os.path.exists(rep):
        full_metrics(rep, os.path.join(P, a.tag + "_ncu_full_metrics.txt"),
                     "# ncu --set full --clock-control none --import-source on, NB200_BENCH_READS=%d python bench.py --steps 1 --warmup 3 "
                     "--no-cpu-baseline --hbm-transcripts 0 (one launch per kernel; see scripts/gpu_*.sh of the round)" % a.reads_per_launch)
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "--workload", "cfg2", "--reads-per-launch",
                        str(a.reads_per_launch), "-o", os.path.join(P, a.tag + "_ncu_summary.json")], check=True)
    so = os.path.join(ROOT, "nimble_b200", "libnimble_b200.so")
    sass(so, r"sw_kernel", os.path.join(P, a.tag + "_sass_sw_kernel.txt"), "banded Smith-Waterman: DPX instructions of the row recurrence")
    sass(so, r"probe_kernel.*Li1ELb0", os.path.join(P, a.tag + "_sass_probe_kernel.txt"), "probe_kernel<1,false>: 256-bit table loads, REDUX intersections")


if __name__ == "__main__":
    main()
