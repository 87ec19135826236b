#!/usr/bin/env python
"""Executed warp-instructions per SOURCE LINE of one kernel: joins `ncu --page source --csv` (per-SASS-instruction
counts, in program order) with `nvdisasm --print-line-info` of the same build (built with -lineinfo).

  python scripts/sass_by_line.py REPORT.ncu-rep probe_kernelILi1E --per 2000000 [--so nimble_b200/libnimble_b200.so]
"""
import argparse, csv, io, os, re, subprocess, tempfile, collections


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report"); ap.add_argument("mangled_substr")
    ap.add_argument("--per", type=float, default=1.0, help="divide counts by this (reads per launch)")
    ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nimble_b200", "libnimble_b200.so"))
    ap.add_argument("--kernel-regex", default=None)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--by-samples", action="store_true", help="order by warp-stall samples (where the time goes) instead of executed instructions")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=d, capture_output=True)
        cub = max((os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")), key=os.path.getsize)
        dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and a.mangled_substr in l)
    lines, cur = [], None        # (source line, sass text) in program order
    for l in dis[start + 1:]:
        if l.startswith("\t.section") or (l.startswith(".text.") and a.mangled_substr not in l):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append((cur, m.group(2).strip()))
    rx = a.kernel_regex or ("regex:" + re.sub(r"ILi(\d)E.*", "", a.mangled_substr))
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # several kernels may match (template instances): split at header rows, pick the one whose length matches
    blocks, curb = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            curb = {"name": r[1], "rows": []}; blocks.append(curb); continue
        if curb is not None and len(r) > 6 and r[0].startswith("0x"):
            curb["rows"].append(r)
    hdr = next(r for r in rows if r and r[0] == "Address")
    ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    isamp = hdr.index("# Samples")
    blk = min(blocks, key=lambda b: abs(len(b["rows"]) - len(lines)))
    n = min(len(blk["rows"]), len(lines))
    print("kernel:", blk["name"][:80], "| sass in report", len(blk["rows"]), "| sass in disassembly", len(lines))
    per = collections.OrderedDict()
    tot = 0
    for i in range(n):
        key = lines[i][0]
        e = int(blk["rows"][i][ie]); t = int(blk["rows"][i][it])
        v = per.setdefault(key, [0, 0, 0, 0])
        v[0] += e; v[1] += t; v[2] += 1; v[3] += int(blk["rows"][i][isamp] or 0)
        tot += e
    print("total warp-inst / unit: %.1f" % (tot / a.per))
    src = {}
    samples = sum(v[3] for v in per.values()) or 1
    order = (lambda kv: -kv[1][3]) if a.by_samples else (lambda kv: -kv[1][0])
    for key, v in sorted(per.items(), key=order)[:a.top]:
        if key is None:
            print("%8.1f  (no line)" % (v[0] / a.per)); continue
        f, ln = key
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(a.so)), "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:110] if 0 < ln <= len(src[f]) else ""
        print("%8.1f  thr/inst %4.1f  sass %3d  stall-samples %4.1f%%  %s:%d  %s" % (v[0] / a.per, v[1] / max(1, v[0]), v[2], 100.0 * v[3] / samples, f, ln, text))


if __name__ == "__main__":
    main()
