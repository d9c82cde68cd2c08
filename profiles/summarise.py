#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (ncu launch list) and *.ncu-rep (full captures) into small text files.

  python profiles/summarise.py <tag>      ->  profiles/<tag>_launches.txt, profiles/<tag>_<kernel>.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled"]


def launches(tag):
    f = OUT / "launches.csv"
    if not f.exists():
        return
    lines = [l for l in f.read_text().splitlines() if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("\n".join(lines)))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name).split("::")[-1][:60]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(ROOT / "profiles" / f"{tag}_launches.txt", "w") as fh:
        fh.write(f"# ncu launch list ({sum(cnt.values())} launches, {total:.1f} ms device time; cold-cache, serialised: compare SHARES)\n")
        fh.write(f"{'kernel':62s} {'launches':>8s} {'ms':>10s} {'share':>7s}\n")
        for k in sorted(tot, key=tot.get, reverse=True):
            fh.write(f"{k:62s} {cnt[k]:8d} {tot[k]:10.3f} {100*tot[k]/total:6.1f}%\n")
    print(open(ROOT / "profiles" / f"{tag}_launches.txt").read())


def full(tag):
    traffic = {}  # workload -> kernel -> mean dram bytes (read + write) per captured launch; bench.py reports it as roofline.traffic
    for rep in sorted(OUT.glob("prof_*.ncu-rep")):
        r = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        out = [f"# {rep.name}: ncu --set full --clock-control none, per launch"]
        per_kernel = defaultdict(list)
        for row in rows[2:]:
            try:
                def val(name):
                    i = hdr.index(name)
                    v = float(row[i].replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
                kname = re.sub(r"<.*|\(.*", "", row[hdr.index("Kernel Name")]).split("::")[-1].replace("void ", "").strip()
                per_kernel[kname].append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
            except (ValueError, IndexError):
                pass
            out.append(f"## launch id {row[0]} {row[hdr.index('Kernel Name')][:80]}")
            for i, h in enumerate(hdr):
                if any(h.startswith(k) for k in KEYS):
                    out.append(f"{h:80s} {row[i]:>18s} {units[i]}")
        (ROOT / "profiles" / f"{tag}_{rep.stem}.txt").write_text("\n".join(out) + "\n")
        print("\n".join(out[:70]))
        # capture files are named prof_<workload>_<kernel pattern>.ncu-rep
        parts = rep.stem.split("_")
        if len(parts) >= 3:
            for k, v in per_kernel.items():
                v = [x for x in v if x == x]  # a launch whose counters ncu could not read comes back as nan
                if v:
                    traffic.setdefault(parts[1], {})[k] = sum(v) / len(v)
    if traffic:
        import json

        f = ROOT / "profiles" / "ncu_traffic.json"
        old = json.load(open(f)) if f.exists() else {}
        for w, d in traffic.items():
            old.setdefault(w, {}).update(d)
        json.dump(old, open(f, "w"), indent=1, sort_keys=True)
        print("wrote", f)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    full(tag)
