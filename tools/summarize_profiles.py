#!/usr/bin/env python3
"""Turns the raw ncu exports of a gpurun call into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.md
    python tools/summarize_profiles.py full gpurun_out/prof_r01.ncu-rep profiles/r01_ncu_full_summary.json

`launches`: per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list.
`full`: selected metrics of every launch in an `ncu --set full` report (read with `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import io
import json
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def launches(src, dst):
    """Launch list with gpu__time_duration.sum and, optionally, smsp__inst_executed.sum per launch."""
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "")
        v = float(r[mv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, 0.0])
        if r[mn] == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += {"ns": v * 1e-6, "us": v * 1e-3, "ms": v, "s": v * 1e3}.get(r[mu], v)
        elif r[mn] == "smsp__inst_executed.sum":
            a[2] += v
    tot = sum(v[1] for v in agg.values())
    toti = sum(v[2] for v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | total ms | share | avg us | M warp-instr | instr share |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for name, (n, ms, ins) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.3f | %.1f | %.1f | %.3f |\n" % (name, n, ms, ms / tot, ms / n * 1e3, ins / 1e6,
                                                                         ins / toti if toti else 0.0))
        f.write("| **total** | %d | %.3f | 1.000 | | %.1f | |\n" % (sum(v[0] for v in agg.values()), tot, toti / 1e6))
    print(open(dst).read())


def full(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].replace("void ", "").rstrip("<")}
        for m in FULL_METRICS:
            if m in hdr:
                d[m] = r[hdr.index(m)]
        d["units"] = {m: units[hdr.index(m)] for m in FULL_METRICS if m in hdr}
        res.append(d)
    json.dump(res, open(dst, "w"), indent=1)
    for d in res:
        print(d["kernel"], {k: v for k, v in d.items() if k not in ("kernel", "units")})


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
