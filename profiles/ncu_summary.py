#!/usr/bin/env python
"""Turn ncu output into the small tracked summaries under profiles/.

  python profiles/ncu_summary.py raw  <ncu --page raw --csv file>  <out.json>
      one object per profiled kernel with the metrics this repo argues from
  python profiles/ncu_summary.py list <ncu --metrics gpu__time_duration.sum --csv log> <out.txt> "<command>"
      per-kernel totals and shares of a launch list
"""
import csv
import json
import re
import sys
from collections import defaultdict

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def raw(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        o = {"Kernel Name": d["Kernel Name"]}
        for k in names:
            if k in KEEP or k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k:
                o[k] = f"{d[k]} {u.get(k, '')}".strip()
        out.append(o)
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(out)} kernels -> {dst}")


def launch_list(src, dst, command):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    names = rows[0]
    kn, mv, mu = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) != len(names):
            continue
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
        k = re.sub(r"\(.*", "", r[kn])
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none   command: {command}\n"
                "(every launch of this library in the run; cold-cache, serialised launch times: compare SHARES, "
                "not absolutes)\n\n")
        for k in sorted(tot, key=lambda x: -tot[x]):
            f.write(f"{tot[k]:10.3f} ms {cnt[k]:5d}x {100 * tot[k] / total:6.1f}%  avg {tot[k] / cnt[k]:9.3f} ms  {k}\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3])
    else:
        launch_list(sys.argv[2], sys.argv[3], sys.argv[4])
