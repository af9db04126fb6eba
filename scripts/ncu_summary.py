#!/usr/bin/env python
"""Summarise `ncu --set full` reports into a small JSON (one entry per profiled launch).

    python scripts/ncu_summary.py <tag> <workload> report.ncu-rep [...]  ->  profiles/<tag>_ncu_<workload>.json

The JSON keeps what the roofline discussion needs (duration, DRAM bytes, issue utilisation, registers,
top warp-stall reasons). `bench.py` reads `dram_bytes` from `profiles/traffic.json` (written by --traffic)
for the `roofline.traffic` field. Runs here (no GPU needed): ncu only imports the report.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "lts__t_bytes.sum": "l2_bytes",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3,
        "nsecond": 1e-3, "ns": 1e-3, "second": 1e6}


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")].replace("b200norm::", "")}
        stalls = []
        for i, h in enumerate(hdr):
            if h in KEEP:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                e[KEEP[h]] = v * UNIT.get(units[i], 1)
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        e["dram_bytes"] = e.get("dram_read", 0) + e.get("dram_write", 0)
        e["stalls_per_issue"] = {k: round(v, 3) for v, k in sorted(stalls, reverse=True)[:8]}
        out.append(e)
    return out


def main():
    tag, workload, reports = sys.argv[1], sys.argv[2], sys.argv[3:]
    entries = []
    for rep in reports:
        entries += summarise(rep)
    dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_{workload}.json")
    with open(dst, "w") as f:
        json.dump({"source": [os.path.basename(r) for r in reports],
                   "command": "ncu --set full --clock-control none --import-source on (scripts/gpu_visit2.sh / gpu_ncu.sh)",
                   "note": "per-launch, cold-cache and serialised; DRAM writes are lower than the algorithmic "
                           "output bytes because part of the output is still dirty in the 126 MB L2 when the kernel ends",
                   "launches": entries}, f, indent=1)
    # traffic table consumed by bench.py: workload -> {fwd|bwd: dram bytes per launch}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for e in entries:
        k = "bwd" if "bwd" in e["kernel"] else "fwd"
        table.setdefault(workload, {})[k] = {"dram_bytes": e["dram_bytes"], "kernel": e["kernel"], "from": os.path.basename(dst)}
    with open(tpath, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps(entries, indent=1))


if __name__ == "__main__":
    main()
