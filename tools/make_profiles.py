"""Turn the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/."""
import csv
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        name, val = r[4], float(r[-1])
        unit = r[-2]
        us = val / 1000.0 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1000.0)
        short = name.split("(")[0].replace("void ", "")[:90]
        per.setdefault(short, []).append(us)
    total = sum(sum(v) for v in per.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)\n")
        f.write("# command: python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras, first 60 kernel launches of the process: "
                "torch input synthesis, then headline launches of 1 048 576 trajectories each (warm-up and timed steps are the same launch)\n")
        f.write(f"{'kernel':92s} {'launches':>8s} {'avg_us':>10s} {'total_us':>10s} {'share':>7s}\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:92s} {len(v):8d} {sum(v)/len(v):10.2f} {sum(v):10.1f} {100*sum(v)/total:6.1f}%\n")
    print("wrote", dst)


def full(rep, dst, traffic_key):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    keep = ("kernel:", "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput",
            "fp64_cycles_active.avg.pct_of_peak_sustained_active", "lsu_wavefronts.avg.pct", "lsu_wavefronts_mem_shared.sum ",
            "wavefronts_mem_lgds", "warps_active", "registers_per_thread ", "issue_active", "issue_stalled", "occupancy_limit",
            "shared_mem_per_block_dynamic", "lts__t_sectors.sum ", "lts__throughput", "lts__t_sector_hit_rate", "inst_executed.sum ",
            "local_op")
    lines = [l for l in out.splitlines() if any(k in l for k in keep)]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch, BASELINE config C3 (262144 x 16 seg)\n")
        f.write("\n".join(lines) + "\n")
    rd = wr = None
    for l in lines:
        if l.startswith("dram__bytes_read.sum "):
            rd = float(l.split()[-1]) * (1e6 if "Mbyte" in l else 1e9 if "Gbyte" in l else 1.0)
        if l.startswith("dram__bytes_write.sum "):
            wr = float(l.split()[-1]) * (1e6 if "Mbyte" in l else 1e9 if "Gbyte" in l else 1.0)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if rd is not None and wr is not None:
        t[traffic_key] = rd + wr
        json.dump(t, open(tpath, "w"), indent=1)
    print("wrote", dst, "traffic", t.get(traffic_key))


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
