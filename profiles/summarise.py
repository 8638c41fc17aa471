"""Turn the ncu outputs of one profiling pass into the tracked summaries under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_v9.csv profiles/r1_ncu_launch_shares_v9.json "<command>"
    python profiles/summarise.py full gpurun_out/prof_v9.ncu-rep profiles/r1_ncu_full_v9_summary.json [profiles/traffic.json C4 8]

`launches`: the `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list -> per-kernel
launch count, total time and share.  `full`: an `ncu --set full` report -> the metrics DESIGN.md quotes
per captured launch, and (optionally) traffic.json = DRAM bytes of the k_rad passes of one batch.
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def short(name):
    name = re.split(r"[(<]", name)[0].strip()
    return name[5:] if name.startswith("void ") else name  # templated kernels print their return type


def launches(path, out, command):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    head = rows[0]
    kn, mn, mv, mu = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    per = {}
    for r in rows[1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu].replace("second", "s").replace("usecond", "us"), 1.0) if r[mu] in ("ns", "us", "ms", "s") else {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(r[mu], 1.0)
        k = per.setdefault(short(r[kn]), {"launches": 0, "total_us": 0.0})
        k["launches"] += 1
        k["total_us"] += v
    tot = sum(k["total_us"] for k in per.values())
    for k in per.values():
        k["total_us"] = round(k["total_us"], 1)
        k["share"] = round(k["total_us"] / tot, 4)
    json.dump({"command": command, "unit": "us (sum over all launches in the capture)", "total_us": round(tot, 1), "kernels": per},
              open(out, "w"), indent=1)
    print(json.dumps({k: v["share"] for k, v in per.items()}))


def full(rep, out, traffic=None, config=None, frames=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    res = []
    for r in rows[2:]:
        e = {"Kernel Name": r[col["Kernel Name"]], "Grid Size": r[col["Grid Size"]], "Block Size": r[col["Block Size"]]}
        for m in KEEP:
            if m in col:
                e[f"{m} [{units[col[m]]}]"] = r[col[m]]
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    print(len(res), "launches summarised")
    if traffic:
        def to_bytes(v, u):
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd = wr = 0.0
        n = 0
        parts = []
        for r in rows[2:]:
            if short(r[col["Kernel Name"]]) != "k_rad" or n >= 3:
                continue
            a = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            b = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            rd += a; wr += b; n += 1
            parts.append((round(a / 1e6, 2), round(b / 1e6, 2)))
        json.dump({"kernel": "k_rad", "config": config, "frames_per_launch": int(frames), "dram_bytes_per_launch": int(rd + wr),
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the three k_rad passes of one batch "
                           f"(read, written MB per pass: {parts}); ncu --set full --clock-control none; summary in {out}",
                   "source": out}, open(traffic, "w"), indent=1)
        print("traffic", int(rd + wr))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(*sys.argv[2:])
