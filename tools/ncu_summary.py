"""Turns gpurun_out/*.ncu-rep and launches.csv into the small text summaries committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1b.ncu-rep profiles/r1_ncu_full.md [frames_in_capture]
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r1_launches.md
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full(rep, out, frames):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s (%d windows in the captured batch)\n\n" % (os.path.basename(rep), frames))
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name)
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            f.write("| DRAM read+write per window | %.0f | byte |\n\n" % ((rd + wr) / frames))
            traffic[name + "_dram_bytes_per_window"] = (rd + wr) / frames
    return traffic


def launches(csvf, out):
    rows = [r for r in csv.reader(open(csvf)) if len(r) > 5]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[1:]:
        try:
            name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
            v = float(r[idx["Metric Value"]])
            unit = r[idx["Metric Unit"]]
        except (ValueError, KeyError, IndexError):
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, a[0], a[1], 100 * a[1] / tot))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        t = full(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
        print(json.dumps(t, indent=1))
