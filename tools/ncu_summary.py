"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/<name>.txt
  python tools/ncu_summary.py full gpurun_out/prof_gemm.ncu-rep           > profiles/<name>.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: gpu__time_duration.sum per kernel (ncu --clock-control none; cold-cache, serialised launches -> compare shares)")
    print(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
    for k, (n, t) in agg.items():
        print(f"{k[:72]:72s} {n:8d} {t:10.1f} {100 * t / tot:6.1f}%")
    print(f"{'TOTAL':72s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    print(f"# {path}: selected metrics of `ncu --set full --clock-control none` (one block per captured launch)")
    for row in r[2:]:
        print("-" * 100)
        print(row[hdr.index("Kernel Name")])
        for i, h in enumerate(hdr):
            if (h in KEEP or any(h.endswith("." + k) for k in KEEP)) and row[i] != "":
                print(f"  {h.split('TriageCompute.')[-1]:92s} {units[i]:16s} {row[i]}")


def traffic(path):
    """JSON {kernel name: {"dram_bytes_per_launch": read + write, "launches_captured": n}} for bench.py's
    roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    agg = {}
    for row in r[2:]:
        name = re.sub(r"\(.*", "", row[ik]).replace("void ", "")
        b = float(row[ir].replace(",", "")) * scale[units[ir]] + float(row[iw].replace(",", "")) * scale[units[iw]]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += b
        a[1] += 1
    print(json.dumps({k: {"dram_bytes_per_launch": v[0] / v[1], "launches_captured": v[1]} for k, v in agg.items()}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
