"""Distil `ncu --page raw --csv` exports and launch lists into the small tracked files under profiles/.
usage: python tools/summarise_ncu.py raw <raw.csv> <out.csv>      (one row per captured launch, key metrics only)
       python tools/summarise_ncu.py launches <launches.csv> <out.csv>   (per-kernel share of summed device time)"""
import csv
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def raw(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        cols = [k for k in KEYS if k in idx]
        w.writerow(["launch", "kernel"] + [f"{k} [{units[idx[k]]}]" for k in cols])
        for n, r in enumerate(rows[2:]):
            w.writerow([n, r[name_i][:90]] + [r[idx[k]] for k in cols])


def launches(src, dst):
    lines = open(src).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = OrderedDict()
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"])
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        k = r["Kernel Name"].split("(")[0][:80]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        f.write("# per-kernel share of summed device time; ncu times are cold-cache and serialised: compare SHARES, not absolutes\n")
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1] / tot:.4f}"])


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
