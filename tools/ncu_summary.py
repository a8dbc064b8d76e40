"""Summaries of ncu captures for profiles/:  python tools/ncu_summary.py full <rep.ncu-rep>  |  launches <launches.csv>"""
import csv, subprocess, sys, io, collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]

def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for val in rows[2:]:
        name = val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:95s} {val[i]} {units[i]}")

def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    iu = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(r[ik], [])
        a.append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:46]:46s} n={len(v):4d} total={sum(v):10.2f} ms  share={100 * sum(v) / tot:5.1f}%  min={min(v):.3f} max={max(v):.3f} ms")

if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
