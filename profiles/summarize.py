"""
Turns the gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

  python profiles/summarize.py launches <launches.csv> <out.txt>      # per-step kernel list + shares
  python profiles/summarize.py kernel <prof.ncu-rep> <out.txt>        # key metrics of one --set full capture
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in rows
             if r["Metric Name"] == "gpu__time_duration.sum"]
    starts = [i for i, (n, _) in enumerate(names) if "k_k8_stats" in n or "k_make_keys" in n]
    out = [f"source: {src}  ({len(names)} launches captured, {len(starts)} kernel_values steps)",
           "ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised launches -- compare SHARES",
           ""]
    if len(starts) >= 2:
        s, e = starts[-2], starts[-1]
        agg, tot = collections.OrderedDict(), 0.0
        for n, t in names[s:e]:
            k = n.split("(")[0][:60]
            agg.setdefault(k, [0, 0.0])
            agg[k][0] += 1
            agg[k][1] += t
            tot += t
        out.append(f"one kernel_values step (launches {s}..{e - 1}): {e - s} launches, {tot / 1e3:.1f} us of kernel time")
        for k, (c, t) in agg.items():
            out.append(f"{t / 1e3:10.1f} us  x{c:2d}  {100 * t / tot:5.1f}%  {k}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    out = [f"source: {src}  (ncu --set full --clock-control none --import-source on)"]
    for row in rows:
        out.append("")
        out.append(f"launch {row[hdr.index('ID')]}: {row[hdr.index('Kernel Name')][:100]}")
        for m in METRICS:
            if m in hdr:
                out.append(f"  {m:75s} {row[hdr.index(m)]:>18s} {units[hdr.index(m)]}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def traffic(src, dst, units="10000000"):
    """profiles/r2_traffic.json: DRAM bytes and executed FP64 flops per launch of k_interp_cells (bench.py reads it)."""
    import json
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, rows = r[0], [x for x in r[2:] if "k_interp_cells" in x[r[0].index("Kernel Name")]]
    f = lambda row, m: float(row[hdr.index(m)].replace(",", ""))
    n, u = len(rows), float(units)
    dram = sum(f(x, "dram__bytes_read.sum") + f(x, "dram__bytes_write.sum") for x in rows) / n
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[1][hdr.index("dram__bytes_read.sum")]]
    def ops(x, op):      # thread-level instruction count (this ncu only exports the per-cycle form of the counter)
        m = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
        if m in hdr:
            return f(x, m)
        return f(x, m + ".per_cycle_elapsed") * f(x, "smsp__cycles_elapsed.avg" if "smsp__cycles_elapsed.avg" in hdr
                                                    else "sm__cycles_elapsed.avg")
    fl = sum(2 * ops(x, "dfma") + ops(x, "dadd") + ops(x, "dmul") for x in rows) / n
    out = {"interp_cells_dram_bytes_per_launch": dram * scale, "executed_flops_per_unit": fl / u, "launches_captured": n,
           "units_per_launch": u, "source": f"ncu --set full --clock-control none capture of bench.py ({src.split('/')[-1]}), "
                                            "dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches"}
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
