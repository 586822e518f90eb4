"""Summaries of ncu captures for profiles/.
usage: python tools/ncu_summary.py full  report.ncu-rep out.json      (selected metrics of one --set full capture)
       python tools/ncu_summary.py list  launches.csv   out.json      (per-kernel totals/shares of a launch list)"""
import collections, csv, json, subprocess, sys

FULL_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "smsp__inst_executed_op_shared_atom.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct",
]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.split("\n")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for k in ("Kernel Name", "Block Size", "Grid Size"):
        if k in hdr:
            d[k] = vals[hdr.index(k)]
    for k in FULL_KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": vals[i], "unit": units[i]}
    json.dump(d, open(out, "w"), indent=1)
    print(out, len(d), "metrics")


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        t = float(r[iv].replace(",", ""))
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        name = r[ik].split("(")[0][:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    res = [{"kernel": k, "launches": v[0], "total_us": round(v[1], 3), "share": v[1] / tot}
           for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump(res, open(out, "w"), indent=1)
    for r in res[:6]:
        print(r)


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
