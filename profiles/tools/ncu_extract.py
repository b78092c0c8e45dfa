"""Reads `ncu -i X.ncu-rep --page raw --csv` from stdin and prints one JSON line per kernel launch with the metrics we keep."""
import csv, json, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    if len(r) < 10:
        continue
    out = {"kernel": r[idx["Kernel Name"]][:90]}
    for k in keep:
        if k in idx:
            out[k] = r[idx[k]] + " " + units[idx[k]]
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i] or 0)) for h, i in idx.items()
          if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    tot = sum(v for _, v in st) or 1.0
    out["stalls_pct"] = {h: round(100 * v / tot, 1) for h, v in sorted(st, key=lambda kv: -kv[1])[:6]}
    for h, i in idx.items():
        if "tensor" in h and "pct" in h and "avg" in h and r[i] not in ("0", "", "n/a") and h not in out:
            out[h] = r[i]
    print(json.dumps(out))
