"""Summarise an `ncu --set full` report (.ncu-rep) as JSON: one entry per captured launch with the metrics the
roofline discussion uses (profiles/README.md).  Runs on the CPU box:

    python tests/ncu_summarize.py gpurun_out/r2_ring.ncu-rep > profiles/r2_ring_kernel_ncu_launches.json
"""
from __future__ import annotations

import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "gpu_time_duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_shared_memory_bytes",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar_per_issue",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    res = []
    for r in rows[hi + 2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in WANT.items():
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    d[name] = float(v)
                    u = units[hdr.index(k)]
                    if u:
                        d[name + "_unit"] = u
                except ValueError:
                    pass
        res.append(d)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
