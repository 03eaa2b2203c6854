#!/usr/bin/env python
"""Turn an `ncu --set full` report into the per-kernel summary CSV kept under profiles/ (one row per
captured launch: duration, DRAM bytes, L2 hit rate, tensor-pipe activity, occupancy limiters).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_ncu_full_x.csv
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    units = rows[1]
    keep = ["ID", "Kernel Name", "Grid Size", "Block Size"] + [m for m in METRICS if m in hdr]
    idx = [hdr.index(k) for k in keep]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if len(r) == len(hdr):
                w.writerow([r[i] for i in idx])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
