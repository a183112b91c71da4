#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline discussion uses.
    python scripts/ncu_summary.py gpurun_out/prof_find.ncu-rep > profiles/r01_find_kernel_tma.csv"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
  out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  cols = [w for w in WANT if w in idx]
  w = csv.writer(sys.stdout)
  w.writerow(cols)
  w.writerow([units[idx[c]] for c in cols])
  for r in rows[2:]:
    w.writerow([r[idx[c]] for c in cols])


if __name__ == "__main__":
  main()
