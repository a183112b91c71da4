#!/usr/bin/env python
"""DRAM traffic of ONE launch of the headline kernel on the headline workload, for bench.py's roofline.traffic:
runs bench.py (3 steps, no e2e / CPU arm / extras) under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum -k regex:find_kernel_tma`, takes the LAST captured launch (warm table) and writes
gpurun_out/r02_traffic.json = {kernel: {dram_read, dram_write, time_ns, src_sha, when, cmd}}.  Copy it to profiles/ and
commit it together with the kernel sources it was taken from: bench.py refuses a capture whose src_sha is stale.
    gpurun -- 'python scripts/ncu_traffic.py'"""
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  out_dir = os.path.join(ROOT, "gpurun_out")
  os.makedirs(out_dir, exist_ok=True)
  log = os.path.join(out_dir, "r02_traffic_ncu.csv")
  cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
         "-k", "regex:find_kernel_tma", "-c", "6", "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "bench.py"),
         "--steps", "3", "--warmup", "3", "--no-e2e", "--no-cpu-baseline", "--no-hard-cases", "--no-c3"]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
  if r.returncode != 0:
    print(r.stdout[-2000:], r.stderr[-2000:])
    sys.exit(1)
  rows, hdr = [], None
  for rec in csv.reader(open(log)):
    if rec and rec[0] == "ID":
      hdr = rec
    elif hdr and len(rec) == len(hdr):
      rows.append(dict(zip(hdr, rec)))
  by_id = {}
  for d in rows:
    by_id.setdefault(d["ID"], {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    by_id[d["ID"]]["unit:" + d["Metric Name"]] = d["Metric Unit"]
  last = by_id[sorted(by_id, key=int)[-1]]

  def to_bytes(name):
    u = last["unit:" + name].lower()
    return last[name] * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
  rec = {"find_kernel_tma<16>": {"dram_read": to_bytes("dram__bytes_read.sum"), "dram_write": to_bytes("dram__bytes_write.sum"),
                                 "time": last["gpu__time_duration.sum"], "time_unit": last["unit:gpu__time_duration.sum"],
                                 "launches_captured": len(by_id), "src_sha": bench.src_sha(),
                                 "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ"), "cmd": " ".join(cmd[:12]) + " ..."}}
  with open(os.path.join(out_dir, "r02_traffic.json"), "w") as f:
    json.dump(rec, f, indent=1)
  print(json.dumps(rec))


if __name__ == "__main__":
  main()
