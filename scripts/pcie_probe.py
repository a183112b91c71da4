#!/usr/bin/env python
"""How fast is this box's PCIe link in each direction alone and in both directions at once (pinned memory)?"""
import json
import time

import torch

dev = torch.device("cuda", 0)
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps):
    if h2d:
      with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    if d2h:
      with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
  torch.cuda.synchronize()
  dt = (time.perf_counter() - t0) / reps
  return dt


for _ in range(2):
  run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print(json.dumps({"h2d_GBs": n / a / 1e9, "d2h_GBs": n / b / 1e9, "both_each_GBs": n / c / 1e9, "both_total_GBs": 2 * n / c / 1e9}))
# chunked, like the table's host pipelines (8 MiB chunks, 3 streams per direction)
ck = 8 << 20
ss = [torch.cuda.Stream() for _ in range(6)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in range(3):
  for i, off in enumerate(range(0, n, ck)):
    with torch.cuda.stream(ss[i % 3]):
      d_in[off:off + ck].copy_(h_in[off:off + ck], non_blocking=True)
    with torch.cuda.stream(ss[3 + i % 3]):
      h_out[off:off + ck].copy_(d_out[off:off + ck], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(json.dumps({"chunked_both_total_GBs": 2 * n / dt / 1e9}))
