#!/usr/bin/env python
"""Probe (torchrun, 2 ranks): can torch symmetric memory (CUDA VMM + handle exchange) be used for the peer-mapped
table planes, and how fast is RANDOM 256 B row access to a peer's buffer through such a mapping?"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
  rank = int(os.environ["RANK"])
  world = int(os.environ["WORLD_SIZE"])
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", device_id=dev)
  out = {"rank": rank}
  try:
    import torch.distributed._symmetric_memory as symm_mem
    rows, dim = 32 * 1024 * 1024, 64  # 8 GiB of fp32 rows
    t0 = time.time()
    buf = symm_mem.empty((rows, dim), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    out["alloc_s"] = round(time.time() - t0, 2)
    out["buffer_ptrs"] = [hex(p) for p in hdl.buffer_ptrs]
    buf.normal_()
    torch.cuda.synchronize()
    dist.barrier()
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (rows, dim), torch.float32)
    idx = torch.randint(0, rows, (1 << 20,), device=dev)
    for name, src in (("local", buf), ("remote", remote)):
      for _ in range(3):
        r = src.index_select(0, idx)
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(5):
        r = src.index_select(0, idx)
      b.record()
      torch.cuda.synchronize()
      ms = a.elapsed_time(b) / 5
      out["gather_%s_ms" % name] = round(ms, 4)
      out["gather_%s_GBs" % name] = round((1 << 20) * dim * 4 / ms / 1e6, 1)
    # streaming copy from the peer buffer
    dst = torch.empty(1 << 26, dtype=torch.float32, device=dev)
    flat = remote.reshape(-1)[: 1 << 26]
    for _ in range(2):
      dst.copy_(flat)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    dst.copy_(flat)
    b.record()
    torch.cuda.synchronize()
    out["stream_remote_GBs"] = round((1 << 28) / a.elapsed_time(b) / 1e6, 1)
    dist.barrier()
  except Exception as e:  # noqa: BLE001
    out["error"] = repr(e)
  print(json.dumps(out), flush=True)
  dist.barrier()
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
