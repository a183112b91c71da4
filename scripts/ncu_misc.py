#!/usr/bin/env python
"""Exercises once each the kernels that the bench does not launch (for one ncu --set full capture):
remove, export, unique, partition, staged Adagrad, and the one-sided peer kernels over two fake shards."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommenders_addons_b200 import dynamic_embedding as de  # noqa: E402
from recommenders_addons_b200.dynamic_embedding import variable as V  # noqa: E402
import bench as B  # noqa: E402

dev = torch.device("cuda", 0)
dim, res, Bn = 64, 8_000_000, 1 << 20
gen = torch.Generator(device=dev).manual_seed(0)
var = de.Variable(dim=dim, init_size=2 * res, initializer=0.0, num_slot_planes=1, name="ncu_misc")
t = var.tables[0]
for b in range(0, res, 1 << 20):
  r = torch.arange(b, b + (1 << 20), dtype=torch.int64, device=dev)
  t.insert(B.rank_to_key_torch(r), torch.randn(r.numel(), dim, device=dev, generator=gen) * 0.01)
k = B.rank_to_key_torch(torch.randperm(res, device=dev, generator=gen)[:Bn])
g = torch.randn(Bn, dim, device=dev, generator=gen) * 0.01
opt = de.FusedAdagrad(0.01, 0.1)
opt.apply_sparse(var, k, g)                       # apply_staged_kernel<0>
t.accum(k, g, torch.ones(Bn, dtype=torch.bool, device=dev))   # accum_kernel
de.unique(torch.cat([k, k[: Bn // 2]]))            # unique_* kernels
V.partition(k, 8, True)                           # partition_* kernels
newk = B.rank_to_key_torch(torch.arange(3 * res, 3 * res + Bn, device=dev))
t.insert(newk, g)                                  # insert of NEW keys
t.remove(newk)                                     # remove_kernel
t.export()                                         # export_* kernels
shards = [de.Variable(dim=dim, init_size=1 << 22, initializer=0.0, name="ncu_peer%d" % i) for i in range(2)]
pv = de.PeerShardedVariable(fake_shards=shards)
pv.upsert(k, g)                                    # peer_insert_kernel
pv.lookup(k)                                       # peer_find_kernel
pv.attach_inbox(Bn)
pv.route(k, g)                                     # peer_route_kernel
torch.cuda.synchronize()
print("ncu_misc done")
