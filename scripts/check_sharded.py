#!/usr/bin/env python
"""Multi-GPU parity on real GPUs: catalog-sharded, user-data-parallel generation (ShardedGenerator over NCCL) must
produce exactly the paths the single-GPU path (IRSNN.get_seq_in_batch, full catalog) produces for the same users.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from types import SimpleNamespace
import influentialrs_b200 as pkg
from influentialrs_b200.dist import ShardedGenerator

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = SimpleNamespace(n_item=300_007, n_user=1000, max_len=201, n_layers=3, n_heads=4, emb_dim=128, u_emb_dim=10, ffn_dim=256,
                      dropout=0.0, lr1=1e-3)
torch.manual_seed(1234)                                   # same weights on every rank
net = pkg.InfluentialNet(cfg).to(dev).eval()
irn = pkg.IRSNN(cfg, net, dev)
B, P, L = 96, 6, cfg.max_len
g = torch.Generator().manual_seed(99)                     # the same global batch on every rank
seqs = torch.zeros((world * B, L), dtype=torch.long)
for b in range(world * B):
    n = L if b % 3 == 0 else int(torch.randint(30, L + 1, (1,), generator=g))
    seqs[b, L - n:] = torch.randperm(cfg.n_item, generator=g)[:n] + 1
users = torch.randint(0, cfg.n_user, (world * B,), generator=g)
mine = slice(rank * B, (rank + 1) * B)
sg = ShardedGenerator(irn, rank, world)
got, _, _, _ = sg.get_seq_in_batch(seqs[mine].to(dev), users[mine].to(dev), seqs[mine, -1].to(dev), max_path_len=P)
want, _, _, _ = irn.get_seq_in_batch(seqs[mine].to(dev), users[mine].to(dev), seqs[mine, -1].to(dev), max_path_len=P, gap_len=0)
same = torch.tensor([int(np.array_equal(np.asarray(got), np.asarray(want)))], device=dev)
n_diff = torch.tensor([int((np.asarray(got) != np.asarray(want)).any(1).sum())], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN); dist.all_reduce(n_diff)
if rank == 0:
    print(f"sharded ({world} GPUs) vs single-GPU paths: {'IDENTICAL' if int(same) else 'DIFFERENT'} "
          f"({world * B} users x {P} steps, rows that differ: {int(n_diff)})")
dist.destroy_process_group()
sys.exit(0 if int(same) else 1)
