#!/usr/bin/env python
"""cfg4 data-parallel: IRN train_batch (incl. gradient all-reduce + Adam) with the batch split over the ranks.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_train_dp.py [--global-batch 4096]
Fixed global batch (strong scaling) by default; --per-gpu-batch B for weak scaling.  One JSON line on rank 0."""
import argparse, json, os, sys
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import influentialrs_b200 as pkg
from influentialrs_b200.dist import make_data_parallel

ap = argparse.ArgumentParser()
ap.add_argument("--global-batch", type=int, default=4096)
ap.add_argument("--per-gpu-batch", type=int, default=0)
ap.add_argument("--n-item", type=int, default=500_000)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = args.per_gpu_batch or args.global_batch // world
L, d = 50, 128
c = SimpleNamespace(n_item=args.n_item, n_user=100_000, max_len=L, n_layers=6, n_heads=4, emb_dim=d, u_emb_dim=10, ffn_dim=256,
                    dropout=0.0, lr1=1e-3)
torch.manual_seed(1234)                      # identical replicas
net = pkg.InfluentialNet(c).to(dev)
irn = pkg.IRSNN(c, net, dev)
if world > 1:
    make_data_parallel(irn)
g = torch.Generator().manual_seed(1234 + rank)
seqs = torch.randint(1, args.n_item + 1, (B, L), generator=g).to(dev)
users = torch.randint(0, c.n_user, (B,), generator=g).to(dev)
losses = [irn.train_batch(seqs, users)]      # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    losses.append(irn.train_batch(seqs, users))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # replicas must stay identical: compare a parameter checksum across ranks
    cs = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    in_sync = bool((hi - lo).abs().item() <= 1e-9 * max(1.0, abs(hi.item())))
else:
    in_sync = True
if rank == 0:
    print(json.dumps({"config": f"cfg4 DP: IRN train_batch, N={args.n_item}, L={L}, d={d}, batch {B} per GPU x {world} GPUs",
                      "metric": "train steps/s", "value": 1e3 / ms, "ms_per_step": ms, "samples_per_s": B * world * 1e3 / ms,
                      "n_gpus": world, "scaling": "weak" if args.per_gpu_batch else "strong", "replicas_in_sync": in_sync,
                      "loss_first_last": [losses[0], losses[-1]], "data": "synthetic"}))
if world > 1:
    dist.destroy_process_group()
