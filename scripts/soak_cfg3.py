"""Full cfg3 job: 100k users x 20 path steps at N = 1M, L = 201, on one GPU or (under torchrun) on N GPUs with the catalog
sharded -- checks the size-independent properties of the generated paths (items in range; the pick of step i is never an item
of the window AT THAT STEP, i.e. of original_history[i:] or an earlier pick -- the i oldest items have slid out and are eligible
again; zeroed after the first hit of the target) and that no kernel watchdog fired.

    python scripts/soak_cfg3.py [users]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/soak_cfg3.py [users]
"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev, timeout=__import__("datetime").timedelta(seconds=180))
cfg = dict(bench.CFG3)
pkg, c, net, irn = bench.build_irn(cfg, dev)
U = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
U_loc = (U + world - 1) // world
U_loc = min(U_loc, max(0, U - rank * U_loc))                 # ragged last rank
gen = torch.Generator(device=dev).manual_seed(99 + rank)
seqs, users = bench.synth_batch(max(U_loc, 1), cfg, gen, dev)
seqs, users = seqs[:U_loc], users[:U_loc]
targets = seqs[:, -1].clone()
api = irn
if world > 1:
    from influentialrs_b200.dist import ShardedGenerator
    api = ShardedGenerator(irn, rank, world)
    api.get_seq_in_batch(seqs[:64], users[:64], targets[:64], max_path_len=2)      # warm-up (weight images, NCCL)
    dist.barrier()
torch.cuda.synchronize(); t0 = time.time()
paths, tg, hist, n_early = api.get_seq_in_batch(seqs, users, targets, max_path_len=20, gap_len=0)
torch.cuda.synchronize(); dt = time.time() - t0
if world > 1:
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
assert paths.shape == (U_loc, 20)
assert int(pkg.ops._error_flag(dev).item()) == 0
N = cfg["n_item"]
assert U_loc == 0 or (paths.min() >= 0 and paths.max() <= N)
win = seqs[:, :-1].cpu().numpy()
bad = 0
for b in range(0, U_loc, max(1, U_loc // 2000)):            # sample of users for the O(P*L) membership checks
    p = paths[b][paths[b] > 0].astype(np.int64)
    assert len(set(p.tolist())) == len(p), "repeated item in a path"
    t = int(tg[b])
    if t in p.tolist():
        k = p.tolist().index(t)
        assert (paths[b][k + 1:] == 0).all()
    for i_, item in enumerate(p.tolist()):                  # the window at step i = original[i:] + the picks so far
        bad += int(item in win[b][i_:])
assert bad == 0, "a generated item was in the user's window at the step it was generated"
if rank == 0:
    print(f"soak ok: {U} users x 20 steps on {world} GPU(s) in {dt:.2f} s = {U * 20 / dt:.0f} user-steps/s end to end "
          f"(host buffers out, max over ranks), rank 0: {n_early} early successes, histories returned: {len(hist)}")
if world > 1:
    dist.destroy_process_group()
