"""Full cfg3 job on one GPU: 100k users x 20 path steps at N = 1M, L = 201 -- checks the size-independent properties of
the generated paths (items in range, never an item of the user's window, no repeats within a path before the target,
zeroed after the first hit) and that no kernel watchdog fired."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda:0")
cfg = dict(bench.CFG3)
pkg, c, net, irn = bench.build_model(cfg, dev)
U = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
gen = torch.Generator(device=dev).manual_seed(99)
seqs, users = bench.synth_batch(U, cfg, gen, dev)
targets = seqs[:, -1].clone()
torch.cuda.synchronize(); t0 = time.time()
paths, tg, hist, n_early = irn.get_seq_in_batch(seqs, users, targets, max_path_len=20, gap_len=0)
torch.cuda.synchronize(); dt = time.time() - t0
assert paths.shape == (U, 20)
assert int(pkg.ops._error_flag(dev).item()) == 0
N = cfg["n_item"]
assert paths.min() >= 0 and paths.max() <= N
win = seqs[:, :-1].cpu().numpy()
bad = 0
for b in range(0, U, max(1, U // 2000)):                    # sample of users for the O(P*L) membership checks
    p = paths[b][paths[b] > 0].astype(np.int64)
    assert len(set(p.tolist())) == len(p), "repeated item in a path"
    t = int(tg[b])
    if t in p.tolist():
        k = p.tolist().index(t)
        assert (paths[b][k + 1:] == 0).all()
        p = p[:k]
    bad += int(np.isin(p, win[b]).sum())
assert bad == 0, "a generated item was in the user's window"
print(f"soak ok: {U} users x 20 steps in {dt:.2f} s = {U * 20 / dt:.0f} user-steps/s end to end (host buffers out), "
      f"{n_early} early successes, histories returned: {len(hist)}")
