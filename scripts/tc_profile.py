"""Run the tcgen05 scorer a few times at the cfg3 shape (for ncu)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from influentialrs_b200 import ops
dev = "cuda:0"
M, N, d, Lx = int(os.environ.get("M", 4096)), int(os.environ.get("N", 1_000_000)), 128, 200
g = torch.Generator().manual_seed(1)
h = torch.randn((M, d), generator=g).to(dev)
W = (torch.randn((N, d), generator=g) / math.sqrt(d)).to(dev)
bias = (torch.randn(N, generator=g) * 0.1).to(dev)
excl = torch.randint(1, N + 1, (M, Lx), generator=g).to(dev)
e = ops.sort_exclusions(excl, N, 1)
prep = ops.scorer_prepare_weights(W)
for _ in range(4):
    ops.score_argmax_tc(h, W, prep, bias, e, 1, 0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.score_argmax_tc(h, W, prep, bias, e, 1, 0)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"tcgen05 scorer M={M} N={N}: {ms:.3f} ms -> {2.0 * M * N * d / ms / 1e9:.1f} TFLOP/s algorithmic, {3 * 2.0 * M * N * d / ms / 1e9:.1f} issued")
