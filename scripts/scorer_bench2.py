"""Scorer main-kernel floor: with / without exclusion lists and bias (variant bit 1 = single MMA, bit 2 = two producer warps, bit 3 = one K chunk per ring slot instead of two in single-MMA mode)."""
import sys, math, torch
sys.path.insert(0, ".")
from influentialrs_b200 import ops
dev = "cuda:0"
N, d, M, Lx = 1_000_000, 128, 4096, 200
g = torch.Generator(device=dev).manual_seed(1)
W = (torch.rand((N, d), generator=g, device=dev) * 2 - 1) / math.sqrt(d)
bias = (torch.rand((N,), generator=g, device=dev) * 2 - 1) / math.sqrt(d)
h = torch.nn.functional.layer_norm(torch.randn((M, d), generator=g, device=dev), (d,))
window = torch.randint(1, N + 1, (M, Lx), generator=g, device=dev)
excl = ops.sort_exclusions(window, N, 1)
prep = ops.scorer_prepare_weights(W)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for v in (2, 6, 10, 14, 0):
    for name, bb, ee in (("bias+excl", bias, excl), ("bias only", bias, None), ("excl only", None, excl), ("neither", None, None)):
        print(f"variant {v} {name:10s}: {t(lambda: ops.score_argmax_tc(h, W, prep, bb, ee, 1, variant=v)):.3f} ms")
