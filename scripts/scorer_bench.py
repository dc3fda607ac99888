import sys, math, torch
sys.path.insert(0, ".")
from influentialrs_b200 import ops
dev = "cuda:0"
N, d, M, Lx = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000), 128, (int(sys.argv[2]) if len(sys.argv) > 2 else 4096), 200
print(f'N={N} M={M}')
g = torch.Generator(device=dev).manual_seed(1)
W = (torch.rand((N, d), generator=g, device=dev) * 2 - 1) / math.sqrt(d)
bias = (torch.rand((N,), generator=g, device=dev) * 2 - 1) / math.sqrt(d)
h = torch.nn.functional.layer_norm(torch.randn((M, d), generator=g, device=dev), (d,))
SPREAD = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # window ids drawn from SPREAD catalogs: a shard sees 1/SPREAD of them
window = torch.randint(1, SPREAD * N + 1, (M, Lx), generator=g, device=dev)
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
for v in (0, 2):
    print(f"variant {v}: {t(lambda: ops.score_argmax_tc(h, W, prep, bias, excl, 1, variant=v)):.3f} ms")
a = ops.score_argmax_tc(h, W, prep, bias, excl, 1, variant=0)[1]
b = ops.score_argmax_tc(h, W, prep, bias, excl, 1, variant=2)[1]
print("same winners:", bool(torch.equal(a, b)))
