"""Bring-up of the tcgen05 scorer on a real B200: compares against the CUDA-core engine (same
arithmetic after re-scoring) for both descriptor-stride variants and a few shapes; prints timings."""
import sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from influentialrs_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)

def case(M, N, d, Lx, variant, time_it=False):
    g = torch.Generator().manual_seed(M + N + d)
    h = torch.randn((M, d), generator=g).to(dev)
    W = (torch.randn((N, d), generator=g) / math.sqrt(d)).to(dev)
    bias = (torch.randn(N, generator=g) * 0.1).to(dev)
    excl = torch.randint(1, N + 1, (M, Lx), generator=g).to(dev)
    e = ops.sort_exclusions(excl, N, 1)
    prep = ops.scorer_prepare_weights(W)
    rv, ri = ops.score_topk(h, W, bias, 1, e, 1)
    torch.cuda.synchronize()
    try:
        tv, ti = ops.score_argmax_tc(h, W, prep, bias, e, 1, variant)
        torch.cuda.synchronize()
    except Exception as ex:
        print(f"M={M} N={N} d={d} variant={variant}: EXCEPTION {ex}")
        return False
    same_i = (ti == ri).float().mean().item()
    same_v = (tv == rv).float().mean().item()
    print(f"M={M} N={N} d={d} Lx={Lx} variant={variant}: items equal {same_i:.4f}, values bit-equal {same_v:.4f}, "
          f"max|dv|={float((tv - rv).abs().max()):.3e}")
    if time_it and same_i == 1.0:
        for fn, name in ((lambda: ops.score_argmax_tc(h, W, prep, bias, e, 1, variant), "tcgen05"),
                         (lambda: ops.score_topk(h, W, bias, 1, e, 1), "cuda-core")):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"   {name}: {ms:.3f} ms  -> {2.0 * M * N * d / ms / 1e9:.1f} TFLOP/s algorithmic")
    return same_i == 1.0

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ok = case(128, 2048, 128, 8, variant)
if ok:
    case(300, 5000, 128, 50, variant)
    case(64, 3415, 64, 59, variant)
    case(33, 1000, 30, 5, variant)
    case(4096, 1_000_000, 128, 200, variant, time_it=True)
