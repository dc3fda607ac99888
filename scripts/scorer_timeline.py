"""Per-tile timeline of CTA 0 of the fused scorer (debug hook irs_scorer_debug_timeline)."""
import ctypes, sys, math, torch
sys.path.insert(0, ".")
from influentialrs_b200 import ops
from influentialrs_b200._lib import lib
dev = "cuda:0"
N, d, M, Lx = 1_000_000, 128, 4096, 200
g = torch.Generator(device=dev).manual_seed(1)
W = (torch.rand((N, d), generator=g, device=dev) * 2 - 1) / math.sqrt(d)
bias = (torch.rand((N,), generator=g, device=dev) * 2 - 1) / math.sqrt(d)
h = torch.nn.functional.layer_norm(torch.randn((M, d), generator=g, device=dev), (d,))
window = torch.randint(1, N + 1, (M, Lx), generator=g, device=dev)
excl = ops.sort_exclusions(window, N, 1)
prep = ops.scorer_prepare_weights(W)
l = lib()
l.irs_scorer_debug_timeline.argtypes = [ctypes.c_void_p]; l.irs_scorer_debug_timeline.restype = None
for v in (2, 0):
    for _ in range(2): ops.score_argmax_tc(h, W, prep, bias, excl, 1, variant=v)
    tl = torch.zeros((3, 64, 4), dtype=torch.int64, device=dev)
    l.irs_scorer_debug_timeline(tl.data_ptr())
    ops.score_argmax_tc(h, W, prep, bias, excl, 1, variant=v); torch.cuda.synchronize()
    l.irs_scorer_debug_timeline(None)
    t = tl.cpu(); t0 = int(t[t > 0].min())
    print(f"--- variant {v}: tile | MMA: wait_tempty got_tempty issued | EPI: wait_tfull got_tfull done   (cycles from start)")
    for it in range(20, 30):
        m = [int(x) - t0 if x else 0 for x in t[0, it, :3]]; e = [int(x) - t0 if x else 0 for x in t[1, it, :3]]
        print(f"{it:3d} | {m[0]:8d} {m[1]:8d} {m[2]:8d} | {e[0]:8d} {e[1]:8d} {e[2]:8d}   epi busy {e[2]-e[1]:5d} wait {e[1]-e[0]:5d}  mma issue {m[2]-m[1]:5d} wait {m[1]-m[0]:5d}")
