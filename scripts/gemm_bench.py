"""Time the tcgen05 linear layers at the cfg3 decoder shapes against cuBLAS fp32."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from influentialrs_b200 import ops
dev = "cuda:0"
R = int(os.environ.get("R", 4096 * 201))
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (K, Nout, epi, name) in [(128, 256, 0, "in_proj q,k"), (128, 128, 0, "in_proj v"), (128, 128, 2, "out_proj+LN1+LN2"),
                             (128, 256, 1, "linear1+relu"), (256, 128, 2, "linear2+LN3")]:
    A = torch.randn((R, K), device=dev)
    W = torch.randn((Nout, K), device=dev) / math.sqrt(K)
    bias = torch.randn(Nout, device=dev)
    resid = torch.randn((R, Nout), device=dev)
    v = [torch.randn(Nout, device=dev) for _ in range(5)]
    prep = ops.linear_prepare(W)
    out = torch.empty((R, Nout), device=dev)
    if epi == 2:
        f = lambda: ops.linear_tc(A, prep, Nout, bias, 2, resid=resid, g1=v[0], b1=v[1], c2=v[2], g2=v[3], b2=v[4], out=out)
        nbytes = R * (K + 2 * Nout) * 4
    else:
        f = lambda: ops.linear_tc(A, prep, Nout, bias, epi, out=out)
        nbytes = R * (K + Nout) * 4
    ms = timeit(f)
    ms_ref = timeit(lambda: torch.nn.functional.linear(A, W, bias))
    print(f"{name:18s} K={K} Nout={Nout}: tcgen05 {ms:.3f} ms ({nbytes / ms / 1e6:.0f} GB/s, {2.0 * R * K * Nout / ms / 1e9:.0f} TF alg) | cuBLAS fp32 {ms_ref:.3f} ms")
print("error flag", int(ops._error_flag(torch.device(dev)).item()))
