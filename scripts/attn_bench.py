"""Time the attention kernels at the cfg3 decoder shape."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from influentialrs_b200 import ops
dev = "cuda:0"
B, L, H, dh = int(os.environ.get("B", 4096)), 201, 4, 32
d = H * dh
qkv = torch.randn((B, L, 3 * d), device=dev)
ids = torch.randint(1, 1000, (B, L), device=dev)
r_u = torch.randn(B, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
from influentialrs_b200._lib import lib
images = ops.qkv_to_images(qkv[..., :d], qkv[..., d:2*d], qkv[..., 2*d:], (3*d, 3*d, 3*d), B, L, H, 0)
ms = timeit(lambda: ops.pim_attention_img(images, ids, r_u, B, L, H, 0))
ms1 = timeit(lambda: ops.pim_attention_img(images, ids, r_u, B, L, H, 0, q_row0=L - 2, n_q=1))
msc = timeit(lambda: ops.qkv_to_images(qkv[..., :d], qkv[..., d:2*d], qkv[..., 2*d:], (3*d, 3*d, 3*d), B, L, H, 0))
print(f"image kernel: full {ms:.3f} ms ({images.numel()/ms/1e6:.0f} GB/s image read), one-row {ms1:.3f} ms, fp32->image conversion {msc:.3f} ms")
for tc in (True, "one-cta-per-head", False):
    ops.USE_IMG_ATTENTION = tc is True
    ops.USE_TC_ATTENTION = bool(tc)
    ms = timeit(lambda: ops.pim_attention(qkv, ids, r_u, H, 0))
    ms1 = timeit(lambda: ops.pim_attention(qkv, ids, r_u, H, 0, q_row0=L - 2, n_q=1))
    fl = 4.0 * B * H * L * L * dh / 2
    print(f"tc={tc}: full {ms:.3f} ms ({fl / ms / 1e9:.1f} TF causal-alg, {B*L*4*d*4/ms/1e6:.0f} GB/s), one-row {ms1:.3f} ms")
print("error flag", int(ops._error_flag(torch.device(dev)).item()))
