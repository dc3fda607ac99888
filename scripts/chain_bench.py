"""Standalone timing of the fused decoder-chain kernel vs the three linear_tc launches + in_proj it replaces."""
import math, sys, torch
sys.path.insert(0, ".")
from influentialrs_b200 import ops
dev = "cuda:0"
d, ffn = 128, 256
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 201
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
P = dict(Wo=rn(d, d) / math.sqrt(d), bo=0.1 * rn(d), g1=1 + 0.1 * rn(d), b1=0.1 * rn(d), c2=0.3 * rn(d), g2=1 + 0.1 * rn(d),
         b2=0.1 * rn(d), W1=rn(ffn, d) / math.sqrt(d), bf1=0.1 * rn(ffn), W2=rn(d, ffn) / math.sqrt(ffn), bf2=0.1 * rn(d),
         g3=1 + 0.1 * rn(d), b3=0.1 * rn(d), Win=rn(3 * d, d) / math.sqrt(d), bin=0.1 * rn(3 * d))
attn, x = rn(R, d), rn(R, d)
prep = ops.decoder_chain_prepare(P["Wo"], P["W1"], P["W2"], P["Win"])
xo = torch.empty_like(x); qkv = torch.empty((R, 3 * d), device=dev)
def fused():
    ops.decoder_chain_tc(attn, x, prep, P["bo"], P["g1"], P["b1"], P["c2"], P["g2"], P["b2"], P["bf1"], P["bf2"], P["g3"], P["b3"],
                         P["bin"], x_out=xo, qkv_out=qkv)
pw = [ops.linear_prepare(P["Wo"]), ops.linear_prepare(P["W1"]), ops.linear_prepare(P["W2"]),
      ops.linear_prepare(P["Win"][:256].contiguous()), ops.linear_prepare(P["Win"][256:].contiguous())]
qkv2 = torch.empty_like(qkv)
def separate():
    y = ops.linear_tc(attn, pw[0], d, P["bo"], 2, resid=x, g1=P["g1"], b1=P["b1"], c2=P["c2"], g2=P["g2"], b2=P["b2"])
    f = ops.linear_tc(y, pw[1], ffn, P["bf1"], 1)
    x2 = ops.linear_tc(f, pw[2], d, P["bf2"], 2, resid=y, g1=P["g3"], b1=P["b3"])
    ops.linear_tc(x2, pw[3], 256, P["bin"][:256].contiguous(), 0, out=qkv2[:, :256])
    ops.linear_tc(x2, pw[4], 128, P["bin"][256:].contiguous(), 0, out=qkv2[:, 256:])
    return x2
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
import ctypes
from influentialrs_b200._lib import lib
L = 201
R = (R // L) * L
attn, x, xo = attn[:R].contiguous(), x[:R].contiguous(), xo[:R].contiguous()
images = ops.qkv_images_buffer(R // L, L, 4, torch.device(dev), slot=5)
def fused_img():
    ops.decoder_chain_tc(attn, x, prep, P["bo"], P["g1"], P["b1"], P["c2"], P["g2"], P["b2"], P["bf1"], P["bf2"], P["g3"], P["b3"],
                         P["bin"], x_out=xo, qkv_images=images, L=L, mask_mode=0)
print(f"image output: {timeit(fused_img):.3f} ms")
qkv = torch.empty((R, 3 * d), device=dev); qkv2 = torch.empty_like(qkv)
tf, ts = timeit(fused), timeit(separate)
x2 = separate(); fused(); torch.cuda.synchronize()
print("max|x diff|", float((x2 - xo).abs().max()), "max|qkv diff|", float((qkv2 - qkv).abs().max()))
byt = R * d * 4 * (2 + 1 + 3)
flop = 2.0 * R * (d * d + d * ffn * 2 + d * 3 * d)
print(f"R={R}: fused {tf:.3f} ms ({byt / tf / 1e6:.0f} GB/s algorithmic, {3 * flop / tf / 1e9:.0f} TF issued), separate {ts:.3f} ms, "
      f"error_flag={int(ops._error_flag(torch.device(dev)).item())}")
