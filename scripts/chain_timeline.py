"""Phase timeline of CTA 0 of the fused decoder-chain kernel (debug hook irs_decoder_chain_debug_timeline)."""
import ctypes, math, sys, torch
sys.path.insert(0, ".")
from influentialrs_b200 import ops
from influentialrs_b200._lib import lib
dev = "cuda:0"
d, ffn = 128, 256
R = 148 * 128 * 8
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
P = dict(Wo=rn(d, d) / math.sqrt(d), bo=0.1 * rn(d), g1=1 + 0.1 * rn(d), b1=0.1 * rn(d), c2=0.3 * rn(d), g2=1 + 0.1 * rn(d),
         b2=0.1 * rn(d), W1=rn(ffn, d) / math.sqrt(d), bf1=0.1 * rn(ffn), W2=rn(d, ffn) / math.sqrt(ffn), bf2=0.1 * rn(d),
         g3=1 + 0.1 * rn(d), b3=0.1 * rn(d), Win=rn(3 * d, d) / math.sqrt(d), bin=0.1 * rn(3 * d))
attn, x = rn(R, d), rn(R, d)
prep = ops.decoder_chain_prepare(P["Wo"], P["W1"], P["W2"], P["Win"])
xo = torch.empty_like(x)
L = 201
R = (R // L) * L
attn, x, xo = attn[:R], x[:R], xo[:R]
images = ops.qkv_images_buffer(R // L, L, 4, torch.device(dev), slot=5)
def fused():
    ops.decoder_chain_tc(attn, x, prep, P["bo"], P["g1"], P["b1"], P["c2"], P["g2"], P["b2"], P["bf1"], P["bf2"], P["g3"], P["b3"],
                         P["bin"], x_out=xo, qkv_images=images, L=L, mask_mode=0)
for _ in range(3): fused()
tl = torch.zeros((8, 3, 32), dtype=torch.int64, device=dev)
l = lib()
l.irs_decoder_chain_debug_timeline.argtypes = [ctypes.c_void_p]
l.irs_decoder_chain_debug_timeline.restype = None
l.irs_decoder_chain_debug_timeline(tl.data_ptr())
fused(); torch.cuda.synchronize()
l.irs_decoder_chain_debug_timeline(None)
t = tl.cpu()
t0 = int(t[t > 0].min())
names = {0: ["wait_a0", "a0_ready", "g1_issued", "a1_ready", "w_a2_0", "a2_0", "w_a2_1", "a2_1", "w_a2_2", "a2_2", "w_a2_3", "a2_3",
             "ffn_issued", "a3_ready", "g4ab_issued", "g4c_issued"],
         1: ["wait_d1", "d1_ready", "e1_done", "e2_0_start", "e2_0_done", "e2_1_start", "e2_1_done", "e2_2_start", "e2_2_done",
             "e2_3_start", "e2_3_done", "d3_ready", "e3_done", "d4ab_ready", "e4ab_done", "d4c_ready", "e4c_done"],
         2: ["loads_issued", "q_free", "a0_written"]}
for it in range(1, 5):
    print(f"--- tile iteration {it}")
    ev = []
    for role in range(3):
        for i, n in enumerate(names[role]):
            v = int(t[it, role, i])
            if v: ev.append((v - t0, "MMA" if role == 0 else ("EPI" if role == 1 else "LOAD"), n))
    ev.sort()
    base = ev[0][0]
    for v, r, n in ev:
        print(f"{v - base:8d}  {r:4s} {n}")
