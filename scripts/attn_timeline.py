import ctypes, sys, torch
sys.path.insert(0, ".")
from influentialrs_b200 import ops
from influentialrs_b200._lib import lib
dev = "cuda:0"
B, L, H, dh = 148 * 3, 201, 4, 32
d = H * dh
qkv = torch.randn((B, L, 3 * d), device=dev); ids = torch.randint(1, 1000, (B, L), device=dev); r_u = torch.randn(B, device=dev)
for _ in range(3): ops.pim_attention(qkv, ids, r_u, H, 0)
tl = torch.zeros((8, 4, 16), dtype=torch.int64, device=dev)
l = lib(); l.irs_pim_attn_debug_timeline.argtypes = [ctypes.c_void_p]; l.irs_pim_attn_debug_timeline.restype = None
l.irs_pim_attn_debug_timeline(tl.data_ptr())
ops.pim_attention(qkv, ids, r_u, H, 0); torch.cuda.synchronize()
l.irs_pim_attn_debug_timeline(None)
t = tl.cpu(); t0 = int(t[t > 0].min())
names = {0: ["wS", "S", "pass1", "P", "O", "epi"], 1: ["wS", "S", "pass1", "P", "O", "epi"],
         2: ["wP0", "P0", "PV0i", "kv", "QK0i", "wP1", "P1", "PV1i", "-", "QK1i"], 3: ["free", "issued"]}
ev = []
for it in range(2, 6):
    for role in range(4):
        for i, n in enumerate(names[role]):
            v = int(t[it, role, i])
            if v: ev.append((v - t0, ["SM0", "SM1", "MMA", "LOAD"][role], f"{n}[{it}]"))
ev.sort()
for v, r, n in ev: print(f"{v:8d}  {r:4s} {n}")
