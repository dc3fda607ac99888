#!/usr/bin/env python
"""Secondary metrics of SURVEY.md section 8d on one B200 (synthetic data, reference default initialisers):
  cfg2  SASRec full-catalog next-item scoring, ml-1m shape, batch 1024            -> scored users/s
  cfg4  IRN training step (gather + PIM attention fwd/bwd + full softmax CE + scatter-add + Adam), 500k items,
        batch 4096, L=50, d=128                                                    -> train steps/s
  cfg5  Evaluator measurements of generated paths (SampleNet d=128, L=60, N=1M, B=8192, path length 21)
        + Caser full-catalog scoring [8192, 256] x [1M, 256]                       -> evaluated users/s, scored users/s
One JSON line per config on stdout.   python scripts/bench_secondary.py [cfg2 cfg4 cfg5] [--small]"""
import json, math, os, sys, time
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import influentialrs_b200 as pkg

dev = torch.device("cuda:0")
small = "--small" in sys.argv
which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg2", "cfg5", "cfg4"]


def timed(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cfg2():
    N, B, T = 3706, 1024, 20
    cfg = SimpleNamespace(n_user=6040, n_item=N, hidden_units=120, max_len=T, dropout_rate=0.2, num_blocks=4, num_heads=3)
    torch.manual_seed(1234)
    net = pkg.SAS(cfg, dev).to(dev).eval()
    g = torch.Generator().manual_seed(1234)
    seqs = torch.randint(1, N + 1, (B, T), generator=g)
    for b in range(B):
        seqs[b, : int(torch.randint(0, 6, (1,), generator=g))] = 0
    rats = (torch.rand((B, T), generator=g) > 0.5).long() * (seqs > 0)
    sd, rd = seqs.to(dev), rats.to(dev)
    ms = timed(lambda: net.predict_topk(sd, rd, top_k=50, hist=sd), warm=3, reps=20)
    return {"config": "cfg2: SASRec next-item scoring, ml-1m shape (N=3706, T=20, C=120, H=3, 4 blocks), batch 1024, top-50 with history filter",
            "metric": "scored users/s", "value": B / (ms / 1e3), "ms_per_batch": ms}


def cfg4():
    N, B, L, d = (20000, 256, 50, 128) if small else (500_000, 4096, 50, 128)
    c = SimpleNamespace(n_item=N, n_user=100_000, max_len=L, n_layers=6, n_heads=4, emb_dim=d, u_emb_dim=10, ffn_dim=256,
                        dropout=0.0, lr1=1e-3)
    torch.manual_seed(1234)
    net = pkg.InfluentialNet(c).to(dev)
    irn = pkg.IRSNN(c, net, dev)
    g = torch.Generator().manual_seed(1234)
    out = {}
    for pad in (0.0, 0.44):
        seqs = torch.randint(1, N + 1, (B, L), generator=g)
        if pad > 0:
            for b in range(B):
                seqs[b, : int(L * pad * 2 * float(torch.rand(1, generator=g)))] = 0
        users = torch.randint(0, c.n_user, (B,), generator=g)
        sd, ud = seqs.to(dev), users.to(dev)
        losses = []
        ms = timed(lambda: losses.append(irn.train_batch(sd, ud)), warm=1, reps=2)
        out[f"pad_{pad}"] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "rows": int((seqs[:, 1:] > 0).sum()), "loss_first_last": [losses[0], losses[-1]]}
    return {"config": f"cfg4: IRN train_batch incl. Adam, N={N}, batch {B}, L={L}, d={d}, 6 layers/4 heads (pad fraction 0 and ~0.44)",
            "metric": "train steps/s", "value": out["pad_0.0"]["steps_per_s"], "detail": out}


def cfg5():
    N, B, L, d, P = (20000, 256, 60, 128, 5) if small else (1_000_000, 8192, 60, 128, 21)
    c = SimpleNamespace(n_item=N, max_len=L, n_layers=6, n_heads=4, emb_dim=d, ffn_dim=256, dropout=0.0, lr1=1e-3)
    torch.manual_seed(1234)
    net = pkg.SampleNet(c).to(dev).eval()
    ev = pkg.Evaluator(c, net, dev)
    g = torch.Generator().manual_seed(1234)
    nh = 30
    hist = torch.zeros((B, L), dtype=torch.long)
    new = torch.zeros((B, L), dtype=torch.long)
    ids = torch.randint(1, N + 1, (B, nh + P + 1), generator=g)
    hist[:, :nh] = ids[:, :nh]
    new[:, :nh + P] = ids[:, :nh + P]
    targets = ids[:, nh + P]
    start = torch.full((B,), nh, dtype=torch.long)
    lp = torch.full((B,), P, dtype=torch.long)
    hd, nd, td, sd, ld = hist.to(dev), new.to(dev), targets.to(dev), start.to(dev), lp.to(dev)
    t_pp = timed(lambda: ev.get_pp_in_batch(nd, sd, ld), warm=1, reps=1)
    t_rr = timed(lambda: ev.get_rr_increase_in_batch(hd, nd, td), warm=0, reps=1)
    t_gr = timed(lambda: ev.get_grad_in_batch(hd.clone(), nd, td, sd, ld), warm=0, reps=1)
    # Caser scoring over the catalog
    x = torch.randn((B, 2 * d), device=dev)
    W2 = torch.randn((N + 1, 2 * d), device=dev) / (2 * d)
    b2 = torch.zeros((N + 1,), device=dev)
    t_ca = timed(lambda: pkg.ops.score_topk(x, W2[1:], b2[1:], 50, None, 1), warm=1, reps=2)
    tot = t_pp + t_rr + t_gr
    return {"config": f"cfg5: Evaluator get_pp + get_rr_increase + get_grad (SampleNet d={d}, L={L}, N={N}, B={B}, path length {P}) + Caser "
                      f"catalog scoring [B,{2*d}]x[N,{2*d}] top-50",
            "metric": "evaluated users/s", "value": B / (tot / 1e3),
            "detail": {"get_pp_ms": t_pp, "get_rr_increase_ms": t_rr, "get_grad_ms": t_gr, "caser_top50_ms": t_ca,
                       "caser_scored_users_per_s": B / (t_ca / 1e3)}}


for name in which:
    t0 = time.time()
    r = {"cfg2": cfg2, "cfg4": cfg4, "cfg5": cfg5}[name]()
    r["wall_s"] = time.time() - t0
    r["data"] = "synthetic"
    r["n_gpus"] = 1
    print(json.dumps(r), flush=True)
    torch.cuda.empty_cache()
