"""Shared test helpers: golden loading and tolerances (tests only)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: logits / attention outputs / gradients within 1e-3 relative error.  Element-wise
# relative error is meaningless next to zero (SURVEY.md section 7 "hard parts"), so the bound used
# everywhere is  |a-b| <= RTOL * max|ref|  over the compared tensor (row-max for logits).
RTOL = 1e-3


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    rest = {k: z[k] for k in z.files if not k.startswith("sd.")}
    return sd, rest


def max_rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def assert_close_rel(a, b, rtol=RTOL, what=""):
    r = max_rel(a, b)
    assert r <= rtol, f"{what}: max|a-b|/max|ref| = {r:.3e} > {rtol:g}"


def cfg3_shape_state(g, module_cls, stored=None):
    """Rebuild the weights of the ``irn_cfg3_shape`` fixture: the reference's default initialisers under the stored seed,
    through ``module_cls`` (the drop-in InfluentialNet creates the same torch.nn modules in the same order), with the
    normal_-initialised embedding tables taken from the fixture (``stored``: torch's CPU normal_ stream depends on the
    host's vector ISA), and check the per-tensor fingerprint the reference-side generator stored -- bit for bit."""
    from oracle.ref_shim import irn_config
    cfg = irn_config(*[int(x) for x in g["cfg"][:7]], u_emb_dim=int(g["cfg"][7]))
    torch.manual_seed(int(g["seed"]))
    net = module_cls(cfg)
    if stored:
        net.load_state_dict(stored, strict=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    keys = [str(k) for k in g["sd_keys"]]
    assert sorted(sd.keys()) == keys
    for i, k in enumerate(keys):
        bits = sd[k].contiguous().reshape(-1).view(torch.int32).to(torch.int64)
        w = torch.arange(bits.numel(), dtype=torch.int64) % 8191 + 1
        got = (int(bits.sum()), int((bits * w).sum()), bits.numel())
        assert got == tuple(int(x) for x in g["sd_fingerprint"][i]), f"weights of {k} differ from the reference's"
    return cfg, net, sd


def baseline_tail_scores(g, name):
    """Score matrix [B,N] (fp64) each classical baseline's predict_next sorts (model/baselines.py), from the stored factors."""
    T = lambda k: torch.from_numpy(g[k]).double()
    hist, users = torch.from_numpy(g["hist"]), torch.from_numpy(g["users"])
    last = hist[:, -1] - 1
    if name == "pop":
        return T("pop_counts").unsqueeze(0).expand(hist.shape[0], -1)
    if name == "mc":
        return (T("mc_gam").float() @ T("mc_eta").float()).double()[last]
    if name == "fpmc":
        a = (T("fpmc_gamU").float() @ T("fpmc_gamI").float())
        b = (T("fpmc_kap").float() @ T("fpmc_eta").float())
        return (a[users] + b[last]).double()
    if name == "tr":
        H, R, r, beta = T("tr_H").float(), T("tr_R").float(), T("tr_r").float(), T("tr_beta").float()
        l = torch.square(H[last].unsqueeze(1) + r.view(1, 1, -1) + R[users].unsqueeze(1) - H.unsqueeze(0))
        return (-l.sum(2) - beta.unsqueeze(0)).double()
    if name == "bpr":
        return (T("bpr_W").float()[users] @ T("bpr_H").float().t()).double()
    raise KeyError(name)


def assert_same_topk_up_to_ties(got, want, scores, hist=None, h=50):
    """Two top-k id lists [B,k] over the same score matrix [B,N] are the same answer up to the order INSIDE groups of
    equal scores: the reference sorts with torch's default (unstable) sort, so the order of tied items -- and which of
    them make a cut that falls inside a tie group -- is implementation-defined there (SURVEY D7 pins 'lower id first'
    here).  Checks: same score sequence position by position, no excluded item, no duplicates, and every item scoring
    strictly above the k-th score present in both."""
    got, want = torch.as_tensor(got).long(), torch.as_tensor(want).long()
    s = torch.as_tensor(scores)
    sg, sw = s.gather(1, got - 1), s.gather(1, want - 1)
    assert torch.equal(sg, sw), "score sequences differ"
    for b in range(got.shape[0]):
        gb, wb = got[b].tolist(), want[b].tolist()
        assert len(set(gb)) == len(gb)
        if hist is not None:
            ex = set(hist[b, -h:].tolist())
            assert not (set(gb) & ex)
        kth = float(sg[b, -1])
        assert {i for i, v in zip(gb, sg[b].tolist()) if v > kth} == {i for i, v in zip(wb, sw[b].tolist()) if v > kth}
