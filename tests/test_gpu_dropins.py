"""GPU parity of the drop-in SampleNet / Evaluator / SAS / Caser classes
against vectors produced by the reference itself (tests/golden, oracle/make_golden.py) and the oracle."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import irn_oracle as O
from tests.helpers import load_golden, assert_close_rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import influentialrs_b200 as p
    return p


def _evaluator(pkg):
    sd, g = load_golden("evaluator_small")
    n_item, max_len, n_layers, n_heads, emb, ffn = (int(v) for v in g["cfg"])
    cfg = SimpleNamespace(n_item=n_item, max_len=max_len, n_layers=n_layers, n_heads=n_heads, emb_dim=emb, ffn_dim=ffn,
                          dropout=0.0, lr1=1e-3)
    net = pkg.SampleNet(cfg)
    net.load_state_dict(sd)                       # strict: the reference's keys and shapes
    net.to(DEV).eval()
    return pkg.Evaluator(cfg, net, torch.device(DEV)), sd, g


def test_samplenet_forward_matches_reference(pkg):
    ev, sd, g = _evaluator(pkg)
    new = torch.from_numpy(g["new_seqs"]).to(DEV)
    with torch.no_grad():
        logits = ev.net(new[:, :-1]).cpu()
    ref = torch.from_numpy(g["logits_new"])
    ok = ~torch.isnan(ref)
    assert torch.equal(torch.isnan(logits), torch.isnan(ref))
    assert_close_rel(logits[ok], ref[ok], 1e-3, "SampleNet.forward")


def test_evaluator_measurements_match_reference(pkg):
    ev, sd, g = _evaluator(pkg)
    D = lambda k: torch.from_numpy(g[k]).to(DEV)
    hist, new, tg, sp, lp = D("histories"), D("new_seqs"), D("targets"), D("start_pos"), D("l_path")
    np.testing.assert_allclose(ev.get_pp_in_batch(new, sp, lp), g["pp"], atol=1e-4)
    irr, ir = ev.get_rr_increase_in_batch(hist, new, tg)
    np.testing.assert_array_equal(ir, g["ir"])
    np.testing.assert_allclose(irr, g["irr"], atol=1e-12)
    h2 = hist.clone()
    tp, pp, avg, ioi = ev.get_grad_in_batch(h2, new, tg, sp, lp)
    np.testing.assert_allclose(tp, g["t_probs"], atol=1e-4)
    np.testing.assert_allclose(pp, g["p_probs"], atol=1e-4)
    np.testing.assert_allclose(avg, g["avg_ps"], atol=1e-4)
    np.testing.assert_allclose(ioi, g["iois"], atol=1e-4)
    assert not torch.equal(h2, hist)              # like the reference, the history window is advanced in place
    assert abs(ev.get_loss_on_eval_data(new) - float(g["eval_loss"])) < 1e-4
    hit, rr = ev.get_accuracy_metrics_in_batch(new, top_k=5, use_h=True)
    assert hit == int(g["acc_hit"])
    np.testing.assert_allclose(rr, g["acc_rr"], rtol=0, atol=0)


def test_evaluator_train_batch_decreases_loss(pkg):
    ev, sd, g = _evaluator(pkg)
    new = torch.from_numpy(g["new_seqs"]).to(DEV)
    l0 = ev.train_batch(new)
    assert abs(l0 - float(g["eval_loss"])) < 1e-4   # dropout = 0: the training loss of step 0 is the eval loss
    for _ in range(5):
        l1 = ev.train_batch(new)
    assert l1 < l0


def _sas(pkg):
    sd, g = load_golden("sas_small")
    n_item, hidden, max_len, blocks, heads = (int(v) for v in g["cfg"])
    cfg = SimpleNamespace(n_user=9, n_item=n_item, hidden_units=hidden, max_len=max_len, dropout_rate=0.2, num_blocks=blocks,
                          num_heads=heads)
    net = pkg.SAS(cfg, torch.device(DEV))
    net.load_state_dict(sd)
    net.to(DEV).eval()
    return net, sd, g


def test_sas_predict_matches_reference(pkg):
    net, sd, g = _sas(pkg)
    with torch.no_grad():
        feats = net.log2feats(g["seqs"], g["rats"]).cpu()
        logits = net.predict(np.arange(len(g["seqs"])), g["seqs"], g["rats"]).cpu()
    assert_close_rel(feats, g["feats"], 1e-3, "SAS.log2feats")
    assert_close_rel(logits, g["logits"], 1e-3, "SAS.predict")


def test_sas_predict_topk_is_sort_filter_slice(pkg):
    net, sd, g = _sas(pkg)
    seqs = torch.from_numpy(g["seqs"])
    k = 10
    vals, items = net.predict_topk(g["seqs"], g["rats"], top_k=k, hist=seqs.to(DEV))
    ref_scores = torch.from_numpy(g["logits"]).double()
    want = O.predict_next_tail(ref_scores, seqs, k)            # sort -> +1 -> delete_item_in_history -> [:k]
    # decisions the fp32 kernels can legitimately flip: neighbours closer than 1e-5 in the reference scores
    got = items.cpu()
    for b in range(seqs.shape[0]):
        if torch.equal(got[b], want[b]):
            continue
        s = ref_scores[b][want[b] - 1]
        assert float((s[:-1] - s[1:]).min()) < 1e-5, f"row {b}: order differs without a near-tie"


def test_caser_scoring_matches_reference(pkg):
    _, g = load_golden("caser_small")
    W2, b2, x = torch.from_numpy(g["W2"]), torch.from_numpy(g["b2"]), torch.from_numpy(g["x"])
    n_items, dims = W2.shape[0] - 1, W2.shape[1] // 2
    args = SimpleNamespace(max_len=g["seqs"].shape[1], d=dims, nh=4, nv=2, drop=0.5, ac_conv="relu", ac_fc="relu")
    net = pkg.Caser(8, n_items, args)
    with torch.no_grad():
        net.W2.weight.copy_(W2)
        net.b2.weight.copy_(b2)
    net.to(DEV).eval()
    net.features = lambda *a: x.to(DEV)                      # the golden file stores x (conv features + user embedding)
    B = x.shape[0]
    seqs = torch.from_numpy(g["seqs"]).to(DEV)
    vals, items = net.predict_topk(seqs, torch.zeros_like(seqs), torch.zeros((B, 1), dtype=torch.long, device=DEV), top_k=n_items,
                                    hist=seqs)
    ref = torch.from_numpy(g["scores"])
    # all items in reference order, minus the history, scores equal
    for b in range(B):
        it = items[b].cpu()
        n_valid = n_items - seqs.shape[1]
        sc = ref[b][it[:n_valid] - 1]
        assert_close_rel(vals[b, :n_valid].cpu(), sc, 1e-4, "caser scores")
        assert not set(it[:n_valid].tolist()) & set(seqs[b].cpu().tolist())
    # the module's own forward(for_pred=True) agrees with the fused scorer (random weights, full module)
    del net.features
    net2 = pkg.Caser(8, n_items, args).to(DEV).eval()
    with torch.no_grad():
        net2.b2.weight.normal_(0, 0.1)
        seq = torch.randint(1, n_items + 1, (3, args.max_len), device=DEV)
        rat = torch.randint(0, 2, (3, args.max_len), device=DEV)
        usr = torch.arange(3, device=DEV).view(3, 1)
        items_all = torch.arange(1, n_items + 1, device=DEV)
        v, it = net2.predict_topk(seq, rat, usr, top_k=5)
        for b in range(3):
            s = net2(seq[b:b + 1], rat[b:b + 1], usr[b:b + 1], items_all, for_pred=True)
            top = torch.topk(s, 5)
            assert_close_rel(v[b].cpu(), top.values.cpu(), 1e-4, "caser fused vs module")
            assert torch.equal(it[b].cpu(), (top.indices + 1).cpu())


def test_predict_next_tail_both_forms(pkg):
    """Baselines' predict_next tail (sort -> +1 -> delete_item_in_history(last 50) -> [:k]) in its factorised and its
    explicit-score form against the oracle restatement."""
    from influentialrs_b200.baselines import predict_next_tail
    g = torch.Generator().manual_seed(21)
    B, N, d, k = 9, 400, 16, 12
    feats, items = torch.randn((B, d), generator=g), torch.randn((N, d), generator=g)
    hist = torch.zeros((B, 70), dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(1, 71, (1,), generator=g))
        hist[b, 70 - n:] = torch.randperm(N, generator=g)[:n] + 1
    scores = feats.double() @ items.double().t()
    want = O.predict_next_tail(scores, hist, k)                     # only the last 50 history entries count
    got = predict_next_tail(k, hist.to(DEV), features=feats.to(DEV), item_matrix=items.to(DEV)).cpu().long()
    assert torch.equal(got, want)
    # explicit scores with heavy ties (popularity counts): stable order = lower id first
    pop = torch.randint(0, 5, (N,), generator=g).float()
    want2 = O.predict_next_tail(pop.double().unsqueeze(0).expand(B, -1), hist, k)
    got2 = predict_next_tail(k, hist.to(DEV), scores=pop.to(DEV)).cpu().long()
    assert torch.equal(got2, want2)


@pytest.mark.parametrize("name", ["pop", "mc", "fpmc", "tr", "bpr"])
def test_predict_next_tail_matches_reference_baselines(pkg, name):
    """predict_next_tail (both input forms) against predict_next of the reference's own POP / MC / FPMC / TransRec / BPR
    (tests/golden/baseline_tails.npz, produced by model/baselines.py through oracle/make_golden.py)."""
    from influentialrs_b200.baselines import predict_next_tail
    from tests.helpers import baseline_tail_scores, assert_same_topk_up_to_ties
    _, g = load_golden("baseline_tails")
    hist = torch.from_numpy(g["hist"])
    users = torch.from_numpy(g["users"])
    k = int(g["top_k"])
    want = torch.from_numpy(g[name + "_preds"]).long()
    scores = baseline_tail_scores(g, name)
    got = predict_next_tail(k, hist.to(DEV), scores=scores.float().to(DEV)).cpu().long()
    if name == "pop":
        assert_same_topk_up_to_ties(got, want, scores, hist)     # tie-heavy counts; the reference's sort is unstable
    else:
        assert torch.equal(got, want)
    T = lambda key: torch.from_numpy(g[key]).float().to(DEV)
    if name == "bpr":                                             # factorised forms through the fused scorer
        got2 = predict_next_tail(k, hist.to(DEV), features=T("bpr_W")[users.to(DEV)], item_matrix=T("bpr_H")).cpu().long()
        assert torch.equal(got2, want)
    if name == "fpmc":
        last = (hist[:, -1] - 1).to(DEV)
        feats = torch.cat([T("fpmc_gamU")[users.to(DEV)], T("fpmc_kap")[last]], 1)
        items = torch.cat([T("fpmc_gamI").t(), T("fpmc_eta").t()], 1).contiguous()
        got2 = predict_next_tail(k, hist.to(DEV), features=feats, item_matrix=items).cpu().long()
        assert torch.equal(got2, want)
