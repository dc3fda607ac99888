"""CPU: the drop-in classes keep the reference's constructor signatures and state_dict keys/shapes
(checked against state dicts saved FROM the reference, tests/golden)."""
from types import SimpleNamespace

import torch

from tests.helpers import load_golden


def test_samplenet_state_dict_is_the_references():
    from influentialrs_b200.urs import SampleNet
    from influentialrs_b200.evaluator import Evaluator
    sd, g = load_golden("evaluator_small")
    n_item, max_len, n_layers, n_heads, emb, ffn = (int(v) for v in g["cfg"])
    cfg = SimpleNamespace(n_item=n_item, max_len=max_len, n_layers=n_layers, n_heads=n_heads, emb_dim=emb, ffn_dim=ffn,
                          dropout=0.1, lr1=1e-3)
    net = SampleNet(cfg)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    ev = Evaluator(cfg, net, torch.device("cpu"))
    for attr in ("net", "optimizer", "pla_lr_scheduler", "loss_function", "softmax", "vocab_size", "PAD_ID"):
        assert hasattr(ev, attr)


def test_sas_state_dict_is_the_references():
    from influentialrs_b200.baselines import SAS
    sd, g = load_golden("sas_small")
    n_item, hidden, max_len, blocks, heads = (int(v) for v in g["cfg"])
    cfg = SimpleNamespace(n_user=9, n_item=n_item, hidden_units=hidden, max_len=max_len, dropout_rate=0.2, num_blocks=blocks,
                          num_heads=heads)
    net = SAS(cfg, torch.device("cpu"))
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)


def test_caser_module_shapes():
    from influentialrs_b200.baselines import Caser
    _, g = load_golden("caser_small")
    dims = g["W2"].shape[1] // 2
    args = SimpleNamespace(max_len=g["seqs"].shape[1], d=dims, nh=4, nv=2, drop=0.5, ac_conv="relu", ac_fc="relu")
    net = Caser(8, g["W2"].shape[0] - 1, args)
    assert tuple(net.W2.weight.shape) == g["W2"].shape and tuple(net.b2.weight.shape) == g["b2"].shape
    # the torch part of the module (feature extractor + per-item scoring) runs on CPU like the reference's
    seq = torch.randint(1, 10, (1, args.max_len))
    out = net(seq, torch.zeros_like(seq), torch.tensor([[0]]), torch.arange(1, 11), for_pred=True)
    assert out.shape == (10,)
