"""Import the UNMODIFIED reference (read-only, /root/reference) in the dev container.

TEST INFRASTRUCTURE.  Used only by oracle/make_golden.py and by bench.py's reference legs
(--impl reference: the reference's own get_seq_in_batch on the host cores; --impl torch_gpu: the same
code on cuda through stock torch).  /root/reference does not exist on the GPU box; there the staged,
unmodified copy oracle/_ref/ (oracle/stage_ref.py, git-ignored) is imported instead.  Nothing in the
-m gpu tests, smoke() or the product package imports this file.

Three non-invasive shims:
  D2  torch>=2.2 ReduceLROnPlateau has no ``verbose`` kwarg -> swallow it.
  D3  model/sas.py, model/caser.py import DatasetNN from utils; it lives in data_provider.
  D1  InfluentialNet.decoding passes pi_factor positionally into ``w_h``
      (model/influentialRS.py:183-184); IntendedNet reroutes a tensor ``w_h`` to ``pi_factor``,
      i.e. the keyword call the signature (:120-124) evidently intends.  No other line changes.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch

# /root/reference exists only in the dev container; oracle/_ref is the staged copy (oracle/stage_ref.py) that travels
# to the GPU box with the snapshot
REF_CANDIDATES = [os.environ.get("IRS_REF", ""), "/root/reference",
                  os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]


def find_reference():
    for p in REF_CANDIDATES:
        if p and os.path.isfile(os.path.join(p, "model", "influentialRS.py")):
            return p
    return None


def load_reference():
    """Returns a namespace with the reference classes, or None if the tree is absent."""
    ref = find_reference()
    if ref is None:
        return None
    sys.dont_write_bytecode = True          # the tree is read-only
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from torch.optim import lr_scheduler as _ls

    if not getattr(_ls.ReduceLROnPlateau, "_irs_shim", False):
        class _RLROP(_ls.ReduceLROnPlateau):                                    # D2
            _irs_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)
        _ls.ReduceLROnPlateau = _RLROP
    import utils as ref_utils
    import data_provider as ref_dp
    ref_utils.DatasetNN = ref_dp.DatasetNN                                       # D3
    from model.influentialRS import InfluentialNet, IRSNN
    from model.uRS import SampleNet
    from model.evaluator import Evaluator
    from model.sas import SAS
    from model.caser import Caser

    class IntendedNet(InfluentialNet):                                           # D1
        def _generate_square_subsequent_mask(self, size, w_h=0.05, w_obj=1, pi_factor=None):
            if torch.is_tensor(w_h):
                pi_factor, w_h = w_h, 0.05
            return super()._generate_square_subsequent_mask(size, w_h, w_obj, pi_factor)

    return SimpleNamespace(path=ref, InfluentialNet=InfluentialNet, IntendedNet=IntendedNet, IRSNN=IRSNN,
                           SampleNet=SampleNet, Evaluator=Evaluator, SAS=SAS, Caser=Caser,
                           utils=ref_utils, data_provider=ref_dp)


def irn_config(n_item, n_user, max_len, n_layers, n_heads, emb_dim, ffn_dim, u_emb_dim=10, dropout=0.0, lr1=1e-3):
    """The argparse Namespace fields the reference constructors read (model/influentialRS.py:36-47,90)."""
    return SimpleNamespace(n_item=n_item, n_user=n_user, max_len=max_len, n_layers=n_layers, n_heads=n_heads,
                           emb_dim=emb_dim, u_emb_dim=u_emb_dim, ffn_dim=ffn_dim, dropout=dropout, lr1=lr1)
