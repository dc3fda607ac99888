"""Drop-in ``SampleNet`` (reference: model/uRS.py:18-71): the user-oriented next-item RS the evaluator
scores influence paths with.  Same constructor, module names, ``state_dict`` keys/shapes and method
signatures as the reference; the arithmetic runs in libirs_b200.so: embedding gather (K1), decoder layers
with the causal + key-padding mask built in-kernel (K3 + the fused decoder chain), cross-attention over the
all-zero memory folded to a constant.  ``forward`` materialises [B,L,N] logits only for API parity -- the
evaluator (evaluator.py) reads single rows through the fused catalog scorer instead."""
from __future__ import annotations

import math

import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .irn import PositionalEncoding, _decoder_stack


class SampleNet(nn.Module):
    """model/uRS.py:18-71."""

    def __init__(self, config):
        super().__init__()
        self.PAD_ID = 0
        self.item_embed_path = None
        self.n_item = config.n_item
        self.max_len = config.max_len
        self.n_layers = config.n_layers
        self.n_heads = config.n_heads
        self.embed_dim = config.emb_dim
        self.ffn_dim = config.ffn_dim
        self.dropout = config.dropout
        # same modules, same order as the reference => same RNG stream => same initial weights
        self.word_embedder = nn.Embedding(self.n_item + 1, self.embed_dim, padding_idx=self.PAD_ID)
        self.pos_embedder = PositionalEncoding(self.embed_dim, self.max_len)
        self.decoder = nn.TransformerDecoder(
            decoder_layer=nn.TransformerDecoderLayer(d_model=self.embed_dim, nhead=self.n_heads,
                                                     dim_feedforward=self.ffn_dim, dropout=self.dropout,
                                                     activation="relu"),
            num_layers=self.n_layers)
        self.project = nn.Linear(self.embed_dim, self.n_item)

    def embed(self, dec_inp_seq):
        """word_embedder(seq)*sqrt(d) + pe, dropout in training (model/uRS.py:55-56)."""
        L = dec_inp_seq.size(1)
        x = ops.embed_gather(dec_inp_seq, self.word_embedder.weight, self.pos_embedder.pe[0, :L], math.sqrt(self.embed_dim))
        return F.dropout(x, self.dropout, self.training)

    def decoding(self, dec_inp_seq, last_row=None):
        """Decoder output [B,L,d] under the causal + key-padding mask (model/uRS.py:52-64)."""
        dec_inp_seq = dec_inp_seq.contiguous()
        return _decoder_stack(self, self.embed(dec_inp_seq), dec_inp_seq, None, ops.MASK_CAUSAL_PAD, last_row)

    def invalidate_prepared(self):
        """Forget every cached weight image of this module (call after editing weights through ``.data``)."""
        self.__dict__.pop("_tc_cache", None)
        ops.invalidate_prepared()

    def forward(self, dec_inp_seq):
        """Logits [B,L,N] (model/uRS.py:66-69); API parity only, the evaluator never materialises them."""
        return self.project(self.decoding(dec_inp_seq))
