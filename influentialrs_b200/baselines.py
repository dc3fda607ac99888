"""Drop-in scoring paths of the SASRec and Caser baselines (reference: model/sas.py:100-228,355-386 and
model/caser.py:62-183,265-299).

Both baselines end in the same operation as IRN: score every catalog item against one feature vector per
user, drop what the user has already seen, keep the best k.  The reference materialises the [B,N] score
matrix, sorts it fully per user on the CPU and filters with an O(N x h) boolean compare
(utils.delete_item_in_history); here the tail is the fused catalog scorer (``ops.score_topk``: scores never
reach HBM, exclusions and top-k in the epilogue).  Module names / ``state_dict`` keys are the reference's,
so its checkpoints load."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

activation_getter = {"iden": lambda x: x, "relu": F.relu, "tanh": torch.tanh, "sigm": torch.sigmoid}


def _history_exclusions(hist, n_item, h=50):
    """utils.delete_item_in_history keeps only the last ``h`` history items (utils.py:8-12)."""
    if hist is None:
        return None
    return ops.sort_exclusions(hist[:, -h:].contiguous(), n_item, 1)


# ----------------------------------------------------------------------------------------------- SASRec
class PointWiseFeedForward(nn.Module):
    """model/sas.py:77-98 (1x1 convolutions == per-position linears)."""

    def __init__(self, hidden_units, dropout_rate):
        super().__init__()
        self.conv1 = nn.Conv1d(hidden_units, hidden_units, kernel_size=1)
        self.dropout1 = nn.Dropout(p=dropout_rate)
        self.relu = nn.ReLU()
        self.conv2 = nn.Conv1d(hidden_units, hidden_units, kernel_size=1)
        self.dropout2 = nn.Dropout(p=dropout_rate)

    def forward(self, inputs):
        y = F.linear(self.dropout1(F.relu(F.linear(inputs, self.conv1.weight[:, :, 0], self.conv1.bias))),
                     self.conv2.weight[:, :, 0], self.conv2.bias)
        return self.dropout2(y) + inputs


class SAS(nn.Module):
    """SASRec network (model/sas.py:106-228): ``log2feats`` / ``predict`` with the reference's signatures
    plus ``predict_topk``, the fused replacement of predict -> sort -> delete_item_in_history -> [:k]
    (model/sas.py:355-386)."""

    def __init__(self, config=None, device=None):
        super().__init__()
        self.user_num = config.n_user
        self.item_num = config.n_item
        self.dev = device
        self.args = config
        self.item_emb = nn.Embedding(self.item_num + 1, config.hidden_units, padding_idx=0)
        self.rat_emb = nn.Embedding(2, config.hidden_units)
        self.pos_emb = nn.Embedding(config.max_len, config.hidden_units)
        self.emb_dropout = nn.Dropout(p=config.dropout_rate)
        self.attention_layernorms = nn.ModuleList()
        self.attention_layers = nn.ModuleList()
        self.forward_layernorms = nn.ModuleList()
        self.forward_layers = nn.ModuleList()
        self.last_layernorm = nn.LayerNorm(config.hidden_units, eps=1e-8)
        for _ in range(config.num_blocks):
            self.attention_layernorms.append(nn.LayerNorm(config.hidden_units, eps=1e-8))
            self.attention_layers.append(nn.MultiheadAttention(config.hidden_units, config.num_heads, config.dropout_rate))
            self.forward_layernorms.append(nn.LayerNorm(config.hidden_units, eps=1e-8))
            self.forward_layers.append(PointWiseFeedForward(config.hidden_units, config.dropout_rate))

    def _ids(self, a):
        return torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).long().to(self.item_emb.weight.device)

    def log2feats(self, log_seqs, rat_seqs):
        """model/sas.py:154-190, inference arithmetic (dropout follows self.training).  Keeps the reference's
        quirk that rating ids are looked up in the ITEM table (:162); the causal attention takes
        queries from LN(x) and keys/values from x and has no key-padding mask."""
        log_seqs, rat_seqs = self._ids(log_seqs), self._ids(rat_seqs)
        E = self.item_emb.weight
        C = E.shape[1]
        T = log_seqs.shape[1]
        H = self.attention_layers[0].num_heads if len(self.attention_layers) else 1
        x = ops.embed_gather(log_seqs, E, self.pos_emb.weight[:T].contiguous(), C ** 0.5)     # E[seq]*sqrt(C) + pos
        x = x + F.embedding(rat_seqs, E)
        x = self.emb_dropout(x)
        keep = log_seqs.ne(0).unsqueeze(-1)
        x = x * keep
        for i in range(len(self.attention_layers)):
            ln, mha = self.attention_layernorms[i], self.attention_layers[i]
            q_in = F.layer_norm(x, (C,), ln.weight, ln.bias, ln.eps)
            in_w, in_b = mha.in_proj_weight, mha.in_proj_bias
            q = F.linear(q_in, in_w[:C], in_b[:C])
            kv = F.linear(x, in_w[C:], in_b[C:])
            if self.training and mha.dropout > 0:
                raise NotImplementedError("attention-probability dropout (training) is outside the built scope")
            o = ops.attention_qkv(q, kv[..., :C].contiguous(), kv[..., C:].contiguous(), H, ops.MASK_CAUSAL)
            fl = self.forward_layernorms[i]
            x = ops.residual_layernorm(q_in.contiguous(), F.linear(o, mha.out_proj.weight), mha.out_proj.bias, fl.weight, fl.bias,
                                       eps=fl.eps) if not torch.is_grad_enabled() else \
                F.layer_norm(q_in + F.linear(o, mha.out_proj.weight, mha.out_proj.bias), (C,), fl.weight, fl.bias, fl.eps)
            x = self.forward_layers[i](x)
            x = x * keep
        return F.layer_norm(x, (C,), self.last_layernorm.weight, self.last_layernorm.bias, self.last_layernorm.eps)

    def forward(self, user_ids, log_seqs, rat_seqs, pos_seqs, neg_seqs):
        """model/sas.py:192-206 (training logits of sampled positives / negatives)."""
        f = self.log2feats(log_seqs, rat_seqs)
        return (f * self.item_emb(self._ids(pos_seqs))).sum(-1), (f * self.item_emb(self._ids(neg_seqs))).sum(-1)

    def predict(self, user_ids, log_seqs, rat_seqs, item_indices=[]):
        """Logits [B,I] = E[items] . f_last (model/sas.py:208-228); materialised, API parity."""
        final = self.log2feats(log_seqs, rat_seqs)[:, -1, :]
        if len(item_indices) == 0:
            return final @ self.item_emb.weight[1:].t()
        return final @ self.item_emb(self._ids(item_indices)).t()

    def predict_topk(self, log_seqs, rat_seqs, top_k=50, hist=None, h=50, scorer=None):
        """Best ``top_k`` items per user (score desc, id asc) among items not in the last ``h`` entries of
        ``hist`` [B,Lh] (0 = pad): SASNN.predict_next without the [B,N] matrix, sort and filter
        (model/sas.py:355-386, utils.py:8-12).  Returns (scores [B,k], items [B,k]).  ``scorer``: a dist.ShardedScorer
        built over ``item_emb.weight[1:]`` scores a catalog sharded across GPUs."""
        with torch.no_grad():
            final = self.log2feats(log_seqs, rat_seqs)[:, -1, :].contiguous()
            if scorer is not None:
                return scorer.topk(final, top_k, None if hist is None else hist[:, -h:].contiguous())
            W = self.item_emb.weight[1:]
            return ops.score_topk_any(final, W, None, top_k, _history_exclusions(hist, self.item_num, h), 1)


# ------------------------------------------------------------------------------------------------ Caser
class Caser(nn.Module):
    """Caser network (model/caser.py:62-183) with the reference's constructor and ``forward``; plus
    ``features`` / ``predict_topk`` for the batched, fused version of the per-user prediction loop
    (model/caser.py:265-299)."""

    def __init__(self, num_users, num_items, model_args):
        super().__init__()
        self.args = model_args
        L = self.args.max_len
        dims = self.args.d
        self.n_h = self.args.nh
        self.n_v = self.args.nv
        self.drop_ratio = self.args.drop
        self.ac_conv = activation_getter[self.args.ac_conv]
        self.ac_fc = activation_getter[self.args.ac_fc]
        self.num_items = num_items
        self.user_embeddings = nn.Embedding(num_users, dims)
        self.item_embeddings = nn.Embedding(num_items + 1, dims, padding_idx=0)
        self.rating_embeddings = nn.Embedding(2, dims)
        self.conv_v = nn.Conv2d(1, self.n_v, (L, 1))
        lengths = [i + 1 for i in range(L)]
        self.conv_h = nn.ModuleList([nn.Conv2d(1, self.n_h, (i, dims)) for i in lengths])
        self.fc1_dim_v = self.n_v * dims
        self.fc1_dim_h = self.n_h * len(lengths)
        self.fc1 = nn.Linear(self.fc1_dim_v + self.fc1_dim_h, dims)
        self.W2 = nn.Embedding(num_items + 1, dims + dims, padding_idx=0)
        self.b2 = nn.Embedding(num_items + 1, 1, padding_idx=0)
        self.dropout = nn.Dropout(self.drop_ratio)
        self.user_embeddings.weight.data.normal_(0, 1.0 / self.user_embeddings.embedding_dim)
        self.item_embeddings.weight.data.normal_(0, 1.0 / self.item_embeddings.embedding_dim)
        self.W2.weight.data.normal_(0, 1.0 / self.W2.embedding_dim)
        self.b2.weight.data.zero_()
        self.cache_x = None

    def features(self, seq_var, rat_var, user_var):
        """x = [relu(fc1(conv features)), user embedding]  [B, 2*dims]  (model/caser.py:146-171)."""
        item_embs = (self.item_embeddings(seq_var) + self.rating_embeddings(rat_var)).unsqueeze(1)
        user_emb = self.user_embeddings(user_var).reshape(seq_var.shape[0], -1)
        outs = []
        if self.n_v:
            outs.append(self.conv_v(item_embs).view(-1, self.fc1_dim_v))
        if self.n_h:
            out_hs = []
            for conv in self.conv_h:
                conv_out = self.ac_conv(conv(item_embs).squeeze(3))
                out_hs.append(F.max_pool1d(conv_out, conv_out.size(2)).squeeze(2))
            outs.append(torch.cat(out_hs, 1))
        z = self.ac_fc(self.fc1(self.dropout(torch.cat(outs, 1))))
        return torch.cat([z, user_emb], 1)

    def forward(self, seq_var, rat_var, user_var, item_var, for_pred=False):
        """model/caser.py:125-183."""
        x = self.features(seq_var, rat_var, user_var)
        w2, b2 = self.W2(item_var), self.b2(item_var)
        if for_pred:
            return (x * w2.squeeze()).sum(1) + b2.squeeze()
        return torch.baddbmm(b2, w2, x.unsqueeze(2)).squeeze()

    def predict_topk(self, seq_var, rat_var, user_var, top_k=50, hist=None, h=50, scorer=None):
        """Best ``top_k`` items per user of score[j] = x . W2[j] + b2[j] over the whole catalog, items in the
        last ``h`` history entries removed: the batched, fused form of Recommender.predict_next
        (model/caser.py:265-299).  Returns (scores [B,k], items [B,k]).  ``scorer``: a dist.ShardedScorer built over
        (W2.weight[1:], b2.weight[1:,0]) scores a catalog sharded across GPUs."""
        with torch.no_grad():
            x = self.features(seq_var, rat_var, user_var).contiguous()
            if scorer is not None:
                return scorer.topk(x, top_k, None if hist is None else hist[:, -h:].contiguous())
            return ops.score_topk_any(x, self.W2.weight[1:], self.b2.weight[1:, 0].contiguous(), top_k,
                                      _history_exclusions(hist, self.num_items, h), 1)


# ------------------------------------------------------------------- classical baselines' predict_next tails
def predict_next_tail(top_k, hist=None, h=50, *, features=None, item_matrix=None, item_bias=None, scores=None):
    """The tail every baseline's ``predict_next`` ends in (model/baselines.py:57-71 POP, :165-179 MC, :280-298 FPMC,
    :411-428 BPR, :527-550; model/sas.py:380-386; model/caser.py:291-298): rank all items by score, add 1, remove the
    items of the user's last ``h`` history entries (utils.delete_item_in_history, utils.py:8-12), keep ``top_k``.

    Two input forms, for the whole batch at once (``hist`` [B,Lh] int64, 0 = pad, or None):
      * factorised scores ``features`` [B,d] x ``item_matrix`` [N,d] (+ ``item_bias`` [N]) -- FPMC / BPR / SASRec / Caser:
        the fused catalog scorer (scores never materialised, exclusions and top-k in the epilogue);
      * an explicit score matrix ``scores`` [B,N] (or [N], shared by every user) -- POP counts, MC transition rows:
        excluded items are masked and a stable descending sort keeps the reference's tie order (lower id first).
    Returns item ids [B, top_k] as a float tensor, like the reference's ``preds``."""
    if features is not None:
        n_item = item_matrix.shape[0]
        _, items = ops.score_topk_any(features.contiguous(), item_matrix, item_bias, top_k,
                                      _history_exclusions(hist, n_item, h), 1)
        return items.float()
    if scores.dim() == 1:
        B = 1 if hist is None else hist.shape[0]
        scores = scores.unsqueeze(0).expand(B, -1)
    scores = scores.float().clone()
    if hist is not None:
        last = hist[:, -h:]
        cols = (last - 1).clamp(min=0)
        scores.scatter_(1, cols, torch.where(last > 0, torch.full_like(scores[:, :1], float("-inf")).expand_as(cols),
                                             scores.gather(1, cols)))
    order = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :top_k]
    return (order + 1).float()


def get_path(hist, users, targets, predict_next, k_c=5, max_path_len=20, fv=None, binary=True):
    """Persuasion-path generation of the classical baselines (main_baselines.get_path, main_baselines.py:93-121), for the
    whole batch at once and without the per-sample Python loop: at every step ask ``predict_next`` for each user's ``k_c``
    best unseen items, take the candidate whose feature vector is closest to the target's (utils.cal_fv_dist, utils.py:202-206:
    Hamming distance for ``binary`` vectors, Euclidean otherwise; first candidate wins ties, like np.argsort()[0] on <= 16
    entries), append it to the user's history; a user stops once it has reached its target.

    hist     [B,Lh] int64, right-aligned (pre-padded with 0) histories; not modified
    predict_next(hist [B,Lh'], users [B], k_c) -> item ids [B,k_c] (e.g. ``predict_next_tail`` over BPR / FPMC factors)
    fv       [n_item+1, F] feature vectors, row = item id
    Returns (paths float64 np [B,P] -- 0 after the target has been reached --, n_success)."""
    import numpy as np
    dev = hist.device
    B = hist.shape[0]
    targets = torch.as_tensor(targets, device=dev).long()
    users = torch.as_tensor(users, device=dev).long()
    fv = torch.as_tensor(fv, device=dev)
    fv = fv.double() if not binary else fv
    hist = torch.cat([torch.zeros((B, max_path_len), dtype=hist.dtype, device=dev), hist], 1)    # room to grow, still right-aligned
    paths = torch.zeros((B, max_path_len), dtype=torch.float64, device=dev)
    stopped = torch.zeros((B,), dtype=torch.bool, device=dev)
    tfv = fv[targets].unsqueeze(1)                                                # [B,1,F]
    for i in range(max_path_len):
        cand = predict_next(hist, users, k_c).long()                                # [B,k_c]
        cfv = fv[cand.clamp(min=0)]
        dist = (cfv != tfv).sum(-1).double() if binary else (cfv - tfv).pow(2).sum(-1).sqrt()
        pick = cand.gather(1, torch.sort(dist, dim=1, stable=True).indices[:, :1])[:, 0]
        act = ~stopped
        paths[:, i] = torch.where(act, pick.double(), paths[:, i])
        grown = torch.cat([hist[:, 1:], pick.view(-1, 1)], 1)
        hist = torch.where(act.view(-1, 1), grown, hist)
        stopped = stopped | (act & (pick == targets))
    return paths.cpu().numpy(), int(stopped.sum())
