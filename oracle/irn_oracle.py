"""CPU oracle for the IRN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (influentialrs_b200/) never does; it fails loudly when the
CUDA extension is missing.

This is a formula-level restatement (explicit matmuls on torch CPU tensors, fp32 by default,
fp64 on request) of the reference's PyTorch path.  It works from a plain ``state_dict`` whose keys
are the reference's own, so the same weights drive the reference, the oracle
and the CUDA path.  Every function cites the reference file:line it follows (paths relative to
/root/reference).

Parity status: the reference ships no tests / golden vectors.  The oracle is
therefore pinned against the *reference itself*, imported and run in the dev container by
oracle/make_golden.py (with the D1/D2/D3 shims of oracle/ref_shim.py); the resulting vectors are
committed under tests/golden/ and tests/test_oracle_golden.py checks this file against them.

Known, deliberate deviations from the shipped reference text:
  D1  the PIM is built through the keyword ``pi_factor=`` branch (model/influentialRS.py:144-151);
      the shipped positional call cannot run with batch > 1.
  D6  gap_len = 0 semantics only (shift-left branch, model/influentialRS.py:442-450).
  D7  "first of top-100 not in window" is extended to "best item not in window".
  ties are broken by lower item id (north_star); torch.topk/sort leave it unspecified.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

W_H_DEFAULT = 0.05   # model/influentialRS.py:122
W_OBJ_DEFAULT = 1.0  # model/influentialRS.py:123
NEG_INF = float("-inf")


# --------------------------------------------------------------------------------------------
# a1: embedding gather + positional encoding
# --------------------------------------------------------------------------------------------
def positional_table(max_len: int, d: int, dtype=torch.float32) -> torch.Tensor:
    """Sinusoidal table [max_len, d].  model/layers.py:21-29 (always built in fp32 there)."""
    pe = torch.zeros((max_len, d), dtype=torch.float32)
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def embed(ids: torch.Tensor, table: torch.Tensor, pe: torch.Tensor) -> torch.Tensor:
    """x[b,l,:] = E[ids[b,l]] * sqrt(d) + pe[l].  model/influentialRS.py:174-175, model/uRS.py:55.

    Two separately rounded fp32 operations (multiply, then add): the CUDA gather must not contract
    them into an FMA if it is to be bit-exact.
    """
    d = table.shape[1]
    L = ids.shape[1]
    return table[ids] * math.sqrt(d) + pe[:L].unsqueeze(0)


# --------------------------------------------------------------------------------------------
# a2/a3: PIF and the Personalized Impressionability Mask
# --------------------------------------------------------------------------------------------
def pif(sd: SD, users: torch.Tensor) -> torch.Tensor:
    """r_u = user_mask_layer(user_embedder(user)) -> [B,1].  model/influentialRS.py:180."""
    u = sd["user_embedder.weight"][users]
    return u @ sd["user_mask_layer.weight"].t() + sd["user_mask_layer.bias"]


def pim_mask(L: int, r_u: torch.Tensor, w_h: float = W_H_DEFAULT, w_obj: float = W_OBJ_DEFAULT) -> torch.Tensor:
    """[B,L,L] additive mask.  model/influentialRS.py:139-151 (keyword branch, D1).

    M[b,i,j] = w_obj*r_u[b] if j == L-1 (every row, incl. above the diagonal) else (w_h if j<=i
    else -inf).  The reference builds it in float64 and casts to float32 (:146,:185); r_u is fp32
    and w_obj = 1 so the cast is exact.
    """
    B = r_u.shape[0]
    i = torch.arange(L).unsqueeze(1)
    j = torch.arange(L).unsqueeze(0)
    base = torch.where(j <= i, torch.tensor(float(w_h), dtype=r_u.dtype), torch.tensor(NEG_INF, dtype=r_u.dtype))
    m = base.unsqueeze(0).repeat(B, 1, 1)
    last = (w_obj * r_u.reshape(B).to(torch.float64)).to(r_u.dtype)
    # keep autograd connectivity to r_u (the reference's in-place copy does, :149)
    m = torch.cat([m[:, :, : L - 1], last.reshape(B, 1, 1).expand(B, L, 1)], dim=2)
    return m


def causal_mask(L: int, dtype=torch.float32) -> torch.Tensor:
    """[L,L] 0 / -inf.  model/uRS.py:47-50."""
    i = torch.arange(L).unsqueeze(1)
    j = torch.arange(L).unsqueeze(0)
    return torch.where(j <= i, torch.tensor(0.0, dtype=dtype), torch.tensor(NEG_INF, dtype=dtype))


# --------------------------------------------------------------------------------------------
# a4: post-norm transformer decoder with an all-zero memory
# --------------------------------------------------------------------------------------------
def _mha(q_in, kv_in, in_w, in_b, out_w, out_b, H, add_mask=None, key_pad=None):
    """torch.nn.functional.multi_head_attention_forward restated, batch-first [B,L,d].

    in_proj split into thirds; heads are contiguous dh-slices; scale 1/sqrt(dh); float attn_mask
    and the 0/-inf key-padding mask are ADDED to the scores (functional.py:6215,6618-6620 in torch
    2.11) before the softmax.
    """
    B, Lq, d = q_in.shape
    Lk = kv_in.shape[1]
    dh = d // H
    wq, wk, wv = in_w[:d], in_w[d:2 * d], in_w[2 * d:]
    bq, bk, bv = in_b[:d], in_b[d:2 * d], in_b[2 * d:]
    q = (q_in @ wq.t() + bq).reshape(B, Lq, H, dh).transpose(1, 2)   # [B,H,Lq,dh]
    k = (kv_in @ wk.t() + bk).reshape(B, Lk, H, dh).transpose(1, 2)
    v = (kv_in @ wv.t() + bv).reshape(B, Lk, H, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))            # [B,H,Lq,Lk]
    if add_mask is not None:
        s = s + (add_mask.unsqueeze(1) if add_mask.dim() == 3 else add_mask)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], NEG_INF)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, d)
    return o @ out_w.t() + out_b


def decoder_layer(sd: SD, prefix: str, x, H, add_mask, key_pad, mem_len: int, fold_cross: bool = False):
    """One nn.TransformerDecoderLayer, norm_first=False, relu (torch/nn/modules/transformer.py
    _sa_block -> norm1 -> _mha_block -> norm2 -> _ff_block -> norm3), as configured at
    model/influentialRS.py:67-74.  ``mem_len`` rows of zero memory (:172-173)."""
    p = lambda n: sd[prefix + n]
    d = x.shape[-1]
    sa = _mha(x, x, p("self_attn.in_proj_weight"), p("self_attn.in_proj_bias"),
              p("self_attn.out_proj.weight"), p("self_attn.out_proj.bias"), H, add_mask, key_pad)
    x = F.layer_norm(x + sa, (d,), p("norm1.weight"), p("norm1.bias"), 1e-5)
    if fold_cross:
        # softmax over identical keys is uniform -> output is W_o b_v + b_o exactly
        bv = p("multihead_attn.in_proj_bias")[2 * d:]
        ca = (bv @ p("multihead_attn.out_proj.weight").t() + p("multihead_attn.out_proj.bias")).expand_as(x)
    else:
        mem = torch.zeros(x.shape[0], mem_len, d, dtype=x.dtype)
        ca = _mha(x, mem, p("multihead_attn.in_proj_weight"), p("multihead_attn.in_proj_bias"),
                  p("multihead_attn.out_proj.weight"), p("multihead_attn.out_proj.bias"), H)
    x = F.layer_norm(x + ca, (d,), p("norm2.weight"), p("norm2.bias"), 1e-5)
    ff = torch.relu(x @ p("linear1.weight").t() + p("linear1.bias")) @ p("linear2.weight").t() + p("linear2.bias")
    x = F.layer_norm(x + ff, (d,), p("norm3.weight"), p("norm3.bias"), 1e-5)
    return x


def n_layers_of(sd: SD) -> int:
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))


def irn_decoding(sd: SD, seqs: torch.Tensor, users: torch.Tensor, n_heads: int,
                 w_h: float = W_H_DEFAULT, w_obj: float = W_OBJ_DEFAULT, fold_cross: bool = False):
    """InfluentialNet.decoding in eval mode (dropout off).  model/influentialRS.py:157-200 (D1).
    Returns (h [B,L,d], r_u [B,1])."""
    pe = sd["pos_embedder.pe"][0]
    x = embed(seqs, sd["item_embedder.weight"], pe)
    r_u = pif(sd, users)
    mask = pim_mask(seqs.shape[1], r_u, w_h, w_obj)
    pad = seqs.eq(0)
    for i in range(n_layers_of(sd)):
        x = decoder_layer(sd, f"decoder.layers.{i}.", x, n_heads, mask, pad, pe.shape[0], fold_cross)
    return x, r_u


def irn_forward(sd: SD, seqs, users, n_heads, **kw) -> torch.Tensor:
    """InfluentialNet.forward -> logits [B,L,N].  model/influentialRS.py:202-216 (use_u False)."""
    h, _ = irn_decoding(sd, seqs, users, n_heads, **kw)
    return h @ sd["project.weight"].t() + sd["project.bias"]


def samplenet_decoding(sd: SD, seqs: torch.Tensor, n_heads: int, fold_cross: bool = False) -> torch.Tensor:
    """SampleNet.decoding: IRN decoder with the plain causal mask, no user.  model/uRS.py:52-64."""
    pe = sd["pos_embedder.pe"][0]
    x = embed(seqs, sd["word_embedder.weight"], pe)
    mask = causal_mask(seqs.shape[1], x.dtype)
    pad = seqs.eq(0)
    for i in range(n_layers_of(sd)):
        x = decoder_layer(sd, f"decoder.layers.{i}.", x, n_heads, mask, pad, pe.shape[0], fold_cross)
    return x


def samplenet_forward(sd: SD, seqs, n_heads, **kw) -> torch.Tensor:
    """SampleNet.forward.  model/uRS.py:66-69."""
    return samplenet_decoding(sd, seqs, n_heads, **kw) @ sd["project.weight"].t() + sd["project.bias"]


# --------------------------------------------------------------------------------------------
# a6: softmax cross-entropy over the catalog
# --------------------------------------------------------------------------------------------
def ce_rows(seqs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row selector of train_batch: rows (b, l<L-1) whose next id is non-pad; class = id-1.
    model/influentialRS.py:294-303."""
    tgt = seqs[:, 1:].reshape(-1)
    sel = tgt.gt(0)
    return sel, tgt[sel] - 1


def irn_loss(sd: SD, seqs, users, n_heads, **kw) -> torch.Tensor:
    """IRSNN.train_batch / get_loss_on_eval_data loss value (mean CE over selected rows).
    model/influentialRS.py:252-276, :278-303."""
    logits = irn_forward(sd, seqs, users, n_heads, **kw)
    N = logits.shape[-1]
    out = logits[:, :-1, :].reshape(-1, N)
    sel, cls = ce_rows(seqs)
    return F.cross_entropy(out[sel], cls)


def irn_loss_and_grads(sd: SD, seqs, users, n_heads, **kw):
    """Loss + autograd gradients w.r.t. every floating-point parameter (the oracle for K2/K4/K5b)."""
    leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(".pe") else v)
            for k, v in sd.items()}
    loss = irn_loss(leaf, seqs, users, n_heads, **kw)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v))
             for k, v in leaf.items() if v.is_floating_point() and v.requires_grad}
    # nn.Embedding(padding_idx=0) never accumulates into row 0 (model/influentialRS.py:111)
    grads["item_embedder.weight"][0].zero_()
    return loss.detach(), grads


def samplenet_loss(sd: SD, target: torch.Tensor, n_heads, **kw) -> torch.Tensor:
    """Evaluator.train_batch / get_loss_on_eval_data loss.  model/evaluator.py:53-92."""
    logits = samplenet_forward(sd, target[:, :-1], n_heads, **kw)
    N = logits.shape[-1]
    tgt = target[:, 1:].reshape(-1)
    sel = tgt.gt(0)
    return F.cross_entropy(logits.reshape(-1, N)[sel], tgt[sel] - 1)


# --------------------------------------------------------------------------------------------
# fused-scorer semantics: exclusion, arg-max / top-k with (-score, id) order, rank, lse+gather
# --------------------------------------------------------------------------------------------
def topk_excluding(scores: torch.Tensor, excl: Optional[torch.Tensor], k: int, item_base: int = 1):
    """Top-k of each row of ``scores`` [M,N] (column j <-> item j+item_base) over items not in
    ``excl`` [M,Lx] (0 entries ignored), ordered by (-score, item id).  Returns (vals, items)."""
    s = scores.clone()
    if excl is not None:
        M, N = s.shape
        col = excl.long() - item_base
        ok = (col >= 0) & (col < N)
        rows = torch.arange(M).unsqueeze(1).expand_as(col)
        s[rows[ok], col[ok]] = NEG_INF
    # stable sort on descending score keeps the lower column first among equals
    order = torch.sort(s, dim=1, descending=True, stable=True).indices[:, :k]
    return torch.gather(s, 1, order), order + item_base


def rank_excluding(scores: torch.Tensor, label: torch.Tensor, excl: Optional[torch.Tensor], item_base: int = 1):
    """1-based position of ``label`` in the (-score, id)-ordered list of non-excluded items;
    0 if the label itself is excluded (the reference then skips the sample,
    model/influentialRS.py:386).  Equals index+1 of ``(indices == label).nonzero()`` (:387-388)."""
    M, N = scores.shape
    col = label.long() - item_base
    sl = scores[torch.arange(M), col].unsqueeze(1)
    j = torch.arange(N).unsqueeze(0)
    ahead = (scores > sl) | ((scores == sl) & (j < col.unsqueeze(1)))
    label_excluded = torch.zeros(M, dtype=torch.bool)
    if excl is not None:
        ecol = excl.long() - item_base
        ok = (ecol >= 0) & (ecol < N)
        keep = torch.ones(M, N, dtype=torch.bool)
        rows = torch.arange(M).unsqueeze(1).expand_as(ecol)
        keep[rows[ok], ecol[ok]] = False
        ahead &= keep
        label_excluded = ~keep[torch.arange(M), col]
    r = ahead.sum(1) + 1
    r[label_excluded] = 0
    return r


def lse_gather(scores: torch.Tensor, sel: torch.Tensor, item_base: int = 1):
    """(logsumexp over the row, raw logits of the selected items).  log-prob = logit - lse
    (model/evaluator.py:194-205)."""
    lse = torch.logsumexp(scores, dim=1)
    col = (sel.long() - item_base).clamp(min=0)
    return lse, torch.gather(scores, 1, col)


def delete_item_in_history(sorted_items: torch.Tensor, hist: torch.Tensor, h: Optional[int] = None):
    """utils.py:8-12 (h=50 window) / model/influentialRS.py:312-323 (h=None)."""
    hh = hist if h is None else hist[-h:]
    return sorted_items[~sorted_items.unsqueeze(1).eq(hh).any(1)]


# --------------------------------------------------------------------------------------------
# a7: influence-path generation
# --------------------------------------------------------------------------------------------
def generate_paths(sd: SD, seqs, users, targets, n_heads, max_path_len=20, return_margins=False, **kw):
    """Batched restatement of IRSNN.get_seq_in_batch for gap_len=0, sample=False
    (model/influentialRS.py:392-470; validated against it):
    only row L-2 is read; the pick is the arg-max raw logit over items not in temp[:, :L-1]
    (softmax is monotone); then temp <- [temp[:,1:L-1], next, target]; finally the path is zeroed
    after the first occurrence of the target (:452-467).

    Returns (paths f32 [B,P], targets i64 [B], actual_history list, n_early_success) as the
    reference does, plus (optionally) the top1-top2 margin of every decision."""
    temp = seqs.clone()
    B, L = temp.shape
    p = L - 2
    paths = torch.zeros((B, max_path_len))
    margins = torch.zeros((B, max_path_len))
    W, beta = sd["project.weight"], sd["project.bias"]
    for i in range(max_path_len):
        h, _ = irn_decoding(sd, temp, users, n_heads, **kw)
        s = h[:, p, :] @ W.t() + beta                         # [B,N]
        vals, items = topk_excluding(s, temp[:, : p + 1], 2)
        nxt = items[:, 0]
        margins[:, i] = vals[:, 0] - vals[:, 1]
        paths[:, i] = nxt.float()
        temp = torch.cat([temp[:, 1: L - 1], nxt.unsqueeze(1), temp[:, L - 1:]], dim=1)
    paths = paths.numpy()
    tg = targets.numpy()
    hist = seqs[:, :-1].numpy()
    n_early = 0
    actual = []
    for b in range(B):
        pos = np.where(paths[b] == tg[b])[0]
        if len(pos):
            n_early += 1
            paths[b, pos[0] + 1:] = 0
        actual.append(hist[b][hist[b] != 0])
    if return_margins:
        return paths, tg, actual, n_early, margins.numpy()
    return paths, tg, actual, n_early


def generate_paths_faithful(sd: SD, seqs, users, targets, n_heads, max_path_len=20, **kw):
    """Per-sample loop exactly as the reference writes it (softmax over [B,L,N], topk(100), window
    filter, first survivor; model/influentialRS.py:412-450), gap_len=0.  Small cases only; this is
    also the shape of work the CPU baseline times."""
    temp = seqs.clone()
    B, L = temp.shape
    p = L - 2
    paths = torch.zeros((B, max_path_len))
    for i in range(max_path_len):
        out = torch.softmax(irn_forward(sd, temp, users, n_heads, **kw), dim=2)
        for b in range(B):
            k = min(100, out.shape[2])
            _, ind = out[b][p].topk(k)
            ind = ind + 1
            keep = ~ind.unsqueeze(1).eq(temp[b][: p + 1]).any(1)
            nxt = ind[keep][0].item()
            paths[b][i] = nxt
            row = torch.zeros(L, dtype=torch.long)
            row[:-2] = temp[b][1:-1]
            row[-2] = nxt
            row[-1] = temp[b][-1]
            temp[b] = row
    paths = paths.numpy()
    tg = targets.numpy()
    n_early = 0
    for b in range(B):
        pos = np.where(paths[b] == tg[b])[0]
        if len(pos):
            n_early += 1
            paths[b, pos[0] + 1:] = 0
    return paths, tg, n_early


# --------------------------------------------------------------------------------------------
# a8: accuracy metrics
# --------------------------------------------------------------------------------------------
def accuracy_metrics(sd: SD, raw: Sequence[torch.Tensor], seqs, users, labels, n_heads,
                     top_k=20, gap_len=0, use_h=True, **kw):
    """IRSNN.get_accuracy_metrics_in_batch (model/influentialRS.py:340-390) by counting:
    rank = 1 + #{items not in raw history ahead of the label}.  Returns (hit_count, rr array)."""
    B, L = seqs.shape
    pos = L - (gap_len + 1) - 1
    h, _ = irn_decoding(sd, seqs, users, n_heads, **kw)
    s = h[:, pos, :] @ sd["project.weight"].t() + sd["project.bias"]
    hit, rr = 0, []
    for b in range(B):
        excl = raw[b].reshape(1, -1) if use_h else None
        r = int(rank_excluding(s[b:b + 1], labels[b:b + 1], excl)[0])
        if r > 0:
            if r <= top_k:
                hit += 1
            rr.append(1.0 / r)
    return hit, np.array(rr)


# --------------------------------------------------------------------------------------------
# a11: evaluator measurements
# --------------------------------------------------------------------------------------------
def _first_zero_minus1(row: torch.Tensor) -> int:
    """Evaluator._get_first_none_zero_index.  model/evaluator.py:136-144."""
    z = (row == 0).nonzero()
    return row.shape[0] - 1 if len(z) == 0 else z[0].item() - 1


def _last_path_index(row: torch.Tensor, target) -> int:
    """Evaluator._get_last_path_index.  model/evaluator.py:146-154."""
    t = (row == target).nonzero()
    return _first_zero_minus1(row) if len(t) == 0 else t[0].item() - 1


def evaluator_pp(sd: SD, new_seqs, start_pos, l_paths, n_heads, **kw) -> List[float]:
    """Evaluator.get_pp_in_batch: per-sequence mean CE over the path rows.  model/evaluator.py:292-323."""
    logits = samplenet_forward(sd, new_seqs[:, :-1], n_heads, **kw)
    out = []
    for i in range(new_seqs.shape[0]):
        l, r = int(start_pos[i]), int(start_pos[i]) + int(l_paths[i])
        tgt = new_seqs[i][l:r]
        m = tgt.gt(0)
        out.append(F.cross_entropy(logits[i][l - 1:r - 1][m], tgt[m] - 1).item())
    return out


def evaluator_rr_increase(sd: SD, histories, new_seqs, targets, n_heads, **kw):
    """Evaluator.get_rr_increase_in_batch.  model/evaluator.py:245-290."""
    def ranks(seqs, end_fn):
        dec = seqs[:, :-1].clone()
        logits = samplenet_forward(sd, dec, n_heads, **kw)
        rs = []
        for i in range(dec.shape[0]):
            end = end_fn(dec[i], targets[i])
            r = rank_excluding(logits[i][end].unsqueeze(0), targets[i:i + 1], dec[i][: end + 1].unsqueeze(0))
            rs.append(int(r[0]))
        return rs
    begin_r = ranks(histories, lambda row, t: _first_zero_minus1(row))
    end_r = ranks(new_seqs, _last_path_index)
    irr = np.array([1 / end_r[i] - 1 / begin_r[i] for i in range(len(end_r))])
    ir = np.array([end_r[i] - begin_r[i] for i in range(len(end_r))])
    return irr, ir


def evaluator_grad(sd: SD, histories, new_seqs, targets, start_pos, l_paths, n_heads, **kw):
    """Evaluator.get_grad_in_batch (model/evaluator.py:162-243): log-prob of the next path item and
    of the target at each path step, history appended/shifted in place.  Returns
    (t_probs [B,S], p_probs [B,S], avg_ps, iois)."""
    B = new_seqs.shape[0]
    paths = [new_seqs[i][int(start_pos[i]): int(start_pos[i]) + int(l_paths[i])].numpy() for i in range(B)]
    S = int(max(int(x) for x in l_paths))
    temp = histories[:, :-1].clone()
    Lh = temp.shape[1]
    t_probs = np.zeros((B, S))
    p_probs = np.zeros((B, S))
    for i in range(S):
        lp = torch.log_softmax(samplenet_forward(sd, temp, n_heads, **kw), dim=2)
        for j in range(B):
            end = _first_zero_minus1(temp[j])
            if i < int(l_paths[j]):
                nxt = int(paths[j][i])
                p_probs[j, i] = lp[j][end][nxt - 1].item()
                t_probs[j, i] = lp[j][end][int(targets[j]) - 1].item()
            else:
                nxt = 0
            if end == Lh - 1:
                row = torch.zeros(Lh, dtype=torch.long)
                row[:-1] = temp[j][1:]
                row[-1] = nxt
                temp[j] = row
            else:
                temp[j][end + 1] = nxt
    avg_ps, iois = [], []
    for i in range(B):
        tp = t_probs[i][t_probs[i] < 0]
        pp = p_probs[i][p_probs[i] < 0]
        iois.append(tp[-1] - tp[0])
        avg_ps.append(sum(pp) / len(pp))
    return t_probs, p_probs, avg_ps, iois


# --------------------------------------------------------------------------------------------
# a12 / a13: baseline scorers
# --------------------------------------------------------------------------------------------
def sas_log2feats(sd: SD, log_seqs: torch.Tensor, rat_seqs: torch.Tensor, n_heads: int) -> torch.Tensor:
    """SAS.log2feats in eval mode.  model/sas.py:154-190.  Keeps the D9 quirk: the rating ids are
    looked up in the ITEM table (:162).  MHA here has a boolean causal mask and no key padding."""
    E = sd["item_emb.weight"]
    C = E.shape[1]
    T = log_seqs.shape[1]
    x = E[log_seqs] * (C ** 0.5)
    x = x + sd["pos_emb.weight"][:T].unsqueeze(0)
    x = x + E[rat_seqs]
    keep = (log_seqs != 0).unsqueeze(-1)
    x = x * keep
    cm = causal_mask(T, x.dtype)
    nb = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("attention_layers."))
    for i in range(nb):
        q = F.layer_norm(x, (C,), sd[f"attention_layernorms.{i}.weight"], sd[f"attention_layernorms.{i}.bias"], 1e-8)
        in_w, in_b = sd[f"attention_layers.{i}.in_proj_weight"], sd[f"attention_layers.{i}.in_proj_bias"]
        # query from LN(x), keys/values from x (model/sas.py:174-177)
        dh = C // n_heads
        B = x.shape[0]
        qq = (q @ in_w[:C].t() + in_b[:C]).reshape(B, T, n_heads, dh).transpose(1, 2)
        kk = (x @ in_w[C:2 * C].t() + in_b[C:2 * C]).reshape(B, T, n_heads, dh).transpose(1, 2)
        vv = (x @ in_w[2 * C:].t() + in_b[2 * C:]).reshape(B, T, n_heads, dh).transpose(1, 2)
        s = (qq @ kk.transpose(-1, -2)) / math.sqrt(dh) + cm
        o = (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B, T, C)
        o = o @ sd[f"attention_layers.{i}.out_proj.weight"].t() + sd[f"attention_layers.{i}.out_proj.bias"]
        x = q + o
        x = F.layer_norm(x, (C,), sd[f"forward_layernorms.{i}.weight"], sd[f"forward_layernorms.{i}.bias"], 1e-8)
        w1, b1 = sd[f"forward_layers.{i}.conv1.weight"][:, :, 0], sd[f"forward_layers.{i}.conv1.bias"]
        w2, b2 = sd[f"forward_layers.{i}.conv2.weight"][:, :, 0], sd[f"forward_layers.{i}.conv2.bias"]
        x = x + (torch.relu(x @ w1.t() + b1) @ w2.t() + b2)           # model/sas.py:90-98
        x = x * keep
    return F.layer_norm(x, (C,), sd["last_layernorm.weight"], sd["last_layernorm.bias"], 1e-8)


def sas_predict(sd: SD, log_seqs, rat_seqs, n_heads) -> torch.Tensor:
    """SAS.predict logits [B,N] = E[1..N] . f_last.  model/sas.py:208-228."""
    f = sas_log2feats(sd, log_seqs, rat_seqs, n_heads)[:, -1, :]
    return f @ sd["item_emb.weight"][1:].t()


def caser_scores(x: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """Caser for_pred scoring over the whole catalog: score[b,j] = x[b].W2[j+1] + b2[j+1].
    model/caser.py:172-179 with item_var = arange(N)+1 (:281)."""
    return x @ W2[1:].t() + b2[1:, 0]


def predict_next_tail(scores: torch.Tensor, hist: Optional[torch.Tensor], top_k: int, h: int = 50):
    """sort(descending) -> +1 -> delete_item_in_history(last h) -> [:top_k].
    model/sas.py:380-386, model/caser.py:291-298, utils.py:8-12."""
    out = []
    for b in range(scores.shape[0]):
        excl = None if hist is None else hist[b][-h:].reshape(1, -1)
        out.append(topk_excluding(scores[b:b + 1], excl, top_k)[1][0])
    return torch.stack(out)


# --------------------------------------------------------------------------------------------
# synthetic weights with the reference's default initialisers (for sizes with no golden file)
# --------------------------------------------------------------------------------------------
def synth_irn_state(n_item, n_user, max_len, d, n_layers, ffn, u_d=10, seed=1234, dtype=torch.float32) -> SD:
    """Random IRN state_dict with the reference's key names/shapes and
    init distributions of nn.Embedding / nn.Linear / nn.MultiheadAttention / nn.LayerNorm."""
    g = torch.Generator().manual_seed(seed)
    def unif(shape, bound):
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
    sd: SD = {}
    E = torch.randn((n_item + 1, d), generator=g, dtype=dtype)
    E[0] = 0
    sd["item_embedder.weight"] = E
    sd["user_embedder.weight"] = torch.randn((n_user, u_d), generator=g, dtype=dtype)
    sd["pos_embedder.pe"] = positional_table(max_len, d, dtype).unsqueeze(0)
    for i in range(n_layers):
        p = f"decoder.layers.{i}."
        for a in ("self_attn", "multihead_attn"):
            sd[p + a + ".in_proj_weight"] = unif((3 * d, d), math.sqrt(6.0 / (4 * d)))   # xavier_uniform
            sd[p + a + ".in_proj_bias"] = torch.zeros(3 * d, dtype=dtype)
            sd[p + a + ".out_proj.weight"] = unif((d, d), 1 / math.sqrt(d))
            sd[p + a + ".out_proj.bias"] = torch.zeros(d, dtype=dtype)
        sd[p + "linear1.weight"] = unif((ffn, d), 1 / math.sqrt(d))
        sd[p + "linear1.bias"] = unif((ffn,), 1 / math.sqrt(d))
        sd[p + "linear2.weight"] = unif((d, ffn), 1 / math.sqrt(ffn))
        sd[p + "linear2.bias"] = unif((d,), 1 / math.sqrt(ffn))
        for n in ("norm1", "norm2", "norm3"):
            sd[p + n + ".weight"] = torch.ones(d, dtype=dtype)
            sd[p + n + ".bias"] = torch.zeros(d, dtype=dtype)
    sd["user_mask_layer.weight"] = unif((1, u_d), 1 / math.sqrt(u_d))
    sd["user_mask_layer.bias"] = unif((1,), 1 / math.sqrt(u_d))
    sd["project.weight"] = unif((n_item, d), 1 / math.sqrt(d))
    sd["project.bias"] = unif((n_item,), 1 / math.sqrt(d))
    return sd


def to_dtype(sd: SD, dtype) -> SD:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
