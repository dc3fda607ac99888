"""Drop-in ``Evaluator`` (reference: model/evaluator.py:22-323): probability / rank measurements of
influence paths with an independent next-item RS (``SampleNet``).

Same constructor and method signatures / return types as the reference.  The reference materialises
[B,L,N] logits (+ LogSoftmax) for every call and then walks the batch in Python with ``.item()`` per
sample; here each measurement is: one decode, a gather of the one decoder row per sample that is read,
and the fused catalog scorer (log-sum-exp + selected logits, or rank by counting) -- the logits never
exist.  All per-sample index arithmetic (end of sequence, end of path, history append / shift) runs as
tensor ops on the device."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.optim import lr_scheduler

from . import ops


def _first_zero_minus1(seqs):
    """Evaluator._get_first_none_zero_index per row (model/evaluator.py:136-144): index of the first 0
    minus one, or len-1 when there is none.  seqs [B,L] -> int64 [B] (can be -1)."""
    is0 = seqs.eq(0)
    first0 = is0.float().argmax(1)
    return torch.where(is0.any(1), first0 - 1, torch.full_like(first0, seqs.shape[1] - 1))


def _last_path_index(seqs, targets):
    """Evaluator._get_last_path_index per row (model/evaluator.py:146-154)."""
    hit = seqs.eq(targets.view(-1, 1))
    first = hit.float().argmax(1)
    return torch.where(hit.any(1), first - 1, _first_zero_minus1(seqs))


def _prefix_ids(seqs, end):
    """Exclusion ids of seqs[b, :end[b]+1], rest zeroed (model/evaluator.py:266,284 ``dec_seqs[i][:end+1]``)."""
    L = seqs.shape[1]
    keep = torch.arange(L, device=seqs.device).view(1, L) <= end.view(-1, 1)
    return torch.where(keep, seqs, torch.zeros_like(seqs)).contiguous()


class Evaluator(nn.Module):
    """model/evaluator.py:22-323."""

    def __init__(self, config, net, device):
        super().__init__()
        self.PAD_ID = 0
        self.vocab_size = config.n_item
        self.net = net
        self.device = device
        self.softmax = nn.LogSoftmax(dim=2)
        self.loss_function = nn.CrossEntropyLoss()
        self.optimizer = optim.Adam(filter(lambda x: x.requires_grad, self.net.parameters()),
                                    betas=(0.9, 0.98), eps=1e-09, lr=config.lr1)
        self.pla_lr_scheduler = lr_scheduler.ReduceLROnPlateau(self.optimizer, factor=0.5, patience=4)
        self.scorer = None     # dist.ShardedScorer: catalog-sharded rank / log-prob across GPUs

    def _rank(self, rows, label, excl_ids):
        """Rank of ``label`` among the items not in ``excl_ids`` [B,Lx] (raw ids, 0 = pad; None = no filter)."""
        if self.scorer is not None:
            return self.scorer.rank(rows, label, excl_ids)
        excl = None if excl_ids is None else ops.sort_exclusions(excl_ids, self.vocab_size, 1)
        return ops.score_rank(rows, self.net.project.weight, self.net.project.bias, label, excl, 1)

    def _lse_gather(self, rows, sel):
        if self.scorer is not None:
            return self.scorer.lse_gather(rows, sel)
        return ops.score_lse_gather(rows, self.net.project.weight, self.net.project.bias, sel, 1)

    # -- loss ---------------------------------------------------------------------------------------
    def _ce(self, target):
        """mean CE over rows (b, l) whose next id is non-pad (model/evaluator.py:53-66,80-92)."""
        h = self.net.decoding(target[:, :-1])                                  # [B,L-1,d]
        d = h.shape[-1]
        tgt = target[:, 1:].reshape(-1)
        rows = torch.nonzero(tgt > self.PAD_ID).squeeze(1)
        hm = h.reshape(-1, d).index_select(0, rows)
        return ops.softmax_ce_mean(hm, self.net.project.weight, self.net.project.bias, tgt.index_select(0, rows) - 1)

    def train_batch(self, target):
        self.net.train()
        loss = self._ce(target)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item()

    def get_loss_on_eval_data(self, eval_data):
        self.net.eval()
        with torch.no_grad():
            return self._ce(eval_data).item()

    # -- accuracy -----------------------------------------------------------------------------------
    def get_accuracy_metrics_in_batch(self, seqs, top_k=20, use_h=True):
        """Hit@top_k count and reciprocal ranks of the label = last item of each post-padded sequence
        (model/evaluator.py:94-133), by counting the items ahead of the label."""
        self.net.eval()
        B, L = seqs.shape
        with torch.no_grad():
            h = self.net.decoding(seqs[:, :-1])                                # [B,L-1,d]
            end_pos = _first_zero_minus1(seqs)                                 # get_end_index (model/layers.py:34-42)
            label = seqs.gather(1, (end_pos % L).view(-1, 1)).squeeze(1)
            rows = h[torch.arange(B, device=seqs.device), (end_pos - 1) % (L - 1)]
            rank = self._rank(rows, label, _prefix_ids(seqs, end_pos - 1) if use_h else None).cpu().numpy()
        found = rank > 0
        return int(((rank <= top_k) & found).sum()), np.reciprocal(rank[found].astype(np.float64))

    # -- influence-path measurements --------------------------------------------------------------------
    def get_grad_in_batch(self, histories, new_seqs, targets, start_pos, l_paths):
        """Log-probability of the next path item and of the target at every path step, with the history
        appended / shifted in place (model/evaluator.py:162-243).  Like the reference, ``histories[:, :-1]``
        is updated in place.  Returns (t_probs [B,S], p_probs [B,S], avg_ps, iois)."""
        self.net.eval()
        dev = histories.device
        B = new_seqs.size(0)
        start_pos = torch.as_tensor(start_pos, device=dev).long()
        l_paths_t = torch.as_tensor(l_paths, device=dev).long()
        targets_t = torch.as_tensor(targets, device=dev).long()
        S = int(l_paths_t.max().item())
        temp = histories[:, :-1]
        Lh = temp.shape[1]
        ar = torch.arange(B, device=dev)
        W, beta = self.net.project.weight, self.net.project.bias
        t_cols, p_cols = [], []
        with torch.no_grad():
            for i in range(S):
                h = self.net.decoding(temp)                                    # [B,Lh,d]
                end = _first_zero_minus1(temp)
                act = l_paths_t > i
                pos = (start_pos + i).clamp(max=new_seqs.shape[1] - 1)
                nxt = torch.where(act, new_seqs.gather(1, pos.view(-1, 1)).squeeze(1), torch.zeros_like(targets_t))
                rows = h[ar, end % Lh]
                sel = torch.stack([nxt, torch.where(act, targets_t, torch.zeros_like(targets_t))], 1)
                lse, logit = self._lse_gather(rows, sel)
                logp = torch.where(act.view(-1, 1), logit - lse.view(-1, 1), torch.zeros_like(logit))
                p_cols.append(logp[:, 0])
                t_cols.append(logp[:, 1])
                full = end == Lh - 1
                shifted = torch.cat([temp[:, 1:], nxt.view(-1, 1)], 1)
                appended = temp.scatter(1, (end + 1).clamp(max=Lh - 1).view(-1, 1), nxt.view(-1, 1))
                temp.copy_(torch.where(full.view(-1, 1), shifted, appended))
        t_probs = torch.stack(t_cols, 1).double().cpu().numpy()
        p_probs = torch.stack(p_cols, 1).double().cpu().numpy()
        avg_ps, iois = [], []
        for i in range(B):
            tp = t_probs[i][t_probs[i] < 0]
            pp = p_probs[i][p_probs[i] < 0]
            iois.append(tp[-1] - tp[0])
            avg_ps.append(sum(pp) / len(pp))
        return t_probs, p_probs, avg_ps, iois

    def _rank_rows(self, dec_seqs, end, targets):
        B, L = dec_seqs.shape
        h = self.net.decoding(dec_seqs)
        rows = h[torch.arange(B, device=dec_seqs.device), end % L]
        rank = self._rank(rows, targets, _prefix_ids(dec_seqs, end)).cpu().numpy()
        bad = bool((rank <= 0).any())
        if self.scorer is not None:                   # catalog-sharded: every rank must raise together (see any_rank)
            bad = self.scorer.any_rank(bad, dec_seqs.device)
        if bad:
            raise IndexError("target item is part of the history (the reference fails on this input too)")
        return rank

    def get_rr_increase_in_batch(self, histories, new_seqs, targets):
        """Rank of the target (history-filtered) before vs after the path (model/evaluator.py:245-290)."""
        targets = torch.as_tensor(targets, device=histories.device).long()
        with torch.no_grad():
            dec = histories[:, :-1].clone()
            begin_r = self._rank_rows(dec, _first_zero_minus1(dec), targets)
            dec = new_seqs[:, :-1].clone()
            end_r = self._rank_rows(dec, _last_path_index(dec, targets), targets)
        irr = np.array([1 / int(end_r[i]) - 1 / int(begin_r[i]) for i in range(len(end_r))])
        ir = np.array([int(end_r[i]) - int(begin_r[i]) for i in range(len(end_r))])
        return irr, ir

    def get_pp_in_batch(self, new_seqs, start_pos, l_paths):
        """Per-sequence mean cross-entropy over the path rows (model/evaluator.py:292-323)."""
        self.net.eval()
        dev = new_seqs.device
        B, L = new_seqs.shape
        start_pos = torch.as_tensor(start_pos, device=dev).long().view(-1, 1)
        l_paths = torch.as_tensor(l_paths, device=dev).long().view(-1, 1)
        with torch.no_grad():
            h = self.net.decoding(new_seqs[:, :-1].clone())                    # [B,L-1,d]
            d = h.shape[-1]
            col = torch.arange(L, device=dev).view(1, L)
            in_path = (col >= start_pos) & (col < start_pos + l_paths) & new_seqs.gt(self.PAD_ID)   # target positions
            b_idx, t_idx = torch.nonzero(in_path, as_tuple=True)
            rows = h[b_idx, t_idx - 1]                                         # logits row that predicts position t
            lse, logit = self._lse_gather(rows, new_seqs[b_idx, t_idx].view(-1, 1))
            ce = (lse - logit[:, 0]).double()
            tot = torch.zeros(B, dtype=torch.float64, device=dev).index_add_(0, b_idx, ce)
            cnt = torch.zeros(B, dtype=torch.float64, device=dev).index_add_(0, b_idx, torch.ones_like(ce))
        return (tot / cnt).cpu().tolist()
