"""CPU, world_size 2 over gloo: the host logic of the multi-GPU paths (catalog shard bounds and
item_base, gather order, candidate merge with (score desc, id asc) ties, window bookkeeping,
gradient weighting) with the oracle's CPU scorer injected in place of the CUDA operators."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from types import SimpleNamespace

from oracle import irn_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gen_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from influentialrs_b200.dist import ShardedGenerator, shard_bounds
        torch.set_num_threads(1)
        N, L, H, P, Bl = 997, 14, 2, 5, 6                       # odd catalog size: uneven shards
        sd = O.synth_irn_state(N, 20, L, 32, 2, 64, seed=9)
        sd["project.weight"][500] = sd["project.weight"][10]    # exact tie straddling the shard boundary
        sd["project.bias"][500] = sd["project.bias"][10]
        g = torch.Generator().manual_seed(3)
        seqs = torch.zeros((world * Bl, L), dtype=torch.long)
        for b in range(world * Bl):
            n = L if b % 3 == 0 else int(torch.randint(3, L + 1, (1,), generator=g))
            seqs[b, L - n:] = torch.randperm(N, generator=g)[:n] + 1
        users = torch.randint(0, 20, (world * Bl,), generator=g)
        stub = SimpleNamespace(n_item=N, net=SimpleNamespace(project=SimpleNamespace(
            weight=sd["project.weight"], bias=sd["project.bias"])))

        def decode(win, us):
            return O.irn_decoding(sd, win, us, H, fold_cross=True)[0][:, L - 2]

        def score(self, h_all, win_all):
            s = h_all @ self.W.t() + self.b
            return O.topk_excluding(s, win_all[:, :-1], 1, item_base=self.lo + 1)

        def merge(vals, items):
            G, M, k = vals.shape
            v = vals.permute(1, 0, 2).reshape(M, G * k)
            it = items.permute(1, 0, 2).reshape(M, G * k)
            order = np.lexsort((it.numpy(), -v.numpy()), axis=1)[:, :k]
            order = torch.from_numpy(order)
            return v.gather(1, order), it.gather(1, order)

        def shift(win_all, nxt, paths, step, row0, n):
            win_all[:, :-2] = win_all[:, 1:-1].clone()
            win_all[:, -2] = nxt
            paths[:, step] = nxt[row0:row0 + n].float()

        sg = ShardedGenerator(stub, rank, world, decode_fn=decode, merge_fn=merge, shift_fn=shift)
        sg.score_fn = lambda h, w: score(sg, h, w)
        assert (sg.lo, sg.hi) == shard_bounds(N, rank, world)
        mine = slice(rank * Bl, (rank + 1) * Bl)
        paths, tg, hist, ne = sg.get_seq_in_batch(seqs[mine], users[mine], seqs[mine, -1], max_path_len=P)
        want, wtg, whist, _, margins = O.generate_paths(sd, seqs, users, seqs[:, -1], H, P, return_margins=True,
                                                        fold_cross=True)
        ok = margins[mine].min(1) > 1e-5
        np.testing.assert_array_equal(paths[ok], want[mine][ok])
        np.testing.assert_array_equal([len(x) for x in hist], [len(x) for x in whist[mine]])
        out.put((rank, "ok", int(ok.sum())))
    except Exception as e:  # pragma: no cover
        out.put((rank, "fail", repr(e)))
    finally:
        dist.destroy_process_group()


def _ragged_gen_worker(rank, world, port, out):
    """Ranks holding DIFFERENT numbers of users (a ragged last batch, one rank with none at all in the second tile):
    ShardedGenerator.generate pads to the largest local batch, get_seq_in_batch walks the same number of device tiles on
    every rank, and begin() refuses unequal batches instead of hanging in the collective."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from influentialrs_b200.dist import ShardedGenerator
        torch.set_num_threads(1)
        N, L, H, P = 400, 10, 2, 3
        sd = O.synth_irn_state(N, 20, L, 32, 2, 64, seed=11)
        g = torch.Generator().manual_seed(4)
        n_loc = [7, 3][rank]                                    # tiles of 4 users: rank 0 -> 4 + 3, rank 1 -> 3 + 0
        seqs_all = torch.zeros((10, L), dtype=torch.long)
        for b in range(10):
            n = L if b % 2 == 0 else int(torch.randint(3, L + 1, (1,), generator=g))
            seqs_all[b, L - n:] = torch.randperm(N, generator=g)[:n] + 1
        users_all = torch.randint(0, 20, (10,), generator=g)
        mine = slice(0, 7) if rank == 0 else slice(7, 10)
        stub = SimpleNamespace(n_item=N, user_tile=4, net=SimpleNamespace(project=SimpleNamespace(
            weight=sd["project.weight"], bias=sd["project.bias"])))

        def decode(win, us):
            h = O.irn_decoding(sd, win, us, H, fold_cross=True)[0][:, L - 2]
            return torch.nan_to_num(h)                          # an all-PAD pad window decodes to NaN; its row is dropped anyway

        def merge(vals, items):
            G, M, k = vals.shape
            v = vals.permute(1, 0, 2).reshape(M, G * k)
            it = items.permute(1, 0, 2).reshape(M, G * k)
            order = torch.from_numpy(np.lexsort((it.numpy(), -v.numpy()), axis=1)[:, :k])
            return v.gather(1, order), it.gather(1, order)

        def shift(win_all, nxt, paths, step, row0, n):
            win_all[:, :-2] = win_all[:, 1:-1].clone()
            win_all[:, -2] = nxt
            paths[:, step] = nxt[row0:row0 + n].float()

        sg = ShardedGenerator(stub, rank, world, decode_fn=decode, merge_fn=merge, shift_fn=shift)
        sg.score_fn = lambda h, w: O.topk_excluding(h @ sg.W.t() + sg.b, w[:, :-1], 1, item_base=sg.lo + 1)
        paths, tg, hist, ne = sg.get_seq_in_batch(seqs_all[mine], users_all[mine], seqs_all[mine, -1], max_path_len=P)
        assert paths.shape == (n_loc, P)
        want, _, _, _, margins = O.generate_paths(sd, seqs_all, users_all, seqs_all[:, -1], H, P, return_margins=True, fold_cross=True)
        ok = margins[mine].min(1) > 1e-5
        np.testing.assert_array_equal(paths[ok], want[mine][ok])
        # step()/begin() with unequal batches must raise on every rank, not hang
        try:
            sg._all = None
            sg.begin(seqs_all[mine].clone())
            raised = False
        except ValueError:
            raised = True
        assert raised
        out.put((rank, "ok", int(ok.sum())))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from influentialrs_b200.dist import make_data_parallel
        torch.set_num_threads(1)
        # two ranks with different row counts: the weighted all-reduce must equal the global mean gradient
        torch.manual_seed(0)
        w = torch.nn.Parameter(torch.randn(5, 3))
        net = torch.nn.ParameterList([w])
        x_all = torch.randn(10, 3)
        rows = [slice(0, 3), slice(3, 10)][rank]
        irn = SimpleNamespace(net=net, last_ce_rows=rows.stop - rows.start, grad_sync=None)
        make_data_parallel(irn)
        loss = (x_all[rows] @ w.t()).pow(2).sum(1).mean()
        loss.backward()
        irn.grad_sync()
        w2 = w.detach().clone().requires_grad_(True)
        (x_all @ w2.t()).pow(2).sum(1).mean().backward()
        torch.testing.assert_close(w.grad, w2.grad, rtol=1e-5, atol=1e-6)
        out.put((rank, "ok", 0))
    except Exception as e:  # pragma: no cover
        out.put((rank, "fail", repr(e)))
    finally:
        dist.destroy_process_group()


def _scorer_worker(rank, world, port, out):
    """ShardedScorer on CPU over gloo with plain-torch compute injected: rank / log-sum-exp + selected
    logits / top-k of a row-sharded catalog equal the unsharded oracle, with uneven shards, an exact tie straddling the
    shard boundary, labels inside the exclusion list, PAD selections, and ranks bringing DIFFERENT numbers of rows."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from influentialrs_b200.dist import ShardedScorer
        torch.set_num_threads(1)
        g = torch.Generator().manual_seed(5)
        N, d, Lx, k = 997, 16, 9, 7
        W = torch.randn((N, d), generator=g)
        bias = 0.1 * torch.randn((N,), generator=g)
        W[500], bias[500] = W[10], bias[10]                      # exact tie: item 11 (shard 0) == item 501 (shard 1)
        n_rows = [5, 8]                                           # ragged: rank 0 brings 5 rows, rank 1 brings 8
        h_all = torch.randn((sum(n_rows), d), generator=g)
        ids_all = torch.zeros((sum(n_rows), Lx), dtype=torch.long)
        for b in range(sum(n_rows)):
            n = int(torch.randint(0, Lx + 1, (1,), generator=g))
            ids_all[b, :n] = torch.randperm(N, generator=g)[:n] + 1
        label_all = torch.randint(1, N + 1, (sum(n_rows),), generator=g)
        label_all[0] = 501
        label_all[1] = 11
        label_all[2] = ids_all[2, 0] if ids_all[2, 0] > 0 else label_all[2]     # a label inside its own exclusion list
        sel_all = torch.randint(1, N + 1, (sum(n_rows), 2), generator=g)
        sel_all[3, 0] = 0                                                        # 0 = PAD selection -> logit 0.0
        sel_all[7, 1] = 0
        row0 = sum(n_rows[:rank])
        mine = slice(row0, row0 + n_rows[rank])

        def scores(self, h):
            return h @ self.W.t() + self.b

        def select(self, h, sel):
            s = scores(self, h)
            c = sel - (self.lo + 1)
            ok = (sel != 0) & (c >= 0) & (c < s.shape[1])
            return torch.where(ok, s.gather(1, c.clamp(0, s.shape[1] - 1)), torch.full_like(s[:, :1], float("-inf")).expand_as(c))

        def count(self, h, label, lab_s, ids):
            s = scores(self, h)
            col = torch.arange(s.shape[1]).unsqueeze(0) + self.lo + 1              # global ids of my columns
            live = torch.ones_like(s, dtype=torch.bool)
            flag = torch.zeros(h.shape[0], dtype=torch.int32)
            if ids is not None:
                for b in range(h.shape[0]):
                    ex = ids[b][(ids[b] > self.lo) & (ids[b] <= self.hi)]
                    live[b, ex - self.lo - 1] = False
                    if self.lo < int(label[b]) <= self.hi and int(label[b]) in ex.tolist():
                        flag[b] = 1
            ahead = live & (col != label.view(-1, 1)) & ((s > lab_s.view(-1, 1)) | ((s == lab_s.view(-1, 1)) & (col < label.view(-1, 1))))
            return ahead.sum(1), flag

        def topk(self, h, kk, ids):
            return O.topk_excluding(scores(self, h), ids, kk, item_base=self.lo + 1)

        def merge(vals, items):
            G, M, kk = vals.shape
            v = vals.permute(1, 0, 2).reshape(M, G * kk)
            it = items.permute(1, 0, 2).reshape(M, G * kk)
            order = torch.from_numpy(np.lexsort((it.numpy(), -v.numpy()), axis=1)[:, :kk])
            return v.gather(1, order), it.gather(1, order)

        sc = ShardedScorer(W, bias, rank, world, merge_fn=merge)
        sc.select_fn = lambda h, sel: select(sc, h, sel)
        sc.count_fn = lambda h, lab, ls, ids: count(sc, h, lab, ls, ids)
        sc.lse_fn = lambda h: torch.logsumexp(scores(sc, h), 1)
        sc.topk_fn = lambda h, kk, ids: topk(sc, h, kk, ids)
        full = h_all @ W.t() + bias
        want_rank = O.rank_excluding(full, label_all, ids_all)
        got_rank = sc.rank(h_all[mine], label_all[mine], ids_all[mine])
        assert torch.equal(got_rank, want_rank[mine]), (got_rank, want_rank[mine])
        assert int(want_rank[2]) == 0 or ids_all[2, 0] == 0
        want_lse, want_logit = O.lse_gather(full, sel_all)
        want_logit = torch.where(sel_all == 0, torch.zeros_like(want_logit), want_logit)
        got_lse, got_logit = sc.lse_gather(h_all[mine], sel_all[mine])
        torch.testing.assert_close(got_lse, want_lse[mine].float(), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(got_logit, want_logit[mine].float(), rtol=1e-6, atol=1e-6)
        # a data error seen by ONE rank is raised by all (otherwise the others wait in the next collective)
        assert sc.any_rank(rank == 1, torch.device("cpu")) is True and sc.any_rank(False, torch.device("cpu")) is False
        want_v, want_i = O.topk_excluding(full, ids_all, k)
        got_v, got_i = sc.topk(h_all[mine], k, ids_all[mine])
        assert torch.equal(got_i, want_i[mine])
        torch.testing.assert_close(got_v, want_v[mine].float(), rtol=1e-6, atol=1e-6)
        out.put((rank, "ok", 0))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run(worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"
    return res


def test_sharded_generation_matches_unsharded_oracle():
    res = _run(_gen_worker)
    assert all(r[2] >= 4 for r in res)


def test_sharded_generation_with_ragged_batches():
    _run(_ragged_gen_worker)


def test_sharded_scorer_matches_unsharded_oracle():
    _run(_scorer_worker)


def test_data_parallel_gradient_weighting():
    _run(_dp_worker)


def test_trim_paths_matches_reference_tail():
    from influentialrs_b200.dist import trim_paths
    paths = np.array([[5, 7, 9, 7, 2], [1, 2, 3, 4, 5], [8, 8, 8, 8, 8]], dtype=np.float32)
    targets = np.array([7, 9, 8])
    hist = np.array([[0, 0, 3, 4], [1, 2, 3, 4], [0, 9, 0, 1]])
    p, t, h, ne = trim_paths(paths.copy(), targets, hist)
    np.testing.assert_array_equal(p, [[5, 7, 0, 0, 0], [1, 2, 3, 4, 5], [8, 0, 0, 0, 0]])
    assert ne == 2 and [x.tolist() for x in h] == [[3, 4], [1, 2, 3, 4], [9, 1]]


def test_candidate_packing_is_bit_exact():
    """The (score, item) exchange packs the fp32 bit pattern into an int64 word: every float must survive, including
    -inf (a shard with no live column), -0.0, denormals and NaN payloads (the overflow marker of irs_score_topk)."""
    import torch
    from influentialrs_b200.dist import pack_candidates, unpack_candidates
    bits = torch.tensor([0x00000000, 0x80000000, 0xff800000, 0x7f800000, 0x7fc00001, 0xffc12345, 0x00000001, 0x807fffff,
                         0x3f800000, 0xc2f6e979], dtype=torch.int64)
    vals = bits.to(torch.int32).view(torch.float32).reshape(5, 2)
    items = torch.tensor([[-1, -2], [0, 1], [2**31 - 1, 2**40], [7, 8], [9, 10]], dtype=torch.int64)
    p = pack_candidates(vals, items)
    assert p.dtype == torch.int64 and p.shape == (5, 2, 2)
    v2, i2 = unpack_candidates(p)
    assert torch.equal(v2.view(torch.int32), vals.view(torch.int32))
    assert torch.equal(i2, items)
    # a strided view (what all_gather_into_tensor hands back after unsqueeze/cat) unpacks the same way
    q = torch.stack([p, p + 0])[1]
    v3, i3 = unpack_candidates(q)
    assert torch.equal(v3.view(torch.int32), vals.view(torch.int32)) and torch.equal(i3, items)
