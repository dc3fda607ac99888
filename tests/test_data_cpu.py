"""CPU: the vectorised batch builders reproduce the reference's row-by-row collate functions
(data_provider.py:591-617 restated below as the checker)."""
import numpy as np
import torch

from influentialrs_b200.data import to_csr, collate_eval_irs, collate_train


def _reference_collate_eval_irs(batch, seq_len, gap_len):
    """Row-by-row restatement of DataLoaderEvalIRS._collate_fn (data_provider.py:591-617)."""
    raw_seqs, seqs, users, targets, labels = [], [], [], [], []
    for row in batch:
        seq, target = row[0], row[2]
        new_seq = np.zeros(seq_len)
        l_history = seq_len - gap_len - 1
        item_list = seq[-l_history:]
        start_history = -len(item_list) - gap_len - 1
        new_seq[start_history:start_history + len(item_list)] = item_list
        new_seq[-1] = target
        raw_seqs.append(torch.LongTensor(seq))
        seqs.append(new_seq)
        users.append(row[1]); targets.append(row[2]); labels.append(row[3])
    return raw_seqs, torch.LongTensor(np.array(seqs)), torch.LongTensor(users), torch.LongTensor(targets), torch.tensor(labels)


def test_collate_eval_irs_matches_reference_loop():
    rng = np.random.default_rng(3)
    hist = [rng.integers(1, 500, size=int(n)).tolist() for n in rng.integers(1, 90, size=40)]
    hist[0] = hist[0][:1]                                   # single-item history
    hist[1] = rng.integers(1, 500, size=300).tolist()       # longer than the window
    users, targets, labels = rng.integers(0, 50, 40), rng.integers(1, 500, 40), rng.integers(1, 500, 40)
    values, offsets = to_csr(hist)
    assert values.dtype == np.int32 and offsets[-1] == sum(len(h) for h in hist)
    for seq_len, gap_len in ((60, 0), (60, 20), (201, 0)):
        rows = rng.permutation(40)[:17]
        batch = [(np.array(hist[i]), users[i], targets[i], labels[i]) for i in rows]
        want = _reference_collate_eval_irs(batch, seq_len, gap_len)
        got = collate_eval_irs(values, offsets, rows, users[rows], targets[rows], labels[rows], seq_len, gap_len, pin=False)
        assert torch.equal(got[1], want[1]) and got[1].dtype == torch.int64
        assert all(torch.equal(a, b) for a, b in zip(got[0], want[0]))
        assert torch.equal(got[2], want[2]) and torch.equal(got[3], want[3]) and torch.equal(got[4], want[4])


def test_collate_train_casts_like_the_reference():
    w = np.array([[0., 0., 3., 7.], [1., 2., 3., 4.]])                # irs_valid_seq rows are float64
    seqs, users = collate_train(w, [5, 6], pin=False)
    assert seqs.dtype == torch.int64 and seqs.tolist() == [[0, 0, 3, 7], [1, 2, 3, 4]] and users.tolist() == [5, 6]
    seqs, users = collate_train(w, [5, 6], rows=[1], pin=False)
    assert seqs.tolist() == [[1, 2, 3, 4]] and users.tolist() == [6]


def test_collate_matches_the_imported_reference_loaders():
    """The same builders against outputs of the reference's own DataLoaderEvalIRS._collate_fn / DataLoaderIRS._collate_fn
    (data_provider.py:591-617, :568-575), produced by oracle/make_golden.py::collate_case."""
    from tests.helpers import load_golden
    _, g = load_golden("collate")
    lens = g["hist_lens"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    hist = [g["hist_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(lens))]
    values, offsets = to_csr(hist)
    for seq_len, gap_len in ((60, 0), (60, 20), (201, 0)):
        tag = f"L{seq_len}_g{gap_len}"
        rows = g[tag + "_rows"]
        raw, seqs, us, tg, lb = collate_eval_irs(values, offsets, rows, g["users"][rows], g["targets"][rows], g["labels"][rows],
                                                 seq_len, gap_len, pin=False)
        assert seqs.dtype == torch.int64 and np.array_equal(seqs.numpy(), g[tag + "_seqs"])
        assert np.array_equal(us.numpy(), g[tag + "_users"]) and np.array_equal(tg.numpy(), g[tag + "_targets"])
        assert np.array_equal(lb.numpy(), g[tag + "_labels"])
        assert [len(r) for r in raw] == g[tag + "_raw_lens"].tolist()
        assert np.array_equal(torch.cat(raw).numpy(), g[tag + "_raw_flat"])
    seqs, us = collate_train(g["train_windows"], g["train_users"], pin=False)
    assert np.array_equal(seqs.numpy(), g["train_seqs"]) and np.array_equal(us.numpy(), g["train_users"])
