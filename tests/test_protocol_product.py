"""The product's host protocol (ader_b200/data.py) against the oracle restatement and the
fixtures minted from the reference's own util.py: same batches, split, quotas, groups (CPU)."""
import json
import os
import random

import numpy as np
import pytest

from ader_b200 import data as D
from oracle import protocol as P

MAXLEN = 50


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "protocol.json")) as f:
        return json.load(f)


def test_product_protocol_matches_reference_fixture(gold, golden_dir):
    random.seed(0)
    np.random.seed(0)
    dl = D.DataLoader(os.path.join(golden_dir, "tiny_data"))
    train_sess, info = dl.train_loader(0)
    assert info == gold["train_info"]
    s = D.Sampler(train_sess, MAXLEN, 16)
    assert s.data_size() == gold["n_rows_before_split"]
    valid, train = s.split_data(0.1, return_train=True)
    assert valid == gold["valid_rows"] and len(train) == gold["n_train_rows"]
    assert s.batch_num() == gold["batch_num"]
    for ref in gold["batches"]:
        seq, pos = s.sampler()
        assert np.array(seq).tolist() == ref["seq"] and [int(p) for p in pos] == ref["pos"]
    test_sess, info = dl.evaluate_loader(1)
    assert info == gold["test_info"]
    assert test_sess == gold["test_sessions"]
    assert dl.max_item() == gold["max_item"]
    ev = D.Sampler(test_sess, MAXLEN, 8, is_subseq=False)
    assert ev.data_size() == gold["test_rows_total"]
    for ref in gold["test_batches"]:
        seq, pos = ev.sampler()
        assert np.array(seq).tolist() == ref["seq"] and [int(p) for p in pos] == ref["pos"]


def _sessions(seed, n=300, vmax=40):
    rng = np.random.RandomState(seed)
    return [rng.randint(1, vmax + 1, rng.randint(1, 9)).tolist() for _ in range(n)]


def test_sampler_stream_equivalence_with_oracle():
    data = _sessions(1)
    def run(cls):
        random.seed(5); np.random.seed(5)
        s = cls(data, MAXLEN, 32)
        v, t = s.split_data(0.1, True) if cls is D.Sampler else s.split_data(0.1)
        out = []
        for _ in range(s.batch_num() * 2 + 1):
            q, p = s.sampler()
            out.append((np.array(q).tolist(), [int(x) for x in p]))
        return v, t, out, random.random(), np.random.rand()
    assert run(D.Sampler) == run(P.RefSampler)


def test_epoch_order_equals_batchwise_iteration():
    data = _sessions(2)
    random.seed(9)
    a = D.Sampler(data, MAXLEN, 32, is_subseq=True)
    order = a.epoch_order()
    ra = random.random()
    random.seed(9)
    b = D.Sampler(data, MAXLEN, 32, is_subseq=True)
    rows = np.concatenate([b.next_indices() for _ in range(b.batch_num())])
    assert np.array_equal(order, rows) and ra == random.random()


def test_exemplar_generator_groups_and_quota_match_oracle():
    data = [s for s in _sessions(3, 400, 25) if len(s) > 1]
    random.seed(4); np.random.seed(4)
    by_item, count = P.group_by_label(data, MAXLEN, 32, 25)
    quota = P.exemplar_quota(count, 60, equal=False)
    ref_rand = [P.random_picks(len(v), quota[k - 1]).tolist() if quota[k - 1] > 0 else None for k, v in by_item.items()]
    state = (random.random(), np.random.rand())
    random.seed(4); np.random.seed(4)
    g = D.ExemplarGenerator(data, 60, False, 32, MAXLEN, 0.0, 25)
    assert g.item_count.tolist() == quota.tolist()
    assert g.items.tolist() == [int(k) for k in by_item.keys()]
    mine = g.sess_by_item
    for k, v in by_item.items():
        assert np.array_equal(np.array(v), np.array(mine[k]))
    # random selection consumes np.random per item in the same order
    got = []
    for gi, it in enumerate(g.items.tolist()):
        n = int(g.seg_off[gi + 1] - g.seg_off[gi]); m = g.item_count[it - 1]
        got.append(np.random.choice(n, min(m, n), replace=False).tolist() if m > 0 else None)
    assert got == ref_rand
    assert state == (random.random(), np.random.rand())


def test_pack_rows_label_generator():
    rows = [[5], [1, 2], list(range(1, 60)), []]
    ids, label, n_in = D.pack_rows(rows, 50)
    assert n_in.tolist() == [0, 1, 50, 0] and label.tolist() == [5, 2, 59, 0]
    for r in (1, 2):
        seq, pos = P.row_to_input(rows[r], 50)
        assert np.array_equal(ids[r], seq) and label[r] == pos


def test_dataloader_binary_cache_matches_text_parse(golden_dir, tmp_path):
    """the cached (session, item) pair file gives the same sessions / infos as parsing the text, is reused on
    the second construction, and is invalidated when the source file changes."""
    import shutil
    src = os.path.join(golden_dir, "tiny_data")
    data = tmp_path / "tiny_data"
    shutil.copytree(src, data)
    cache = tmp_path / "cache"

    def load(cache_dir):
        dl = D.DataLoader(str(data), cache_dir=cache_dir)
        tr, i1 = dl.train_loader(0)
        te, i2 = dl.evaluate_loader(1)
        return tr, i1, te, i2, dl.max_item()

    plain = load(None)
    first = load(str(cache))
    files = sorted(os.listdir(cache))
    assert len(files) == 2 and all(f.endswith(".pairs.npy") for f in files)
    stamp = {f: os.stat(cache / f).st_mtime_ns for f in files}
    second = load(str(cache))
    assert plain == first == second
    assert {f: os.stat(cache / f).st_mtime_ns for f in files} == stamp          # reused, not rewritten
    # reference semantics of the grouping: sessions by first appearance, items in file order
    by = {}
    for line in open(data / "period_0.txt"):
        s, i = line.split()
        by.setdefault(int(s), []).append(int(i))
    assert plain[0] == list(by.values())
    # a changed source invalidates the cache entry
    with open(data / "period_0.txt", "a") as f:
        f.write("999999 1\n999999 2\n")
    third = load(str(cache))
    assert third[0][:-1] == plain[0] and third[0][-1] == [1, 2]
