"""Mint golden fixtures from the reference's OWN host code (util.py, unmodified).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes small fixtures next to this file:
  tiny_data/period_{0..3}.txt  synthetic sessions in the reference's "<sess> <item>" format
  protocol.json                DataLoader / Sampler / split / ExemplarGenerator grouping +
                               multinomial quota / Evaluator.results outputs, seed 0
  herding.npz                  reps + picks of ExemplarGenerator.herding for 40 cases

The reference has no tests or golden vectors of its own; these are outputs
of the reference code itself, which is the strongest pin available for the host protocol.
"""
import json
import os
import random
import sys
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refstub  # noqa: E402

MAXLEN = 50


def write_tiny_data(root):
    """4 period files; ids assigned in first-appearance order like data/preprocessing.py."""
    rng = np.random.RandomState(1234)
    os.makedirs(root, exist_ok=True)
    next_item, sess_id = 1, 0
    for p in range(4):
        lines = []
        for _ in range(260):
            sess_id += 1
            n = int(min(1 + rng.geometric(0.28), 14))
            if rng.rand() < 0.01:
                n = 60                                   # exercises maxlen truncation
            if rng.rand() < 0.05:
                n = 1                                    # length-1 sessions are skipped at batch time
            for _ in range(n):
                if rng.rand() < 0.12 or next_item < 20:
                    item = next_item
                    next_item += 1
                else:
                    item = int(rng.zipf(1.3)) % (next_item - 1) + 1
                lines.append("%d %d\n" % (sess_id, item))
        with open(os.path.join(root, "period_%d.txt" % p), "w") as f:
            f.writelines(lines)


def main():
    util = refstub.load_reference_util()
    data_root = os.path.join(HERE, "tiny_data")
    write_tiny_data(data_root)

    # util.DataLoader resolves ../../data/<dataset> relative to cwd (util.py:28)
    import tempfile
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "data"))
    os.symlink(data_root, os.path.join(tmp, "data", "TINY"))
    os.makedirs(os.path.join(tmp, "results", "x"))
    cwd = os.getcwd()
    os.chdir(os.path.join(tmp, "results", "x"))
    out = {}
    try:
        random.seed(0)
        np.random.seed(0)
        dl = util.DataLoader("TINY")
        train_sess, info = dl.train_loader(0)
        out["train_info"] = info
        out["n_train_sessions"] = len(train_sess)
        sampler = util.Sampler(train_sess, MAXLEN, 16)
        out["n_rows_before_split"] = sampler.data_size()
        valid, train = sampler.split_data(valid_portion=0.1, return_train=True)
        out["valid_rows"] = valid
        out["train_rows_head"] = train[:20]
        out["n_train_rows"] = len(train)
        out["batch_num"] = sampler.batch_num()
        batches = []
        for _ in range(sampler.batch_num() + 2):         # crosses the wrap + reshuffle
            seq, pos = sampler.sampler()
            batches.append({"seq": np.array(seq).tolist(), "pos": np.array(pos).tolist()})
        out["batches"] = batches
        test_sess, info = dl.evaluate_loader(1)
        out["test_info"] = info
        out["test_sessions"] = test_sess
        out["max_item"] = dl.max_item()

        # test Evaluator's sampler (prefix augmentation, is_subseq=False) -- first 2 batches
        ev = util.Sampler(test_sess, MAXLEN, 8, is_subseq=False)
        out["test_rows_total"] = ev.data_size()
        eb = []
        for _ in range(2):
            seq, pos = ev.sampler()
            eb.append({"seq": np.array(seq).tolist(), "pos": np.array(pos).tolist()})
        out["test_batches"] = eb

        # ExemplarGenerator grouping + multinomial quota (util.py:366-399)
        cand = list(train) + list(valid)
        gen = util.ExemplarGenerator(cand, 120, False, 16, MAXLEN, 0.0, dl.max_item())
        out["quota"] = gen.item_count.tolist()
        out["label_order"] = [int(k) for k in gen.sess_by_item.keys()]
        out["group_sizes"] = [len(v) for v in gen.sess_by_item.values()]
        first = next(iter(gen.sess_by_item.values()))
        out["first_group"] = np.array(first).tolist()
        gen_eq = util.ExemplarGenerator(cand, 120, True, 16, MAXLEN, 0.0, dl.max_item())
        out["quota_equal"] = gen_eq.item_count.tolist()
        # random selection picks per item (util.py:506-512), consumes np.random in label order
        rnd = []
        for item, seqs in gen_eq.sess_by_item.items():
            m = gen_eq.item_count[item - 1]
            if m > 0:
                rnd.append(np.random.choice(len(seqs), min(m, len(seqs)), replace=False).tolist())
            else:
                rnd.append([])
        out["random_picks"] = rnd

        # exemplar sampler (add_exemplar / exemplar_sampler, util.py:173-186, :241-263)
        exemplars = [[r, [float(len(r)), 0.5]] for r in train[:37]]
        ex = util.Sampler([], MAXLEN, 5)
        ex.add_exemplar(exemplars)
        exb = []
        for _ in range(ex.batch_num() + 1):
            seq, pos, lg = ex.exemplar_sampler()
            exb.append({"seq": np.array(seq).tolist(), "pos": np.array(pos).tolist(),
                        "logits": [list(x) for x in lg]})
        out["exemplar_batches"] = exb

        # Evaluator.results (util.py:329-339)
        evalr = util.Evaluator.__new__(util.Evaluator)
        evalr.ranks = [0, 3, 19, 20, 9, 10, 500, 1, 7, 25, 0, 11]
        out["ranks"] = evalr.ranks
        out["metrics"] = list(evalr.results())
        out["display"] = (evalr.__setattr__("mode", "valid") or evalr.display(3))
    finally:
        os.chdir(cwd)

    with open(os.path.join(HERE, "protocol.json"), "w") as f:
        json.dump(out, f)

    # herding KATs from the reference's own ExemplarGenerator.herding (util.py:401-434)
    gen = util.ExemplarGenerator.__new__(util.ExemplarGenerator)
    gen.exemplars = defaultdict(list)
    rng = np.random.RandomState(7)
    reps, picks, ms, offs = [], [], [], [0]
    cases = [(1, 1), (2, 1), (2, 2), (3, 2), (5, 5), (8, 3), (13, 13), (20, 7), (33, 20), (64, 10),
             (98, 40), (98, 98), (150, 30), (350, 100)]
    cases += [(int(rng.randint(2, 120)), 0) for _ in range(26)]
    for ci, (n, m) in enumerate(cases):
        if m == 0:
            m = int(rng.randint(1, n + 1))
        d = 150
        center = rng.randn(d).astype(np.float32)
        rep = (center[None] + 0.6 * rng.randn(n, d)).astype(np.float32)
        if ci % 5 == 4 and n > 3:
            rep[n // 2] = rep[0]                         # exact duplicate -> first-max tie-break
        seq = np.concatenate([rng.randint(1, 9, (n, MAXLEN)), np.full((n, 1), 3)], 1)
        seq[:, :MAXLEN - 3] = 0
        for i in range(n):
            seq[i, -2] = i + 1                           # identify the candidate from its stored session
        logits = np.zeros((n, 1), np.float32)
        saved = gen.herding(rep, logits, seq, ci, m)
        chosen = [int(e[0][-2]) - 1 for e in gen.exemplars[ci]]
        assert saved == len(chosen)
        reps.append(rep)
        picks.append(np.array(chosen + [-1] * (m - len(chosen)), np.int32))
        ms.append(m)
        offs.append(offs[-1] + n)
    np.savez_compressed(os.path.join(HERE, "herding.npz"),
                        rep=np.concatenate(reps, 0), seg_off=np.array(offs, np.int64),
                        m=np.array(ms, np.int32),
                        picks=np.concatenate(picks), pick_off=np.cumsum([0] + ms).astype(np.int64))
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
