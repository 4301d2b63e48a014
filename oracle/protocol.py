"""NumPy / pure-Python restatement of the reference's host protocol (TEST INFRASTRUCTURE).

Follows util.py and main.py of doublemul/ADER (paths relative to /root/reference).  It
consumes Python's ``random`` and ``numpy.random`` global streams in the same order as the
reference, so with equal seeds it yields the same batches, valid split,
quotas and picks.  Pinned against fixtures minted from the reference's own ``util.py``
(tests/golden/make_golden.py -> tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
import os
import random
from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import numpy as np


# ---- util.py:17-107 DataLoader ----------------------------------------------------------
class PeriodFiles:
    def __init__(self, data_dir: str):
        self.data_dir = data_dir
        self.seen = set()                     # util.py:26 item_set

    def _read(self, period: int):
        with open(os.path.join(self.data_dir, "period_%d.txt" % period)) as f:
            for line in f:
                s, i = line.rstrip().split(" ")
                yield int(s), int(i)

    def train(self, period: int) -> Tuple[List[List[int]], int]:
        """util.py:32-58: sessions in first-appearance order; every item becomes 'seen'."""
        by_sess: Dict[int, List[int]] = {}
        for s, i in self._read(period):
            self.seen.add(i)
            by_sess.setdefault(s, []).append(i)
        sessions = list(by_sess.values())
        return sessions, sum(len(x) for x in sessions)

    def evaluate(self, period: int) -> Tuple[List[List[int]], int, int]:
        """util.py:60-102: drop unseen items, then sessions left with a single item."""
        by_sess: Dict[int, List[int]] = {}
        total = removed = 0
        for s, i in self._read(period):
            total += 1
            if i not in self.seen:
                removed += 1
                continue
            by_sess.setdefault(s, []).append(i)
        kept = []
        for s, items in by_sess.items():
            if len(items) == 1:
                removed += 1
            else:
                kept.append(items)
        return kept, total, removed

    def max_item(self) -> int:                # util.py:104-107
        return max(self.seen)


# ---- util.py:110-273 Sampler -------------------------------------------------------------
def prefix_rows(data: Sequence[Sequence[int]], is_subseq: bool) -> List[List[int]]:
    """util.py:136-145: a session [i1..ik] yields itself, then session[:-1], ... down to
    length 2 (only when k > 2); length-1/2 sessions yield just themselves."""
    if is_subseq:
        return [list(s) for s in data]
    rows = []
    for s in data:
        rows.append(list(s))
        for cut in range(1, len(s) - 1):
            rows.append(list(s[:-cut]))
    return rows


def row_to_input(session: Sequence[int], maxlen: int) -> Tuple[np.ndarray, int]:
    """util.py:151-171: label = last item; input = last <=maxlen items before it,
    right-aligned, zero left-padded."""
    seq = np.zeros(maxlen, np.int32)
    body = list(session[:-1])[-maxlen:]
    if body:
        seq[maxlen - len(body):] = body
    return seq, int(session[-1])


class RefSampler:
    def __init__(self, data, maxlen: int, batch_size: int, is_subseq: bool = False):
        self.maxlen, self.batch_size = maxlen, batch_size
        self.rows = prefix_rows(data, is_subseq)
        self.logits: list = []
        self.cursor = 0
        self._reshuffle()

    def _reshuffle(self):
        self.order = list(range(len(self.rows)))
        random.shuffle(self.order)                                   # util.py:148-149

    def batch_num(self) -> int:                                      # util.py:270-273
        return math.ceil(len(self.rows) * 1.0 / self.batch_size)

    def data_size(self) -> int:
        return len(self.rows)

    def add_exemplar(self, exemplar):                                # util.py:173-186
        self.logits = []
        for session, logits in exemplar:
            self.rows.append(session)
            self.logits.append(logits)
        self._reshuffle()

    def split_data(self, valid_portion: float):                      # util.py:188-216
        n = len(self.rows)
        sidx = np.arange(n, dtype="int32")
        np.random.shuffle(sidx)                                      # util.py:203
        n_train = int(np.round(n * (1.0 - valid_portion)))
        valid = [self.rows[s] for s in sidx[n_train:]]
        train = [self.rows[s] for s in sidx[:n_train]]
        self.rows = train
        self._reshuffle()                                            # util.py:210-211
        return valid, train

    def _take(self):
        lo = self.cursor * self.batch_size
        idx = [self.order[i] for i in range(lo, min(lo + self.batch_size, len(self.rows)))]
        idx = [i for i in idx if len(self.rows[i]) > 1]              # util.py:228-229
        self.cursor += 1
        if self.cursor == self.batch_num():                          # util.py:234-237
            self.cursor = 0
            random.shuffle(self.order)
        return idx

    def sampler(self):                                               # util.py:218-239
        idx = self._take()
        pairs = [row_to_input(self.rows[i], self.maxlen) for i in idx]
        seqs = [p[0] for p in pairs]
        pos = [p[1] for p in pairs]
        return seqs, pos

    def exemplar_sampler(self):                                      # util.py:241-263
        idx = self._take()
        pairs = [row_to_input(self.rows[i], self.maxlen) for i in idx]
        return [p[0] for p in pairs], [p[1] for p in pairs], [self.logits[i] for i in idx]


# ---- util.py:366-434 ExemplarGenerator ---------------------------------------------------
def group_by_label(data, maxlen: int, batch_size: int, max_item: int):
    """util.py:382-393: iterate the shuffled sampler once; per label keep [seq(maxlen), label]
    in encounter order, count per label."""
    s = RefSampler(data, maxlen, batch_size, is_subseq=True)
    by_item: Dict[int, List[np.ndarray]] = defaultdict(list)
    count = np.zeros(max_item)
    for _ in range(s.batch_num()):
        seqs, pos = s.sampler()
        for q, item in zip(seqs, pos):
            by_item[int(item)].append(np.append(q, item))
            count[item - 1] += 1
    return by_item, count


def exemplar_quota(count: np.ndarray, m: int, equal: bool) -> np.ndarray:
    """util.py:395-399."""
    if equal:
        count = np.ones_like(count)
    prob = count / count.sum()
    return np.int32(np.random.multinomial(n=m, pvals=prob, size=1)[0])


def herding_picks(rep: np.ndarray, m: int) -> List[int]:
    """util.py:419-432 in float32 NumPy: D = rep^T / ||rep^T||_2 (per candidate),
    mu = mean over candidates, w = mu; at most ceil(1.1 m) steps of
    i = argmax(w.D) (first max), w += mu - D[:, i]; keep first occurrences."""
    D = rep.T / np.linalg.norm(rep.T, axis=0)
    mu = D.mean(axis=1)
    w = mu
    picks: List[int] = []
    step = 0
    while len(picks) != m and step < 1.1 * m:
        i = int(np.argmax(np.dot(w, D)))
        w = w + mu - D[:, i]
        step += 1
        if i not in picks:
            picks.append(i)
    return picks


def loss_picks(n: int, m: int) -> List[int]:
    """util.py:482-489 as executed: ``model.loss`` is a scalar mean, so
    ``np.array(scalar).argsort()[:k]`` is ``[0]`` for any k >= 1."""
    return [0] if min(m, n) >= 1 else []


def random_picks(n: int, m: int) -> np.ndarray:
    """util.py:512."""
    return np.random.choice(n, min(m, n), replace=False)


def stored_session(seq_with_label: np.ndarray) -> List[int]:
    """util.py:433 -- non-zero entries of [input(maxlen), label]."""
    return seq_with_label[seq_with_label != 0].tolist()


# ---- main.py --------------------------------------------------------------------------
def flatten_exemplars(exemplar_pre: dict) -> list:
    """main.py:54-65."""
    out = []
    for item in exemplar_pre.values():
        if isinstance(item, list):
            out.extend([i for i in item if i])
    return out


def adaptive_lambda(lambda0: float, item_prev: int, item_cur: int, n_exemplar: int, n_train: int) -> float:
    """main.py:199-200."""
    return lambda0 * math.sqrt((item_prev / item_cur) * (n_exemplar / n_train))


def exemplar_rows_per_step(n_exemplar: int, batch_num: int) -> int:
    """main.py:186-187."""
    return int(n_exemplar / batch_num)
