"""Flat parameter layout of the SASRec network (reference: ADER.py:25-96, EWC.py:90).

All trainable state is ONE flat fp32 vector in the reference's variable-creation order; the 32
reference tensors are views into it.  This file restates the layout independently of the C
library (``ader_param_offset``) -- tests check that both agree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass(frozen=True)
class Hyper:
    """Network-shaping flags of main.py:75-108."""
    item_num: int             # main.py:133-138; the item table has item_num + 1 rows
    hidden_units: int = 150   # main.py:103
    maxlen: int = 50          # main.py:104
    num_blocks: int = 2       # main.py:99
    num_heads: int = 1        # main.py:100

    @property
    def v_tab(self) -> int:
        return self.item_num + 1


_BLOCK = ["ln1.beta", "ln1.gamma", "wq", "bq", "wk", "bk", "wv", "bv",
          "ln2.beta", "ln2.gamma", "w1", "b1", "w2", "b2"]


class ParamLayout:
    def __init__(self, hp: Hyper):
        self.hp = hp
        d = hp.hidden_units
        self.entries: List[Tuple[str, Tuple[int, ...], int]] = []
        off = 0

        def add(name, shape):
            nonlocal off
            self.entries.append((name, shape, off))
            off += int(np.prod(shape))

        add("item_table", (hp.v_tab, d))                       # modules.py:118-122
        add("pos_table", (hp.maxlen, d))                       # ADER.py:41-51
        for b in range(hp.num_blocks):
            for n in _BLOCK:
                if n.startswith("w"):
                    add("b%d.%s" % (b, n), (d, d))             # modules.py:172-174, 254-261
                else:
                    add("b%d.%s" % (b, n), (d,))
        add("lnf.beta", (d,))                                  # ADER.py:82
        add("lnf.gamma", (d,))
        self.total = off
        self.dense_offset = hp.v_tab * d
        self.dense_count = self.total - self.dense_offset

    def names(self) -> List[str]:
        return [e[0] for e in self.entries]

    def offset(self, idx: int) -> int:
        return self.entries[idx][2]

    def views(self, flat):
        """Split a flat tensor/array into the reference's per-variable tensors (views)."""
        out = []
        for _, shape, off in self.entries:
            n = int(np.prod(shape))
            out.append(flat[off:off + n].reshape(shape))
        return out

    def init_flat(self, seed: int = 0) -> np.ndarray:
        """Seeded Glorot-uniform init (TF get_variable / dense / conv1d default), zeros for biases
        and LN beta, ones for LN gamma.  TF's init stream cannot be reproduced outside
        TF; parity runs share this init (same draw order as the test oracle's)."""
        rng = np.random.RandomState(seed)
        flat = np.zeros(self.total, np.float32)
        for name, shape, off in self.entries:
            n = int(np.prod(shape))
            if name.endswith("gamma"):
                flat[off:off + n] = 1.0
            elif len(shape) == 1:
                pass
            else:
                limit = math.sqrt(6.0 / (shape[0] + shape[1]))
                flat[off:off + n] = rng.uniform(-limit, limit, size=shape).astype(np.float32).ravel()
        return flat
