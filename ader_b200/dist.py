"""Multi-GPU plumbing for the two ways the path shards; torch.distributed is plumbing only.

* Data parallel (configs 1-4): every rank runs the step on its own rows with the GLOBAL mean
  denominators (AderLossArgs.n_train_global / n_ex_global), then ONE all-reduce (sum) of the flat
  gradient; Adam is identical on every rank.  Row order inside a rank stays [train; exemplar].
* Vocab parallel (config 5, 1 M items): the table rows (and Adam state) are sharded by vocabulary
  range; each rank produces per-row (max, sumexp) partials over its columns -- exactly what
  k_tc_logits<FWD> emits per vocabulary chunk -- and the log-sum-exp is merged with an all-reduce(max)
  + all-reduce(sum).  `merge_lse` is that merge; the kernel wiring is the next round's work.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced shard [lo, hi) of n units: the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(n_train: int, n_ex: int, rank: int, world: int):
    """Row shards of one step: train rows and exemplar rows are split SEPARATELY so that every rank
    keeps the [train; exemplar] layout (main.py:229).  Returns ((t_lo, t_hi), (e_lo, e_hi))."""
    return shard_range(n_train, rank, world), shard_range(n_ex, rank, world)


def vocab_shard(V: int, rank: int, world: int, tile: int = 128) -> Tuple[int, int]:
    """Vocabulary columns [lo, hi) of a rank, aligned to the 128-row T128 operand tiles."""
    tiles = (V + tile - 1) // tile
    lo, hi = shard_range(tiles, rank, world)
    return min(lo * tile, V), min(hi * tile, V)


def merge_lse(local_max: torch.Tensor, local_sumexp: torch.Tensor, group=None) -> torch.Tensor:
    """log-sum-exp over vocabulary shards from per-rank (max, sum exp(x - max)) partials."""
    gmax = local_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    scaled = local_sumexp * torch.exp(local_max - gmax)
    scaled = torch.where(torch.isfinite(local_max), scaled, torch.zeros_like(scaled))
    dist.all_reduce(scaled, op=dist.ReduceOp.SUM, group=group)
    return gmax + torch.log(scaled)


class DataParallel:
    """Wrap an `Ader` / `Ewc` model: shard each step's rows over the ranks, sum the flat gradient."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        model.grad_sync = self._sync

    def _sync(self):
        dist.all_reduce(self.model.grad, op=dist.ReduceOp.SUM, group=self.group)

    def train_step(self, seq, pos, max_item, lr=None, dropout_rate=None, exemplar_logits=None, exemplar_pos=None,
                   teacher_rows=None):
        """Same feed as Ader.train_step with the GLOBAL batch on every rank; each rank keeps its shard."""
        n_train = len(pos)
        n_ex = len(seq) - n_train
        (tl, th), (el, eh) = shard_rows(n_train, n_ex, self.rank, self.world)
        rows = list(range(tl, th)) + list(range(n_train + el, n_train + eh))
        if isinstance(seq, torch.Tensor):
            seq_l = seq[torch.as_tensor(rows, device=seq.device)]
        else:
            seq_l = [seq[i] for i in rows]
        kw = {}
        if exemplar_logits is not None:
            if teacher_rows is not None:
                kw["exemplar_logits"], kw["teacher_rows"] = exemplar_logits, teacher_rows[el:eh]
            else:
                kw["exemplar_logits"] = exemplar_logits[el:eh]
        if exemplar_pos is not None:
            kw["exemplar_pos"] = exemplar_pos[el:eh]
        m = self.model
        m.global_counts = (n_train, n_ex)
        loss = m.train_step(seq_l, pos[tl:th], max_item, lr, dropout_rate, **kw)
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss
