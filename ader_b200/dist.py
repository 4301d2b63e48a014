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


class VocabParallelLoss:
    """Vocab-parallel logits + softmax CE + distillation (BASELINE config 5): every rank holds all `rep`
    rows and the table rows of its vocabulary shard [v_lo, v_hi); the tcgen05 kernels run on the shard and
    two tiny collectives stitch the softmax together:
        all-reduce(max) of the per-row maxima, all-reduce(sum) of (rescaled sumexp, label logit, KD dot),
    and one all-reduce(sum) of the partial d_rep [M, d] in backward.  dE needs no communication (each rank
    owns its rows).  With `group=None` and world size 1 it degenerates to the single-GPU kernel."""

    def __init__(self, model, group=None, rank=None, world=None):
        from . import ops
        self.ops, self.model, self.group = ops, model, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.ws = ops.Workspace(model.device)

    def shard(self, V: int):
        return vocab_shard(V, self.rank, self.world)

    def local_forward(self, rep, a, v_lo, v_hi):
        ops, m = self.ops, self.model
        ws = self.ws.get(ops.loss_tc_vp_ws_bytes(m.ms, a, v_lo, v_hi))
        stats = torch.empty((a.M, 4), dtype=torch.float32, device=m.device)
        ops.loss_tc_vp_fwd(m.ms, m.theta, rep, a, v_lo, v_hi, ws, stats)
        return ws, stats

    @staticmethod
    def merge(stats_list_or_tensor, n_train: int, n_ex: int, lambda_: float, kd: bool, reduce_max=None, reduce_sum=None):
        """(max, sumexp, label, dot) partials -> lse, row_loss, loss.  `reduce_*` are the collectives
        (identity when the partials of all shards are passed as a list)."""
        if isinstance(stats_list_or_tensor, (list, tuple)):
            st = torch.stack(list(stats_list_or_tensor))                       # [S, M, 4]
            gmax = st[..., 0].max(dim=0).values
            scale = torch.where(torch.isfinite(st[..., 0]), torch.exp(st[..., 0] - gmax), torch.zeros_like(st[..., 0]))
            rest = torch.stack([(st[..., 1] * scale).sum(0), st[..., 2].sum(0), st[..., 3].sum(0)], dim=1)
        else:
            st = stats_list_or_tensor
            gmax = st[:, 0].clone()
            reduce_max(gmax)
            scale = torch.where(torch.isfinite(st[:, 0]), torch.exp(st[:, 0] - gmax), torch.zeros_like(gmax))
            rest = torch.stack([st[:, 1] * scale, st[:, 2], st[:, 3]], dim=1).contiguous()
            reduce_sum(rest)
        lse = gmax + torch.log(rest[:, 0])
        row_loss = lse - rest[:, 1]
        if kd and n_ex > 0:
            row_loss[n_train:] = lse[n_train:] - rest[n_train:, 2]
        loss = row_loss[:n_train].mean() if n_train > 0 else row_loss.new_zeros(())
        if n_ex > 0:
            loss = loss + lambda_ * row_loss[n_train:].mean()
        return lse.contiguous(), row_loss, loss

    def fwd_bwd(self, rep, pos, max_item, lambda_=0.0, mode=0, teacher=None, teacher_rows=None, ex_pos=None, grad=None):
        """rep [M, d] (replicated) -> (loss, row_loss, d_rep); writes this rank's rows of `grad` (flat)."""
        ops, m = self.ops, self.model
        M, n_train = rep.shape[0], pos.numel()
        n_ex = M - n_train
        v_prev = teacher.shape[1] if teacher is not None else 0
        a = ops.make_loss_args(M, n_train, n_ex, max_item, v_prev, mode if n_ex > 0 else 0, lambda_, pos, ex_pos, teacher, teacher_rows)
        v_lo, v_hi = self.shard(max_item)
        ws, stats = self.local_forward(rep, a, v_lo, v_hi)
        lse, row_loss, loss = self.merge(stats, n_train, n_ex, lambda_, mode == 1,
                                         lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group),
                                         lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group))
        d_rep = torch.empty_like(rep)
        ops.loss_tc_vp_bwd(m.ms, m.theta, rep, a, v_lo, v_hi, ws, lse, d_rep, m.grad if grad is None else grad)
        dist.all_reduce(d_rep, op=dist.ReduceOp.SUM, group=self.group)
        self._keep = (a, pos, teacher, teacher_rows, ex_pos, stats, lse)
        return loss, row_loss, d_rep
