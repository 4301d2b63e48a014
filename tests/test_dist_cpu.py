"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, vocab-parallel LSE merge and
the data-parallel gradient identity (global-mean denominators + all-reduce SUM == full-batch gradient),
with the oracle standing in for the device step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ader_b200 import dist as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fn, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def test_shard_ranges_partition():
    for n in (0, 1, 7, 512, 650, 693):
        for w in (1, 2, 3, 8):
            parts = [D.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    lo, hi = D.vocab_shard(1000000, 3, 8)
    assert lo % 128 == 0 and 0 < hi - lo <= 125056
    assert D.vocab_shard(1000000, 7, 8)[1] == 1000000


def _lse_job(rank, world):
    torch.manual_seed(0)
    x = torch.randn(37, 1001, dtype=torch.float64) * 3
    x[5] = -float("inf"); x[5, 700] = 1.0                       # a row whose mass sits in one shard
    lo, hi = D.vocab_shard(1001, rank, world, tile=128)
    xs = x[:, lo:hi]
    mx = xs.max(dim=1).values
    se = torch.where(torch.isfinite(mx), torch.exp(xs - mx[:, None]).sum(1), torch.zeros_like(mx))
    got = D.merge_lse(mx, se)
    return float((got - torch.logsumexp(x, 1)).abs().max())


def test_vocab_parallel_lse_merge_gloo():
    assert max(_run(_lse_job)) < 1e-12


def _dp_job(rank, world):
    from oracle import sasrec as S
    hp = S.Hyper(item_num=60, hidden_units=12, maxlen=10, num_blocks=1, num_heads=1)
    params = S.randomize_params(S.init_params(hp, 0, torch.float64), 1, 0.3)
    rng = np.random.RandomState(0)
    n_train, n_ex, V, Vp = 7, 4, 50, 40
    ids = np.zeros((n_train + n_ex, 10), np.int64)
    for r in range(len(ids)):
        n = rng.randint(1, 11); ids[r, 10 - n:] = rng.randint(1, V + 1, n)
    pos = rng.randint(1, V + 1, n_train)
    teacher = torch.tensor(rng.randn(n_ex, Vp))
    full = lambda ps: S.loss_ader(ps, torch.tensor(ids), torch.tensor(pos), V, hp, 0.7, exemplar_logits=teacher)
    loss_full, g_full = S.grads_of(full, params)
    (tl, th), (el, eh) = D.shard_rows(n_train, n_ex, rank, world)
    rows = list(range(tl, th)) + list(range(n_train + el, n_train + eh))

    def local(ps):      # local rows, GLOBAL denominators (AderLossArgs.n_train_global / n_ex_global)
        rep = S.forward_rep(ps, torch.tensor(ids[rows]), hp)
        lg = S.logits_of(rep, ps[0], V)
        ce = S.ce_rows(lg[:th - tl], torch.tensor(pos[tl:th])).sum() / n_train
        s = lg[th - tl:, :Vp]
        t = torch.softmax(teacher[el:eh], 1)
        kd = -(t * torch.log_softmax(s, 1)).sum() / n_ex
        return ce + 0.7 * kd
    loss_l, g_l = S.grads_of(local, params)
    flat = torch.cat([g.reshape(-1) for g in g_l])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    lt = torch.tensor([loss_l], dtype=torch.float64)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    want = torch.cat([g.reshape(-1) for g in g_full])
    return float((flat - want).abs().max()), abs(float(lt) - loss_full)


def test_data_parallel_gradient_identity_gloo():
    for gerr, lerr in _run(_dp_job):
        assert gerr < 1e-12 and lerr < 1e-12
