"""Data-parallel step on 2 GPUs (both back ends: peer memory over NVLink, NCCL all-reduce) == the same step on one
GPU with the full batch (needs >= 2 GPUs: run with `gpurun --gpus 2`; skipped on a single-GPU box -- there
tests/test_gpu_dp.py runs the same comparison with emulated ranks)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _args():
    return type("Args", (), dict(hidden_units=150, maxlen=50, num_blocks=2, num_heads=1, random_seed=0, lr=5e-4,
                                 dropout_rate=0.0, disable_distillation=False, loss_impl="exact"))()


def _batch():
    rng = np.random.RandomState(3)
    M, Bt, V, Vp = 37, 25, 380, 300
    ids = np.zeros((M, 50), np.int32)
    for r in range(M):
        n = int(rng.randint(1, 20)); ids[r, 50 - n:] = rng.randint(1, V + 1, n)
    pos = rng.randint(1, V + 1, Bt).astype(np.int32)
    teacher = (rng.randn(M - Bt, Vp) * 2).astype(np.float32)
    return ids, pos, teacher, V


def _worker(rank, world, port, out, backend):
    import torch.distributed as dist
    from ader_b200.dist import DataParallel
    from ader_b200.model import Ader
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        m = Ader(400, _args(), device=torch.device("cuda", rank), init_seed=0)
        m.theta.add_(torch.randn(m.theta.shape, generator=torch.Generator().manual_seed(1)).to(m.device) * 0.05)
        m.update_loss(0.7)
        ids, pos, teacher, V = _batch()
        dp = DataParallel(m, backend=backend)
        if backend == "nvls" and m.dp.kind != "nvls":
            out[rank] = "no multicast"                   # fabric without NVLS: nothing to test
            return
        assert m.dp.kind == backend, "back end %s unavailable (fell back to %s)" % (backend, m.dp.kind)
        loss = dp.train_step(ids, pos, V, 5e-4, 0.0, exemplar_logits=teacher)
        l2 = dp.train_step(ids, pos, V, 5e-4, 0.0, exemplar_logits=teacher)
        torch.cuda.synchronize()
        m.dp.check()
        dist.barrier()
        out[rank] = (float(loss.item()), m.theta.cpu().numpy(), float(l2.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("backend", ["p2p", "nvls", "nccl"])
def test_data_parallel_step_matches_single_gpu(backend):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ader_b200.model import Ader
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out, backend), nprocs=2, join=True)
    if out[0] == "no multicast":
        pytest.skip("no NVLS multicast support on this box")
    m = Ader(400, _args(), init_seed=0)
    m.theta.add_(torch.randn(m.theta.shape, generator=torch.Generator().manual_seed(1)).to(m.device) * 0.05)
    m.update_loss(0.7)
    ids, pos, teacher, V = _batch()
    loss = float(m.train_step(ids, pos, V, 5e-4, 0.0, exemplar_logits=teacher).item())
    loss2 = float(m.train_step(ids, pos, V, 5e-4, 0.0, exemplar_logits=teacher).item())
    ref = m.theta.cpu().numpy()
    for r in range(2):
        assert out[r][0] == pytest.approx(loss, rel=1e-5)
        assert out[r][2] == pytest.approx(loss2, rel=1e-5)
        # one Adam step moves weights by ~lr; gradients agree to ~1e-6 relative
        assert np.abs(out[r][1] - ref).max() < 3e-5
    assert np.array_equal(out[0][1], out[1][1])          # replicas stay bit-identical


def _graph_worker(rank, world, port, out, backend="p2p"):
    import torch.distributed as dist
    from ader_b200.dist import DataParallel, shard_rows
    from ader_b200.model import Ader
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        args = _args(); args.loss_impl = "tc"
        res = {}
        ids, pos, teacher, V = _batch()
        n_train, n_ex = len(pos), len(ids) - len(pos)
        (tl, th), (el, eh) = shard_rows(n_train, n_ex, rank, world)
        rows = list(range(tl, th)) + list(range(n_train + el, n_train + eh))
        for mode in ("eager", "graph"):
            m = Ader(400, args, device=torch.device("cuda", rank), init_seed=0)
            m.theta.add_(torch.randn(m.theta.shape, generator=torch.Generator().manual_seed(1)).to(m.device) * 0.05)
            m.update_loss(0.7)
            DataParallel(m, backend=backend)
            if backend == "nvls" and m.dp.kind != "nvls":
                out[rank] = "no multicast"
                return
            assert m.dp.kind == backend
            m.global_counts = (n_train, n_ex)
            t_dev = torch.from_numpy(teacher).to(m.device)
            trows = np.arange(el, eh, dtype=np.int32)
            gs = m.graph_step(th - tl, eh - el, V, 5e-4, 0.0, teacher=t_dev) if mode == "graph" else None
            for _ in range(3):
                if gs is None:
                    m.train_step(ids[rows], pos[tl:th], V, 5e-4, 0.0, exemplar_logits=t_dev, teacher_rows=trows)
                else:
                    gs.run_rows(ids[rows], pos[tl:th], trows)
            torch.cuda.synchronize()
            m.dp.check()
            dist.barrier()
            res[mode] = m.theta.cpu().numpy()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("backend", ["p2p", "nvls"])
def test_dp_graph_replay_equals_eager_step_across_gpus(backend):
    """The peer-memory data-parallel step captured as a CUDA graph (what bench.py and the period loop replay) gives the
    same bits as the eager step, on every rank (peer loads / stores, and the in-switch NVLS form)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_graph_worker, args=(2, _free_port(), out, backend), nprocs=2, join=True)
    if out[0] == "no multicast":
        pytest.skip("no NVLS multicast support on this box")
    for r in range(2):
        assert np.array_equal(out[r]["eager"], out[r]["graph"])
    assert np.array_equal(out[0]["graph"], out[1]["graph"])
