"""Debug: where do the emulated-DP parameters leave the single-replica ones (tests/test_gpu_dp.py, tc path)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_gpu_dp import _batches, _model, _warm
from ader_b200.dist import local_peer_group, shard_rows
world, steps = 2, 3
batches, V = _batches(steps)
ref = _model("tc")
models = [_model("tc") for _ in range(world)]
comms = local_peer_group(models)
streams = [torch.cuda.Stream() for _ in range(world)]
torch.cuda.synchronize()
for r, m in enumerate(models):
    ids, pos, teacher = batches[0]
    (tl, th), (el, eh) = shard_rows(len(pos), len(ids) - len(pos), r, world)
    rows = list(range(tl, th)) + list(range(len(pos) + el, len(pos) + eh))
    m.global_counts = (len(pos), len(ids) - len(pos))
    _warm(m, streams[r], ids[rows], pos[tl:th], V, exemplar_logits=teacher[el:eh])
d = 150
for s, (ids, pos, teacher) in enumerate(batches):
    n_train, n_ex = len(pos), len(ids) - len(pos)
    ref.train_step(ids, pos, V, 5e-4, 0.0, exemplar_logits=teacher)
    for r, m in enumerate(models):
        (tl, th), (el, eh) = shard_rows(n_train, n_ex, r, world)
        rows = list(range(tl, th)) + list(range(n_train + el, n_train + eh))
        m.global_counts = (n_train, n_ex)
        with torch.cuda.stream(streams[r]):
            m.train_step(ids[rows], pos[tl:th], V, 5e-4, 0.0, exemplar_logits=teacher[el:eh])
    torch.cuda.synchronize()
    g_ref = ref.grad.cpu().numpy(); g_dp = sum(m.grad.cpu().numpy() for m in models)
    dth = np.abs(models[0].theta.cpu().numpy() - ref.theta.cpu().numpy())
    i = int(dth.argmax())
    print("step %d: max |dtheta| %.3e at flat %d (row %d col %d); grad ref %.4e dp %.4e | max |dgrad| %.3e at %d (ref %.4e dp %.4e), max|g| %.3e" % (
        s, dth.max(), i, i // d, i % d, g_ref[i], g_dp[i], np.abs(g_ref - g_dp).max(), int(np.abs(g_ref - g_dp).argmax()),
        g_ref[int(np.abs(g_ref - g_dp).argmax())], g_dp[int(np.abs(g_ref - g_dp).argmax())], np.abs(g_ref).max()))
    row = i // d
    print("   row %d: in ids? %s  is label? %s  |g_ref row| %.3e  |g_dp row| %.3e  rel diff of row %.3e" % (
        row, bool((ids == row).any()), bool((pos == row).any()), np.abs(g_ref[row * d:(row + 1) * d]).max(), np.abs(g_dp[row * d:(row + 1) * d]).max(),
        np.abs(g_ref[row * d:(row + 1) * d] - g_dp[row * d:(row + 1) * d]).max() / max(np.abs(g_ref[row * d:(row + 1) * d]).max(), 1e-30)))
