"""GPU micro-benchmark: time the loss group (tc vs exact) with CUDA events at the bench shape."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ader_b200 import ops
from ader_b200.model import Ader
args = type("A", (), dict(hidden_units=150, maxlen=50, num_blocks=2, num_heads=1, random_seed=0, lr=5e-4, dropout_rate=0.0, disable_distillation=False))()
shapes = [(650, 512, 18661, 17421, 25958)]
if "--big" in sys.argv:
    shapes.append((4096, 4096, 200000, 0, 200001))
for (M, Bt, V, Vp, item_num) in shapes:
    m = Ader(item_num, args, init_seed=0)
    rep = torch.randn(M, 150, device="cuda")
    pos = torch.randint(1, V + 1, (Bt,), device="cuda", dtype=torch.int32)
    teacher = torch.randn(max(M - Bt, 1), (max(Vp, 1) + 3) // 4 * 4, device="cuda")[:, :max(Vp, 1)] if M > Bt else None
    a = ops.make_loss_args(M, Bt, M - Bt, V, Vp if M > Bt else 0, 1 if M > Bt else 0, 0.8, pos, None, teacher, None)
    loss = torch.zeros(1, device="cuda"); row_loss = torch.zeros(M, device="cuda"); d_rep = torch.zeros(M, 150, device="cuda")
    for impl in ("tc", "exact"):
        if impl == "exact" and V > 50000: continue
        fn, wsb = (ops.loss_fwd_bwd_tc, ops.loss_tc_ws_bytes) if impl == "tc" else (ops.loss_fwd_bwd, ops.loss_ws_bytes)
        ws = torch.empty(wsb(m.ms, a), dtype=torch.uint8, device="cuda")
        for _ in range(3): fn(m.ms, m.theta, rep, a, ws, loss, row_loss, d_rep, m.grad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): fn(m.ms, m.theta, rep, a, ws, loss, row_loss, d_rep, m.grad)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 6.0 * M * 150 * V
        print("shape M=%d V=%d impl=%s: %.1f us per fwd+bwd, %.1f TFLOP/s algorithmic, loss %.4f" % (M, V, impl, ms * 1e3, fl / ms / 1e9, float(loss)))
