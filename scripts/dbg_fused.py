"""Diagnostic: rel-L2 error of every forward slot and backward scratch buffer, fused (bf16) vs exact encoder."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_gpu_parity import _ids, _model
from test_gpu_encoder_fused import _slots
from ader_b200 import ops

rng = np.random.RandomState(3)
ids = _ids(rng, 45, 50, 380)
pos = rng.randint(1, 381, 45).astype(np.int32)
ntok = int((ids != 0).sum())
res = {}
for impl in ("exact", "tc"):
    m, hp, params = _model(400, encoder_impl=impl)
    m.loss_and_grad(ids, pos, 380, n_tokens=ntok)
    torch.cuda.synchronize()
    fw = [_slots(m, 45, ntok, b)[1] for b in range(2)]
    d = 150
    step = (ntok * d * 4 + 255) // 256 * 256
    bw = [m._bwd_ws.buf[i * step:i * step + ntok * d * 4].view(torch.float32).clone() for i in range(10)]
    res[impl] = (fw, bw, m.grad.clone())
names = {0: "x", 1: "q1", 2: "Q", 3: "K", 4: "V", 5: "y", 6: "z", 7: "h", 9: "probs"}
rl2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for b in range(2):
    for s, nm in names.items():
        print("fwd block %d %-6s rel-L2 %.3e" % (b, nm, rl2(res["tc"][0][b][s], res["exact"][0][b][s])))
bn = ["gX(in of b0 = out)", "gXin", "gO", "gH", "gZ", "gY", "gQ", "gK", "gV", "gQ1"]
for i in range(10):
    print("bwd(b0) %-20s rel-L2 %.3e  norm %.3e" % (bn[i], rl2(res["tc"][1][i], res["exact"][1][i]), float(res["exact"][1][i].norm())))
