"""1-GPU probe: the peer-memory optimiser kernel with world = 1 (no link traffic) against the plain Adam kernel, grid sweep."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for g in ("0", "148", "296", "592", "1184", "2368"):
        env = dict(os.environ); 
        if g != "0": env["ADER_B200_DP_GRID"] = g
        subprocess.run([sys.executable, __file__, g], env=env)
    sys.exit(0)
import torch
import bench
from ader_b200 import ops
from ader_b200.dist import local_peer_group
from ader_b200.model import Ader
WL = bench.WL
m = Ader(WL["item_num"], bench.make_args(), init_seed=0)
local_peer_group([m])
m.grad.normal_()
V = WL["V"]
def t(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def dp():
    m.dp.begin_step(m); m.dp.apply(m, V, 5e-4)
print("grid %s: dp step world=1 %.1f us | plain adam %.1f us" % (sys.argv[1], t(dp), t(lambda: ops.adam_step(m.ms, m.theta, m.adam_m, m.adam_v, m.grad, m.adam_state, V, 5e-4))))
