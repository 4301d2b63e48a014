"""Debug: per-CTA pipeline timeline of the tcgen05 loss kernels run ALONE on one stream (ader_loss_fwd_bwd_tc at the
bench shape).  Needs a library built with ADER_B200_DEFINES=-DADER_TC_TIMELINE."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ader_b200 import ops, _lib
from ader_b200.model import Ader
WL = bench.WL
dev = torch.device("cuda", 0)
B, Me, V, Vp = WL["B"], WL["M_e"], WL["V"], WL["V_prev"]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    B, Me, V, Vp = 4096, 0, 200000, 0
M = B + Me
model = Ader(max(WL["item_num"], V + 1), bench.make_args(), device=dev, init_seed=0)
rep = torch.randn(M, 150, device=dev) * 0.5
pos = torch.randint(1, V + 1, (B,), device=dev, dtype=torch.int32)
teacher = (torch.randn((max(Me, 1), (max(Vp, 1) + 3) // 4 * 4), device=dev) * 2)[:, :max(Vp, 1)] if Me else None
a = ops.make_loss_args(M, B, Me, V, Vp if Me else 0, 1 if Me else 0, 1.0, pos, None, teacher, None)
ws = torch.empty(ops.loss_tc_ws_bytes(model.ms, a), dtype=torch.uint8, device=dev)
loss = torch.zeros(1, device=dev); rl = torch.zeros(M, device=dev); dr = torch.zeros(M, 150, device=dev)
for _ in range(4):
    ops.loss_fwd_bwd_tc(model.ms, model.theta, rep, a, ws, loss, rl, dr, model.grad)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.debug_loss_tc_kernels(model.ms, model.theta, a, ws, model.grad)
e1.record(); torch.cuda.synchronize()
print("three tc kernels alone: %.1f us per (FWD + DREP + DE)" % (e0.elapsed_time(e1) / 20 * 1e3))
lib = _lib.load()
buf = np.zeros((3, 192, 64), np.int64)
lib.ader_debug_tc_timeline.argtypes = [C.c_void_p]
lib.ader_debug_tc_timeline(buf.ctypes.data_as(C.c_void_p))
for mode, name in enumerate(["FWD", "DREP", "DE"]):
    t = buf[mode]
    used = np.nonzero(t[:, 1])[0]
    if not len(used):
        continue
    g0 = t[used, 0].min()
    dur = (t[used, 41] - t[used, 1]) / 1e3
    print("== %s: %d CTAs; CTA start spread %.1f us; CTA durations (us): min %.1f median %.1f max %.1f; kernel span %.1f us" % (
        name, len(used), (t[used, 0].max() - g0) / 1e3, dur.min(), np.median(dur), dur.max(), (t[used, 41].max() - g0) / 1e3))
    order = used[np.argsort(-dur)]
    for c in [order[0], order[len(order) // 2], order[-1]]:
        r = lambda s: (t[c, s] - t[c, 1]) / 1e3 if t[c, s] else float("nan")
        f = lambda b: " ".join("%5.1f" % r(b + i) for i in range(7))
        print(" CTA %3d start+%.1fus  setup %.1f  Xissue %.1f  end %.1f / %.1f" % (c, (t[c, 0] - g0) / 1e3, r(2), r(3), r(40), r(41)))
        print("    tile 2 epilogue: Tfull %.2f  first-ld %.2f  computed %.2f  dsempty %.2f  done %.2f" % (r(26), r(4), r(5), r(6), r(34)))
        print("    Yissue   %s\n    Yfull    %s\n    Sissue   %s\n    Tfull    %s\n    epi_done %s\n    P2issue  %s" % (f(8), f(48), f(16), f(24), f(32), f(56)))
