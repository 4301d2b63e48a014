"""Debug: per-CTA pipeline timeline of the tcgen05 logits kernels (build with ADER_B200_DEFINES=-DADER_TC_TIMELINE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ader_b200 import ops, _lib
from ader_b200.model import Ader
WL = bench.WL
dev = torch.device("cuda", 0)
B, Me, V, Vp = WL["B"], WL["M_e"], WL["V"], WL["V_prev"]
M = B + Me
model = Ader(WL["item_num"], bench.make_args(), device=dev, init_seed=0)
model.update_loss(WL["lam"])
rng = np.random.RandomState(100)
ids, lab, lens = bench.synth_rows(rng, M, V)
teacher = torch.randn((WL["exemplars"], (Vp + 3) // 4 * 4), device=dev)[:, :Vp] * 2
rows = torch.from_numpy(rng.randint(0, WL["exemplars"], Me).astype(np.int32)).to(dev)
for _ in range(5):
    model.train_step(ids, lab[:B], V, WL["lr"], 0.0, exemplar_logits=teacher, teacher_rows=rows, n_tokens=int(lens.sum()))
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((3, 192, 64), np.int64)
lib.ader_debug_tc_timeline.argtypes = [C.c_void_p]
lib.ader_debug_tc_timeline(buf.ctypes.data_as(C.c_void_p))
MHZ = 1000.0   # stamps are %globaltimer nanoseconds
for mode, name in enumerate(["FWD", "DREP", "DE"]):
    t = buf[mode]
    used = np.nonzero(t[:, 1])[0]
    g0 = t[used, 0].min()
    print("== %s: %d CTAs; CTA start spread %.1f us; CTA durations (us): min %.1f median %.1f max %.1f" % (
        name, len(used), (t[used, 0].max() - g0) / 1e3, ((t[used, 41] - t[used, 1]) / MHZ).min(),
        np.median((t[used, 41] - t[used, 1]) / MHZ), ((t[used, 41] - t[used, 1]) / MHZ).max()))
    slow = used[np.argsort(-(t[used, 41] - t[used, 1]))[:2]]
    fast = used[np.argsort((t[used, 41] - t[used, 1]))[:1]]
    for c in list(slow) + list(fast):
        r = lambda s: (t[c, s] - t[c, 1]) / MHZ if t[c, s] else float("nan")
        print(" CTA %3d start+%.1fus: setup %.1f | Xissue %.1f | Yissue %s | Yfull %s | Sissue %s | Tfull %s | epi_done %s | P2issue %s | end %.1f / %.1f" % (
            c, (t[c, 0] - g0) / 1e3, r(2), r(3), ["%.1f" % r(8 + i) for i in range(7)], ["%.1f" % r(48 + i) for i in range(7)],
            ["%.1f" % r(16 + i) for i in range(7)], ["%.1f" % r(24 + i) for i in range(7)], ["%.1f" % r(32 + i) for i in range(7)],
            ["%.1f" % r(56 + i) for i in range(7)], r(40), r(41)))
