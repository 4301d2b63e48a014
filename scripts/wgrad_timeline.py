"""Debug: phase stamps of k_wgrad (library built with `python -m ader_b200.build --timeline`, run with
ADER_B200_LIB=ader_b200/lib/libader_b200_tl.so): entry, tile k ready (k = 0..5), loop end, stores issued -- of the LAST launch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ader_b200 import _lib
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
    model.train_step(ids, lab[:B], V, WL["lr"], 0.3, exemplar_logits=teacher, teacher_rows=rows, n_tokens=int(lens.sum()))
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((8, 160, 16), np.int64)
lib.ader_debug_fz_timeline.argtypes = [C.c_void_p]
lib.ader_debug_fz_timeline(buf.ctypes.data_as(C.c_void_p))
t = buf[7]
used = np.nonzero(t[:, 0])[0]
print("k_wgrad, problem 0 of the last GEMM launch (CTAs 0..) and of the last LayerNorm-only launch (CTAs 64..): %d CTAs, T = %d" % (len(used), int(lens.sum())))
labels = ["entry", "tile0", "tile1", "tile2", "tile3", "tile4", "tile5", "-", "loop end", "stored"]
for c in used:
    row = t[c]
    base = row[0]
    print("  cta %2d: " % c + ", ".join("%s %.2f" % (labels[i], (row[i] - base) / 1e3) for i in range(1, 10) if row[i] >= base and row[i] != 0))
print("  start spread %.2f us" % ((t[used, 0].max() - t[used, 0].min()) / 1e3))
