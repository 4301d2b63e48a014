"""In-pipeline (warm L2, back-to-back) kernel durations of the bench step via CUPTI (torch.profiler).
Usage: python scripts/kernel_times.py [steps]   -> table of kernel, launches/step, avg us, us/step."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ader_b200 import ops
from ader_b200.model import Ader
from torch.profiler import ProfilerActivity, profile

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
drop = float(os.environ.get("DROPOUT", "0.0"))
WL = bench.WL
dev = torch.device("cuda", 0)
B, Me, V, Vp = WL["B"], WL["M_e"], WL["V"], WL["V_prev"]
M = B + Me
model = Ader(WL["item_num"], bench.make_args(), device=dev, init_seed=0)
model.update_loss(WL["lam"])
rng = np.random.RandomState(100)
t_ids, t_lab, t_len = bench.synth_rows(rng, 32768, V)
e_ids, e_lab, e_len = bench.synth_rows(rng, WL["exemplars"], Vp)
d_t_ids, d_t_lab, d_e_ids = torch.from_numpy(t_ids).to(dev), torch.from_numpy(t_lab).to(dev), torch.from_numpy(e_ids).to(dev)
teacher = torch.randn((WL["exemplars"], (Vp + 3) // 4 * 4), device=dev)[:, :Vp] * 2
ids_buf = torch.empty((M, 50), dtype=torch.int32, device=dev)

def step():
    ti = rng.randint(0, 32768, B).astype(np.int32); ei = rng.randint(0, WL["exemplars"], Me).astype(np.int32)
    dti, dei = torch.from_numpy(ti).to(dev), torch.from_numpy(ei).to(dev)
    ops.gather_rows_i32(d_t_ids, dti, ids_buf[:B]); ops.gather_rows_i32(d_e_ids, dei, ids_buf[B:])
    model.train_step(ids_buf, d_t_lab[dti.long()], V, WL["lr"], drop, exemplar_logits=teacher, teacher_rows=dei,
                     n_tokens=int(t_len[ti].sum() + e_len[ei].sum()))

for _ in range(10):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    a = agg.setdefault(e.name.split("(")[0][:60], [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("kernel time per step: %.1f us over %d steps (dropout %.2f)" % (tot / steps, steps, drop))
print("| kernel | launches/step | avg us | us/step | share |\n|---|---:|---:|---:|---:|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| `%s` | %.1f | %.2f | %.1f | %.1f%% |" % (k, c / steps, t / c, t / steps, 100 * t / tot))
