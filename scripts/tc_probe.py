"""Debug probe (GPU): tcgen05 loss path vs the exact fp32 path for each descriptor variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from ader_b200.model import Ader
args = type("A", (), dict(hidden_units=150, maxlen=50, num_blocks=2, num_heads=1, random_seed=0, lr=5e-4, dropout_rate=0.0, disable_distillation=False))()
M, Bt, V, Vp, item_num = %d, %d, %d, %d, %d
m = Ader(item_num, args, init_seed=0)
m.theta.add_(torch.randn_like(m.theta) * 0.05)
rng = np.random.RandomState(0)
ids = np.zeros((M, 50), np.int32)
for r in range(M):
    n = int(rng.randint(1, 12)); ids[r, 50 - n:] = rng.randint(1, V + 1, n)
pos = rng.randint(1, V + 1, Bt).astype(np.int32)
teacher = torch.randn(M - Bt, Vp, device="cuda") * 2
m.update_loss(0.8)
out = {}
for impl in ("exact", "tc"):
    m.loss_impl = impl
    m.grad.zero_()
    loss = float(m.loss_and_grad(ids, pos, V, exemplar_logits=teacher).item())
    torch.cuda.synchronize()
    out[impl] = (loss, m.last_row_loss.clone(), m._keep[-1].clone(), m.grad.clone())
le, re_, de, ge = out["exact"]; lt, rt, dt, gt = out["tc"]
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-30))
lay = m.layout
tab_e, tab_t = lay.views(ge)[0][1:V+1], lay.views(gt)[0][1:V+1]
print("loss exact %%.6f tc %%.6f | row_loss rel %%.3e | d_rep rel %%.3e | table grad rel %%.3e | dense grad rel %%.3e" %% (
    le, lt, rel(rt, re_), rel(dt, de), rel(tab_t, tab_e), rel(gt[lay.dense_offset:], ge[lay.dense_offset:])))
'''

def main():
    shapes = [(24, 16, 450, 400, 500), (300, 200, 5000, 4000, 6000)]
    variants = sys.argv[1:] or ["0", "1", "2", "3"]
    for v in variants:
        for shp in shapes:
            env = dict(os.environ, ADER_TC_VARIANT=v)
            try:
                r = subprocess.run([sys.executable, "-c", CHILD % ((ROOT,) + shp)], env=env, capture_output=True, text=True, timeout=120)
                tail = (r.stdout.strip().splitlines() or ["<no stdout>"])[-1]
                err = r.stderr.strip().splitlines()[-1] if r.returncode else ""
                print("variant %s shape %s rc=%d: %s %s" % (v, shp, r.returncode, tail, err), flush=True)
            except subprocess.TimeoutExpired:
                print("variant %s shape %s: TIMEOUT" % (v, shp), flush=True)

if __name__ == "__main__":
    main()
