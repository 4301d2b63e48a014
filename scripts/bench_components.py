"""Kernel-group timings on one B200 (CUDA events, warm-up 3, inputs > L2 or L2 flushed where stated):
eval ranking, herding selection, EWC Fisher, Adam, embedding scatter -- the non-training rows of
SURVEY 8(a) (configs C and D of BASELINE.json).  Prints one JSON object per measurement."""
import json, math, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ader_b200 import ops
from ader_b200.model import Ader, Ewc
sys.path.insert(0, ROOT)
import bench as B

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=10, warm=3, flush_l2=True):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush_l2: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

def out(**kw): print(json.dumps(kw), flush=True)

args = B.make_args()
# ---- DIGINETICA shapes (period 10) ----
item_num, V = 43136, 40135
m = Ader(item_num, args, init_seed=0)
rng = np.random.RandomState(0)

# eval: rank of gt among V items (rank only, and rank + top-20), R rows
R = 16384
ids, lab, lens = B.synth_rows(rng, R, V)
d_ids = torch.from_numpy(ids).to(dev); d_lab = torch.from_numpy(lab).to(dev)
for k in (0, 20):
    ms = timed(lambda: m.rank_topk(d_ids, d_lab, V, k, n_tokens=int(lens.sum())), reps=5, flush_l2=False)
    out(what="eval_rank_topk", k=k, rows=R, V=V, ms=ms, rows_per_s=R / ms * 1e3,
        algorithmic_tflops=2.0 * R * 150 * V / ms / 1e9, note="exact fp32 scores (SIMT GEMM) + rank/top-k row kernel; [R,V] chunk materialised")

# herding: N candidates in ~17k label segments (DIGINETICA: median 2, p99 20, max 98), quota 30000
N = 96000
seg = np.minimum(98, np.maximum(1, rng.zipf(1.6, 40000))).astype(np.int64)
seg = seg[np.cumsum(seg) <= N]; N = int(seg.sum())
seg_off = np.zeros(len(seg) + 1, np.int32); np.cumsum(seg, out=seg_off[1:])
quota = np.minimum(seg, np.maximum(0, np.round(seg * 30000.0 / N + rng.rand(len(seg)) - 0.5))).astype(np.int32)
steps = np.array([int(math.ceil(1.1 * q)) for q in quota], np.int32)
rep = torch.randn(N, 150, device=dev)
cand = torch.arange(N, dtype=torch.int32, device=dev)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
picks = torch.zeros(N, dtype=torch.int32, device=dev); n_picked = torch.zeros(len(seg), dtype=torch.int32, device=dev)
ws = torch.empty(ops.herding_ws_bytes(m.ms, N), dtype=torch.uint8, device=dev)
a_seg, a_q, a_s = t(seg_off), t(quota), t(steps)
ms = timed(lambda: ops.herding_segmented(m.ms, rep, cand, a_seg, a_q, a_s, ws, picks, n_picked), reps=5)
out(what="herding_segmented", N=N, segments=len(seg), selected=int(n_picked.sum().item()), ms=ms,
    compulsory_bytes=N * 150 * 4, achieved_gbs=N * 150 * 4 / ms / 1e6, hbm_peak_gbs=PEAKS["hbm_gbs"],
    max_dependent_steps=int(steps.max()), note="critical path is the dependent arg-max chain, not HBM")
ms_rep = timed(lambda: m.rep(d_ids, n_tokens=int(lens.sum())), reps=5, flush_l2=False)
out(what="encoder_rep_pass", rows=R, tokens=int(lens.sum()), ms=ms_rep, rows_per_s=R / ms_rep * 1e3)

# Adam over table rows 1..V + dense params: 28 B/param
n_par = V * 150 + m.layout.dense_count
ms = timed(lambda: m.apply_gradients(V, 5e-4), reps=10)
out(what="adam_tf1", params=n_par, ms=ms, bytes=28 * n_par, achieved_gbs=28 * n_par / ms / 1e6, hbm_peak_gbs=PEAKS["hbm_gbs"],
    frac=28 * n_par / ms / 1e6 / PEAKS["hbm_gbs"], note="L2 flushed before each launch")

# EWC Fisher: per-sample gradient + fp64 accumulate (EWC.py:126-164), 200 samples
e = Ewc(item_num, args, init_seed=0)
data = [rng.randint(1, V + 1, rng.randint(2, 12)).tolist() for _ in range(200)]
torch.cuda.synchronize(); t0 = time.time(); e.compute_fisher(None, data, 50, V); torch.cuda.synchronize(); dt = time.time() - t0
out(what="ewc_fisher", samples=200, seconds=dt, samples_per_s=200 / dt, note="batch-of-one backward per sample (reference semantics), fp64 accumulation on device")
