"""BASELINE config 5: synthetic 1 M-item vocabulary, batch 4096 (+1024 KD rows, V_prev = 900 k):
vocab-parallel logits + softmax CE + distillation forward+backward over N ranks (torchrun).
Prints one JSON line (rank 0): device time per fwd+bwd (max over ranks), algorithmic TFLOP/s."""
import json, os, sys, numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ader_b200.dist import VocabParallelLoss, env_rank
from ader_b200.model import Ader
rank, world, local = env_rank()
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
else:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
V, Vp, B, Me = 1000000, 900000, 4096, 1024
if "--small" in sys.argv: V, Vp, B, Me = 200000, 180000, 4096, 1024
args = type("A", (), dict(hidden_units=150, maxlen=50, num_blocks=2, num_heads=1, random_seed=0, lr=5e-4, dropout_rate=0.0, disable_distillation=False))()
m = Ader(V, args, device=dev, init_seed=0)
g = torch.Generator(device=dev).manual_seed(1)             # identical inputs on every rank
rep = torch.randn(B + Me, 150, device=dev, generator=g)
pos = torch.randint(1, V + 1, (B,), device=dev, dtype=torch.int32, generator=g)
teacher = torch.randn(Me, Vp, device=dev, generator=g)      # Vp % 4 == 0 -> 16-byte aligned rows
vp = VocabParallelLoss(m)
def step(): return vp.fwd_bwd(rep, pos, V, lambda_=0.8, mode=1, teacher=teacher)
for _ in range(3): loss, _, _ = step()
dist.barrier(); torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): loss, _, d_rep = step()
e1.record(); dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item()); M = B + Me
    flops = 6.0 * 150 * (B * V + Me * Vp)
    print(json.dumps({"what": "vocab_parallel_logits_ce_kd_fwd_bwd", "n_gpus": world, "M": M, "V": V, "V_prev": Vp, "ms": ms,
                      "sessions_per_s": M / ms * 1e3, "algorithmic_tflops_total": flops / ms / 1e9,
                      "algorithmic_tflops_per_gpu": flops / ms / 1e9 / world, "loss": float(loss),
                      "d_rep_checksum": float(d_rep.double().abs().sum())}))
sys.stdout.flush(); torch.cuda.synchronize(); os._exit(0)   # no NCCL teardown (can wedge at exit)
