"""2-GPU probe (torchrun): the peer-memory optimiser step alone (arrive + reduce/Adam/all-gather + wait) at the bench
shape, and a plain peer copy of the same bytes for reference."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from ader_b200 import ops
from ader_b200.dist import make_comm
from ader_b200.model import Ader
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
WL = bench.WL
m = Ader(WL["item_num"], bench.make_args(), device=dev, init_seed=0)
m.dp = make_comm(m, backend="p2p")
m.grad.normal_()
V = WL["V"]
def run(n):
    for _ in range(n):
        m.dp.begin_step(m)
        m.dp.apply(m, V, 5e-4)
for _ in range(5):
    run(1)
torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run(1)
for _ in range(5):
    g.replay()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    g.replay()
e1.record(); torch.cuda.synchronize()
t_dp = e0.elapsed_time(e1) / 200 * 1e3
m.dp.check()
# where the time goes: the same step with the peer pointers replaced by local ones (results are meaningless, timing is not)
import ctypes as C
def variant(local_grad, local_theta):
    c = m.dp.comm
    th = [c.theta[r] for r in range(world)]; gr = [c.grad[r] for r in range(world)]; fl = [c.flags[r] for r in range(world)]
    if local_grad: gr = [gr[rank]] * world
    if local_theta: th = [th[rank]] * world
    c2 = ops.dp_comm(rank, world, th, gr, fl, separate_arrive=bool(c.separate_arrive))
    def run2():
        ops.dp_wait(c2)
        ops.dp_adam_step(m.ms, c2, m.adam_m, m.adam_v, m.adam_state, V, 5e-4)
    for _ in range(3): run2()
    torch.cuda.synchronize(); dist.barrier()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        run2()
    for _ in range(5): g2.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(200): g2.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 200 * 1e3
tv = {k: variant(*k) for k in ((True, False), (False, True), (True, True))}
if rank == 0:
    print("variants (us): peer loads only %.1f | peer stores only %.1f | no peer traffic (flags only) %.1f" % (tv[(False, True)], tv[(True, False)], tv[(True, True)]))
# reference: single-GPU Adam over the same index space
e0.record()
for _ in range(200):
    ops.adam_step(m.ms, m.theta, m.adam_m, m.adam_v, m.grad, m.adam_state, V, 5e-4)
e1.record(); torch.cuda.synchronize()
t_adam = e0.elapsed_time(e1) / 200 * 1e3
n_bytes = (V * 150 + 235500) * 4
dist.barrier()
if rank == 0:
    print("world %d: dp step (wait + arrive + reduce/Adam/all-gather) %.1f us; plain Adam %.1f us; slice bytes/rank %.2f MB, link bytes in = out = %.2f MB"
          % (world, t_dp, t_adam, n_bytes / world / 1e6, n_bytes * (world - 1) / world / 1e6))
# NCCL all-reduce of the same ranges for comparison
d = 150
a, b = m.grad[d:(V + 1) * d], m.grad[m.layout.offset(1):]
for _ in range(5):
    dist.all_reduce(a); dist.all_reduce(b)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(100):
    dist.all_reduce(a); dist.all_reduce(b)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print("NCCL all-reduce of the two live ranges: %.1f us" % (e0.elapsed_time(e1) / 100 * 1e3))
dist.barrier()
os._exit(0)
