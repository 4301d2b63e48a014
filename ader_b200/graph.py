"""The training step as CUDA graphs (one per token-capacity bucket).

After the kernel fusion a YOOCHOOSE-shaped step is ~45 launches of 5-25 us each: issued one by one from
Python the host, not the GPU, sets the pace.  ``GraphStep`` captures ``Ader.train_step`` (encoder forward,
logits + CE + distillation, backward, scatter, [NCCL all-reduce], Adam) once per token-capacity bucket into
a ``torch.cuda.CUDAGraph`` and replays it.  Everything a step needs is either

  * a static device buffer (``ids`` / ``pos`` / ``aux`` or the gather indices ``ti`` / ``ei``), filled before
    the replay by stream-ordered copies from rotating pinned host buffers,
  * device-resident state the kernels read themselves: the exact token count T (``row_off[M]``), the Adam step
    (also the dropout counter: ``d_step``), theta / m / v / grad, or
  * a constant of the period (lr, lambda, dropout rate, max_item, the stored teacher matrix).

The token capacity only sizes grids and workspaces (surplus CTAs exit at once), so a handful of buckets
covers every batch; the largest one (M x maxlen) is always safe.  Reference semantics are unchanged: this is
the same sequence of C-ABI calls as the eager ``train_step`` (main.py:233-256).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


class _PinnedRing:
    """Rotating pinned staging buffers for stream-ordered H2D copies (a slot is reused only after its copy ran)."""

    def __init__(self, shape, n=4):
        self.bufs = [torch.empty(shape, dtype=torch.int32).pin_memory() for _ in range(n)]
        self.evs = [None] * n
        self.i = 0

    def stage(self, src: np.ndarray, dst: torch.Tensor):
        k = self.i
        self.i = (k + 1) % len(self.bufs)
        if self.evs[k] is not None:
            self.evs[k].synchronize()
        self.bufs[k].numpy()[...] = src
        dst.copy_(self.bufs[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.evs[k] = ev

    def stage_parts(self, dst: torch.Tensor, parts, keep_tail: bool = False):
        """Concatenate the host arrays `parts` into one pinned slot and copy it to the front of `dst` in one transfer."""
        k = self.i
        self.i = (k + 1) % len(self.bufs)
        if self.evs[k] is not None:
            self.evs[k].synchronize()
        buf = self.bufs[k].numpy()
        o = 0
        for a in parts:
            a = np.asarray(a, dtype=np.int32).reshape(-1)
            buf[o:o + a.size] = a
            o += a.size
        if not keep_tail and o != buf.size:
            raise ValueError("batch parts hold %d values, the step takes %d" % (o, buf.size))
        dst[:o].copy_(self.bufs[k][:o], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.evs[k] = ev


class PendingLoss:
    """Loss of a replayed step on its way to the host (pinned 4-byte D2H copy issued right behind the step)."""

    def __init__(self, buf: torch.Tensor, ev: torch.cuda.Event):
        self.buf, self.ev = buf, ev

    def result(self) -> float:
        self.ev.synchronize()
        return float(self.buf[0])


class GraphStep:
    """Captured ``train_step`` for a fixed batch geometry (n_train + n_ex rows) of one period.

    sources: optional ``(t_ids [Nt, L], t_lab [Nt], e_ids [Ne, L], e_aux [Ne])`` device int32 matrices; when
    given the graph starts by gathering the batch rows from them with the static index buffers ``ti`` / ``ei``
    (``run_indices``); ``e_aux`` is the teacher row of each exemplar (distillation) or its label (one-hot replay).
    """

    def __init__(self, model, n_train: int, n_ex: int, max_item: int, lr: float, dropout_rate: float = 0.0,
                 teacher: Optional[torch.Tensor] = None, sources=None, tcaps: Optional[Sequence[int]] = None, queue=None):
        if dropout_rate > 0.0 and model.encoder_impl != "tc":
            raise ValueError("graph replay with dropout needs the tc encoder (device-side dropout counter)")
        self.model, self.n_train, self.n_ex = model, int(n_train), int(n_ex)
        self.max_item, self.lr, self.p = int(max_item), float(lr), float(dropout_rate)
        self.teacher = teacher
        dev, L = model.device, model.hp.maxlen
        M = self.n_train + self.n_ex
        self.M = M
        # one static int32 buffer [ids | pos | aux]: a host-fed step is ONE pinned staging copy + ONE H2D transfer
        n_aux = max(self.n_ex, 1)
        self.batch = torch.zeros(M * L + self.n_train + n_aux, dtype=torch.int32, device=dev)
        self.ids = self.batch[:M * L].view(M, L)
        self.pos = self.batch[M * L:M * L + self.n_train]
        self.aux = self.batch[M * L + self.n_train:]
        self.pos.fill_(1)
        if self.n_ex > 0 and model.mode == model.ER:
            self.aux.fill_(1)
        self.sources = sources
        # epoch-resident index queue (q int32, q_off int64, counter int32[1]) shared by the GraphSteps of a period: the
        # "queued" graph form gathers step `counter` from it and advances the counter -- no host-to-device copy per step
        self.queue = queue
        self.graphs_q = {}
        if sources is not None:
            self.ti = torch.zeros(self.n_train, dtype=torch.int32, device=dev)
            self.ei = torch.zeros(max(self.n_ex, 1), dtype=torch.int32, device=dev)
            self._ring_ti = _PinnedRing((self.n_train,))
            self._ring_ei = _PinnedRing((max(self.n_ex, 1),))
        self._ring_batch = _PinnedRing((self.batch.numel(),))
        self._ring_ids = _PinnedRing((M, L))
        self._loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(8)]
        self._loss_i = 0
        self._ring_pos = _PinnedRing((self.n_train,))
        self._ring_aux = _PinnedRing((max(self.n_ex, 1),))
        full = M * L
        caps = sorted({min(max(int(t), 1), full) for t in (tcaps or [])} | {full})
        self.tcaps = caps
        self.graphs = {}
        self.use_count = {}                    # replays per token-capacity bucket
        self._row_loss = {}
        self._capture_all()
        # the graphs address these buffers: keep them alive even if the model later grows its workspaces
        self._held = (model._enc_ws.buf, model._bwd_ws.buf, model._loss_ws.buf, model.theta, model.grad, model.adam_m,
                      model.adam_v, model.adam_state, model._loss, teacher, sources)

    # ---- capture ------------------------------------------------------------------------------------
    def _gather(self):
        """Batch assembly from the GPU-resident row matrices (its own small graph: run_rows skips it)."""
        t_ids, t_lab, e_ids, e_aux = self.sources
        if self.n_ex > 0:
            ops.gather_batch(t_ids, t_lab, self.ti, e_ids, e_aux, self.ei[:self.n_ex], self.ids, self.pos, self.aux[:self.n_ex])
        else:
            ops.gather_batch(t_ids, t_lab, self.ti, None, None, None, self.ids, self.pos, None)

    def _gather_q(self):
        t_ids, t_lab, e_ids, e_aux = self.sources
        q, q_off, counter = self.queue
        ops.gather_batch_q(t_ids, t_lab, self.n_train, e_ids if self.n_ex > 0 else None, e_aux if self.n_ex > 0 else None,
                           self.n_ex, q, q_off, counter, self.ids, self.pos, self.aux[:self.n_ex] if self.n_ex > 0 else None)
        ops.queue_advance(counter)

    def _eager(self, tcap: int, device_step: bool, indexed: bool = False):
        m = self.model
        kw = {}
        if self.n_ex > 0:
            if m.mode == m.KD:
                kw = dict(exemplar_logits=self.teacher, teacher_rows=self.aux[:self.n_ex])
            elif m.mode == m.ER:
                kw = dict(exemplar_pos=self.aux[:self.n_ex])
        loss = m.train_step(self.ids, self.pos, self.max_item, self.lr, self.p, n_tokens=tcap,
                            _device_step=device_step, **kw)
        self._row_loss[(tcap, indexed)] = m.last_row_loss
        return loss

    def _capture_all(self):
        m = self.model
        sd = m.state_dict()
        gs = m.global_step
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up on the largest capacity: sizes every workspace once
            if self.sources is not None:
                self._gather()
            self._eager(self.tcaps[-1], False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        m.load_state_dict(sd)                  # the warm-up step must not count
        m.global_step = gs
        torch.cuda.synchronize()
        # graphs are captured on first use, one per (token-capacity bucket, feed form): a period that is over after a few
        # hundred steps pays for the two or three it replays, not for all of them
        self._pool = None
        self._held_more = []
        self.graphs_idx = {}                   # index-fed form: batch gather + step in ONE graph (no gap between two replays)

    def _graph_for(self, tcap: int, indexed, nsteps: int = 1) -> torch.cuda.CUDAGraph:
        """indexed: False = rows fed into the static batch buffer, True = index-fed (gather inside the graph),
        "q" = gather from the epoch-resident queue inside the graph.  nsteps > 1 (queue form only): that many
        consecutive steps in ONE graph - every step reads its rows, the token count, the Adam step and the dropout
        counter from the device, so the captured launches are the same for each; the launch gap between two graph
        replays (~20 us) is paid once per nsteps."""
        table = self.graphs_q if indexed == "q" else (self.graphs_idx if indexed else self.graphs)
        key = tcap if nsteps == 1 else (tcap, nsteps)
        g = table.get(key)
        if g is None:
            if nsteps > 1 and indexed != "q":
                raise ValueError("multi-step graphs need the epoch-resident queue")
            m = self.model
            gs = m.global_step                 # capture records launches, it runs nothing: only the host counter moves
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool):
                for _ in range(nsteps):
                    if indexed == "q":
                        self._gather_q()
                    elif indexed:
                        self._gather()
                    self._eager(tcap, True, indexed=indexed)
            self._pool = g.pool()
            m.global_step = gs
            table[key] = g
            # the graph addresses the workspaces as they are NOW (another pass may have grown them since construction)
            self._held_more.append((m._enc_ws.buf, m._bwd_ws.buf, m._loss_ws.buf))
        return g

    def precapture(self, indexed: Optional[bool] = None):
        """Capture every bucket now (benchmarks: keep capture time out of the timed region)."""
        for tcap in self.tcaps:
            for form in ((False, True) if indexed is None else (indexed,)):
                if form and self.sources is None:
                    continue
                self._graph_for(tcap, form)

    # ---- replay ---------------------------------------------------------------------------------------
    def cap_for(self, n_tokens: Optional[int]) -> int:
        """Smallest token-capacity bucket that holds n_tokens."""
        if n_tokens is not None:
            for t in self.tcaps:
                if t >= n_tokens:
                    return t
        return self.tcaps[-1]

    def _replay(self, n_tokens: Optional[int], indexed=False, nsteps: int = 1):
        cap = self.cap_for(n_tokens)
        self.use_count[cap] = self.use_count.get(cap, 0) + nsteps
        self._graph_for(cap, indexed, nsteps).replay()
        self.model._last_enc = (self.M, cap)           # geometry of the pass whose overflow flag token_overflow() reads
        self.model.global_step += nsteps
        self.model.last_row_loss = self._row_loss[(cap, indexed)]
        return self.model._loss

    @staticmethod
    def _put(ring, dst, src):
        if isinstance(src, torch.Tensor):
            dst.copy_(src.to(torch.int32).view(dst.shape), non_blocking=True)
        else:
            ring.stage(np.asarray(src, dtype=np.int32).reshape(tuple(dst.shape)), dst)

    def run_rows(self, ids, pos, aux=None, n_tokens: Optional[int] = None):
        """ids [M, L], pos [n_train], aux [n_ex] (teacher rows or exemplar labels): host arrays or device tensors."""
        host = not isinstance(ids, torch.Tensor) and not isinstance(pos, torch.Tensor) and not isinstance(aux, torch.Tensor)
        if host and (self.n_ex == 0 or aux is not None):
            self._ring_batch.stage_parts(self.batch, (ids, pos) if self.n_ex == 0 else (ids, pos, aux),
                                         keep_tail=(self.n_ex == 0))
            return self._replay(n_tokens)
        self._put(self._ring_ids, self.ids, ids)
        self._put(self._ring_pos, self.pos, pos)
        if self.n_ex > 0 and aux is not None:
            self._put(self._ring_aux, self.aux[:self.n_ex], aux)
        return self._replay(n_tokens)

    def fetch_loss(self) -> PendingLoss:
        """Queue the device->host read of the step just replayed; ``.result()`` blocks until it has landed.  Lets a host
        loop that wants every step's loss (main.py:233-256 returns it from sess.run) keep one step in flight: feed step
        i+1, then read the loss of step i."""
        k = self._loss_i
        self._loss_i = (k + 1) % len(self._loss_host)
        self._loss_host[k].copy_(self.model._loss.view(-1)[:1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return PendingLoss(self._loss_host[k], ev)

    def run_queued(self, n_tokens: Optional[int] = None, nsteps: int = 1):
        """Replay the step(s) whose rows are the next entries of the epoch-resident queue (no copies: the host only
        launches).  n_tokens: the LARGEST token count among the nsteps steps."""
        if self.queue is None or self.sources is None:
            raise ValueError("GraphStep was built without an index queue")
        return self._replay(n_tokens, indexed="q", nsteps=nsteps)

    def run_indices(self, ti, ei=None, n_tokens: Optional[int] = None):
        """Row indices into the ``sources`` matrices (host arrays or device tensors)."""
        if self.sources is None:
            raise ValueError("GraphStep was built without row sources")
        self._put(self._ring_ti, self.ti, ti)
        if self.n_ex > 0 and ei is not None:
            self._put(self._ring_ei, self.ei[:self.n_ex], ei)
        return self._replay(n_tokens, indexed=True)
