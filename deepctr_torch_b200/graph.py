"""CUDA-graph capture of one training step (forward -> loss -> backward).

The hot path is ~30 short kernel launches per step; at B200 speeds the Python/launch overhead of
issuing them one by one is comparable to their run time.  Every launch of libctr_b200.so goes to
torch's current stream, allocates nothing and never synchronises, so the whole step is capturable:
``GraphedStep`` records it once into a ``torch.cuda.CUDAGraph`` and replays it with new inputs copied
into static buffers.  Gradients land in the same ``.grad`` tensors on every replay (dense tower: dense
tensors; tables: dense or per-unique-row sparse COO, depending on ``model.table_grad``).
"""
from __future__ import annotations

import torch

from . import ops


class _PrefetchMixin:
    """Input pipeline shared by the graphed steps: the next batch travels host -> device on a copy stream while the
    current step computes.  Needs ``self.X`` / ``self.y`` (the graph's static inputs) and ``self._replay()``."""

    # ---- input pipeline: the next batch travels host -> device while the current step computes ----
    def enable_prefetch(self):
        """Two device staging buffers + a copy stream.  ``prefetch(X, y)`` (pinned host tensors) enqueues
        the H2D copy of a FUTURE step on the copy stream; ``step_prefetched()`` consumes the oldest
        prefetched batch (device-to-device into the graph's static inputs, ~10 us) and replays."""
        dev = self.X.device
        cur = torch.cuda.current_stream(dev)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._stage = [(torch.empty_like(self.X), torch.empty_like(self.y)) for _ in range(2)]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]    # H2D into staging[k] finished (copy stream)
        self._free = [torch.cuda.Event(), torch.cuda.Event()]     # staging[k] consumed (compute stream)
        for e in self._free:
            e.record(cur)
        self._pf = 0        # next staging slot to fill
        self._use = 0       # next staging slot to consume
        self._pending = 0

    def prefetch(self, X, y):
        if self._pending >= 2:
            raise RuntimeError("GraphedStep.prefetch: both staging buffers are full")
        k = self._pf
        self._pf ^= 1
        self._pending += 1
        cs = self._copy_stream
        cs.wait_event(self._free[k])
        with torch.cuda.stream(cs):
            self._stage[k][0].copy_(X, non_blocking=True)
            self._stage[k][1].copy_(y.reshape(-1), non_blocking=True)
            self._ready[k].record(cs)

    def step_prefetched(self):
        if self._pending <= 0:
            raise RuntimeError("GraphedStep.step_prefetched: nothing was prefetched")
        k = self._use
        self._use ^= 1
        self._pending -= 1
        cur = torch.cuda.current_stream(self.X.device)
        cur.wait_event(self._ready[k])
        self.X.copy_(self._stage[k][0], non_blocking=True)
        self.y.copy_(self._stage[k][1], non_blocking=True)
        self._free[k].record(cur)
        return self._replay()


class GraphedStep(_PrefetchMixin):
    def __init__(self, model, batch_size, loss_fn=None, warmup=3, with_reg=False):
        self.model = model
        dev = torch.device(model.device)
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs a CUDA model")
        n_cols = max(e for _, e in model.feature_index.values())
        self.X = torch.zeros(batch_size, n_cols, device=dev, dtype=torch.float32)
        self.y = torch.zeros(batch_size, device=dev, dtype=torch.float32)
        self.loss_fn = loss_fn or ops.binary_cross_entropy
        self.with_reg = with_reg
        model.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):     # sizes workspaces, sets kernel attributes
                model.zero_grad(set_to_none=True)
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.check_ids()
        model.zero_grad(set_to_none=True)
        plan = model._plan
        if plan is not None:
            plan.pending.clear()
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib
        l0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss, self.y_pred = self._body()
        # kernels of libctr_b200.so recorded as graph nodes: each replay launches exactly these
        self.launches_per_replay = _lib.launch_count() - l0
        self.replays = 0
        # row gradients parked for the fused optimizer (plan.keep_rowgrads): the captured buffers are
        # rewritten by every replay, so every replay re-publishes them
        self._parked = list(plan.pending) if plan is not None else []

    def _body(self):
        y_pred = self.model(self.X)
        loss = self.loss_fn(y_pred.squeeze(1), self.y, reduction="sum")
        total = loss
        if self.with_reg:
            total = loss + self.model.get_regularization_loss().sum()
        total.backward()
        return loss, y_pred

    def __call__(self, X, y):
        """Copy one batch in (host or device tensors) and replay; returns the (device) loss tensor."""
        self.X.copy_(X, non_blocking=True)
        self.y.copy_(y.reshape(-1), non_blocking=True)
        return self._replay()

    def _replay(self):
        self.graph.replay()
        self.replays += 1
        if self._parked:
            self.model._plan.pending[:] = self._parked
        return self.loss


class ShardedGraphedStep(_PrefetchMixin):
    """The row-sharded multi-GPU step (forward with the row exchange, loss, backward with the row-gradient
    push, dense-gradient all-reduce, receive-list bookkeeping) as TWO alternating CUDA graphs — one per
    step parity, because the receive lists are double-buffered by parity.  NCCL collectives are captured
    like any other kernel; all ranks must construct and call this object in lock-step."""

    def __init__(self, model, batch_size, loss_fn=None, warmup=4):
        if getattr(model, "sharded", None) is None:
            raise RuntimeError("ShardedGraphedStep needs a model prepared by sharded.attach_shards")
        self.model = model
        dev = torch.device(model.device)
        n_cols = max(e for _, e in model.feature_index.values())
        self.X = torch.zeros(batch_size, n_cols, device=dev, dtype=torch.float32)
        self.y = torch.zeros(batch_size, device=dev, dtype=torch.float32)
        self.loss_fn = loss_fn or ops.binary_cross_entropy
        model.train()
        if warmup % 2:
            warmup += 1                                  # leave the parity where it started
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.check_ids()
        from . import _lib
        self.graphs, self.losses = [], []
        self.first_parity = model._plan.step_parity
        l0 = _lib.launch_count()
        for _ in range(2):                               # parity p, then parity p ^ 1
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = self._body()
            self.graphs.append(g)
            self.losses.append(loss)
        self.launches_per_replay = (_lib.launch_count() - l0) // 2
        self.replays = 0

    def _body(self):
        m = self.model
        m.zero_grad(set_to_none=True)
        y_pred = m(self.X)
        loss = self.loss_fn(y_pred.squeeze(1), self.y, reduction="sum")
        loss.backward()
        done = m.sharded.finish_step()
        m.sharded.combine_received(done)       # owner-side dedup-sum: the same (uniq, rowgrad) contract as 1 GPU
        m.sharded.clear_received(done)
        return loss

    def __call__(self, X, y):
        self.X.copy_(X, non_blocking=True)
        self.y.copy_(y.reshape(-1), non_blocking=True)
        return self._replay()

    def _replay(self):
        i = self.replays & 1
        self.graphs[i].replay()
        self.replays += 1
        return self.losses[i]
