"""Row-sharded embedding tables across the GPUs of one node (one process per GPU).

Replaces the reference's single-process ``torch.nn.DataParallel`` (reference
``deepctr_torch/models/basemodel.py:206-209``), which replicates every table on every GPU each
step.  Here every table is split by rows: rank ``s`` owns ``{id : id % G == s}`` at local index
``id // G``; the dense tower is replicated and its gradients are all-reduced with NCCL.

Sparse data path (no NCCL collective on it — see ``csrc/p2p.cu``):

* every rank allocates ONE peer-shareable arena (``ctr_p2p_alloc``) with an identical layout:
  its table shards followed by two receive lists for row gradients; CUDA IPC handles are exchanged
  once (``all_gather_object``) and opened with ``ctr_p2p_open``;
* forward: the fused gather kernel reads remote rows through the peer pointers (NVLink P2P loads);
* backward: local duplicate-free row gradients, then ``ctr_rowgrad_push`` appends each
  (local row, gradient) to its owner's receive list (remote atomic slot claim + P2P stores).
  The dense-gradient all-reduce that follows doubles as the barrier that makes the lists complete;
  the lists are double-buffered by step parity so a fast rank can never overwrite a list its owner
  is still consuming.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops


# ----------------------------------------------------------------------------------------------
# pure host logic (exercised on CPU with gloo in tests/test_sharded_host.py)
# ----------------------------------------------------------------------------------------------
def local_rows(vocab, rank, world):
    """Number of rows of a [vocab, *] table owned by `rank` (ids with id % world == rank)."""
    return (vocab - rank + world - 1) // world if vocab > rank else 0


def max_local_rows(vocab, world):
    return (vocab + world - 1) // world


def shard_rows(full, rank, world):
    """Rows of `full` owned by `rank`, in local order (local index = id // world)."""
    return full[rank::world].contiguous()


def unshard_rows(shards, vocab):
    """Inverse of shard_rows: list of per-rank shards -> the full [vocab, *] table."""
    world = len(shards)
    out = shards[0].new_empty((vocab,) + tuple(shards[0].shape[1:]))
    for r, sh in enumerate(shards):
        n = local_rows(vocab, r, world)
        out[r::world] = sh[:n]
    return out


class ArenaLayout:
    """Identical on every rank: byte offsets of table shards and receive lists inside the arena."""

    def __init__(self, emb_vocabs, lin_vocabs, dim, world, batch, align=256, n_id_cols=None):
        self.world, self.D = world, dim
        self.n_emb, self.n_lin = len(emb_vocabs), len(lin_vocabs)
        self.cap = batch * world                       # worst case: every rank sends `batch` rows
        cursor = 0

        def take(nbytes):
            nonlocal cursor
            off = cursor
            cursor = (cursor + nbytes + align - 1) // align * align
            return off

        self.emb_rows = [max_local_rows(v, world) for v in emb_vocabs]
        self.lin_rows = [max_local_rows(v, world) for v in lin_vocabs]
        self.emb_off = [take(r * dim * 4) for r in self.emb_rows]
        self.lin_off = [take(r * 4) for r in self.lin_rows]
        nf = self.n_emb + self.n_lin
        self.recv = []
        for _ in range(2):                              # double-buffered by step parity
            self.recv.append({"count": take(max(nf, 1) * 4), "ids": take(max(nf, 1) * self.cap * 4),
                              "emb": take(max(self.n_emb, 1) * self.cap * dim * 4),
                              "lin": take(max(self.n_lin, 1) * self.cap * 4)})
        # forward exchange (ctr_shard_request / ctr_shard_serve): request lists and response rows, one slice
        # per peer, sized for the worst case (every id of the batch owned by one peer): about 1 GB per rank at
        # BASELINE config #5, nothing on a 180 GB part, and a skewed id distribution can never overflow
        n_cols = n_id_cols if n_id_cols is not None else max(self.n_emb, self.n_lin, 1)
        self.xcap = batch * n_cols
        self.flags = take(max(world, 1) * 4)            # device-side barrier flags (ctr_p2p_barrier)
        self.x = {"req_cnt": take(world * 4), "req": take(world * self.xcap * 8),
                  "resp_emb": take(world * self.xcap * max(dim, 1) * 4), "resp_lin": take(world * self.xcap * 4)}
        self.nbytes = cursor


def gather_full_state_dict(local_state, table_vocabs, group=None):
    """Gather-on-save: turn a rank-local state_dict (sharded tables, replicated dense params) into
    the reference-compatible full state_dict.  `table_vocabs`: {state key: logical vocab}."""
    world = dist.get_world_size(group)
    full = {}
    for k, v in local_state.items():
        if k in table_vocabs:
            rows = max_local_rows(table_vocabs[k], world)
            pad = v.new_zeros((rows,) + tuple(v.shape[1:]))
            pad[:v.shape[0]] = v
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
            full[k] = unshard_rows(parts, table_vocabs[k])
        else:
            full[k] = v
    return full


def scatter_full_state_dict(full_state, table_vocabs, rank, world):
    """Scatter-on-load: the rank-local view of a reference-compatible full state_dict."""
    out = {}
    for k, v in full_state.items():
        if k in table_vocabs:
            sh = shard_rows(v, rank, world)
            rows = max_local_rows(table_vocabs[k], world)
            pad = sh.new_zeros((rows,) + tuple(sh.shape[1:]))
            pad[:sh.shape[0]] = sh
            out[k] = pad
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------------------------
# device side
# ----------------------------------------------------------------------------------------------
class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can alias memory from ctr_p2p_alloc."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class P2PArena:
    def __init__(self, nbytes, device, group=None):
        self.device = torch.device(device)
        self.nbytes = nbytes
        lib = _lib.load()
        p = ctypes.c_void_p()
        rc = lib.ctr_p2p_alloc(nbytes, ctypes.byref(p))
        if rc != 0:
            raise _lib.CtrLibraryError("ctr_p2p_alloc(%d) failed: %s" % (nbytes, lib.ctr_last_error().decode()))
        self.ptr = p.value
        self._raw = _RawCuda(self.ptr, nbytes)
        self.bytes = torch.as_tensor(self._raw, device=self.device)
        handle = (ctypes.c_ubyte * 64)()
        rc = lib.ctr_p2p_export(ctypes.c_void_p(self.ptr), handle)
        if rc != 0:
            raise _lib.CtrLibraryError("ctr_p2p_export failed: %s" % lib.ctr_last_error().decode())
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_ptr = []
        for r in range(world):
            if r == rank:
                self.peer_ptr.append(self.ptr)
                continue
            q = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            rc = lib.ctr_p2p_open(buf, ctypes.byref(q))
            if rc != 0:
                raise _lib.CtrLibraryError("ctr_p2p_open(rank %d) failed: %s" % (r, lib.ctr_last_error().decode()))
            self.peer_ptr.append(q.value)

    def view(self, offset, shape, dtype=torch.float32):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        return self.bytes[offset:offset + nbytes].view(dtype).view(*shape)


class ShardedPlan(ops.GatherPlan):
    """GatherPlan whose table pointers address every rank's shard and which pushes row gradients."""

    def setup_shards(self, arena, layout, rank, world, logical_emb_vocab, logical_lin_vocab):
        self.n_shards = world
        self.rank, self.world = rank, world
        self.arena, self.layout = arena, layout
        dev = self.device
        i32 = dict(dtype=torch.int32, device=dev)
        i64 = dict(dtype=torch.int64, device=dev)
        self.emb_vocab = torch.tensor(logical_emb_vocab, **i32)
        self.lin_vocab = torch.tensor(logical_lin_vocab, **i32)
        vocab_of = {}
        for col, v in list(zip(self.emb_cols.tolist(), logical_emb_vocab)) + list(zip(self.lin_cols.tolist(), logical_lin_vocab)):
            vocab_of[col] = max(vocab_of.get(col, 0), v)
        self.plan_vocab = torch.tensor([vocab_of[c] for c in self.plan_cols_host], **i32)
        self._emb_ptrs = torch.tensor([arena.peer_ptr[s] + layout.emb_off[f]
                                       for f in range(layout.n_emb) for s in range(world)], **i64)
        self._lin_ptrs = torch.tensor([arena.peer_ptr[s] + layout.lin_off[f]
                                       for f in range(layout.n_lin) for s in range(world)], **i64)
        self.recv_ptrs = []
        for par in range(2):
            r = layout.recv[par]
            self.recv_ptrs.append({k: torch.tensor([arena.peer_ptr[s] + r[k] for s in range(world)], **i64)
                                   for k in ("count", "ids", "emb", "lin")})
        self.step_parity = 0
        self._setup_exchange(arena, layout, rank, world)

    def _setup_exchange(self, arena, layout, rank, world):
        """Pointer tables of the forward row exchange; disabled (direct peer loads) when an id column feeds
        more than one embedding or linear slot, or with CTR_SHARD_EXCHANGE=0."""
        import os
        dev = self.device
        i64 = dict(dtype=torch.int64, device=dev)
        n_cols = len(self.plan_cols_host)
        emb_of, lin_of = [0] * n_cols, [0] * n_cols
        ok = os.environ.get("CTR_SHARD_EXCHANGE", "1") != "0" and self.D % 4 == 0
        for f, pc in enumerate(self.emb_plan_col.tolist()):
            ok = ok and emb_of[pc] == 0
            emb_of[pc] = arena.peer_ptr[rank] + layout.emb_off[f]
        for f, pc in enumerate(self.lin_plan_col.tolist()):
            ok = ok and lin_of[pc] == 0
            lin_of[pc] = arena.peer_ptr[rank] + layout.lin_off[f]
        self.exchange = bool(ok)
        if not self.exchange:
            return
        x, cap, D = layout.x, layout.xcap, max(layout.D, 1)
        me = arena.peer_ptr[rank]
        self.x_cap = cap
        self.x_emb_of_col = torch.tensor(emb_of, **i64)
        self.x_lin_of_col = torch.tensor(lin_of, **i64)
        self.x_cnt_to = torch.zeros(world, dtype=torch.int32, device=dev)
        self.x_inbox_req = torch.tensor([arena.peer_ptr[o] + x["req"] + rank * cap * 8 for o in range(world)], **i64)
        self.x_inbox_cnt = torch.tensor([arena.peer_ptr[o] + x["req_cnt"] for o in range(world)], **i64)
        self.x_req = me + x["req"]
        self.x_req_cnt = me + x["req_cnt"]
        self.x_resp_emb_remote = torch.tensor([arena.peer_ptr[q] + x["resp_emb"] + rank * cap * D * 4 for q in range(world)], **i64)
        self.x_resp_lin_remote = torch.tensor([arena.peer_ptr[q] + x["resp_lin"] + rank * cap * 4 for q in range(world)], **i64)
        self.x_resp_emb_local = torch.tensor([me + x["resp_emb"] + o * cap * D * 4 for o in range(world)], **i64)
        self.x_resp_lin_local = torch.tensor([me + x["resp_lin"] + o * cap * 4 for o in range(world)], **i64)
        self.x_token = torch.zeros(1, device=dev)
        self._x_where = {}
        self.bar_flags = torch.tensor([arena.peer_ptr[s] + layout.flags for s in range(world)], **i64)
        self.bar_epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.use_nccl_barrier = os.environ.get("CTR_SHARD_BARRIER", "p2p") == "nccl"

    def exchange_rows(self, X, B, group=None):
        """Requests -> barrier -> owners serve -> barrier; returns where[B, n_cols] for the gather."""
        import ctypes
        where = self._x_where.get(B)
        if where is None:
            where = torch.empty(B, len(self.plan_cols_host), dtype=torch.int32, device=self.device)
            self._x_where = {B: where}
        _lib.call("ctr_shard_request", ops._ptr(X), X.stride(0), B, len(self.plan_cols_host), ops._ptr(self.plan_cols),
                  ops._ptr(self.plan_vocab), self.world, self.rank, ops._ptr(self.x_cnt_to), ops._ptr(self.x_inbox_req),
                  ops._ptr(self.x_inbox_cnt), ops._ptr(where), self.x_cap, ops._ptr(self.err_flag), self.id_mode,
                  ops._stream())
        self.barrier(group)                                # every request list is complete
        _lib.call("ctr_shard_serve", self.world, self.rank, self.D, ctypes.c_void_p(self.x_req_cnt), ctypes.c_void_p(self.x_req),
                  self.x_cap, ops._ptr(self.x_emb_of_col), ops._ptr(self.x_lin_of_col), ops._ptr(self.x_resp_emb_remote),
                  ops._ptr(self.x_resp_lin_remote), ops._stream())
        self.barrier(group)                                # every response buffer is complete
        return where

    def barrier(self, group=None):
        """All ranks' work enqueued so far (incl. their stores into peer memory) precedes everything enqueued
        after it: a flag exchange over NVLink inside one tiny kernel (CTR_SHARD_BARRIER=nccl: an all-reduce)."""
        if self.use_nccl_barrier:
            dist.all_reduce(self.x_token, group=group)
        else:
            _lib.call("ctr_p2p_barrier", ops._ptr(self.bar_flags), ops._ptr(self.bar_epoch), self.world, self.rank,
                      ops._ptr(self.err_flag), ops._stream())

    def table_ptrs(self):
        return self._emb_ptrs, self._lin_ptrs

    def recv_views(self, parity):
        """This rank's receive lists: (count [n_fields], ids [n_fields, cap], emb [n_emb, cap, D], lin [n_lin, cap]).
        A list belongs to an id COLUMN of the plan (``plan_cols_host``): only the first ``len(plan_cols_host)``
        entries of count / ids are used, and every field reading column pc finds its rows in the slots of list pc."""
        L, r = self.layout, self.layout.recv[parity]
        nf = max(L.n_emb + L.n_lin, 1)
        return (self.arena.view(r["count"], (nf,), torch.int32),
                self.arena.view(r["ids"], (nf, L.cap), torch.int32),
                self.arena.view(r["emb"], (max(L.n_emb, 1), L.cap, L.D)),
                self.arena.view(r["lin"], (max(L.n_lin, 1), L.cap)))


def push_row_grads(plan, ws, rg_emb, rg_lin, B, n_emb):
    """Called from the fused-input backward in sharded mode."""
    par = plan.step_parity
    rp = plan.recv_ptrs[par]
    _lib.call("ctr_rowgrad_push", B, plan.world, len(plan.plan_cols_host), ops._ptr(ws["n_uniq"]), ops._ptr(ws["uniq"]),
              n_emb, plan.D, ops._ptr(rg_emb), B * max(plan.D, 1), ops._ptr(plan.emb_plan_col),
              plan.n_lin, ops._ptr(rg_lin), B, ops._ptr(plan.lin_plan_col),
              ops._ptr(rp["count"]), ops._ptr(rp["ids"]), ops._ptr(rp["emb"]), ops._ptr(rp["lin"]),
              plan.layout.cap, ops._ptr(plan.err_flag), ops._stream())


class ShardedRuntime:
    """Attached to a model as ``model.sharded``: dense-gradient all-reduce, owner-side combine of the
    received row gradients and step bookkeeping."""

    def __init__(self, model, plan, group=None):
        self.model, self.plan, self.group = model, plan, group
        table_ids = set(id(p) for p in plan.emb_params + plan.lin_params)
        self.dense_params = [p for p in model.parameters() if id(p) not in table_ids]
        self.last_done = None          # parity of the receive lists completed by the last finish_step()
        self._cws = None

    def finish_step(self):
        """All-reduce the replicated parameters' gradients (this also orders every rank's
        ``ctr_rowgrad_push`` before any owner reads its receive list), then flip the parity.  The
        completed lists stay readable (``received_row_grads`` / ``combine_received``) until
        ``clear_received`` — the fused optimizer consumes them in ``optim.step()``."""
        grads = [p.grad for p in self.dense_params if p.grad is not None]
        if grads:
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, group=self.group)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        else:
            dist.barrier(group=self.group)
        done = self.plan.step_parity
        self.plan.step_parity ^= 1
        self.last_done = done
        return done

    def received_row_grads(self, parity):
        """(counts, ids, emb_rows, lin_rows) delivered to this rank in the step of `parity`."""
        return self.plan.recv_views(parity)

    def clear_received(self, parity):
        self.plan.recv_views(parity)[0].zero_()

    def _combine_workspace(self):
        if self._cws is None:
            plan, L = self.plan, self.plan.layout
            n_plan = max(len(plan.plan_cols_host), 1)
            dev = plan.device
            ws = plan.alloc_workspace(L.cap, n_cols=n_plan)
            i32 = dict(dtype=torch.int32, device=dev)
            ws["cols"] = torch.arange(n_plan, **i32) * L.cap               # list pc's ids start at pc * cap
            # rows this rank holds of the table(s) fed by plan column pc
            rows_of = [1] * n_plan
            for f, pc in enumerate(plan.emb_plan_col_host):
                rows_of[pc] = max(rows_of[pc], L.emb_rows[f])
            for f, pc in enumerate(plan.lin_plan_col_host):
                rows_of[pc] = max(rows_of[pc], L.lin_rows[f])
            ws["vocab"] = torch.tensor(rows_of, **i32)
            ws["n_plan"] = n_plan
            ws["emb_pc"] = plan.emb_plan_col if L.n_emb else torch.zeros(1, **i32)
            ws["lin_pc"] = plan.lin_plan_col if L.n_lin else torch.zeros(1, **i32)
            ws["comb_emb"] = torch.empty(max(L.n_emb, 1), L.cap, max(L.D, 1), device=dev)
            ws["comb_lin"] = torch.empty(max(L.n_lin, 1), L.cap, device=dev)
            self._cws = ws
        return self._cws

    def combine_received(self, parity):
        """Owner side of the backward: up to G senders deliver a gradient for the same
        local row; build a duplicate-free plan over each field's receive list and sum the duplicates.
        Returns the workspace: ``n_uniq [nf]``, ``uniq [nf, cap]`` (local rows), ``comb_emb [n_emb, cap, D]``,
        ``comb_lin [n_lin, cap]`` — the same (uniq, rowgrad) contract as the single-GPU backward."""
        plan, L = self.plan, self.plan.layout
        ws = self._combine_workspace()
        counts, ids, emb, lin = plan.recv_views(parity)
        nf = ws["n_plan"]
        if L.n_emb + L.n_lin == 0:
            return ws
        _lib.call("ctr_unique_plan", ops._ptr(ids.view(torch.float32)), 1, L.cap, nf, ops._ptr(ws["cols"]),
                  ops._ptr(ws["vocab"]), ops._ptr(ws["keys"]), ops._ptr(ws["vals"]), ws["H"], ops._ptr(ws["n_uniq"]),
                  ops._ptr(ws["uniq"]), ops._ptr(ws["inv"]), ops._ptr(ws["cnt"]), ops._ptr(plan.err_flag), 1,
                  ops._ptr(counts), ops._stream())
        if L.n_emb:
            _lib.call("ctr_rowgrad_combine", L.cap, L.n_emb, L.D, ops._ptr(counts), ops._ptr(ws["n_uniq"]),
                      ops._ptr(ws["inv"]), nf, ops._ptr(ws["emb_pc"]), ops._ptr(emb), L.cap * L.D,
                      ops._ptr(ws["comb_emb"]), L.cap * L.D, ops._stream())
        if L.n_lin:
            _lib.call("ctr_rowgrad_combine", L.cap, L.n_lin, 1, ops._ptr(counts), ops._ptr(ws["n_uniq"]),
                      ops._ptr(ws["inv"]), nf, ops._ptr(ws["lin_pc"]), ops._ptr(lin), L.cap,
                      ops._ptr(ws["comb_lin"]), L.cap, ops._stream())
        return ws

    def apply_received(self, optimizer):
        """Called by ``RowwiseOptimizer.step``: combine the lists of the last finished step, update the
        local shards with the fused optimizer kernels, release the lists."""
        if self.last_done is None:
            return
        L = self.plan.layout
        ws = self.combine_received(self.last_done)
        optimizer.apply_rows(L.cap, ws["n_uniq"], ws["uniq"], ws["comb_emb"], ws["comb_lin"], L.n_emb,
                             emb_plan_col=ws["emb_pc"], lin_plan_col=ws["lin_pc"])
        self.clear_received(self.last_done)
        self.last_done = None


def localize_cfg(cfg, world):
    """Copy of an oracle-style cfg whose sparse vocabularies are the per-rank row counts.  The same
    column dict may be listed under both `linear_columns` and `dnn_columns` (deepcopy keeps that
    sharing): every distinct column is localised exactly once."""
    import copy
    local_cfg = copy.deepcopy(cfg)
    seen = set()
    for col in local_cfg["linear_columns"] + local_cfg["dnn_columns"]:
        if col["type"] == "sparse" and id(col) not in seen:
            seen.add(id(col))
            col["vocab"] = max_local_rows(col["vocab"], world)
    return local_cfg


def build_sharded(cfg, device, rank, world, batch=65536, group=None):
    """Build the model described by an oracle-style cfg with row-sharded tables on this rank."""
    from .config import model_from_cfg as build_model

    local_cfg = localize_cfg(cfg, world)
    model = build_model(local_cfg, device, table_grad="rowwise")
    attach_shards(model, cfg, rank, world, batch, group)
    return model, "dp%d: tower data-parallel (NCCL all-reduce), tables row-sharded (NVLink P2P gather/push)" % world


def attach_shards(model, logical_cfg, rank, world, batch, group=None):
    """Move the (local-size) tables of `model` into a peer-shared arena and switch its fused input
    to the sharded plan.  `logical_cfg` carries the full vocabulary sizes."""
    from .inputs import split_columns
    dev = next(model.parameters()).device
    sparse, dense, varlen = split_columns(model.dnn_feature_columns)
    lsparse, ldense, lvarlen = split_columns(model.linear_feature_columns)
    if varlen or lvarlen:
        raise NotImplementedError("VarLenSparseFeat tables are not sharded")
    logical = {c["name"]: c["vocab"] for c in logical_cfg["linear_columns"] + logical_cfg["dnn_columns"]
               if c["type"] == "sparse"}
    emb_vocab = [logical[c.name] for c in sparse]
    lin_vocab = [logical[c.name] for c in lsparse]
    D = sparse[0].embedding_dim if sparse else 1
    fidx = model.feature_index
    n_id_cols = len(set([fidx[c.name][0] for c in sparse] + [fidx[c.name][0] for c in lsparse]))
    layout = ArenaLayout(emb_vocab, lin_vocab, D, world, batch, n_id_cols=n_id_cols)
    arena = P2PArena(layout.nbytes, dev, group)
    with torch.no_grad():
        for f, c in enumerate(sparse):
            w = model.embedding_dict[c.embedding_name].weight
            view = arena.view(layout.emb_off[f], (layout.emb_rows[f], D))
            view.copy_(w.data[:layout.emb_rows[f]])
            w.data = view
        for f, c in enumerate(lsparse):
            w = model.linear_model.embedding_dict[c.embedding_name].weight
            view = arena.view(layout.lin_off[f], (layout.lin_rows[f], 1))
            view.copy_(w.data[:layout.lin_rows[f]])
            w.data = view
    base = model._gather_plan(dev)
    fi = model.feature_index
    emb_slots = [(model.embedding_dict[c.embedding_name].weight, fi[c.name][0], logical[c.name]) for c in sparse]
    lin_slots = [(model.linear_model.embedding_dict[c.embedding_name].weight, fi[c.name][0], logical[c.name])
                 for c in lsparse]
    plan = ShardedPlan(emb_slots, lin_slots, base.dense_cols.tolist(), base.lin_dense_cols.tolist(), D, dev,
                       id_mode=base.id_mode)
    plan.varlen, plan.lin_varlen, plan.n_sparse = [], [], len(sparse)
    plan.setup_shards(arena, layout, rank, world, emb_vocab, lin_vocab)
    model._plan = plan
    model.table_grad = "sharded"
    model.sharded = ShardedRuntime(model, plan, group)
    dist.barrier(group=group)       # every arena is mapped before the first remote access
    return model
