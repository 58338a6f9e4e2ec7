"""autograd wrappers over the C ABI (include/ctr_b200.h).

Every function here launches hand-written sm_100a kernels from libctr_b200.so on torch's
current CUDA stream; torch only owns the memory and the autograd graph.  There is no CPU or
stock-torch fallback: a non-CUDA tensor raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import functools
import os

import torch

from . import _lib

ACT_CODES = {"linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3}

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else None


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _first_cuda_device(args):
    for a in args:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
        if isinstance(a, (list, tuple)):
            d = _first_cuda_device(a)
            if d is not None:
                return d
    return None


def _on_device(fn):
    """Run an autograd forward/backward with the CUDA device of its tensor arguments current: the C ABI
    launches on cudaGetDevice()'s device and `_stream()` is that device's current stream, so a model built
    on cuda:1 (the reference API accepts any device string) must not launch on device 0."""
    @functools.wraps(fn)
    def wrapped(ctx, *args):
        dev = _first_cuda_device(args)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)
    return wrapped


def _require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("deepctr_torch_b200: %s must be a CUDA tensor — the hot path runs only on "
                           "the GPU kernels of libctr_b200.so (no CPU fallback)" % what)
    if t.dtype != torch.float32:
        raise TypeError("deepctr_torch_b200: %s must be float32, got %s" % (what, t.dtype))


def _rowmajor(t):
    """Return t if its last dim is unit-stride (2-D), else a contiguous copy."""
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


def _round4(n):
    return (n + 3) // 4 * 4


# ----------------------------------------------------------------------------------------------
# scratch of the tensor-core GEMM engine (csrc/gemm_pk.cu): one caller-owned buffer per device,
# grown to the largest requirement seen; the library only ever borrows it (ctr_set_scratch)
# ----------------------------------------------------------------------------------------------
_scratch_buf = {}
_scratch_need = {}
_scratch_retired = []      # outgrown buffers stay alive: a captured CUDA graph may still address them


def ensure_gemm_scratch(dev, B, K, N):
    """Make sure the registered scratch covers the three GEMMs of a [B,K] x [N,K]^T layer
    (forward, input gradient, weight gradient)."""
    key = (int(B), int(K), int(N))
    need = _scratch_need.get(key)
    if need is None:
        fn = _lib.load().ctr_gemm_scratch_bytes
        need = max(fn(B, N, K), fn(B, K, N), fn(N, K, B), fn(K, N, B))
        _scratch_need[key] = need
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    buf = _scratch_buf.get(idx)
    if buf is None or buf.numel() < need:
        if buf is not None:
            _scratch_retired.append(buf)
        buf = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=dev)
        _scratch_buf[idx] = buf
        with torch.cuda.device(idx):
            _lib.call("ctr_set_scratch", _ptr(buf), buf.numel())


def ensure_scratch_bytes(dev, need):
    """Grow the registered scratch to at least `need` bytes (packed weights of the tensor-core kernels)."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    buf = _scratch_buf.get(idx)
    if buf is None or buf.numel() < need:
        if buf is not None:
            _scratch_retired.append(buf)
        buf = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=dev)
        _scratch_buf[idx] = buf
        with torch.cuda.device(idx):
            _lib.call("ctr_set_scratch", _ptr(buf), buf.numel())


def act_code(name):
    if isinstance(name, str) and name.lower() in ACT_CODES:
        return ACT_CODES[name.lower()]
    raise NotImplementedError("activation %r is not implemented by the CUDA tower "
                              "(available: %s)" % (name, sorted(ACT_CODES)))


# ----------------------------------------------------------------------------------------------
# fused input: gather + linear + FM + dnn_input assembly
# ----------------------------------------------------------------------------------------------
class _WsLease:
    """One unique-plan workspace checked out of a GatherPlan's pool.  It returns to the pool when the last
    holder (the autograd node of the forward that built the plan, or the optimizer that still has to consume
    the row gradients) lets go of it — never while a backward may still read it."""
    __slots__ = ("plan", "ws")

    def __init__(self, plan, ws):
        self.plan, self.ws = plan, ws

    def __del__(self):
        try:
            self.plan._ws_free.setdefault(self.ws["B"], []).append(self.ws)
        except Exception:       # interpreter shutdown
            pass


class GatherPlan:
    """Device-side slot metadata for one model (built once, refreshed if storages move).

    emb / lin: lists of (parameter, X column, vocab, plan column) per field; dense_cols: X columns
    copied behind the embedding block; lin_dense_cols: X columns of Linear's dense weight.
    id_mode: how the id cells of X are encoded (include/ctr_b200.h: CTR_IDS_F32 / CTR_IDS_I32BITS)."""

    def __init__(self, emb_slots, lin_slots, dense_cols, lin_dense_cols, dim, device, id_mode=0):
        self.device = torch.device(device)
        self.id_mode = int(id_mode)
        self.emb_params = [s[0] for s in emb_slots]
        self.lin_params = [s[0] for s in lin_slots]
        self.n_emb, self.n_lin = len(emb_slots), len(lin_slots)
        self.D = int(dim) if self.n_emb else 0
        self.n_dense, self.n_lin_dense = len(dense_cols), len(lin_dense_cols)
        self.width = self.n_emb * self.D + self.n_dense
        self.ld = _round4(max(self.width, 1))
        i32 = dict(dtype=torch.int32, device=self.device)
        self.emb_cols = torch.tensor([s[1] for s in emb_slots], **i32)
        self.emb_vocab = torch.tensor([s[2] for s in emb_slots], **i32)
        self.lin_cols = torch.tensor([s[1] for s in lin_slots], **i32)
        self.lin_vocab = torch.tensor([s[2] for s in lin_slots], **i32)
        self.dense_cols = torch.tensor(list(dense_cols), **i32)
        self.lin_dense_cols = torch.tensor(list(lin_dense_cols), **i32)
        # the unique plan is per distinct id column of X
        plan_cols = []
        for s in list(emb_slots) + list(lin_slots):
            if s[1] not in plan_cols:
                plan_cols.append(s[1])
        self.plan_cols_host = plan_cols
        vocab_of = {}
        for s in list(emb_slots) + list(lin_slots):
            vocab_of[s[1]] = max(vocab_of.get(s[1], 0), s[2])
        self.plan_cols = torch.tensor(plan_cols, **i32)
        self.plan_vocab = torch.tensor([vocab_of[c] for c in plan_cols], **i32)
        self.emb_plan_col_host = [plan_cols.index(s[1]) for s in emb_slots]
        self.lin_plan_col_host = [plan_cols.index(s[1]) for s in lin_slots]
        self.emb_plan_col = torch.tensor(self.emb_plan_col_host, **i32)
        self.lin_plan_col = torch.tensor(self.lin_plan_col_host, **i32)
        self.err_flag = torch.zeros(1, **i32)
        self.n_shards = 1
        self._ptr_key = None
        self._emb_ptrs = self._lin_ptrs = None
        self._ws_free = {}
        # set by optim.RowwiseOptimizer: the backward parks (lease, rg_emb, rg_lin, n_emb) in `pending`
        # for the fused optimizer kernels instead of handing sparse COO gradients to autograd
        self.keep_rowgrads = False
        self.pending = []

    def table_ptrs(self):
        key = tuple(p.data_ptr() for p in self.emb_params) + tuple(p.data_ptr() for p in self.lin_params)
        if key != self._ptr_key:
            i64 = dict(dtype=torch.int64, device=self.device)
            self._emb_ptrs = torch.tensor(list(key[:self.n_emb]), **i64)
            self._lin_ptrs = torch.tensor(list(key[self.n_emb:]), **i64)
            self._ptr_key = key
        return self._emb_ptrs, self._lin_ptrs

    def alloc_workspace(self, B, n_cols=None):
        n = len(self.plan_cols_host) if n_cols is None else n_cols
        H = int(_lib.load().ctr_unique_plan_hash_slots(B))
        i32 = dict(dtype=torch.int32, device=self.device)
        return {"B": B, "H": H, "keys": torch.empty(n * H, **i32), "vals": torch.empty(n * H, **i32),
                "n_uniq": torch.empty(n, **i32), "uniq": torch.empty(n, B, **i32),
                "inv": torch.empty(B, n, **i32), "cnt": torch.empty(n, B, **i32)}

    def lease_workspace(self, B):
        """Buffers of one row-wise backward for batch size B, owned by the caller until the lease dies
        (two grad-enabled forwards before a backward therefore never share a plan)."""
        free = self._ws_free.get(B)
        ws = free.pop() if free else self.alloc_workspace(B)
        return _WsLease(self, ws)

    def side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def build_unique_plan(self, X, B, ws):
        _lib.call("ctr_unique_plan", _ptr(X), X.stride(0), B, len(self.plan_cols_host),
                  _ptr(self.plan_cols), _ptr(self.plan_vocab), _ptr(ws["keys"]), _ptr(ws["vals"]),
                  ws["H"], _ptr(ws["n_uniq"]), _ptr(ws["uniq"]), _ptr(ws["inv"]), _ptr(ws["cnt"]),
                  _ptr(self.err_flag), self.id_mode, None, _stream())

    def check_ids(self):
        """Raise IndexError (like nn.Embedding on CPU) if any id was out of range.  Synchronises."""
        flag = int(self.err_flag.item())
        if flag != 0:
            self.err_flag.zero_()
            if flag & 1:
                raise IndexError("index out of range in self (sparse feature id outside [0, vocabulary_size))")
            raise RuntimeError("deepctr_torch_b200: an exchange / receive list of the row-sharded tables overflowed "
                               "(err_flag=%d); raise the batch capacity passed to sharded.attach_shards" % flag)


def _dev_ptr_table(ptrs, dev):
    """Device array of pointers written by a kernel from by-value arguments (capturable in a CUDA graph)."""
    out = torch.empty(max(len(ptrs), 1), dtype=torch.int64, device=dev)
    arr = (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)
    _lib.call("ctr_write_ptrs", arr, len(ptrs), _ptr(out), _stream())
    return out


class _FusedInput(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, X, plan, lin_dense_w, want_blk, want_fm, grad_mode, *tables):
        _require_cuda(X, "X")
        X = _rowmajor(X)
        B = X.shape[0]
        dev = X.device
        emb_ptrs, lin_ptrs = plan.table_ptrs()
        blk = torch.empty(B, plan.ld, device=dev, dtype=torch.float32) if want_blk else None
        lin = torch.empty(B, device=dev, dtype=torch.float32)
        fm = torch.empty(B, device=dev, dtype=torch.float32) if want_fm else None
        n_emb = plan.n_emb if (want_blk or want_fm) else 0
        if plan.n_shards > 1 and getattr(plan, "exchange", False) and B > 0:
            # rows of other shards: requests -> owners gather locally -> contiguous delivery (sharded.py)
            where = plan.exchange_rows(X, B)
            _lib.call("ctr_gather_fwd_exchanged", _ptr(X), X.stride(0), B,
                      n_emb, plan.D, _ptr(emb_ptrs), _ptr(plan.emb_cols), _ptr(plan.emb_vocab),
                      plan.n_lin, _ptr(lin_ptrs), _ptr(plan.lin_cols), _ptr(plan.lin_vocab),
                      plan.n_dense if want_blk else 0, _ptr(plan.dense_cols),
                      plan.n_lin_dense if lin_dense_w is not None else 0, _ptr(plan.lin_dense_cols),
                      _ptr(lin_dense_w), _ptr(blk), plan.ld, _ptr(lin), _ptr(fm), _ptr(plan.err_flag),
                      plan.n_shards, plan.rank, _ptr(where), len(plan.plan_cols_host), _ptr(plan.emb_plan_col),
                      _ptr(plan.lin_plan_col), _ptr(plan.x_resp_emb_local), _ptr(plan.x_resp_lin_local),
                      plan.id_mode, _stream())
        else:
            _lib.call("ctr_gather_fwd", _ptr(X), X.stride(0), B,
                      n_emb, plan.D, _ptr(emb_ptrs), _ptr(plan.emb_cols), _ptr(plan.emb_vocab),
                      plan.n_lin, _ptr(lin_ptrs), _ptr(plan.lin_cols), _ptr(plan.lin_vocab),
                      plan.n_dense if want_blk else 0, _ptr(plan.dense_cols),
                      plan.n_lin_dense if lin_dense_w is not None else 0, _ptr(plan.lin_dense_cols),
                      _ptr(lin_dense_w), _ptr(blk), plan.ld, _ptr(lin), _ptr(fm), _ptr(plan.err_flag),
                      plan.n_shards, plan.id_mode, _stream())
        ctx.plan_event = None
        ctx.lease = None
        if grad_mode in ("rowwise", "sharded") and torch.is_grad_enabled() and B > 0 and plan.plan_cols_host:
            # the duplicate-free plan of the backward depends only on X: build it now on a side stream
            # so that it overlaps the forward tower instead of sitting on the backward's critical path.
            # The workspace belongs to THIS autograd node (a second forward must not overwrite it).
            ctx.lease = plan.lease_workspace(B)
            side = plan.side_stream()
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            X.record_stream(side)
            with torch.cuda.stream(side):
                plan.build_unique_plan(X, B, ctx.lease.ws)
                ctx.plan_event = torch.cuda.Event()
                ctx.plan_event.record(side)
        ctx.plan, ctx.grad_mode = plan, grad_mode
        ctx.want_blk, ctx.want_fm = want_blk, want_fm
        ctx.has_ldw = lin_dense_w is not None
        ctx.n_tables = len(tables)
        ctx.save_for_backward(X, blk if blk is not None else X.new_empty(0))
        outs = (blk if blk is not None else X.new_empty(0), lin, fm if fm is not None else X.new_empty(0))
        if blk is None:
            ctx.mark_non_differentiable(outs[0])
        if fm is None:
            ctx.mark_non_differentiable(outs[2])
        return outs

    @staticmethod
    @_on_device
    def backward(ctx, d_blk, d_lin, d_fm):
        plan = ctx.plan
        X, blk = ctx.saved_tensors
        B = X.shape[0]
        dev = X.device
        if blk.numel() == 0:
            blk = None
        if not ctx.want_blk:
            d_blk = None
        if not ctx.want_fm:
            d_fm = None
        if d_blk is not None:
            d_blk = _rowmajor(d_blk)
            if d_blk.stride(0) % 4 != 0 or d_blk.data_ptr() % 16 != 0:
                tmp = torch.empty(B, plan.ld, device=dev, dtype=torch.float32)
                tmp[:, :d_blk.shape[1]].copy_(d_blk)
                d_blk = tmp
        d_lin = d_lin.contiguous() if d_lin is not None else torch.zeros(B, device=dev)
        d_fm = d_fm.contiguous() if d_fm is not None else None
        emb_live = plan.n_emb > 0 and (d_blk is not None or d_fm is not None)
        n_emb = plan.n_emb if emb_live else 0
        grads_emb, grads_lin = [None] * plan.n_emb, [None] * plan.n_lin
        d_ldw = None
        if ctx.has_ldw:
            d_ldw = torch.empty(plan.n_lin_dense, 1, device=dev, dtype=torch.float32)
            _lib.call("ctr_lin_dense_wgrad", _ptr(X), X.stride(0), B, plan.n_lin_dense,
                      _ptr(plan.lin_dense_cols), _ptr(d_lin), _ptr(d_ldw), _stream())
        if B == 0:
            if ctx.grad_mode == "dense":
                grads_emb = [torch.zeros_like(p) for p in plan.emb_params]
                grads_lin = [torch.zeros_like(p) for p in plan.lin_params]
            return (None, None, d_ldw, None, None, None) + tuple(grads_emb) + tuple(grads_lin)
        if ctx.grad_mode == "dense":
            grads_emb = [torch.zeros_like(p) for p in plan.emb_params] if emb_live else [None] * plan.n_emb
            grads_lin = [torch.zeros_like(p) for p in plan.lin_params]
            eg = _dev_ptr_table([g.data_ptr() for g in grads_emb], dev) if emb_live else None
            lg = _dev_ptr_table([g.data_ptr() for g in grads_lin], dev)
            _lib.call("ctr_scatter_bwd_dense", _ptr(X), X.stride(0), B,
                      n_emb, plan.D, _ptr(eg), _ptr(plan.emb_cols), _ptr(plan.emb_vocab),
                      plan.n_lin, _ptr(lg), _ptr(plan.lin_cols), _ptr(plan.lin_vocab),
                      _ptr(blk), plan.ld, _ptr(d_blk), d_blk.stride(0) if d_blk is not None else 0,
                      _ptr(d_fm), _ptr(d_lin), plan.id_mode, _stream())
            return (None, None, d_ldw, None, None, None) + tuple(grads_emb) + tuple(grads_lin)
        # row-wise: (unique ids, summed row grads) per table
        if not plan.plan_cols_host:            # no sparse feature at all: nothing to scatter
            return (None, None, d_ldw, None, None, None) + tuple(grads_emb) + tuple(grads_lin)
        lease = ctx.lease
        n_plan = len(plan.plan_cols_host)
        if lease is not None:
            torch.cuda.current_stream(dev).wait_event(ctx.plan_event)
        else:                       # the forward ran without grad mode bookkeeping: build the plan here
            lease = plan.lease_workspace(B)
            plan.build_unique_plan(X, B, lease.ws)
        ws = lease.ws
        rg_emb = torch.empty(max(n_emb, 1), B, max(plan.D, 1), device=dev, dtype=torch.float32)
        rg_lin = torch.empty(max(plan.n_lin, 1), B, device=dev, dtype=torch.float32)
        _lib.call("ctr_scatter_bwd_rowwise", B, n_plan, _ptr(ws["inv"]), _ptr(ws["cnt"]),
                  _ptr(ws["n_uniq"]), n_emb, plan.D, _ptr(rg_emb), B * max(plan.D, 1),
                  _ptr(plan.emb_plan_col), plan.n_lin, _ptr(rg_lin), B, _ptr(plan.lin_plan_col),
                  _ptr(blk), plan.ld, _ptr(d_blk), d_blk.stride(0) if d_blk is not None else 0,
                  _ptr(d_fm), _ptr(d_lin), _stream())
        if ctx.grad_mode == "sharded":
            # deliver every unique (row, gradient) to the rank that owns the row (csrc/p2p.cu);
            # table gradients then live in the owners' receive lists, not in autograd
            from . import sharded
            sharded.push_row_grads(plan, ws, rg_emb, rg_lin, B, n_emb)
        elif plan.keep_rowgrads:
            # consumed in place by the fused row-wise optimizer (optim.RowwiseOptimizer.step)
            plan.pending.append((lease, rg_emb, rg_lin, n_emb))
        else:                       # hand autograd padded sparse COO gradients (any torch optimizer that takes them)
            uniq = ws["uniq"].to(torch.int64)
            for f, p in enumerate(plan.emb_params):
                if emb_live:
                    grads_emb[f] = torch.sparse_coo_tensor(uniq[plan.emb_plan_col_host[f]].unsqueeze(0), rg_emb[f],
                                                           size=p.shape, check_invariants=False)
            for f, p in enumerate(plan.lin_params):
                grads_lin[f] = torch.sparse_coo_tensor(uniq[plan.lin_plan_col_host[f]].unsqueeze(0),
                                                       rg_lin[f].unsqueeze(1), size=p.shape,
                                                       check_invariants=False)
        # a table shared by several fields appears several times in `tables`; autograd sums them
        return (None, None, d_ldw, None, None, None) + tuple(grads_emb) + tuple(grads_lin)


def fused_input(X, plan, lin_dense_w, want_blk=True, want_fm=False, grad_mode="dense"):
    """(blk [B, ld] | None, lin [B], fm [B] | None) — see ctr_gather_fwd in include/ctr_b200.h."""
    tables = tuple(plan.emb_params) + tuple(plan.lin_params)
    blk, lin, fm = _FusedInput.apply(X, plan, lin_dense_w, want_blk, want_fm, grad_mode, *tables)
    return (blk if want_blk else None), lin, (fm if want_fm else None)


# ----------------------------------------------------------------------------------------------
# FM on an assembled block (generic composition; the fused gather computes FM itself)
# ----------------------------------------------------------------------------------------------
class _FM(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, E):
        _require_cuda(E, "FM input")
        B, F, D = E.shape
        E2 = E.reshape(B, F * D)
        E2 = _rowmajor(E2)
        out = torch.empty(B, device=E.device, dtype=torch.float32)
        _lib.call("ctr_fm_fwd", _ptr(E2), E2.stride(0), B, F, D, _ptr(out), _stream())
        ctx.save_for_backward(E2)
        ctx.shape = (B, F, D)
        return out.unsqueeze(1)

    @staticmethod
    @_on_device
    def backward(ctx, g):
        (E2,) = ctx.saved_tensors
        B, F, D = ctx.shape
        dE = torch.zeros(B, F * D, device=E2.device, dtype=torch.float32)
        gg = g.reshape(B).contiguous()
        _lib.call("ctr_fm_bwd", _ptr(E2), E2.stride(0), B, F, D, _ptr(gg), _ptr(dE), F * D, _stream())
        return dE.view(B, F, D)


def fm(E):
    return _FM.apply(E)


# ----------------------------------------------------------------------------------------------
# dense tower
# ----------------------------------------------------------------------------------------------
class _DnnLayer(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, x, W, bias, act, w_kn):
        """w_kn: W is stored [K, N] (CrossNetMix factors) instead of nn.Linear's [N, K]."""
        _require_cuda(x, "DNN input")
        x = _rowmajor(x)
        B, K = x.shape
        Wc = W if W.is_contiguous() else W.contiguous()
        N = Wc.shape[1] if w_kn else Wc.shape[0]
        Kw = Wc.shape[0] if w_kn else Wc.shape[1]
        if not w_kn and Kw % 4 != 0:
            # rows of an [N, K] weight with K % 4 != 0 are not 16-byte aligned: stage a zero-padded
            # copy ([N, round_up(K,4)], a few hundred KB) so the GEMM producers can use 128-bit loads
            Wc = torch.nn.functional.pad(Wc, (0, _round4(Kw) - Kw))
        if K != Kw:
            # zero-padded input block [B, round_up(K,4)] (the fused gather writes zeros behind the
            # dense columns): contract over the padded width against the zero-padded weight, so
            # neither the forward nor the backward needs a strided slice of the 113 MB block
            if w_kn or K != _round4(Kw):
                raise ValueError("dnn_layer: input width %d does not match weight width %d" % (K, Kw))
        swn, swk = (1, N) if w_kn else (Wc.shape[1], 1)
        y = torch.empty(B, N, device=x.device, dtype=torch.float32)
        ensure_gemm_scratch(x.device, B, K, N)
        _lib.call("ctr_dnn_layer_fwd", _ptr(x), x.stride(0), _ptr(Wc), swn, swk, _ptr(bias),
                  _ptr(y), N, B, K, N, act, _stream())
        ctx.act, ctx.w_kn, ctx.has_bias = act, w_kn, bias is not None
        ctx.wshape = tuple(W.shape)
        ctx.save_for_backward(x, Wc, y)
        return y

    @staticmethod
    @_on_device
    def backward(ctx, dy):
        x, Wc, y = ctx.saved_tensors
        B, K = x.shape
        N = y.shape[1]
        dy = _rowmajor(dy)
        swn, swk = (1, N) if ctx.w_kn else (Wc.shape[1], 1)
        Kw = ctx.wshape[0] if ctx.w_kn else ctx.wshape[1]
        padded_in = (K != Kw)
        sdwn, sdwk = (1, N) if ctx.w_kn else (K, 1)
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ldx = _round4(K)
        dx_full = torch.empty(B, ldx, device=x.device, dtype=torch.float32) if need_dx else None
        dW = (torch.empty((N, K) if padded_in else ctx.wshape, device=x.device, dtype=torch.float32)
              if need_dw else None)
        db = torch.empty(N, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        _lib.call("ctr_dnn_layer_bwd", _ptr(x), x.stride(0), _ptr(Wc), swn, swk, _ptr(y), N,
                  _ptr(dy), dy.stride(0), _ptr(dx_full), ldx, 0, _ptr(dW), sdwn, sdwk, _ptr(db),
                  B, K, N, ctx.act, _stream())
        dx = dx_full[:, :K] if need_dx else None
        if padded_in and dW is not None:
            dW = dW[:, :Kw].contiguous()
        return dx, dW, db, None, None


def dnn_layer(x, W, bias, act="relu", w_kn=False):
    return _DnnLayer.apply(x, W, bias, act_code(act) if isinstance(act, str) else int(act), w_kn)


# ---- deferred first-layer weight gradient -------------------------------------------------------------------
# The first layer's (dW, db) is the last GEMM of the tower backward and nothing in the backward pass consumes it,
# while its input gradient heads the longest remaining chain (embedding scatter, and on several GPUs the row-gradient
# push over NVLink).  When the weight-gradient engine needs no shared scratch it is therefore issued on a second
# stream; the calling stream joins it in an autograd-engine callback at the end of the backward pass (before an
# optimizer or the dense all-reduce can look at the gradients).  CTR_DEFER_WGRAD=0 keeps everything on one stream.
_DEFER_WGRAD = os.environ.get("CTR_DEFER_WGRAD", "1") != "0"
_wgrad_streams = {}
_wgrad_joins = []


def _wgrad_stream(dev):
    s = _wgrad_streams.get(dev)
    if s is None:
        s = _wgrad_streams[dev] = torch.cuda.Stream(device=dev)
    return s


def _join_deferred_wgrads():
    while _wgrad_joins:
        dev, ev = _wgrad_joins.pop()
        torch.cuda.current_stream(dev).wait_event(ev)


class _DnnTower(torch.autograd.Function):
    """A stack of Linear(+bias) -> activation layers as ONE autograd node (reference core.py:120-134
    without batch-norm / dropout).  Forward = one fused GEMM+bias+activation launch per layer; the
    backward chains ctr_dnn_layer_bwd_chain so that every layer writes the dZ of the layer below
    directly (input gradient times act'(input) in the GEMM epilogue): only the top layer reads an
    activation mask, and no dY (.) act'(Y) product is ever re-derived from two tensors."""

    @staticmethod
    @_on_device
    def forward(ctx, x, act, *params):
        _require_cuda(x, "DNN input")
        x = _rowmajor(x)
        L = len(params) // 2
        xs, Ws, ys, kws = [], [], [], []
        cur = x
        for l in range(L):
            W, b = params[2 * l], params[2 * l + 1]
            B, K = cur.shape
            Wc = W if W.is_contiguous() else W.contiguous()
            N, Kw = Wc.shape
            if Kw % 4 != 0 and K != Kw:
                # the input carries zero columns up to a multiple of 4: the weight rows must match.  (K == Kw: the
                # unpadded weight is used as it is — the pack kernels take any row stride — no pad kernels per step.)
                Wc = torch.nn.functional.pad(Wc, (0, _round4(Kw) - Kw))
            if K != Kw and K != _round4(Kw):
                raise ValueError("dnn_tower: input width %d does not match weight width %d" % (K, Kw))
            y = torch.empty(B, N, device=x.device, dtype=torch.float32)
            ensure_gemm_scratch(x.device, B, K, N)
            _lib.call("ctr_dnn_layer_fwd", _ptr(cur), cur.stride(0), _ptr(Wc), Wc.shape[1], 1, _ptr(b),
                      _ptr(y), N, B, K, N, act, _stream())
            xs.append(cur)
            Ws.append(Wc)
            ys.append(y)
            kws.append(Kw)
            cur = y
        ctx.act, ctx.L, ctx.kws = act, L, kws
        ctx.w0 = params[0]                 # the leaf itself: a pending .grad forbids the deferred weight gradient
        ctx.wshapes = [tuple(params[2 * l].shape) for l in range(L)]
        ctx.save_for_backward(*xs, *Ws, *ys)
        return cur

    @staticmethod
    @_on_device
    def backward(ctx, dy):
        L = ctx.L
        saved = ctx.saved_tensors
        xs, Ws, ys = saved[:L], saved[L:2 * L], saved[2 * L:3 * L]
        dev = dy.device
        cur = _rowmajor(dy)
        grads = [None] * (2 * L)
        dx_out = None
        for l in range(L - 1, -1, -1):
            x, Wc, y = xs[l], Ws[l], ys[l]
            B, K = x.shape
            N = y.shape[1]
            Kc = Wc.shape[1]                      # contraction width actually used (padded or not)
            need_dx = l > 0 or ctx.needs_input_grad[0]
            ldx = _round4(K)
            dx_full = torch.empty(B, ldx, device=dev, dtype=torch.float32) if need_dx else None
            Kd = _round4(Kc)                       # dW rows padded to 16 bytes (vector reductions of the split-K partials)
            dW = torch.empty(N, Kd, device=dev, dtype=torch.float32)
            db = torch.empty(N, device=dev, dtype=torch.float32)
            dy_is_dz = 1 if l < L - 1 else 0
            defer = (l == 0 and need_dx and _DEFER_WGRAD and B > 0 and getattr(ctx.w0, "grad", None) is None
                     and _lib.load().ctr_dnn_wgrad_is_scratch_free(
                         _ptr(x), x.stride(0), _ptr(y), N, _ptr(cur), cur.stride(0), _ptr(dW), Kd, 1, B, K, N,
                         ctx.act, dy_is_dz) == 1)
            if defer:
                main = torch.cuda.current_stream(dev)
                _lib.call("ctr_dnn_layer_bwd_chain", _ptr(x), x.stride(0), _ptr(Wc), Kc, 1, _ptr(y), N,
                          _ptr(cur), cur.stride(0), _ptr(dx_full), ldx, None, Kc, 1, None,
                          B, K, N, ctx.act, dy_is_dz, 0, _stream())
                side = _wgrad_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    _lib.call("ctr_dnn_layer_bwd_chain", _ptr(x), x.stride(0), _ptr(Wc), Kc, 1, _ptr(y), N,
                              _ptr(cur), cur.stride(0), None, 0, _ptr(dW), Kd, 1, _ptr(db),
                              B, K, N, ctx.act, dy_is_dz, 0, _stream())
                    grads[0] = dW[:, :ctx.kws[0]].contiguous() if Kd != ctx.kws[0] else dW
                    ev = torch.cuda.Event()
                    ev.record(side)
                for t in (x, y, cur, dW, db, grads[0]):
                    t.record_stream(side)
                _wgrad_joins.append((dev, ev))
                torch.autograd.Variable._execution_engine.queue_callback(_join_deferred_wgrads)
                grads[1] = db
            else:
                _lib.call("ctr_dnn_layer_bwd_chain", _ptr(x), x.stride(0), _ptr(Wc), Kc, 1, _ptr(y), N,
                          _ptr(cur), cur.stride(0), _ptr(dx_full), ldx, _ptr(dW), Kd, 1, _ptr(db),
                          B, K, N, ctx.act, dy_is_dz,
                          ctx.act if l > 0 else 0, _stream())
                grads[2 * l] = dW[:, :ctx.kws[l]].contiguous() if Kd != ctx.kws[l] else dW
                grads[2 * l + 1] = db
            if need_dx:
                cur = dx_full[:, :K] if ldx != K else dx_full
                dx_out = cur
        return (dx_out if ctx.needs_input_grad[0] else None, None) + tuple(grads)


def dnn_tower(x, act, weights, biases):
    """act(...act(x W0^T + b0)... W_{L-1}^T + b_{L-1}); every layer needs a bias (nn.Linear default)."""
    params = []
    for W, b in zip(weights, biases):
        params += [W, b]
    return _DnnTower.apply(x, act_code(act) if isinstance(act, str) else int(act), *params)


class _RowDot(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, H, w):
        _require_cuda(H, "row-dot input")
        H = _rowmajor(H)
        B, N = H.shape
        wv = w.reshape(-1).contiguous()
        out = torch.empty(B, device=H.device, dtype=torch.float32)
        _lib.call("ctr_rowdot_fwd", _ptr(H), H.stride(0), _ptr(wv), B, N, _ptr(out), 0, _stream())
        ctx.save_for_backward(H, wv)
        ctx.wshape = w.shape
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        H, wv = ctx.saved_tensors
        B, N = H.shape
        g = g.contiguous()
        ld = _round4(N)
        dH = torch.empty(B, ld, device=H.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        dw = torch.empty(N, device=H.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        _lib.call("ctr_rowdot_bwd", _ptr(H), H.stride(0), _ptr(wv), _ptr(g), B, N, _ptr(dH), ld, 0,
                  _ptr(dw), _stream())
        return (dH[:, :N] if dH is not None else None), (dw.view(ctx.wshape) if dw is not None else None)


def rowdot(H, w):
    """[B] = H[B,N] @ w (the bias-free 1-unit Linear heads)."""
    return _RowDot.apply(H, w)


class _Predict(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, bias, binary, *terms):
        terms = [t.reshape(-1).contiguous() for t in terms]
        _require_cuda(terms[0], "logit term")
        B = terms[0].shape[0]
        y = torch.empty(B, device=terms[0].device, dtype=torch.float32)
        arr = (ctypes.c_void_p * len(terms))(*[t.data_ptr() for t in terms])
        _lib.call("ctr_predict_fwd", arr, len(terms), _ptr(bias), B, 1 if binary else 0, None,
                  _ptr(y), _stream())
        ctx.binary, ctx.n_terms, ctx.has_bias = binary, len(terms), bias is not None
        ctx.save_for_backward(y)
        return y.unsqueeze(1)

    @staticmethod
    @_on_device
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        B = y.shape[0]
        dy = dy.reshape(-1).contiguous()
        dlogit = torch.empty(B, device=y.device, dtype=torch.float32)
        dbias = torch.empty(1, device=y.device, dtype=torch.float32) if ctx.has_bias else None
        _lib.call("ctr_predict_bwd", _ptr(y), _ptr(dy), B, 1 if ctx.binary else 0, _ptr(dlogit),
                  _ptr(dbias), _stream())
        return (dbias, None) + (dlogit,) * ctx.n_terms


def predict(terms, bias, binary=True):
    """y[B,1] = sigmoid(sum(terms) + bias) — PredictionLayer (reference layers/core.py:154-160)."""
    if len(terms) > 8:
        raise ValueError("at most 8 logit terms")
    return _Predict.apply(bias, binary, *terms)


# ----------------------------------------------------------------------------------------------
# CIN
# ----------------------------------------------------------------------------------------------
class _CIN(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, E, act, split_half, layer_size, *params):
        """E [B,M,D] (any batch stride); params = (W0, b0, W1, b1, ...) with W [N, H*M, 1]."""
        _require_cuda(E, "CIN input")
        B, M, D = E.shape
        if E.stride(2) != 1 or E.stride(1) != D:
            E = E.contiguous()
        dev = E.device
        n_layers = len(layer_size)
        plan = []
        H = M
        total = 0
        for k, N in enumerate(layer_size):
            last = k == n_layers - 1
            if split_half:
                n_hidden, dstart = (0, 0) if last else (N // 2, N // 2)
            else:
                n_hidden, dstart = N, 0
            plan.append((H, N, n_hidden, dstart, total))
            total += N - dstart
            H = n_hidden
        out = torch.empty(B, total, device=dev, dtype=torch.float32)
        Ys = []
        xp, sxp = E, E.stride(0)
        for k, (Hk, N, n_hidden, dstart, off) in enumerate(plan):
            W = params[2 * k].reshape(N, Hk * M)
            W = W if W.is_contiguous() else W.contiguous()
            b = params[2 * k + 1]
            Y = torch.empty(B, N, D, device=dev, dtype=torch.float32)
            # packed (hi, lo) copy of W for the TMA-fed forward: ceil(N/128) x 2*ceil(HM/32) tiles of 16 KB
            ensure_scratch_bytes(dev, ((N + 127) // 128) * 2 * ((Hk * M + 31) // 32) * 16384 + 256)
            _lib.call("ctr_cin_layer_fwd", _ptr(xp), sxp, Hk, _ptr(E), E.stride(0), M, D, _ptr(W),
                      _ptr(b), N, dstart, act, _ptr(Y), _vp(out.data_ptr() + 4 * off), total, B,
                      _stream())
            Ys.append(Y)
            xp, sxp = Y, N * D
        ctx.plan, ctx.act, ctx.dims = plan, act, (B, M, D, total)
        ctx.save_for_backward(E, *Ys, *params)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        B, M, D, total = ctx.dims
        plan = ctx.plan
        L = len(plan)
        saved = ctx.saved_tensors
        E, Ys, params = saved[0], saved[1:1 + L], saved[1 + L:]
        dev = E.device
        dout = _rowmajor(dout)
        dE = torch.zeros(B, M, D, device=dev, dtype=torch.float32)
        grads = [None] * (2 * L)
        dYh, sdyh = None, 0
        for k in range(L - 1, -1, -1):
            Hk, N, n_hidden, dstart, off = plan[k]
            W = params[2 * k].reshape(N, Hk * M)
            W = W if W.is_contiguous() else W.contiguous()
            first = k == 0
            xp = E if first else Ys[k - 1]
            sxp = E.stride(0) if first else plan[k - 1][1] * D
            dZ = torch.empty(B, N, D, device=dev, dtype=torch.float32)
            dW = torch.empty(N, Hk * M, device=dev, dtype=torch.float32)
            db = torch.empty(N, device=dev, dtype=torch.float32)
            dXp = None if first else torch.empty(B, Hk, D, device=dev, dtype=torch.float32)
            _lib.call("ctr_cin_layer_bwd", _ptr(xp), sxp, Hk, _ptr(E), E.stride(0), M, D, _ptr(W), N,
                      n_hidden if k < L - 1 else 0, dstart, ctx.act, _ptr(Ys[k]), _ptr(dYh), sdyh,
                      _vp(dout.data_ptr() + 4 * off), dout.stride(0), _ptr(dZ), _ptr(dW), _ptr(db),
                      _ptr(dXp), Hk * D, _ptr(dE), M * D, B, _stream())
            grads[2 * k] = dW.view(params[2 * k].shape)
            grads[2 * k + 1] = db
            dYh, sdyh = dXp, Hk * D
        return (dE, None, None, None) + tuple(grads)


def cin(E, layer_size, split_half, act, params):
    return _CIN.apply(E, act_code(act), bool(split_half), tuple(int(s) for s in layer_size), *params)


# ----------------------------------------------------------------------------------------------
# CrossNet
# ----------------------------------------------------------------------------------------------
class _CrossVector(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, x, kernels, bias):
        _require_cuda(x, "CrossNet input")
        x = _rowmajor(x)
        B, n = x.shape
        L = kernels.shape[0]
        kv = kernels.reshape(L, n).contiguous()
        bv = bias.reshape(L, n).contiguous()
        ldo = _round4(n)                  # 16-byte aligned rows: the register kernels move 128 bits per lane
        out_full = torch.empty(B, ldo, device=x.device, dtype=torch.float32)
        s = torch.empty(B, max(L, 1), device=x.device, dtype=torch.float32)
        _lib.call("ctr_cross_vector_fwd", _ptr(x), x.stride(0), _ptr(kv), _ptr(bv), L, n, _ptr(out_full), ldo,
                  _ptr(s), B, _stream())
        ctx.save_for_backward(x, kv, bv, s)
        ctx.shapes = (kernels.shape, bias.shape)
        return out_full[:, :n] if ldo != n else out_full

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        x, kv, bv, s = ctx.saved_tensors
        B, n = x.shape
        L = kv.shape[0]
        dout = _rowmajor(dout)
        ld = _round4(n)
        dx = torch.empty(B, ld, device=x.device, dtype=torch.float32)
        dk = torch.empty(L, n, device=x.device, dtype=torch.float32)
        db = torch.empty(L, n, device=x.device, dtype=torch.float32)
        _lib.call("ctr_cross_vector_bwd", _ptr(x), x.stride(0), _ptr(kv), _ptr(bv), L, n, _ptr(s),
                  _ptr(dout), dout.stride(0), _ptr(dx), ld, 0, _ptr(dk), _ptr(db), B, _stream())
        return dx[:, :n], dk.view(ctx.shapes[0]), db.view(ctx.shapes[1])


class _CrossMatrix(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, x, kernels, bias):
        _require_cuda(x, "CrossNet input")
        x = _rowmajor(x)
        B, n = x.shape
        L = kernels.shape[0]
        W = kernels.contiguous()
        bv = bias.reshape(L, n).contiguous()
        xs, Us = [x], []
        ensure_gemm_scratch(x.device, B, n, n)
        for l in range(L):
            U = torch.empty(B, n, device=x.device, dtype=torch.float32)
            nxt = torch.empty(B, n, device=x.device, dtype=torch.float32)
            _lib.call("ctr_cross_matrix_layer_fwd", _ptr(x), x.stride(0), _ptr(xs[-1]), xs[-1].stride(0),
                      _ptr(W[l]), _ptr(bv[l]), n, _ptr(U), n, _ptr(nxt), n, B, _stream())
            xs.append(nxt)
            Us.append(U)
        ctx.L = L
        ctx.bshape = bias.shape
        ctx.save_for_backward(W, *xs[:-1], *Us)
        return xs[-1] if L > 0 else x.clone()

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        L = ctx.L
        saved = ctx.saved_tensors
        W, xs, Us = saved[0], saved[1:1 + L], saved[1 + L:]
        x0 = xs[0] if L else None
        if L == 0:
            return dout, torch.zeros_like(W), None
        B, n = x0.shape
        dev = x0.device
        g = _rowmajor(dout)
        dx0 = torch.zeros(B, n, device=dev, dtype=torch.float32)
        dW = torch.empty(L, n, n, device=dev, dtype=torch.float32)
        db = torch.empty(L, n, device=dev, dtype=torch.float32)
        dU = torch.empty(B, n, device=dev, dtype=torch.float32)
        for l in range(L - 1, -1, -1):
            gprev = torch.empty(B, n, device=dev, dtype=torch.float32)
            _lib.call("ctr_cross_matrix_layer_bwd", _ptr(x0), x0.stride(0), _ptr(xs[l]), xs[l].stride(0),
                      _ptr(W[l]), _ptr(Us[l]), n, _ptr(g), g.stride(0), n, _ptr(dU), n, _ptr(dx0), n,
                      _ptr(dW[l]), _ptr(db[l]), _ptr(gprev), n, B, _stream())
            g = gprev
        dx0 += g                      # x_0 is also the x_l of layer 0
        return dx0, dW, db.view(ctx.bshape)


def crossnet(x, kernels, bias, parameterization="vector"):
    if parameterization == "vector":
        return _CrossVector.apply(x, kernels, bias)
    if parameterization == "matrix":
        return _CrossMatrix.apply(x, kernels, bias)
    raise ValueError("parameterization should be 'vector' or 'matrix'")


class _CrossMixCombine(torch.autograd.Function):
    """x_{l+1} = sum_e softmax(gate)_e * x0 (.) (uv_e + bias) + x_l  (one layer)."""

    @staticmethod
    @_on_device
    def forward(ctx, x0, xl, uv, gate, bias):
        E, B, n = uv.shape
        x0, xl = _rowmajor(x0), _rowmajor(xl)
        uv, gate = uv.contiguous(), gate.contiguous()
        bv = bias.reshape(n).contiguous()
        out = torch.empty(B, n, device=x0.device, dtype=torch.float32)
        _lib.call("ctr_cross_mix_fwd", _ptr(x0), x0.stride(0), _ptr(xl), xl.stride(0), _ptr(uv), n,
                  _ptr(gate), _ptr(bv), E, n, _ptr(out), n, B, _stream())
        ctx.save_for_backward(x0, uv, gate, bv)
        ctx.bshape = bias.shape
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        x0, uv, gate, bv = ctx.saved_tensors
        E, B, n = uv.shape
        g = _rowmajor(g)
        duv = torch.empty_like(uv)
        dgate = torch.empty_like(gate)
        dx0 = torch.zeros(B, n, device=x0.device, dtype=torch.float32)
        dbias = torch.empty(n, device=x0.device, dtype=torch.float32)
        _lib.call("ctr_cross_mix_bwd", _ptr(x0), x0.stride(0), _ptr(uv), n, _ptr(gate), _ptr(bv), _ptr(g),
                  g.stride(0), E, n, _ptr(duv), _ptr(dgate), _ptr(dx0), n, _ptr(dbias), B, _stream())
        return dx0, g, duv, dgate, dbias.view(ctx.bshape)


def cross_mix_combine(x0, xl, uv, gate, bias):
    return _CrossMixCombine.apply(x0, xl, uv, gate, bias)


# ----------------------------------------------------------------------------------------------
# FiBiNET pieces
# ----------------------------------------------------------------------------------------------
class _SENET(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, E, W1, W2):
        _require_cuda(E, "SENET input")
        B, F, D = E.shape
        if E.stride(2) != 1 or E.stride(1) != D:
            E = E.contiguous()
        W1c, W2c = W1.contiguous(), W2.contiguous()
        R = W1c.shape[0]
        V = torch.empty(B, F, D, device=E.device, dtype=torch.float32)
        _lib.call("ctr_senet_fwd", _ptr(E), E.stride(0), F, D, _ptr(W1c), _ptr(W2c), R, _ptr(V), F * D, B,
                  _stream())
        ctx.save_for_backward(E, W1c, W2c)
        return V

    @staticmethod
    @_on_device
    def backward(ctx, dV):
        E, W1c, W2c = ctx.saved_tensors
        B, F, D = E.shape
        R = W1c.shape[0]
        dV = dV.contiguous()
        dE = torch.empty(B, F, D, device=E.device, dtype=torch.float32)
        dW1, dW2 = torch.empty_like(W1c), torch.empty_like(W2c)
        _lib.call("ctr_senet_bwd", _ptr(E), E.stride(0), F, D, _ptr(W1c), _ptr(W2c), R, _ptr(dV), F * D,
                  _ptr(dE), F * D, 0, _ptr(dW1), _ptr(dW2), B, _stream())
        return dE, dW1, dW2


def senet(E, W1, W2):
    return _SENET.apply(E, W1, W2)


BILINEAR_SEL = {"all": 0, "each": 1, "interaction": 2}


class _Bilinear(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, E, W, wsel):
        """E [B,F,D]; W [n_w, D, D] stacked weights; returns [B, P, D]."""
        _require_cuda(E, "bilinear input")
        B, F, D = E.shape
        if E.stride(2) != 1 or E.stride(1) != D:
            E = E.contiguous()
        Wc = W.contiguous()
        P = F * (F - 1) // 2
        out = torch.empty(B, P, D, device=E.device, dtype=torch.float32)
        ensure_scratch_bytes(E.device, 8 << 20)      # packed weight blocks of the GEMM formulation
        _lib.call("ctr_bilinear_fwd", _ptr(E), E.stride(0), F, D, _ptr(Wc), wsel, _ptr(out), P * D, B,
                  _stream())
        ctx.wsel = wsel
        ctx.save_for_backward(E, Wc)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        E, Wc = ctx.saved_tensors
        B, F, D = E.shape
        P = F * (F - 1) // 2
        dout = dout.contiguous()
        dE = torch.zeros(B, F, D, device=E.device, dtype=torch.float32)
        dW = torch.empty_like(Wc)
        _lib.call("ctr_bilinear_bwd", _ptr(E), E.stride(0), F, D, _ptr(Wc), ctx.wsel, _ptr(dout), P * D,
                  _ptr(dE), F * D, _ptr(dW), B, _stream())
        return dE, dW, None


class _FibinetInput(torch.autograd.Function):
    """FiBiNET's DNN input [bilinear(SENET(E)) | bilinear(E) | dense] (reference models/fibinet.py:82-87) written
    ONCE, in place: both bilinear passes store straight into their column ranges of the padded [B, ld] buffer the
    tower consumes (batch stride `so` of ctr_bilinear_fwd), the backward reads the tower's input gradient through
    strided views — no [B,650,D] intermediates, no 2.7 GB torch.cat at BASELINE config #4."""

    @staticmethod
    @_on_device
    def forward(ctx, E_se, E, W, wsel, dense):
        _require_cuda(E, "bilinear input")
        B, F, D = E.shape
        E, E_se = _block3(E), _block3(E_se)
        Wc = W.contiguous()
        P = F * (F - 1) // 2
        nd = dense.shape[1] if dense is not None else 0
        width = 2 * P * D + nd
        ld = _round4(width)
        buf = torch.empty(B, ld, device=E.device, dtype=torch.float32)
        if ld > 2 * P * D:
            tail = buf[:, 2 * P * D:]
            if nd:
                tail[:, :nd].copy_(dense)
            if ld > width:
                tail[:, nd:].zero_()
        ensure_scratch_bytes(E.device, 8 << 20)      # packed weight blocks of the GEMM formulation
        _lib.call("ctr_bilinear_fwd", _ptr(E_se), E_se.stride(0), F, D, _ptr(Wc), wsel, _ptr(buf), ld, B, _stream())
        _lib.call("ctr_bilinear_fwd", _ptr(E), E.stride(0), F, D, _ptr(Wc), wsel,
                  _vp(buf.data_ptr() + 4 * P * D), ld, B, _stream())
        ctx.wsel, ctx.dims = wsel, (B, F, D, P, width)
        ctx.save_for_backward(E_se, E, Wc)
        return buf

    @staticmethod
    @_on_device
    def backward(ctx, dbuf):
        E_se, E, Wc = ctx.saved_tensors
        B, F, D, P, width = ctx.dims
        dbuf = _rowmajor(dbuf)
        sdo = dbuf.stride(0)
        outs = []
        for k, src in enumerate((E_se, E)):
            dE = torch.zeros(B, F, D, device=E.device, dtype=torch.float32)
            dW = torch.empty_like(Wc)
            _lib.call("ctr_bilinear_bwd", _ptr(src), src.stride(0), F, D, _ptr(Wc), ctx.wsel,
                      _vp(dbuf.data_ptr() + 4 * k * P * D), sdo, _ptr(dE), F * D, _ptr(dW), B, _stream())
            outs.append((dE, dW))
        return outs[0][0], outs[1][0], outs[0][1] + outs[1][1], None, None


def fibinet_dnn_input(E_se, E, W, bilinear_type, dense=None):
    """Padded [B, round_up(2*P*D + n_dense, 4)] DNN input of FiBiNET (SENET branch first)."""
    return _FibinetInput.apply(E_se, E, W, BILINEAR_SEL[bilinear_type], dense)


def bilinear(E, W, bilinear_type):
    return _Bilinear.apply(E, W, BILINEAR_SEL[bilinear_type])


# ----------------------------------------------------------------------------------------------
# VarLen pooled lookup and the whole-table L2 term
# ----------------------------------------------------------------------------------------------
POOL_MODES = {"sum": 0, "mean": 1, "max": 2}


class _VarlenPool(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, X, table, col, T, len_col, mode, err_flag, id_mode):
        _require_cuda(X, "X")
        X = _rowmajor(X)
        B = X.shape[0]
        V, D = table.shape
        out = torch.empty(B, D, device=X.device, dtype=torch.float32)
        _lib.call("ctr_varlen_pool_fwd", _ptr(X), X.stride(0), B, col, T, len_col, _ptr(table), V, D,
                  mode, _ptr(out), D, _ptr(err_flag), id_mode, _stream())
        ctx.args = (col, T, len_col, mode, id_mode)
        ctx.save_for_backward(X, table)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, dout):
        X, table = ctx.saved_tensors
        col, T, len_col, mode, id_mode = ctx.args
        V, D = table.shape
        dout = dout.contiguous()
        dtable = torch.zeros_like(table)
        _lib.call("ctr_varlen_pool_bwd", _ptr(X), X.stride(0), X.shape[0], col, T, len_col, _ptr(table),
                  V, D, mode, _ptr(dout), D, _ptr(dtable), id_mode, _stream())
        return None, dtable, None, None, None, None, None, None


def varlen_pool(X, table, col, maxlen, len_col, combiner, err_flag, id_mode=0):
    if combiner not in POOL_MODES:
        raise ValueError("parameter mode should in [sum, mean, max]")
    return _VarlenPool.apply(X, table, int(col), int(maxlen), -1 if len_col is None else int(len_col),
                             POOL_MODES[combiner], err_flag, int(id_mode))


class _SumSq(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, scale, no_grad_ids, *weights):
        dev = weights[0].device
        out = torch.zeros(1, device=dev, dtype=torch.float32)
        for w in weights:
            _require_cuda(w, "regularised weight")
            wc = w if w.is_contiguous() else w.contiguous()
            _lib.call("ctr_sumsq_acc", _ptr(wc), wc.numel(), float(scale), _ptr(out), _stream())
        ctx.scale = float(scale)
        ctx.skip = [id(w) in no_grad_ids for w in weights]
        ctx.save_for_backward(*weights)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        # d/dw (scale * sum w^2) = 2*scale*w — elementwise scaling of parameters already resident
        return (None, None) + tuple(None if skip else w * (2.0 * ctx.scale * g)
                                    for w, skip in zip(ctx.saved_tensors, ctx.skip))


def l2_penalty(weights, scale, no_grad_ids=()):
    """scale * sum_w sum(w^2) via the streaming sum-of-squares kernel (basemodel.py:412-428).
    Parameters whose id() is in no_grad_ids contribute their value but receive no gradient here (the
    fused row-wise optimizer applies their L2 term to the rows it touches)."""
    weights = [w for w in weights if w.numel() > 0]
    if not weights or scale == 0:
        return None
    return _SumSq.apply(scale, no_grad_ids, *weights)


# ----------------------------------------------------------------------------------------------
# adjacent models: bi-interaction pooling, input-aware refinement, AFM attention
# ----------------------------------------------------------------------------------------------
def _block3(E):
    B, F, D = E.shape
    if E.stride(2) != 1 or E.stride(1) != D:
        E = E.contiguous()
    return E


class _BiPool(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, E):
        _require_cuda(E, "BiInteractionPooling input")
        E = _block3(E)
        B, F, D = E.shape
        out = torch.empty(B, D, device=E.device, dtype=torch.float32)
        _lib.call("ctr_bipool_fwd", _ptr(E), E.stride(0), F, D, _ptr(out), D, B, _stream())
        ctx.save_for_backward(E)
        return out.unsqueeze(1)

    @staticmethod
    @_on_device
    def backward(ctx, g):
        (E,) = ctx.saved_tensors
        B, F, D = E.shape
        g = g.reshape(B, D).contiguous()
        dE = torch.empty(B, F, D, device=E.device, dtype=torch.float32)
        _lib.call("ctr_bipool_bwd", _ptr(E), E.stride(0), F, D, _ptr(g), D, _ptr(dE), F * D, B, _stream())
        return dE


def bi_interaction_pooling(E):
    """[B,F,D] -> [B,1,D] (reference layers/interaction.py:54-61)."""
    return _BiPool.apply(E)


class _Refine(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, P, E, L, softmax):
        _require_cuda(P, "refine factor")
        E = _block3(E)
        B, F, D = E.shape
        P = P.contiguous()
        Lc = L.contiguous() if L is not None else None
        m = torch.empty(B, F, device=E.device, dtype=torch.float32)
        Er = torch.empty(B, F, D, device=E.device, dtype=torch.float32)
        lin = torch.empty(B, device=E.device, dtype=torch.float32) if Lc is not None else None
        _lib.call("ctr_refine_fwd", _ptr(P), _ptr(E), E.stride(0), _ptr(Lc), F, D, 1 if softmax else 0, _ptr(m),
                  _ptr(Er), F * D, _ptr(lin), B, _stream())
        ctx.softmax, ctx.has_L = softmax, Lc is not None
        ctx.save_for_backward(m, E, Lc if Lc is not None else m.new_empty(0))
        if lin is None:
            lin = m.new_empty(0)
            ctx.mark_non_differentiable(lin)
        return Er, lin

    @staticmethod
    @_on_device
    def backward(ctx, dEr, dlin):
        m, E, L = ctx.saved_tensors
        B, F, D = E.shape
        L = L if ctx.has_L else None
        dEr = dEr.contiguous() if dEr is not None else None
        dlin = dlin.contiguous() if (dlin is not None and ctx.has_L) else None
        dP = torch.empty(B, F, device=E.device, dtype=torch.float32)
        dE = torch.empty(B, F, D, device=E.device, dtype=torch.float32)
        dL = torch.empty(B, F, device=E.device, dtype=torch.float32) if L is not None else None
        _lib.call("ctr_refine_bwd", _ptr(m), _ptr(E), E.stride(0), _ptr(L if dlin is not None else None), F, D,
                  1 if ctx.softmax else 0, _ptr(dEr), F * D, _ptr(dlin), _ptr(dP), _ptr(dE), F * D,
                  _ptr(dL if dlin is not None else None), B, _stream())
        if dL is not None and dlin is None:
            dL.zero_()
        return dP, dE, dL, None


def refine(P, E, L=None, softmax=True):
    """(Er [B,F,D], lin [B] | None): input-aware re-weighting of IFM (softmax=True: m = F*softmax(P)) and
    DIFM (softmax=False: m = P); L [B,F] = the sample's per-field linear weights."""
    Er, lin = _Refine.apply(P, E, L, bool(softmax))
    return Er, (lin if L is not None else None)


class _AFM(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, E, W, b, h):
        _require_cuda(E, "AFM input")
        E = _block3(E)
        B, F, D = E.shape
        Wc, bc, hc = W.contiguous(), b.contiguous(), h.reshape(-1).contiguous()
        A = Wc.shape[1]
        out = torch.empty(B, D, device=E.device, dtype=torch.float32)
        _lib.call("ctr_afm_fwd", _ptr(E), E.stride(0), F, D, A, _ptr(Wc), _ptr(bc), _ptr(hc), _ptr(out), B, _stream())
        ctx.save_for_backward(E, Wc, bc, hc)
        ctx.shapes = (W.shape, b.shape, h.shape)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        E, Wc, bc, hc = ctx.saved_tensors
        B, F, D = E.shape
        A = Wc.shape[1]
        g = g.contiguous()
        dE = torch.empty(B, F, D, device=E.device, dtype=torch.float32)
        dW, db, dh = (torch.empty_like(t) for t in (Wc, bc, hc))
        _lib.call("ctr_afm_bwd", _ptr(E), E.stride(0), F, D, A, _ptr(Wc), _ptr(bc), _ptr(hc), _ptr(g),
                  _ptr(dE), F * D, _ptr(dW), _ptr(db), _ptr(dh), B, _stream())
        s = ctx.shapes
        return dE, dW.view(s[0]), db.view(s[1]), dh.view(s[2])


def afm_attention(E, attention_W, attention_b, projection_h):
    """[B,F,D] -> attention output [B,D] (reference layers/interaction.py:307-326; the caller applies the
    dropout and the projection p)."""
    return _AFM.apply(E, attention_W, attention_b, projection_h)


class _FieldAttn(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, Q, K, V, R, heads, scale):
        _require_cuda(Q, "attention input")
        Q, K, V, R = (t.contiguous() for t in (Q, K, V, R))
        B, F, D = Q.shape
        Y = torch.empty_like(Q)
        _lib.call("ctr_fieldattn_fwd", _ptr(Q), _ptr(K), _ptr(V), _ptr(R), F, D, heads, ctypes.c_float(scale), _ptr(Y), B,
                  _stream())
        ctx.heads, ctx.scale = heads, scale
        ctx.save_for_backward(Q, K, V, Y)
        return Y

    @staticmethod
    @_on_device
    def backward(ctx, dY):
        Q, K, V, Y = ctx.saved_tensors
        B, F, D = Q.shape
        dY = dY.contiguous()
        dQ, dK, dV, dR = (torch.empty_like(Q) for _ in range(4))
        _lib.call("ctr_fieldattn_bwd", _ptr(Q), _ptr(K), _ptr(V), _ptr(Y), _ptr(dY), F, D, ctx.heads,
                  ctypes.c_float(ctx.scale), _ptr(dQ), _ptr(dK), _ptr(dV), _ptr(dR), B, _stream())
        return dQ, dK, dV, dR, None, None


def field_attention(Q, K, V, R, heads, scale):
    """relu(multi-head softmax(QK^T * scale) V + R) over the field axis, all [B,F,D]."""
    return _FieldAttn.apply(Q, K, V, R, int(heads), float(scale))


# ----------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------
class _BceSum(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, y_pred, y):
        p = y_pred.reshape(-1).contiguous()
        t = y.reshape(-1).to(torch.float32).contiguous()
        out = torch.empty(1, device=p.device, dtype=torch.float32)
        _lib.call("ctr_bce_sum_fwd", _ptr(p), _ptr(t), p.numel(), _ptr(out), _stream())
        ctx.save_for_backward(p, t)
        ctx.shape = y_pred.shape
        return out.reshape(())

    @staticmethod
    @_on_device
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        g = g.reshape(1).contiguous()
        dp = torch.empty_like(p)
        _lib.call("ctr_bce_sum_bwd", _ptr(p), _ptr(t), _ptr(g), p.numel(), _ptr(dp), _stream())
        return dp.view(ctx.shape), None


def binary_cross_entropy(y_pred, y, reduction="mean", **kw):
    """Drop-in for ``torch.nn.functional.binary_cross_entropy``: the summed loss of ``fit`` (reference
    basemodel.py:254) runs on the library's kernels; other reductions / CPU tensors go to torch."""
    if reduction == "sum" and not kw and isinstance(y_pred, torch.Tensor) and y_pred.is_cuda \
            and y_pred.dtype == torch.float32 and y_pred.numel() == y.numel():
        return _BceSum.apply(y_pred, y)
    return torch.nn.functional.binary_cross_entropy(y_pred, y, reduction=reduction, **kw)


class _PReLU(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, z, alpha):
        _require_cuda(z, "PReLU input")
        zc = z.contiguous()
        y = torch.empty_like(zc)
        _lib.call("ctr_prelu_fwd", _ptr(zc), _ptr(alpha), zc.numel(), _ptr(y), _stream())
        ctx.save_for_backward(zc, alpha)
        return y

    @staticmethod
    @_on_device
    def backward(ctx, dy):
        zc, alpha = ctx.saved_tensors
        dy = dy.contiguous()
        dz = torch.empty_like(zc)
        dalpha = torch.empty_like(alpha)
        _lib.call("ctr_prelu_bwd", _ptr(zc), _ptr(alpha), _ptr(dy), zc.numel(), _ptr(dz), _ptr(dalpha), _stream())
        return dz, dalpha


def prelu(z, alpha):
    """nn.PReLU() with a single slope (reference layers/activation.py:61-62)."""
    return _PReLU.apply(z, alpha)
