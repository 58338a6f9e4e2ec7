"""Fused row-wise optimizer for the embedding / linear tables.

The reference calls ``optim.step()`` on dense ``[V, D]`` table gradients after adding the whole-table L2
term to the loss (reference ``deepctr_torch/models/basemodel.py:262, 412-428, 447-461``): every step
sweeps every table several times.  With ``table_grad="rowwise"`` the backward already produces one summed
gradient per distinct row of the batch; ``RowwiseOptimizer`` hands those ``(uniq, rowgrad)`` buffers to
``ctr_rowopt_step`` (csrc/rowopt.cu), which updates the touched rows and their optimizer state in place,
and drives a stock torch optimizer of the same family for the (small, dense) remaining parameters.

Semantics ("lazy" rows): identical to the dense torch optimizer for ``sgd`` / ``adagrad`` when the table
L2 is zero (rows that are not in the batch have zero gradient and do not move); ``adam`` / ``rmsprop`` keep
per-row moments that only advance when the row occurs (the torch.optim.SparseAdam convention, with
torch.optim.Adam's update formula); L2 regularisation is applied to the rows of the batch
(``g += 2 * l2 * w``) instead of to the whole table.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops

KINDS = {"sgd": 0, "adagrad": 1, "adam": 2, "rmsprop": 3}
# torch defaults of the reference's string shortcuts (basemodel.py:447-461)
DEFAULTS = {
    "sgd": dict(lr=0.01),
    "adagrad": dict(lr=0.01, eps=1e-10),
    "adam": dict(lr=0.001, betas=(0.9, 0.999), eps=1e-8),
    "rmsprop": dict(lr=0.01, alpha=0.99, eps=1e-8),
}


def _dense_optimizer(kind, params, hp):
    params = list(params)
    if not params:
        return None
    if kind == "sgd":
        return torch.optim.SGD(params, lr=hp["lr"])
    if kind == "adagrad":
        return torch.optim.Adagrad(params, lr=hp["lr"], eps=hp["eps"])
    if kind == "adam":
        return torch.optim.Adam(params, lr=hp["lr"], betas=hp["betas"], eps=hp["eps"])
    if kind == "rmsprop":
        return torch.optim.RMSprop(params, lr=hp["lr"], alpha=hp["alpha"], eps=hp["eps"])
    raise NotImplementedError(kind)


class RowwiseOptimizer:
    """``optim.zero_grad()`` / ``optim.step()`` for a model with ``table_grad="rowwise"`` (or row-sharded
    tables): fused kernels for the tables, a torch optimizer of the same family for everything else."""

    def __init__(self, model, kind="adam", l2_embedding=None, l2_linear=None, **hyper):
        if kind not in KINDS:
            raise NotImplementedError("RowwiseOptimizer: optimizer %r (available: %s)" % (kind, sorted(KINDS)))
        self.model, self.kind = model, kind
        self.hp = dict(DEFAULTS[kind])
        self.hp.update(hyper)
        dev = torch.device(model.device)
        plan = model._gather_plan(dev)
        if plan.varlen or plan.lin_varlen:
            raise NotImplementedError("RowwiseOptimizer: VarLenSparseFeat tables use the dense gradient path")
        self.plan = plan
        plan.keep_rowgrads = True
        table_ids = set(id(p) for p in plan.emb_params + plan.lin_params)
        self.dense = _dense_optimizer(kind, [p for p in model.parameters() if id(p) not in table_ids], self.hp)
        self.l2x2_emb = 2.0 * float(model.l2_reg_embedding if l2_embedding is None else l2_embedding)
        self.l2x2_lin = 2.0 * float(model.l2_reg_linear if l2_linear is None else l2_linear)
        b1, b2 = (self.hp.get("betas") or (self.hp.get("alpha", 0.0), 0.0))
        self.hp_dev = torch.tensor([0.0, self.hp["lr"], b1, b2, self.hp.get("eps", 0.0), 1.0, 1.0, 0.0],
                                   dtype=torch.float32, device=dev)
        need1, need2 = kind != "sgd", kind == "adam"
        self.state1 = [[torch.zeros_like(p) for p in ps] if need1 else None for ps in (plan.emb_params, plan.lin_params)]
        self.state2 = [[torch.zeros_like(p) for p in ps] if need2 else None for ps in (plan.emb_params, plan.lin_params)]
        self._ptr_key = None

    # ---- device pointer tables (tables and state), refreshed if a storage moved ------------------
    def _ptrs(self):
        plan = self.plan
        key = tuple(p.data_ptr() for p in plan.emb_params + plan.lin_params)
        if key != self._ptr_key:
            dev = plan.device

            def table(ts):
                if ts is None or len(ts) == 0:
                    return None
                return torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
            self._tab = [table(plan.emb_params), table(plan.lin_params)]
            self._s1 = [table(self.state1[0]), table(self.state1[1])]
            self._s2 = [table(self.state2[0]), table(self.state2[1])]
            self._ptr_key = key
        return self._tab, self._s1, self._s2

    def zero_grad(self, set_to_none=True):
        if self.dense is not None:
            self.dense.zero_grad(set_to_none=set_to_none)
        self.plan.pending.clear()

    def apply_rows(self, cap, n_uniq, uniq, rg_emb, rg_lin, n_emb, emb_plan_col=None, lin_plan_col=None):
        """One fused update of the touched rows from (uniq [n_cols, cap], rowgrads)."""
        plan = self.plan
        tab, s1, s2 = self._ptrs()
        kind = KINDS[self.kind]
        epc = plan.emb_plan_col if emb_plan_col is None else emb_plan_col
        lpc = plan.lin_plan_col if lin_plan_col is None else lin_plan_col
        if n_emb > 0:
            _lib.call("ctr_rowopt_step", kind, cap, ops._ptr(n_uniq), ops._ptr(uniq), n_emb, plan.D, ops._ptr(rg_emb),
                      cap * plan.D, ops._ptr(epc), ops._ptr(tab[0]), ops._ptr(s1[0]), ops._ptr(s2[0]),
                      ops._ptr(self.hp_dev), ctypes.c_float(self.l2x2_emb), ops._stream())
        if plan.n_lin > 0:
            _lib.call("ctr_rowopt_step", kind, cap, ops._ptr(n_uniq), ops._ptr(uniq), plan.n_lin, 1, ops._ptr(rg_lin),
                      cap, ops._ptr(lpc), ops._ptr(tab[1]), ops._ptr(s1[1]), ops._ptr(s2[1]),
                      ops._ptr(self.hp_dev), ctypes.c_float(self.l2x2_lin), ops._stream())

    def step(self):
        plan = self.plan
        with torch.cuda.device(plan.device):
            if self.dense is not None:
                self.dense.step()
            _lib.call("ctr_rowopt_tick", ops._ptr(self.hp_dev), ops._stream())
            sharded = getattr(self.model, "sharded", None)
            if sharded is not None:
                sharded.apply_received(self)
            for lease, rg_emb, rg_lin, n_emb in plan.pending:
                ws = lease.ws
                self.apply_rows(ws["B"], ws["n_uniq"], ws["uniq"], rg_emb, rg_lin, n_emb)
            plan.pending.clear()

    # torch-optimizer surface that callbacks / user code may poke
    @property
    def param_groups(self):
        return self.dense.param_groups if self.dense is not None else []

    def state_dict(self):
        return {"kind": self.kind, "hp": self.hp, "hp_dev": self.hp_dev.cpu(),
                "dense": self.dense.state_dict() if self.dense is not None else None,
                "state1": [[t.cpu() for t in ts] if ts else None for ts in self.state1],
                "state2": [[t.cpu() for t in ts] if ts else None for ts in self.state2]}
