"""BaseModel: the reference's Keras-like runtime (compile / fit / evaluate / predict) and the
sparse half of the hot path, re-hosted on the fused CUDA gather.

Reference: ``deepctr_torch/models/basemodel.py`` — ``Linear`` :34-92, ``BaseModel`` :95-527.
Differences that matter:

* the 52 per-feature ``nn.Embedding`` calls + ``Linear.forward`` + ``combined_dnn_input`` (+ FM)
  are ONE kernel launch (``ops.fused_input``); ``embedding_dict`` / ``linear_model.embedding_dict``
  still hold one ``nn.Embedding`` per ``embedding_name`` as parameter containers, so
  ``state_dict`` keys/shapes are the reference's and checkpoints round-trip both ways;
* ``table_grad="dense"`` (default, drop-in: dense ``[V,D]`` ``.grad`` like ``sparse=False``) or
  ``"rowwise"`` (B200-native: per-unique-row gradients delivered as sparse COO, no table-sized
  traffic — use with SGD / Adagrad / SparseAdam);
* no tensorflow: callbacks come from ``deepctr_torch_b200.callbacks``;
* CUDA only: running the model on a CPU tensor raises (no fallback path exists).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.data as Data
from torch.utils.data import DataLoader

try:
    from tqdm import tqdm
except ImportError:  # pragma: no cover
    tqdm = None

from .. import ops
from ..callbacks import CallbackList, History
from ..inputs import (DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features,
                      compute_input_dim, split_columns)
from ..layers import PredictionLayer


def _take(arrays, lo, hi):
    """Rows [lo, hi) of every array of a feature list (what `validation_split` needs)."""
    return [np.asarray(a)[lo:hi] for a in arrays]


def create_embedding_matrix(feature_columns, init_std=0.0001, linear=False, sparse=False, device="cpu"):
    """``nn.ModuleDict{embedding_name: nn.Embedding(V, D or 1)}``, ``N(0, init_std)`` — same
    container, init order and RNG consumption as reference inputs.py:158-180."""
    sparse_cols, _, varlen_cols = split_columns(feature_columns)
    table = nn.ModuleDict({feat.embedding_name: nn.Embedding(feat.vocabulary_size,
                                                             feat.embedding_dim if not linear else 1,
                                                             sparse=sparse)
                           for feat in sparse_cols + varlen_cols})
    for emb in table.values():
        nn.init.normal_(emb.weight, mean=0, std=init_std)
    return table.to(device)


class Linear(nn.Module):
    """Parameter container of the "wide" part (reference basemodel.py:34-61): dim-1 tables in
    ``embedding_dict`` and the dense ``weight [sum(dim), 1]``.  The arithmetic of
    ``Linear.forward`` (:63-92) is folded into the fused gather kernel."""

    def __init__(self, feature_columns, feature_index, init_std=0.0001, device="cpu"):
        super().__init__()
        self.feature_index = feature_index
        self.device = device
        self.sparse_feature_columns, self.dense_feature_columns, self.varlen_sparse_feature_columns = \
            split_columns(feature_columns)
        self.embedding_dict = create_embedding_matrix(feature_columns, init_std, linear=True, sparse=False,
                                                      device=device)
        for emb in self.embedding_dict.values():  # the reference initialises these twice (:55-56)
            nn.init.normal_(emb.weight, mean=0, std=init_std)
        if len(self.dense_feature_columns) > 0:
            self.weight = nn.Parameter(
                torch.Tensor(sum(fc.dimension for fc in self.dense_feature_columns), 1).to(device))
            torch.nn.init.normal_(self.weight, mean=0, std=init_std)


class BaseModel(nn.Module):
    # how the id cells of X are encoded: "float32" = the reference's fp32-encoded ids (exact below 2^24,
    # models/basemodel.py:242,369); "int32" = the 4-byte cell holds the int32 id itself:
    # `fit/predict/evaluate` pack the matrix accordingly, `forward` expects what `pack_inputs` returns.
    id_dtype = "float32"

    def __init__(self, linear_feature_columns, dnn_feature_columns, l2_reg_linear=1e-5, l2_reg_embedding=1e-5,
                 init_std=0.0001, seed=1024, task="binary", device="cpu", gpus=None, table_grad="dense"):
        super().__init__()
        torch.manual_seed(seed)
        self.l2_reg_linear, self.l2_reg_embedding = l2_reg_linear, l2_reg_embedding
        if table_grad not in ("dense", "rowwise"):
            raise ValueError("table_grad must be 'dense' or 'rowwise'")
        self.table_grad = table_grad
        self.dnn_feature_columns = dnn_feature_columns
        self.linear_feature_columns = linear_feature_columns
        self.device = device
        self.gpus = gpus
        if gpus and str(self.gpus[0]) not in self.device:
            raise ValueError("`gpus[0]` should be the same gpu with `device`")
        if gpus and len(gpus) > 1:
            raise NotImplementedError(
                "single-process DataParallel (reference basemodel.py:206-209) is replaced by one process "
                "per GPU with row-sharded tables: launch with torchrun and use "
                "deepctr_torch_b200.sharded.shard_model(model)")
        self.feature_index = build_input_features(list(linear_feature_columns) + list(dnn_feature_columns))
        self.embedding_dict = create_embedding_matrix(dnn_feature_columns, init_std, sparse=False, device=device)
        self.linear_model = Linear(linear_feature_columns, self.feature_index, device=device)
        self.regularization_weight = []
        self.add_regularization_weight(self.embedding_dict.parameters(), l2=l2_reg_embedding)
        self.add_regularization_weight(self.linear_model.parameters(), l2=l2_reg_linear)
        self.out = PredictionLayer(task)
        self.reg_loss = torch.zeros((1,), device=device)
        self.aux_loss = torch.zeros((1,), device=device)
        self.to(device)
        self._is_graph_network = True
        self._ckpt_saved_epoch = False
        self.history = History()
        self.stop_training = False
        self._plan = None

    # ------------------------------------------------------------------------------------------
    # the sparse half of the hot path
    # ------------------------------------------------------------------------------------------
    def _gather_plan(self, device):
        """Slot metadata of the fused gather (built lazily on the model's CUDA device)."""
        if self._plan is not None and self._plan.device == torch.device(device):
            return self._plan
        sparse, dense, varlen = split_columns(self.dnn_feature_columns)
        lsparse, ldense, lvarlen = split_columns(self.linear_feature_columns)
        dims = set(c.embedding_dim for c in sparse)
        if len(dims) > 1:
            raise ValueError("embedding_dim of SparseFeat must be the same for the fused gather "
                             "(got %s)" % sorted(dims))
        dim = dims.pop() if dims else 0
        fi = self.feature_index
        emb_slots = [(self.embedding_dict[c.embedding_name].weight, fi[c.name][0], c.vocabulary_size)
                     for c in sparse]
        lin_slots = [(self.linear_model.embedding_dict[c.embedding_name].weight, fi[c.name][0],
                      c.vocabulary_size) for c in lsparse]
        dense_cols = [k for c in dense for k in range(fi[c.name][0], fi[c.name][1])]
        lin_dense_cols = [k for c in ldense for k in range(fi[c.name][0], fi[c.name][1])]
        self._plan = ops.GatherPlan(emb_slots, lin_slots, dense_cols, lin_dense_cols, dim, device,
                                    id_mode=1 if self.id_dtype == "int32" else 0)
        self._plan.varlen = varlen
        self._plan.lin_varlen = lvarlen
        self._plan.n_sparse = len(sparse)
        return self._plan

    def embed(self, X, want_fm=False, want_blk=True):
        """Fused lookup for this batch.

        Returns ``(E, dnn_input, lin, fm, blk)``: ``blk [B, ld]`` is the zero-padded block
        (``ld = round_up(width, 4)``) that the DNN tower consumes without slicing (None when pooled
        VarLen fields force a generic assembly); ``E [B,F,D]`` embedding block (sparse fields in column
        order, then pooled VarLen fields — reference basemodel.py:368-380), ``dnn_input
        [B, F*D + n_dense]`` (= ``combined_dnn_input``, inputs.py:126-138), ``lin [B]`` the
        linear logit (basemodel.py:63-92), ``fm [B]`` the FM term or None."""
        if not X.is_cuda:
            raise RuntimeError("deepctr_torch_b200 models run on CUDA only (construct with device='cuda:0'); "
                               "there is no CPU implementation of the hot path")
        plan = self._gather_plan(X.device)
        ldw = self.linear_model.weight if len(self.linear_model.dense_feature_columns) > 0 else None
        varlen, lin_varlen = plan.varlen, plan.lin_varlen
        fused_fm = want_fm and not varlen
        blk, lin, fm = ops.fused_input(X, plan, ldw, want_blk=want_blk or bool(varlen), want_fm=fused_fm,
                                       grad_mode=self.table_grad)
        B = X.shape[0]
        D, F0 = plan.D, plan.n_emb
        E = dnn_input = None
        if not varlen:
            if blk is not None:
                E = blk[:, :F0 * D].view(B, F0, D) if F0 else None
                dnn_input = blk[:, :plan.width] if plan.width else None
        else:
            # pooled VarLen fields are appended behind the sparse fields (generic composition)
            pooled = []
            for c in varlen:
                s, e = self.feature_index[c.name]
                lcol = self.feature_index[c.length_name][0] if c.length_name is not None else None
                pooled.append(ops.varlen_pool(X, self.embedding_dict[c.embedding_name].weight, s, e - s,
                                              lcol, c.combiner, plan.err_flag, plan.id_mode))
            Dv = pooled[0].shape[1]
            if F0 and Dv != D:
                raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
            parts = ([blk[:, :F0 * D]] if F0 else []) + pooled
            E = torch.cat(parts, dim=1).view(B, F0 + len(pooled), Dv)
            dense_part = [blk[:, F0 * D:plan.width]] if plan.n_dense else []
            dnn_input = torch.cat([E.reshape(B, -1)] + dense_part, dim=1)
            if want_fm:
                fm = ops.fm(E).squeeze(1)
        for c in lin_varlen:
            s, e = self.feature_index[c.name]
            lcol = self.feature_index[c.length_name][0] if c.length_name is not None else None
            lin = lin + ops.varlen_pool(X, self.linear_model.embedding_dict[c.embedding_name].weight, s, e - s,
                                        lcol, c.combiner, plan.err_flag, plan.id_mode).squeeze(1)
        return E, dnn_input, lin, fm, (blk if not varlen else None)

    def linear_field_terms(self, X):
        """``(L [B, n_lin] | None, dense_term [B] | None)``: the per-field linear weights ``w_f[id_f]`` of every
        sample (IFM / DIFM multiply them by an input-aware factor before summing — the
        ``sparse_feat_refine_weight`` branch of reference basemodel.py:82-84) and the dense part of the linear
        logit.  One launch of the fused gather with the dim-1 tables in the embedding slots."""
        if self.table_grad != "dense":
            raise NotImplementedError("per-field linear terms (IFM / DIFM) need table_grad='dense'")
        main = self._gather_plan(X.device)
        lp = getattr(self, "_lin_plan", None)
        if lp is None or lp.device != X.device or lp.id_mode != main.id_mode:
            fi = self.feature_index
            lsparse, ldense, _ = split_columns(self.linear_feature_columns)
            slots = [(self.linear_model.embedding_dict[c.embedding_name].weight, fi[c.name][0], c.vocabulary_size)
                     for c in lsparse]
            lin_dense_cols = [k for c in ldense for k in range(fi[c.name][0], fi[c.name][1])]
            lp = ops.GatherPlan(slots, [], [], lin_dense_cols, 1, X.device, id_mode=main.id_mode)
            lp.err_flag = main.err_flag
            self._lin_plan = lp
        ldw = self.linear_model.weight if lp.n_lin_dense > 0 else None
        blk, lin, _ = ops.fused_input(X, lp, ldw, want_blk=lp.n_emb > 0, want_fm=False, grad_mode="dense")
        L = blk[:, :lp.n_emb] if lp.n_emb > 0 else None
        if main.lin_varlen:        # pooled dim-1 rows of the VarLen linear columns follow the sparse ones (basemodel.py:74-77)
            parts = [L] if L is not None else []
            for c in main.lin_varlen:
                s, e = self.feature_index[c.name]
                lcol = self.feature_index[c.length_name][0] if c.length_name is not None else None
                parts.append(ops.varlen_pool(X, self.linear_model.embedding_dict[c.embedding_name].weight, s, e - s,
                                             lcol, c.combiner, main.err_flag, main.id_mode))
            L = torch.cat(parts, dim=1)
        return L, (lin if ldw is not None else None)

    def input_from_feature_columns(self, X, feature_columns, embedding_dict, support_dense=True):
        """API-compatible view of the fused lookup (reference basemodel.py:354-380): a list of
        ``[B,1,D]`` embeddings and a list of dense ``[B,dim]`` slices."""
        _, dense, _ = split_columns(feature_columns)
        if not support_dense and len(dense) > 0:
            raise ValueError("DenseFeat is not supported in dnn_feature_columns")
        E = self.embed(X)[0]
        emb_list = [E[:, f:f + 1, :] for f in range(E.shape[1])] if E is not None else []
        dense_list = [X[:, self.feature_index[c.name][0]:self.feature_index[c.name][1]] for c in dense]
        return emb_list, dense_list

    def compute_input_dim(self, feature_columns, include_sparse=True, include_dense=True, feature_group=False):
        return compute_input_dim(feature_columns, include_sparse, include_dense, feature_group)

    def check_ids(self):
        if self._plan is not None:
            self._plan.check_ids()

    # ------------------------------------------------------------------------------------------
    # regularisation (reference basemodel.py:402-431)
    # ------------------------------------------------------------------------------------------
    def add_regularization_weight(self, weight_list, l1=0.0, l2=0.0):
        if isinstance(weight_list, torch.nn.parameter.Parameter):
            weight_list = [weight_list]
        else:
            weight_list = list(weight_list)
        self.regularization_weight.append((weight_list, l1, l2))

    def get_regularization_loss(self):
        total = torch.zeros((1,), device=self.device)
        for weight_list, l1, l2 in self.regularization_weight:
            params = [w[1] if isinstance(w, tuple) else w for w in weight_list]
            if l1 > 0:
                for p in params:
                    total = total + torch.sum(l1 * torch.abs(p))
            if l2 > 0:
                if params and params[0].is_cuda:
                    # tables whose gradient is consumed row-wise by the fused optimizer contribute their VALUE
                    # here; their L2 gradient is applied to the touched rows inside ctr_rowopt_step
                    lazy = self._lazy_l2_ids()
                    pen = ops.l2_penalty(params, l2, no_grad_ids=lazy)
                    if pen is not None:
                        total = total + pen
                else:
                    for p in params:
                        total = total + torch.sum(l2 * torch.square(p))
        return total

    def _lazy_l2_ids(self):
        plan = self._plan
        if plan is None or not (plan.keep_rowgrads or getattr(self, "sharded", None) is not None):
            return ()
        return set(id(p) for p in plan.emb_params + plan.lin_params)

    def use_int_ids(self, on=True):
        """Switch the id encoding of X (see `id_dtype`); rebuilds the launch metadata lazily."""
        self.id_dtype = "int32" if on else "float32"
        if self._plan is not None and getattr(self, "sharded", None) is None:
            self._plan = None
        elif self._plan is not None:
            self._plan.id_mode = 1 if on else 0
        return self

    def add_auxiliary_loss(self, aux_loss, alpha):
        self.aux_loss = aux_loss * alpha

    # ------------------------------------------------------------------------------------------
    # compile / fit / evaluate / predict (reference basemodel.py:137-352, 433-512)
    # ------------------------------------------------------------------------------------------
    def compile(self, optimizer, loss=None, metrics=None):
        self.metrics_names = ["loss"]
        self.optim = self._get_optim(optimizer)
        self.loss_func = self._get_loss_func(loss)
        self.metrics = self._get_metrics(metrics)

    def _get_optim(self, optimizer):
        if isinstance(optimizer, str):
            if self.table_grad in ("rowwise", "sharded") and torch.device(self.device).type == "cuda":
                # tables: fused row-wise kernels on the row-gradient stream; everything else: the torch
                # optimizer of the same family with the reference's defaults (deepctr_torch_b200/optim.py)
                from ..optim import KINDS, RowwiseOptimizer
                if optimizer not in KINDS:
                    raise NotImplementedError
                return RowwiseOptimizer(self, optimizer)
            if optimizer == "sgd":
                return torch.optim.SGD(self.parameters(), lr=0.01)
            if optimizer == "adam":
                return torch.optim.Adam(self.parameters())
            if optimizer == "adagrad":
                return torch.optim.Adagrad(self.parameters())
            if optimizer == "rmsprop":
                return torch.optim.RMSprop(self.parameters())
            raise NotImplementedError
        return optimizer

    def _get_loss_func(self, loss):
        if isinstance(loss, str):
            return self._get_loss_func_single(loss)
        if isinstance(loss, list):
            return [self._get_loss_func_single(l) for l in loss]
        return loss

    def _get_loss_func_single(self, loss):
        if loss == "binary_crossentropy":
            return ops.binary_cross_entropy
        if loss == "mse":
            return F.mse_loss
        if loss == "mae":
            return F.l1_loss
        raise NotImplementedError

    @staticmethod
    def _accuracy_score(y_true, y_pred):
        from sklearn.metrics import accuracy_score
        return accuracy_score(y_true, np.where(y_pred > 0.5, 1, 0))

    def _get_metrics(self, metrics, set_eps=False):
        from sklearn.metrics import log_loss, mean_squared_error, roc_auc_score
        metrics_ = {}
        if metrics:
            for metric in metrics:
                if metric in ("binary_crossentropy", "logloss"):
                    metrics_[metric] = log_loss
                if metric == "auc":
                    metrics_[metric] = roc_auc_score
                if metric == "mse":
                    metrics_[metric] = mean_squared_error
                if metric in ("accuracy", "acc"):
                    metrics_[metric] = self._accuracy_score
                self.metrics_names.append(metric)
        return metrics_

    # ------------------------------------------------------------------------------------------
    # input pipeline (reference basemodel.py:191-198, 242-243: ONE [N, C] matrix per call)
    # ------------------------------------------------------------------------------------------
    def _id_feature_names(self):
        names = set()
        for c in list(self.linear_feature_columns or []) + list(self.dnn_feature_columns or []):
            if isinstance(c, (SparseFeat, VarLenSparseFeat)):
                names.add(c.name)
                if isinstance(c, VarLenSparseFeat) and c.length_name is not None:
                    names.add(c.length_name)
        return names

    def pack_inputs(self, x):
        """dict / list of per-feature arrays -> ONE float32 ``[N, C]`` matrix in ``feature_index`` order.

        ``id_dtype == "float32"``: what the reference builds (ids converted to fp32, exact below 2^24).
        ``id_dtype == "int32"``: id cells carry the int32 id itself (bit pattern) — the matrix stays one
        float32-typed buffer so a batch is still one contiguous H2D copy."""
        if isinstance(x, dict):
            x = [x[feature] for feature in self.feature_index]
        arrs = []
        for a in x:
            a = np.asarray(a)
            arrs.append(a[:, None] if a.ndim == 1 else a)
        n = arrs[0].shape[0] if arrs else 0
        width = sum(a.shape[1] for a in arrs)
        out = np.empty((n, width), dtype=np.float32)
        as_int = self.id_dtype == "int32"
        bits = out.view(np.int32)
        id_names = self._id_feature_names() if as_int else ()
        cursor = 0
        names = list(self.feature_index.keys())
        for k, a in enumerate(arrs):
            w = a.shape[1]
            if as_int and k < len(names) and names[k] in id_names:
                ai = a.astype(np.int64)
                if ai.size and (ai.min() < -2 ** 31 or ai.max() >= 2 ** 31):
                    raise IndexError("index out of range in self (id does not fit int32)")
                bits[:, cursor:cursor + w] = ai
            else:
                out[:, cursor:cursor + w] = a
            cursor += w
        return out

    def _batches(self, X_all, y_all, batch_size, shuffle):
        """Yield device batches (X [b, C], y) with the NEXT batch's host->device copy already in flight.

        Small data sets (<= a quarter of the free HBM) are made device-resident once and batches are
        device-side row gathers; larger ones stream through two pinned staging buffers on a copy stream
        (the reference copies synchronously per batch, basemodel.py:242-243).  The index order is produced
        by the same ``DataLoader(shuffle=...)`` machinery as the reference, so the RNG consumption and
        the batch composition are identical."""
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("deepctr_torch_b200 models run on CUDA only (construct with device='cuda:0'); "
                               "there is no CPU implementation of the hot path")
        n = X_all.shape[0]
        order = DataLoader(Data.TensorDataset(torch.arange(n)), shuffle=shuffle, batch_size=batch_size)
        shard = getattr(self, "sharded", None)

        def my_part(idx):
            if shard is None:
                return idx
            chunk = (idx.numel() + shard.plan.world - 1) // shard.plan.world
            return idx[shard.plan.rank * chunk:(shard.plan.rank + 1) * chunk]

        free, _total = torch.cuda.mem_get_info(dev)
        resident = X_all.numel() * 4 + (y_all.numel() * 4 if y_all is not None else 0) <= free // 4
        if resident:
            Xd = X_all.to(dev, non_blocking=True)
            yd = y_all.to(dev, non_blocking=True) if y_all is not None else None
            for (idx,) in order:
                idx = my_part(idx)
                if not shuffle and idx.numel() > 0:
                    lo, hi = int(idx[0]), int(idx[-1]) + 1
                    yield Xd[lo:hi], (yd[lo:hi] if yd is not None else None)
                else:
                    idx_d = idx.to(dev, non_blocking=True)
                    yield Xd.index_select(0, idx_d), (yd.index_select(0, idx_d) if yd is not None else None)
            return
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        C = X_all.shape[1]
        ywidth = tuple(y_all.shape[1:]) if y_all is not None else ()
        stage = [(torch.empty(batch_size, C).pin_memory(), torch.empty((batch_size,) + ywidth).pin_memory())
                 for _ in range(2)]
        devbuf = [(torch.empty(batch_size, C, device=dev), torch.empty((batch_size,) + ywidth, device=dev))
                  for _ in range(2)]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        for e in consumed:
            e.record(cur)

        def enqueue(k, idx):
            m = idx.numel()
            copied[k].synchronize()              # the previous H2D out of this staging buffer has finished
            torch.index_select(X_all, 0, idx, out=stage[k][0][:m])
            if y_all is not None:
                torch.index_select(y_all, 0, idx, out=stage[k][1][:m])
            copy_stream.wait_event(consumed[k])  # the step that read devbuf[k] has finished
            with torch.cuda.stream(copy_stream):
                devbuf[k][0][:m].copy_(stage[k][0][:m], non_blocking=True)
                if y_all is not None:
                    devbuf[k][1][:m].copy_(stage[k][1][:m], non_blocking=True)
                copied[k].record(copy_stream)
            return m

        it = iter(order)
        k = 0
        nxt = next(it, None)
        m_next = enqueue(k, my_part(nxt[0])) if nxt is not None else 0
        while nxt is not None:
            m, kk = m_next, k
            nxt = next(it, None)
            if nxt is not None:
                k ^= 1
                m_next = enqueue(k, my_part(nxt[0]))
            cur.wait_event(copied[kk])
            yield devbuf[kk][0][:m], (devbuf[kk][1][:m] if y_all is not None else None)
            consumed[kk].record(cur)

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, initial_epoch=0, validation_split=0.,
            validation_data=None, shuffle=True, callbacks=None):
        """Keras-style training loop with the reference's arguments, log keys and History
        (reference basemodel.py:137-309).  Differences, all on the input/bookkeeping side:
        batches are prefetched, the epoch loss is accumulated on the device and read once per epoch, the
        per-batch training metrics are computed at the end of the epoch from predictions kept on the
        device, and out-of-range ids are reported once per epoch."""
        if isinstance(x, dict):
            x = [x[feature] for feature in self.feature_index]
        val_x, val_y = [], []
        if validation_data:
            if len(validation_data) not in (2, 3):
                raise ValueError("When passing a `validation_data` argument, it must contain either 2 items "
                                 "(x_val, y_val), or 3 items (x_val, y_val, val_sample_weights). "
                                 "However we received `validation_data=%s`" % (validation_data,))
            val_x, val_y = validation_data[0], validation_data[1]
        elif validation_split and 0. < validation_split < 1.:
            n0 = len(x[0])
            cut = int(n0 * (1. - validation_split))
            x, val_x = _take(x, 0, cut), _take(x, cut, n0)
            y, val_y = np.asarray(y)[:cut], np.asarray(y)[cut:]
        do_validation = len(val_y) > 0
        X_all = torch.from_numpy(self.pack_inputs(x))
        y_all = torch.from_numpy(np.asarray(y)).float()
        batch_size = 256 if batch_size is None else batch_size
        sample_num = X_all.shape[0]
        steps_per_epoch = (sample_num - 1) // batch_size + 1
        model = self.train()
        loss_func, optim = self.loss_func, self.optim
        shard = getattr(self, "sharded", None)
        dev = torch.device(self.device)

        callbacks = CallbackList((callbacks or []) + [self.history])
        callbacks.set_model(self)
        callbacks.on_train_begin()
        self.stop_training = False
        if verbose:
            print("Train on {0} samples, validate on {1} samples, {2} steps per epoch".format(
                sample_num, len(val_y), steps_per_epoch))
        for epoch in range(initial_epoch, epochs):
            callbacks.on_epoch_begin(epoch)
            start_time = time.time()
            loss_acc = torch.zeros((), device=dev)
            kept = []
            batches = self._batches(X_all, y_all, batch_size, shuffle)
            bar = tqdm(batches, total=steps_per_epoch, disable=verbose != 1) if tqdm is not None else batches
            try:
                for xb, yb in bar:
                    y_pred = model(xb).squeeze()
                    optim.zero_grad()
                    if isinstance(loss_func, list):
                        loss = sum(loss_func[i](y_pred[:, i], yb[:, i], reduction="sum")
                                   for i in range(self.num_tasks))
                    else:
                        loss = loss_func(y_pred, yb.squeeze(), reduction="sum")
                    if shard is None:
                        total_loss = loss + self.get_regularization_loss() + self.aux_loss
                        total_loss.backward()
                    else:
                        # data term first; its dense gradients are summed over the ranks, then the
                        # (replicated) regulariser adds its gradient once
                        loss.backward()
                        shard.finish_step()
                        reg = self.get_regularization_loss() + self.aux_loss
                        if reg.requires_grad:
                            reg.sum().backward()
                        total_loss = loss + reg
                    loss_acc += total_loss.detach().sum()
                    optim.step()
                    if verbose > 0 and self.metrics:
                        kept.append((yb.detach(), y_pred.detach()))
            finally:
                if tqdm is not None:
                    bar.close()
            if shard is not None:
                torch.distributed.all_reduce(loss_acc, group=shard.group)
            self.check_ids()
            epoch_logs = {"loss": float(loss_acc.item()) / sample_num}      # the only host sync of the epoch
            if kept:
                host = [(t.cpu().numpy(), p.cpu().numpy().astype("float64")) for t, p in kept]
                for name, metric_fun in self.metrics.items():
                    epoch_logs[name] = np.sum([metric_fun(t, p) for t, p in host]) / steps_per_epoch
            if do_validation:
                for name, result in self.evaluate(val_x, val_y, batch_size).items():
                    epoch_logs["val_" + name] = result
            if verbose > 0:
                msg = "{0}s - loss: {1: .4f}".format(int(time.time() - start_time), epoch_logs["loss"])
                for name in self.metrics:
                    if name in epoch_logs:
                        msg += " - " + name + ": {0: .4f}".format(epoch_logs[name])
                if do_validation:
                    for name in self.metrics:
                        msg += " - val_" + name + ": {0: .4f}".format(epoch_logs["val_" + name])
                print("Epoch {0}/{1}".format(epoch + 1, epochs))
                print(msg)
            callbacks.on_epoch_end(epoch, epoch_logs)
            if self.stop_training:
                break
        callbacks.on_train_end()
        return self.history

    def evaluate(self, x, y, batch_size=256):
        pred_ans = self.predict(x, batch_size)
        return {name: metric_fun(y, pred_ans) for name, metric_fun in self.metrics.items()}

    def predict(self, x, batch_size=256):
        """float64 ``[N, 1]`` predictions (reference basemodel.py:325-352); results stay on the device until
        the last batch, one device->host copy at the end."""
        model = self.eval()
        X_all = torch.from_numpy(self.pack_inputs(x))
        shard, self_sharded = getattr(self, "sharded", None), None
        if shard is not None:
            raise NotImplementedError("predict() on a row-sharded model: gather the tables with "
                                      "sharded.gather_full_state_dict and predict on one GPU")
        outs = []
        with torch.no_grad():
            for xb, _ in self._batches(X_all, None, batch_size, False):
                outs.append(model(xb))
        self.check_ids()
        if not outs:
            return np.zeros((0, 1), dtype="float64")
        return torch.cat(outs).cpu().numpy().astype("float64")

    def make_graphed_step(self, batch_size, loss_fn=None, with_reg=False):
        """One forward+loss+backward captured as a CUDA graph (see deepctr_torch_b200.graph)."""
        from ..graph import GraphedStep, ShardedGraphedStep
        if getattr(self, "sharded", None) is not None:
            return ShardedGraphedStep(self, batch_size, loss_fn=loss_fn)
        return GraphedStep(self, batch_size, loss_fn=loss_fn, with_reg=with_reg)

    def _in_multi_worker_mode(self):
        return None

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_plan"] = None      # device-side launch metadata is rebuilt lazily after unpickling
        state["_lin_plan"] = None
        return state

    @property
    def embedding_size(self):
        cols = [c for c in (self.dnn_feature_columns or []) if isinstance(c, (SparseFeat, VarLenSparseFeat))]
        sizes = set(c.embedding_dim for c in cols)
        if len(sizes) > 1:
            raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
        return list(sizes)[0]
