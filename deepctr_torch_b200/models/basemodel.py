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


def slice_arrays(arrays, start=None, stop=None):
    """Keras-style slicing of an array or list of arrays (reference layers/utils.py:19-70)."""
    if arrays is None:
        return [None]
    if isinstance(arrays, np.ndarray):
        arrays = [arrays]
    if isinstance(start, list) and stop is not None:
        raise ValueError("The stop argument has to be None if the value of start is a list.")
    if isinstance(arrays, list):
        if hasattr(start, "__len__"):
            if hasattr(start, "shape"):
                start = start.tolist()
            return [None if x is None else x[start] for x in arrays]
        if len(arrays) == 1:
            return arrays[0][start:stop]
        return [None if x is None else x[start:stop] for x in arrays]
    if hasattr(start, "__len__"):
        if hasattr(start, "shape"):
            start = start.tolist()
        return arrays[start]
    if hasattr(start, "__getitem__"):
        return arrays[start:stop]
    return [None]


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
    def __init__(self, linear_feature_columns, dnn_feature_columns, l2_reg_linear=1e-5, l2_reg_embedding=1e-5,
                 init_std=0.0001, seed=1024, task="binary", device="cpu", gpus=None, table_grad="dense"):
        super().__init__()
        torch.manual_seed(seed)
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
        self._plan = ops.GatherPlan(emb_slots, lin_slots, dense_cols, lin_dense_cols, dim, device)
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
                                              lcol, c.combiner, plan.err_flag))
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
                                        lcol, c.combiner, plan.err_flag).squeeze(1)
        return E, dnn_input, lin, fm, (blk if not varlen else None)

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
                    pen = ops.l2_penalty(params, l2)
                    if pen is not None:
                        total = total + pen
                else:
                    for p in params:
                        total = total + torch.sum(l2 * torch.square(p))
        return total

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
            if optimizer == "sgd":
                return torch.optim.SGD(self.parameters(), lr=0.01)
            if optimizer == "adam":
                if self.table_grad == "rowwise":
                    raise NotImplementedError("torch.optim.Adam does not accept the sparse row gradients of "
                                              "table_grad='rowwise'; use 'sgd' or 'adagrad'")
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
            return F.binary_cross_entropy
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

    def _to_matrix(self, x):
        """dict / list of arrays -> ONE [N, C] array in feature_index order (reference :155-156,191-198)."""
        if isinstance(x, dict):
            x = [x[feature] for feature in self.feature_index]
        x = [np.asarray(a) for a in x]
        for i in range(len(x)):
            if len(x[i].shape) == 1:
                x[i] = np.expand_dims(x[i], axis=1)
        return np.concatenate(x, axis=-1)

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, initial_epoch=0, validation_split=0.,
            validation_data=None, shuffle=True, callbacks=None):
        if isinstance(x, dict):
            x = [x[feature] for feature in self.feature_index]
        do_validation = False
        if validation_data:
            do_validation = True
            if len(validation_data) == 2:
                val_x, val_y = validation_data
            elif len(validation_data) == 3:
                val_x, val_y, _ = validation_data
            else:
                raise ValueError("When passing a `validation_data` argument, it must contain either 2 items "
                                 "(x_val, y_val), or 3 items (x_val, y_val, val_sample_weights). "
                                 "However we received `validation_data=%s`" % (validation_data,))
            if isinstance(val_x, dict):
                val_x = [val_x[feature] for feature in self.feature_index]
        elif validation_split and 0. < validation_split < 1.:
            do_validation = True
            n0 = x[0].shape[0] if hasattr(x[0], "shape") else len(x[0])
            split_at = int(n0 * (1. - validation_split))
            x, val_x = slice_arrays(x, 0, split_at), slice_arrays(x, split_at)
            y, val_y = slice_arrays(y, 0, split_at), slice_arrays(y, split_at)
        else:
            val_x, val_y = [], []
        X_all = torch.from_numpy(self._to_matrix(x))
        y_all = torch.from_numpy(np.asarray(y))
        train_tensor_data = Data.TensorDataset(X_all, y_all)
        if batch_size is None:
            batch_size = 256
        model = self.train()
        loss_func, optim = self.loss_func, self.optim
        train_loader = DataLoader(dataset=train_tensor_data, shuffle=shuffle, batch_size=batch_size)
        sample_num = len(train_tensor_data)
        steps_per_epoch = (sample_num - 1) // batch_size + 1

        callbacks = CallbackList((callbacks or []) + [self.history])
        callbacks.set_model(self)
        callbacks.on_train_begin()
        self.stop_training = False
        if verbose:
            print("Train on {0} samples, validate on {1} samples, {2} steps per epoch".format(
                len(train_tensor_data), len(val_y), steps_per_epoch))
        for epoch in range(initial_epoch, epochs):
            callbacks.on_epoch_begin(epoch)
            epoch_logs = {}
            start_time = time.time()
            total_loss_epoch = 0.0
            train_result = {}
            it = enumerate(train_loader)
            bar = tqdm(it, disable=verbose != 1) if tqdm is not None else it
            try:
                for _, (x_train, y_train) in bar:
                    xb = x_train.to(self.device).float()
                    yb = y_train.to(self.device).float()
                    y_pred = model(xb).squeeze()
                    optim.zero_grad()
                    if isinstance(loss_func, list):
                        loss = sum(loss_func[i](y_pred[:, i], yb[:, i], reduction="sum")
                                   for i in range(self.num_tasks))
                    else:
                        loss = loss_func(y_pred, yb.squeeze(), reduction="sum")
                    total_loss = loss + self.get_regularization_loss() + self.aux_loss
                    total_loss_epoch += total_loss.item()      # host sync, as in the reference (:259)
                    self.check_ids()
                    total_loss.backward()
                    optim.step()
                    if verbose > 0:
                        for name, metric_fun in self.metrics.items():
                            train_result.setdefault(name, []).append(metric_fun(
                                yb.cpu().data.numpy(), y_pred.cpu().data.numpy().astype("float64")))
            finally:
                if tqdm is not None:
                    bar.close()
            epoch_logs["loss"] = total_loss_epoch / sample_num
            for name, result in train_result.items():
                epoch_logs[name] = np.sum(result) / steps_per_epoch
            if do_validation:
                for name, result in self.evaluate(val_x, val_y, batch_size).items():
                    epoch_logs["val_" + name] = result
            if verbose > 0:
                epoch_time = int(time.time() - start_time)
                print("Epoch {0}/{1}".format(epoch + 1, epochs))
                eval_str = "{0}s - loss: {1: .4f}".format(epoch_time, epoch_logs["loss"])
                for name in self.metrics:
                    eval_str += " - " + name + ": {0: .4f}".format(epoch_logs[name])
                if do_validation:
                    for name in self.metrics:
                        eval_str += " - val_" + name + ": {0: .4f}".format(epoch_logs["val_" + name])
                print(eval_str)
            callbacks.on_epoch_end(epoch, epoch_logs)
            if self.stop_training:
                break
        callbacks.on_train_end()
        return self.history

    def evaluate(self, x, y, batch_size=256):
        pred_ans = self.predict(x, batch_size)
        return {name: metric_fun(y, pred_ans) for name, metric_fun in self.metrics.items()}

    def predict(self, x, batch_size=256):
        model = self.eval()
        tensor_data = Data.TensorDataset(torch.from_numpy(self._to_matrix(x)))
        test_loader = DataLoader(dataset=tensor_data, shuffle=False, batch_size=batch_size)
        pred_ans = []
        with torch.no_grad():
            for _, x_test in enumerate(test_loader):
                xb = x_test[0].to(self.device).float()
                pred_ans.append(model(xb).cpu().data.numpy())
                self.check_ids()
        return np.concatenate(pred_ans).astype("float64")

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
        return state

    @property
    def embedding_size(self):
        cols = [c for c in (self.dnn_feature_columns or []) if isinstance(c, (SparseFeat, VarLenSparseFeat))]
        sizes = set(c.embedding_dim for c in cols)
        if len(sizes) > 1:
            raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
        return list(sizes)[0]
