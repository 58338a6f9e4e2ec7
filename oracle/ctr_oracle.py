"""CPU ORACLE for the CTR embedding + interaction hot path — TEST INFRASTRUCTURE ONLY.

This file is a functional torch-CPU restatement of the reference's algorithm for the path
``X[B, C] -> multi-slot embedding lookup -> {Linear, FM, CIN, CrossNet, CrossNetMix, SENET,
Bilinear, DNN} -> logit -> sigmoid`` (forward) with the backward obtained by autograd over the
same restatement — exactly how the reference obtains its own backward
(reference ``deepctr_torch/models/basemodel.py:261``).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg.  The product package ``deepctr_torch_b200`` never imports it and has
no CPU compute path at all.

Pinning status: PINNED.  The reference's own tests hold no golden vectors for this path
(SURVEY.md §4/§8c), so the oracle is pinned against outputs of the unmodified reference run in
the build container: ``tests/golden/make_golden.py`` imports ``/root/reference/deepctr_torch``
and records inputs, weights, pre-sigmoid logits, losses and every parameter gradient into
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them through this file
(bit-exact gather, logits/grads to fp32 round-off).

All functions take ``params``: a dict ``{reference state_dict key -> tensor}`` (key names and
shapes are the reference's) and ``cfg``: a plain-dict model description
(see ``make_cfg``).  dtype follows the params (fp32 for parity, fp64 for the noise floor).
"""
from __future__ import annotations

import itertools
from collections import OrderedDict

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# configuration helpers (plain dicts so they can be stored next to the golden vectors as JSON)
# ----------------------------------------------------------------------------------------------
def sparse_col(name, vocab, dim, embedding_name=None):
    return {"type": "sparse", "name": name, "vocab": int(vocab), "dim": int(dim),
            "embedding_name": embedding_name or name}


def dense_col(name, dimension=1):
    return {"type": "dense", "name": name, "dimension": int(dimension)}


def varlen_col(name, vocab, dim, maxlen, combiner="mean", length_name=None, embedding_name=None):
    return {"type": "varlen", "name": name, "vocab": int(vocab), "dim": int(dim),
            "maxlen": int(maxlen), "combiner": combiner, "length_name": length_name,
            "embedding_name": embedding_name or name}


def make_cfg(model, linear_columns, dnn_columns, **kwargs):
    return {"model": model, "linear_columns": list(linear_columns),
            "dnn_columns": list(dnn_columns), "kwargs": dict(kwargs)}


def feature_index(cfg):
    """Column map of X — restates reference inputs.py:99-123 over linear + dnn columns
    (reference basemodel.py:111-112)."""
    index = OrderedDict()
    cursor = 0
    for col in cfg["linear_columns"] + cfg["dnn_columns"]:
        if col["name"] in index:
            continue
        if col["type"] == "sparse":
            index[col["name"]] = (cursor, cursor + 1)
            cursor += 1
        elif col["type"] == "dense":
            index[col["name"]] = (cursor, cursor + col["dimension"])
            cursor += col["dimension"]
        elif col["type"] == "varlen":
            index[col["name"]] = (cursor, cursor + col["maxlen"])
            cursor += col["maxlen"]
            ln = col.get("length_name")
            if ln is not None and ln not in index:
                index[ln] = (cursor, cursor + 1)
                cursor += 1
        else:
            raise TypeError("Invalid feature column type,got", col["type"])
    return index


def num_input_columns(cfg):
    idx = feature_index(cfg)
    return max(e for _, e in idx.values()) if idx else 0


# ----------------------------------------------------------------------------------------------
# L2: ids, embedding rows, pooled sequences, linear term
# ----------------------------------------------------------------------------------------------
def decode_ids(X, start, end):
    """fp32-encoded ids -> int64 by truncation (reference basemodel.py:369 ``.long()``)."""
    return X[:, start:end].long()


def _pool_sequence(seq_emb, X, col, findex):
    """Masked sum / mean / max over the time axis — restates reference
    inputs.py:141-155 + layers/sequence.py:49-77."""
    s, e = findex[col["name"]]
    T = seq_emb.shape[1]
    if col.get("length_name") is None:
        mask = (decode_ids(X, s, e) != 0).to(seq_emb.dtype)                  # [B, T]
        length = mask.sum(dim=-1, keepdim=True)                               # [B, 1]
    else:
        ls, le = findex[col["length_name"]]
        length_i = decode_ids(X, ls, le)                                      # [B, 1]
        mask = (torch.arange(T)[None, :] < length_i).to(seq_emb.dtype)
        length = length_i.to(seq_emb.dtype)
    mask3 = mask.unsqueeze(2).expand(-1, -1, seq_emb.shape[2])
    mode = col["combiner"]
    if mode == "max":
        return (seq_emb - (1 - mask3) * 1e9).max(dim=1)[0]
    pooled = (seq_emb * mask3).sum(dim=1)
    if mode == "mean":
        # the reference adds eps as a float32 tensor (sequence.py:37,72-73)
        pooled = pooled / (length.to(torch.float32) + torch.tensor([1e-8])).to(pooled.dtype)
    elif mode != "sum":
        raise ValueError("parameter mode should in [sum, mean, max]")
    return pooled


def embedding_rows(params, prefix, X, columns, findex):
    """Per-field row lookups: sparse fields first (list order), then pooled varlen fields —
    the order of reference basemodel.py:368-380.  Returns a list of ``[B, D]`` tensors."""
    out = []
    for col in columns:
        if col["type"] == "sparse":
            s, e = findex[col["name"]]
            table = params[prefix + col["embedding_name"] + ".weight"]
            out.append(F.embedding(decode_ids(X, s, e)[:, 0], table))
    for col in columns:
        if col["type"] == "varlen":
            s, e = findex[col["name"]]
            table = params[prefix + col["embedding_name"] + ".weight"]
            seq = F.embedding(decode_ids(X, s, e), table)                     # [B, T, D]
            out.append(_pool_sequence(seq, X, col, findex))
    return out


def dense_values(X, columns, findex):
    vals = [X[:, findex[c["name"]][0]:findex[c["name"]][1]] for c in columns if c["type"] == "dense"]
    return vals


def linear_logit(params, X, cfg, findex, refine_weight=None):
    """Σ_f w_f[id_f] + X_dense · w  — restates reference basemodel.py:63-92."""
    cols = cfg["linear_columns"]
    dt = _param_dtype(params)
    logit = torch.zeros(X.shape[0], 1, dtype=dt)
    rows = embedding_rows(params, "linear_model.embedding_dict.", X, cols, findex)  # each [B,1]
    if rows:
        cat = torch.cat(rows, dim=-1)                                         # [B, F]
        if refine_weight is not None:
            cat = cat * refine_weight
        logit = logit + cat.sum(dim=-1, keepdim=True)
    dvals = dense_values(X, cols, findex)
    if dvals:
        logit = logit + torch.cat(dvals, dim=-1).to(dt).matmul(params["linear_model.weight"])
    return logit


def combined_dnn_input(rows, dvals, dtype):
    """flatten(cat(E)) ++ cat(dense) — restates reference inputs.py:126-138."""
    parts = []
    if rows:
        parts.append(torch.cat(rows, dim=-1))
    if dvals:
        parts.append(torch.cat(dvals, dim=-1).to(dtype))
    if not parts:
        raise NotImplementedError
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------------------------
# L1: interaction layers
# ----------------------------------------------------------------------------------------------
def fm(E):
    """0.5 Σ_d[(Σ_f e)^2 − Σ_f e^2] — restates reference layers/interaction.py:26-34. E [B,F,D]."""
    s = E.sum(dim=1)
    return 0.5 * (s * s - (E * E).sum(dim=1)).sum(dim=1, keepdim=True)


def bi_interaction(E):
    """0.5 ((Σ_f e)^2 − Σ_f e^2) keeping the embedding axis — restates reference
    layers/interaction.py:54-61.  E [B,F,D] -> [B,1,D]."""
    s = E.sum(dim=1, keepdim=True)
    return 0.5 * (s * s - (E * E).sum(dim=1, keepdim=True))


def afm_layer(params, prefix, rows):
    """Attentional FM — restates reference layers/interaction.py:307-331 (dropout 0).
    rows: list of F tensors [B,D] -> [B,1]."""
    W, b = params[prefix + "attention_W"], params[prefix + "attention_b"]
    h, p = params[prefix + "projection_h"], params[prefix + "projection_p"]
    left, right = [], []
    for r, c in itertools.combinations(rows, 2):
        left.append(r.unsqueeze(1))
        right.append(c.unsqueeze(1))
    ip = torch.cat(left, dim=1) * torch.cat(right, dim=1)                      # [B,P,D]
    att = torch.relu(torch.tensordot(ip, W, dims=([-1], [0])) + b)
    score = torch.softmax(torch.tensordot(att, h, dims=([-1], [0])), dim=1)    # [B,P,1]
    out = (score * ip).sum(dim=1)
    return torch.tensordot(out, p, dims=([-1], [0]))


def interacting_layer(params, prefix, E, head_num, use_res=True, scaling=True):
    """Multi-head self-attention over the fields + residual + relu — restates reference
    layers/interaction.py:352-394 (AutoInt's InteractingLayer as used by DIFM).  E [B,F,D]."""
    D = E.shape[-1]
    dh = D // head_num
    q = torch.tensordot(E, params[prefix + "W_Query"], dims=([-1], [0]))
    k = torch.tensordot(E, params[prefix + "W_key"], dims=([-1], [0]))
    v = torch.tensordot(E, params[prefix + "W_Value"], dims=([-1], [0]))
    q = torch.stack(torch.split(q, dh, dim=2))
    k = torch.stack(torch.split(k, dh, dim=2))
    v = torch.stack(torch.split(v, dh, dim=2))
    ip = torch.einsum("bnik,bnjk->bnij", q, k)
    if scaling:
        ip = ip / dh ** 0.5
    att = torch.softmax(ip, dim=-1)
    res = torch.matmul(att, v)
    res = torch.cat(torch.split(res, 1), dim=-1).squeeze(0)
    if use_res:
        res = res + torch.tensordot(E, params[prefix + "W_Res"], dims=([-1], [0]))
    return torch.relu(res)


def _activation(name, x):
    name = (name or "linear").lower()
    if name == "relu":
        return torch.relu(x)
    if name == "linear":
        return x
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    raise NotImplementedError(name)


def dnn(params, prefix, x, activation="relu"):
    """Linear(+bias) -> activation stack — restates reference layers/core.py:120-134 with
    ``use_bn=False`` and ``dropout_rate=0`` (the parity configuration)."""
    i = 0
    while (prefix + "linears.%d.weight" % i) in params:
        w = params[prefix + "linears.%d.weight" % i]
        b = params[prefix + "linears.%d.bias" % i]
        z = F.linear(x, w, b)
        if (activation or "").lower() == "prelu":        # nn.PReLU() per layer (reference activation.py:61-62)
            x = F.prelu(z, params[prefix + "activation_layers.%d.weight" % i])
        else:
            x = _activation(activation, z)
        i += 1
    return x


def cin(params, prefix, E, layer_size, split_half=True, activation="relu"):
    """Compressed interaction network — restates reference layers/interaction.py:207-248.
    Returns ``[B, featuremap_num]``."""
    B, M, D = E.shape
    hidden = [E]
    direct = []
    n_layers = len(layer_size)
    for k, size in enumerate(layer_size):
        prev = hidden[-1]
        outer = prev.unsqueeze(2) * E.unsqueeze(1)                           # [B,H,M,D]  (h-major)
        P = outer.reshape(B, prev.shape[1] * M, D)
        W = params[prefix + "conv1ds.%d.weight" % k]                           # [N, H*M, 1]
        b = params[prefix + "conv1ds.%d.bias" % k]
        Z = torch.einsum("nk,bkd->bnd", W[:, :, 0], P) + b[None, :, None]
        Y = _activation(activation, Z)
        if split_half:
            if k != n_layers - 1:
                nxt, dc = Y[:, : size // 2], Y[:, size // 2:]
            else:
                nxt, dc = None, Y
        else:
            nxt, dc = Y, Y
        direct.append(dc)
        hidden.append(nxt)
    return torch.cat(direct, dim=1).sum(dim=-1)


def crossnet(params, prefix, x, parameterization="vector"):
    """restates reference layers/interaction.py:438-453.  x [B, in]."""
    kernels = params[prefix + "kernels"]
    bias = params[prefix + "bias"]
    x0 = x
    xl = x
    for l in range(kernels.shape[0]):
        if parameterization == "vector":
            s = xl.matmul(kernels[l])                                         # [B,1]
            xl = x0 * s + bias[l][:, 0][None, :] + xl
        elif parameterization == "matrix":
            xl = x0 * (xl.matmul(kernels[l].t()) + bias[l][:, 0][None, :]) + xl
        else:
            raise ValueError("parameterization should be 'vector' or 'matrix'")
    return xl


def crossnet_mix(params, prefix, x):
    """restates reference layers/interaction.py:499-534.  x [B, in]."""
    U = params[prefix + "U_list"]
    V = params[prefix + "V_list"]
    C = params[prefix + "C_list"]
    bias = params[prefix + "bias"]
    L, E = U.shape[0], U.shape[1]
    x0 = x
    xl = x
    for l in range(L):
        outs, gates = [], []
        for e in range(E):
            gates.append(xl.matmul(params[prefix + "gating.%d.weight" % e].t()))     # [B,1]
            v = torch.tanh(xl.matmul(V[l, e]))                                # [B,r]
            v = torch.tanh(v.matmul(C[l, e].t()))
            uv = v.matmul(U[l, e].t())                                        # [B,in]
            outs.append(x0 * (uv + bias[l][:, 0][None, :]))
        outs = torch.stack(outs, dim=2)                                       # [B,in,E]
        g = torch.softmax(torch.stack(gates, dim=1), dim=1)                   # [B,E,1]
        xl = torch.matmul(outs, g)[:, :, 0] + xl
    return xl


def senet(params, prefix, E):
    """restates reference layers/interaction.py:93-101."""
    Z = E.mean(dim=-1)
    A = torch.relu(F.linear(Z, params[prefix + "excitation.0.weight"]))
    A = torch.relu(F.linear(A, params[prefix + "excitation.2.weight"]))
    return E * A.unsqueeze(2)


def bilinear(params, prefix, E, bilinear_type="interaction"):
    """restates reference layers/interaction.py:140-156.  Returns [B, F(F-1)/2, D]."""
    Fn = E.shape[1]
    outs = []
    for p, (i, j) in enumerate(itertools.combinations(range(Fn), 2)):
        if bilinear_type == "all":
            W = params[prefix + "bilinear.weight"]
        elif bilinear_type == "each":
            W = params[prefix + "bilinear.%d.weight" % i]
        elif bilinear_type == "interaction":
            W = params[prefix + "bilinear.%d.weight" % p]
        else:
            raise NotImplementedError
        outs.append(F.linear(E[:, i], W) * E[:, j])
    return torch.stack(outs, dim=1)


# ----------------------------------------------------------------------------------------------
# L3: the five drop-in models (pre-sigmoid logit)
# ----------------------------------------------------------------------------------------------
def _param_dtype(params):
    for v in params.values():
        if v.is_floating_point():
            return v.dtype
    return torch.float32


def model_logit(cfg, params, X):
    """Pre-sigmoid, pre-``out.bias`` sum of branch logits, then ``+ out.bias``.

    Restates reference models/deepfm.py:67-86, xdeepfm.py:79-107, fibinet.py:76-102,
    dcn.py:74-96, dcnmix.py:80-102 and layers/core.py:154-158."""
    kw = cfg["kwargs"]
    model = cfg["model"]
    findex = feature_index(cfg)
    dt = _param_dtype(params)
    Xd = X.to(dt)
    rows = embedding_rows(params, "embedding_dict.", X, cfg["dnn_columns"], findex)
    dvals = dense_values(Xd, cfg["dnn_columns"], findex)
    lin = linear_logit(params, Xd, cfg, findex)
    act = kw.get("dnn_activation", "relu")
    hidden = tuple(kw.get("dnn_hidden_units", ()))
    has_dnn_cols = len(cfg["dnn_columns"]) > 0

    if model == "DeepFM":
        logit = lin
        if kw.get("use_fm", True) and rows:
            logit = logit + fm(torch.stack(rows, dim=1))
        if has_dnn_cols and len(hidden) > 0:
            h = dnn(params, "dnn.", combined_dnn_input(rows, dvals, dt), act)
            logit = logit + F.linear(h, params["dnn_linear.weight"])
    elif model == "xDeepFM":
        logit = lin
        cin_size = tuple(kw.get("cin_layer_size", (256, 128)))
        if len(cin_size) > 0 and has_dnn_cols:
            c = cin(params, "cin.", torch.stack(rows, dim=1), cin_size,
                    kw.get("cin_split_half", True), kw.get("cin_activation", "relu"))
            logit = logit + F.linear(c, params["cin_linear.weight"])
        if has_dnn_cols and len(hidden) > 0:
            h = dnn(params, "dnn.", combined_dnn_input(rows, dvals, dt), act)
            logit = logit + F.linear(h, params["dnn_linear.weight"])
    elif model == "FiBiNET":
        E = torch.stack(rows, dim=1)
        btype = kw.get("bilinear_type", "interaction")
        se = senet(params, "SE.", E)
        p1 = bilinear(params, "Bilinear.", se, btype)
        p2 = bilinear(params, "Bilinear.", E, btype)
        both = torch.cat((p1, p2), dim=1)                                     # SENET branch first
        flat = [both.reshape(both.shape[0], -1)]
        h = dnn(params, "dnn.", combined_dnn_input(flat, dvals, dt), act)
        dnn_logit = F.linear(h, params["dnn_linear.weight"])
        if len(cfg["linear_columns"]) > 0 and has_dnn_cols:
            logit = lin + dnn_logit
        elif len(cfg["linear_columns"]) == 0:
            logit = dnn_logit
        else:
            logit = lin
    elif model in ("DCN", "DCNMix"):
        logit = lin
        x = combined_dnn_input(rows, dvals, dt)
        cross_num = kw.get("cross_num", 2)
        if model == "DCN":
            cross = lambda t: crossnet(params, "crossnet.", t, kw.get("cross_parameterization", "vector"))
        else:
            cross = lambda t: crossnet_mix(params, "crossnet.", t)
        if len(hidden) > 0 and cross_num > 0:
            stack = torch.cat((cross(x), dnn(params, "dnn.", x, act)), dim=-1)
            logit = logit + F.linear(stack, params["dnn_linear.weight"])
        elif len(hidden) > 0:
            logit = logit + F.linear(dnn(params, "dnn.", x, act), params["dnn_linear.weight"])
        elif cross_num > 0:
            logit = logit + F.linear(cross(x), params["dnn_linear.weight"])
    elif model == "WDL":
        # restates reference models/wdl.py:62-78
        logit = lin
        if has_dnn_cols and len(hidden) > 0:
            h = dnn(params, "dnn.", combined_dnn_input(rows, dvals, dt), act)
            logit = logit + F.linear(h, params["dnn_linear.weight"])
    elif model == "NFM":
        # restates reference models/nfm.py:62-80 (bi_dropout = 0)
        bi = bi_interaction(torch.stack(rows, dim=1))[:, 0]
        h = dnn(params, "dnn.", combined_dnn_input([bi], dvals, dt), act)
        logit = lin + F.linear(h, params["dnn_linear.weight"])
    elif model == "AFM":
        # restates reference models/afm.py:55-70
        logit = lin
        if rows:
            if kw.get("use_attention", True):
                logit = logit + afm_layer(params, "fm.", rows)
            else:
                logit = logit + fm(torch.stack(rows, dim=1))
    elif model in ("IFM", "DIFM"):
        # restates reference models/ifm.py:69-91 and models/difm.py:82-112
        E = torch.stack(rows, dim=1)
        nf = len(rows)
        flat = combined_dnn_input(rows, [], dt)
        if model == "IFM":
            h = dnn(params, "factor_estimating_net.", flat, act)
            m = nf * torch.softmax(F.linear(h, params["transform_weight_matrix_P.weight"]), dim=1)
        else:
            att = interacting_layer(params, "vector_wise_net.", E, kw.get("att_head_num", 4), kw.get("att_res", True))
            m_vec = F.linear(att.reshape(att.shape[0], -1), params["transform_matrix_P_vec.weight"])
            m_bit = F.linear(dnn(params, "bit_wise_net.", flat, act), params["transform_matrix_P_bit.weight"])
            m = m_vec + m_bit
        logit = linear_logit(params, Xd, cfg, findex, refine_weight=m) + fm(E * m.unsqueeze(-1))
    else:
        raise NotImplementedError(model)
    return logit + params["out.bias"]


def model_forward(cfg, params, X):
    logit = model_logit(cfg, params, X)
    return torch.sigmoid(logit) if cfg["kwargs"].get("task", "binary") == "binary" else logit


def regularization_loss(cfg, params):
    """Σ_groups l2·Σw² — restates reference basemodel.py:412-428 for the groups each model
    registers (basemodel.py:126-127, deepfm.py:62-64, xdeepfm.py:57-60,74-75, dcn.py:68-71,
    dcnmix.py:71-76)."""
    kw = cfg["kwargs"]
    model = cfg["model"]
    total = torch.zeros(1, dtype=_param_dtype(params))

    def add(keys, l2):
        nonlocal total
        if l2 and l2 > 0:
            for k in keys:
                total = total + (l2 * params[k] * params[k]).sum()

    add([k for k in params if k.startswith("embedding_dict.")], kw.get("l2_reg_embedding", 1e-5))
    # DCN / DCNMix never forward l2_reg_linear to BaseModel (dcn.py:49-51) -> default 1e-5
    base_l2_lin = 1e-5 if model in ("DCN", "DCNMix") else kw.get("l2_reg_linear", 1e-5)
    add([k for k in params if k.startswith("linear_model.")], base_l2_lin)
    dnn_w = [k for k in params if k.startswith("dnn.") and "weight" in k and "bn" not in k]
    if model == "AFM" and kw.get("use_attention", True):
        add(["fm.attention_W"], kw.get("l2_reg_att", 1e-5))
    if model == "IFM":
        add([k for k in params if k.startswith("factor_estimating_net.") and "weight" in k and "bn" not in k] +
            ["transform_weight_matrix_P.weight"], kw.get("l2_reg_dnn", 0))
    if model == "DIFM":
        add([k for k in params if (k.startswith("vector_wise_net.") or k.startswith("bit_wise_net.")) and
             "weight" in k and "bn" not in k] + ["transform_matrix_P_vec.weight", "transform_matrix_P_bit.weight"],
            kw.get("l2_reg_dnn", 0))
    if model in ("DeepFM", "xDeepFM", "WDL", "NFM") and "dnn_linear.weight" in params:
        add(dnn_w, kw.get("l2_reg_dnn", 0))
        add(["dnn_linear.weight"], kw.get("l2_reg_dnn", 0))
    if model == "xDeepFM":
        add([k for k in params if k.startswith("cin.") and "weight" in k], kw.get("l2_reg_cin", 0))
    if model in ("DCN", "DCNMix"):
        add(dnn_w, kw.get("l2_reg_dnn", 0))
        add(["dnn_linear.weight"], kw.get("l2_reg_linear", 1e-5))
        if model == "DCN":
            add(["crossnet.kernels"], kw.get("l2_reg_cross", 1e-5))
        else:
            add(["crossnet.U_list", "crossnet.V_list", "crossnet.C_list"], kw.get("l2_reg_cross", 1e-5))
    return total


def loss_and_grads(cfg, params, X, y, with_reg=False):
    """BCE(sum) forward + autograd backward (reference basemodel.py:245-261).

    Returns ``(logit, y_pred, loss, {key: grad})`` — grads are dense, like the reference's."""
    leaves = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
    logit = model_logit(cfg, leaves, X)
    task = cfg["kwargs"].get("task", "binary")
    y_pred = torch.sigmoid(logit) if task == "binary" else logit
    yv = y.reshape(-1).to(y_pred.dtype)
    if task == "binary":
        loss = F.binary_cross_entropy(y_pred.squeeze(-1), yv, reduction="sum")
    else:
        loss = F.mse_loss(y_pred.squeeze(-1), yv, reduction="sum")
    total = loss + regularization_loss(cfg, leaves).sum() if with_reg else loss
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()
             if v.requires_grad}
    return logit.detach(), y_pred.detach(), loss.detach(), grads


# ----------------------------------------------------------------------------------------------
# synthetic Criteo-shaped inputs shared by tests / bench
# ----------------------------------------------------------------------------------------------
def synthetic_batch(cfg, batch, seed=2026, zipf_alpha=None):
    """``X[B, C]`` with fp32-encoded ids (uniform, or Zipf when ``zipf_alpha``) and dense
    values in [0,1), labels Bernoulli(0.25)."""
    g = torch.Generator().manual_seed(seed)
    findex = feature_index(cfg)
    C = num_input_columns(cfg)
    X = torch.zeros(batch, C, dtype=torch.float32)
    seen = set()
    for col in cfg["linear_columns"] + cfg["dnn_columns"]:
        if col["name"] in seen:
            continue
        seen.add(col["name"])
        s, e = findex[col["name"]]
        if col["type"] == "dense":
            X[:, s:e] = torch.rand(batch, e - s, generator=g)
        else:
            V = col["vocab"]
            if zipf_alpha:
                # inverse CDF of the continuous power law p(k) ~ k^-alpha on [1, V+1)
                u = torch.rand(batch, e - s, generator=g, dtype=torch.float64)
                a1 = 1.0 - zipf_alpha
                k = (1.0 + u * (float(V + 1) ** a1 - 1.0)).pow(1.0 / a1)
                ids = (k.floor() - 1).clamp(0, V - 1).float()
            else:
                ids = torch.randint(0, V, (batch, e - s), generator=g).float()
            if col["type"] == "varlen":
                lens = torch.randint(1, e - s + 1, (batch, 1), generator=g)
                pos = torch.arange(e - s)[None, :]
                ids = torch.where(pos < lens, ids.clamp_min(1), torch.zeros_like(ids))
                ln = col.get("length_name")
                if ln is not None:
                    ls, le = findex[ln]
                    X[:, ls:le] = lens.float()
            X[:, s:e] = ids
    y = (torch.rand(batch, generator=g) < 0.25).float()
    return X, y
