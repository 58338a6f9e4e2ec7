"""Plain-dict model descriptions -> model instances (shared by bench.py, the tests and sharded.py).

A description is ``{"model": "DeepFM", "linear_columns": [...], "dnn_columns": [...], "kwargs": {...}}``
with columns ``{"type": "sparse"|"dense"|"varlen", "name", "vocab", "dim", ...}`` — JSON-serialisable,
stored next to the golden vectors."""
from __future__ import annotations

from .inputs import DenseFeat, SparseFeat, VarLenSparseFeat


def columns_from_cfg(cols):
    out = []
    for c in cols:
        if c["type"] == "sparse":
            out.append(SparseFeat(c["name"], c["vocab"], embedding_dim=c["dim"], embedding_name=c["embedding_name"]))
        elif c["type"] == "dense":
            out.append(DenseFeat(c["name"], c["dimension"]))
        elif c["type"] == "varlen":
            sf = SparseFeat(c["name"], c["vocab"], embedding_dim=c["dim"], embedding_name=c["embedding_name"])
            out.append(VarLenSparseFeat(sf, maxlen=c["maxlen"], combiner=c["combiner"], length_name=c["length_name"]))
        else:
            raise TypeError("Invalid feature column type,got", c["type"])
    return out


def model_from_cfg(cfg, device="cpu", **extra):
    from . import models
    cls = getattr(models, cfg["model"])
    kw = dict(cfg["kwargs"])
    for k in ("dnn_hidden_units", "cin_layer_size"):
        if k in kw:
            kw[k] = tuple(kw[k])
    kw.update(extra)
    return cls(columns_from_cfg(cfg["linear_columns"]), columns_from_cfg(cfg["dnn_columns"]), device=device, **kw)
