"""Feature-column API (the boundary the hot path is entered through).

Mirrors the reference's column types and helpers so user code written against
``deepctr_torch.inputs`` keeps working (reference ``deepctr_torch/inputs.py``):

* ``SparseFeat``        — reference inputs.py:20-38 (``embedding_dim="auto"`` -> ``6*int(V**0.25)``,
                          ``embedding_name`` defaults to ``name``, hashing by ``name``)
* ``VarLenSparseFeat``  — reference inputs.py:41-77
* ``DenseFeat``         — reference inputs.py:80-87
* ``get_feature_names`` — reference inputs.py:90-92
* ``build_input_features`` — reference inputs.py:99-123: column map of the single ``X[B, C]``
  matrix every model consumes (SparseFeat = 1 column, DenseFeat = ``dimension`` columns,
  VarLenSparseFeat = ``maxlen`` columns followed by an optional length column).

What is *not* here on purpose: ``create_embedding_matrix`` returning 52 ``nn.Embedding`` modules
and the per-feature ``embedding_lookup`` loops.  Tables live in
``deepctr_torch_b200.embedding.TableDict`` and every lookup goes through one fused CUDA gather
(``deepctr_torch_b200.ops.fused_input``).
"""
from __future__ import annotations

from collections import OrderedDict, namedtuple

DEFAULT_GROUP_NAME = "default_group"


class SparseFeat(namedtuple("SparseFeat", ["name", "vocabulary_size", "embedding_dim", "use_hash",
                                           "dtype", "embedding_name", "group_name"])):
    """One categorical column: ids in ``[0, vocabulary_size)`` stored in one column of ``X``."""
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32",
                embedding_name=None, group_name=DEFAULT_GROUP_NAME):
        if embedding_name is None:
            embedding_name = name
        if embedding_dim == "auto":
            embedding_dim = 6 * int(pow(vocabulary_size, 0.25))
        if use_hash:
            print("Notice! Feature Hashing on the fly currently is not supported in torch version,"
                  "you can use tensorflow version!")
        return super().__new__(cls, name, vocabulary_size, embedding_dim, use_hash, dtype,
                               embedding_name, group_name)

    def __hash__(self):
        return hash(self.name)


class VarLenSparseFeat(namedtuple("VarLenSparseFeat", ["sparsefeat", "maxlen", "combiner",
                                                       "length_name"])):
    """A padded multi-value categorical column pooled with ``combiner`` (sum | mean | max)."""
    __slots__ = ()

    def __new__(cls, sparsefeat, maxlen, combiner="mean", length_name=None):
        return super().__new__(cls, sparsefeat, maxlen, combiner, length_name)

    name = property(lambda self: self.sparsefeat.name)
    vocabulary_size = property(lambda self: self.sparsefeat.vocabulary_size)
    embedding_dim = property(lambda self: self.sparsefeat.embedding_dim)
    use_hash = property(lambda self: self.sparsefeat.use_hash)
    dtype = property(lambda self: self.sparsefeat.dtype)
    embedding_name = property(lambda self: self.sparsefeat.embedding_name)
    group_name = property(lambda self: self.sparsefeat.group_name)

    def __hash__(self):
        return hash(self.name)


class DenseFeat(namedtuple("DenseFeat", ["name", "dimension", "dtype"])):
    """``dimension`` float columns of ``X`` used as they are."""
    __slots__ = ()

    def __new__(cls, name, dimension=1, dtype="float32"):
        return super().__new__(cls, name, dimension, dtype)

    def __hash__(self):
        return hash(self.name)


def build_input_features(feature_columns):
    """``OrderedDict{feature_name: (start, end)}`` — the column map of ``X``.

    Duplicate names are laid out once (first occurrence wins), exactly like the reference
    (inputs.py:107-108); an unknown column type raises ``TypeError`` (inputs.py:122).
    """
    features = OrderedDict()
    cursor = 0
    for feat in feature_columns:
        if isinstance(feat, SparseFeat):
            width, extra = 1, None
        elif isinstance(feat, DenseFeat):
            width, extra = feat.dimension, None
        elif isinstance(feat, VarLenSparseFeat):
            width, extra = feat.maxlen, feat.length_name
        else:
            raise TypeError("Invalid feature column type,got", type(feat))
        if feat.name in features:
            continue
        features[feat.name] = (cursor, cursor + width)
        cursor += width
        if extra is not None and extra not in features:
            features[extra] = (cursor, cursor + 1)
            cursor += 1
    return features


def get_feature_names(feature_columns):
    return list(build_input_features(feature_columns).keys())


def split_columns(feature_columns):
    """(sparse, dense, varlen) sub-lists in their original order."""
    cols = list(feature_columns) if feature_columns else []
    sparse = [c for c in cols if isinstance(c, SparseFeat)]
    dense = [c for c in cols if isinstance(c, DenseFeat)]
    varlen = [c for c in cols if isinstance(c, VarLenSparseFeat)]
    return sparse, dense, varlen


def compute_input_dim(feature_columns, include_sparse=True, include_dense=True, feature_group=False):
    """Width of ``combined_dnn_input`` (reference basemodel.py:382-400)."""
    sparse, dense, varlen = split_columns(feature_columns)
    emb_cols = [c for c in (feature_columns or []) if isinstance(c, (SparseFeat, VarLenSparseFeat))]
    dense_dim = sum(c.dimension for c in dense)
    sparse_dim = len(emb_cols) if feature_group else sum(c.embedding_dim for c in emb_cols)
    return (sparse_dim if include_sparse else 0) + (dense_dim if include_dense else 0)
