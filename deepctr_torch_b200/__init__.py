"""deepctr_torch_b200 — the DeepCTR-Torch embedding + interaction hot path on B200 (sm_100a).

Drop-in surface (same names as ``deepctr_torch``): ``inputs.{SparseFeat, DenseFeat,
VarLenSparseFeat, get_feature_names, build_input_features}``, ``models.{DeepFM, xDeepFM,
FiBiNET, DCN, DCNMix}``, ``callbacks.{EarlyStopping, ModelCheckpoint, History}``.
All arithmetic of the path runs in hand-written CUDA kernels behind the C ABI of
``libctr_b200.so`` (``include/ctr_b200.h``); importing this package does not load the library,
the first kernel call does (and raises if it is missing).
"""
from . import callbacks, inputs  # noqa: F401
from .inputs import DenseFeat, SparseFeat, VarLenSparseFeat, get_feature_names  # noqa: F401

__version__ = "0.1.0"
