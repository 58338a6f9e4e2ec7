"""Host-side (CPU) logic of the drop-in boundary: column API, column map, state_dict layout,
initialisation parity, error conventions, callbacks.  No kernel is launched here."""
import json
import os

import numpy as np
import pytest
import torch

from deepctr_torch_b200 import callbacks as cb
from deepctr_torch_b200.inputs import (DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features,
                                       get_feature_names)
from deepctr_torch_b200.models import DCN, DeepFM, xDeepFM
from helpers import GOLDEN_DIR, MODEL_CASES, build_model, load_case
from oracle import ctr_oracle as O


def test_sparsefeat_defaults_and_hash():
    f = SparseFeat("c1", 1000, embedding_dim="auto")
    assert f.embedding_dim == 6 * int(pow(1000, 0.25))
    assert f.embedding_name == "c1" and f.group_name == "default_group" and f.dtype == "int32"
    assert hash(f) == hash("c1")
    assert SparseFeat("a", 10).embedding_dim == 4
    v = VarLenSparseFeat(SparseFeat("s", 50, 8), maxlen=5, combiner="sum", length_name="s_len")
    assert (v.name, v.vocabulary_size, v.embedding_dim, v.maxlen, v.combiner) == ("s", 50, 8, 5, "sum")
    d = DenseFeat("x", 3)
    assert d.dimension == 3 and d.dtype == "float32"


def test_build_input_features_layout():
    cols = [SparseFeat("a", 10), DenseFeat("x", 3), VarLenSparseFeat(SparseFeat("s", 50, 8), 4, length_name="sl"),
            SparseFeat("a", 10), SparseFeat("b", 5)]
    idx = build_input_features(cols)
    assert list(idx.items()) == [("a", (0, 1)), ("x", (1, 4)), ("s", (4, 8)), ("sl", (8, 9)), ("b", (9, 10))]
    assert get_feature_names(cols) == ["a", "x", "s", "sl", "b"]
    with pytest.raises(TypeError):
        build_input_features(["not a column"])


@pytest.mark.parametrize("name", MODEL_CASES)
def test_feature_index_and_state_dict_match_reference(name):
    c = load_case(name)
    m = build_model(c["cfg"], "cpu")
    assert list(m.feature_index.items()) == list(O.feature_index(c["cfg"]).items())
    sd = m.state_dict()
    assert list(sd.keys()) == list(c["state"].keys())          # same keys, same order
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(c["state"][k].shape), k
    m.load_state_dict(c["state"])                              # reference checkpoint loads as is


def test_initialisation_is_bit_identical_to_reference():
    """Same seed -> same RNG consumption order -> identical initial weights (reference
    basemodel.py:100-127, deepfm.py:52-64)."""
    z = np.load(os.path.join(GOLDEN_DIR, "fit_criteo_sample.npz"))
    cfg = json.loads(str(z["cfg"]))
    m = build_model(cfg, "cpu")
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), z["init/" + k]), k


def test_error_conventions():
    cols = [SparseFeat("a", 10, 4), SparseFeat("b", 10, 4), DenseFeat("x", 1)]
    with pytest.raises(ValueError):
        xDeepFM(cols, cols, cin_layer_size=(7, 4), cin_split_half=True)      # odd size with split_half
    with pytest.raises(ValueError):
        DCN(cols, cols, cross_parameterization="bogus")
    with pytest.raises(ValueError):
        DCN(cols, cols, dnn_hidden_units=())                                 # DNN: hidden_units is empty
    with pytest.raises(ValueError):
        DeepFM(cols, cols, device="cpu", gpus=[0])                           # gpus[0] must match device
    m = DeepFM(cols, cols)
    with pytest.raises(NotImplementedError):
        m.compile("bogus_optimizer", "binary_crossentropy")
    with pytest.raises(NotImplementedError):
        m.compile("adam", "bogus_loss")
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(4, 3))                                                 # no CPU path, fails loudly
    mixed = [SparseFeat("a", 10, 4), SparseFeat("b", 10, 8)]
    from deepctr_torch_b200.models import FiBiNET
    with pytest.raises(ValueError):
        FiBiNET(mixed, mixed)


def test_regularization_groups_match_reference():
    for name in ("deepfm_small", "dcn_vector", "dcnmix_small", "xdeepfm_small"):
        c = load_case(name)
        m = build_model(c["cfg"], "cpu")
        m.load_state_dict(c["state"])
        reg = float(m.get_regularization_loss())
        assert abs(reg - c["reg_loss"]) <= 1e-6 * max(1.0, abs(c["reg_loss"])), name


def test_callbacks_early_stopping_and_history(tmp_path):
    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.stop_training = False

    model = Dummy()
    es = cb.EarlyStopping(monitor="val_loss", patience=1, mode="min")
    hist = cb.History()
    ck = cb.ModelCheckpoint(str(tmp_path / "m_{epoch:02d}.ckpt"), monitor="val_loss", save_best_only=True,
                            save_weights_only=True)
    cl = cb.CallbackList([es, hist, ck])
    cl.set_model(model)
    cl.on_train_begin()
    for epoch, v in enumerate([1.0, 0.8, 0.9, 0.95]):
        cl.on_epoch_begin(epoch)
        cl.on_epoch_end(epoch, {"val_loss": v})
        if model.stop_training:
            break
    assert model.stop_training and es.stopped_epoch == 2
    assert hist.history["val_loss"] == [1.0, 0.8, 0.9]
    assert sorted(p.name for p in tmp_path.iterdir()) == ["m_01.ckpt", "m_02.ckpt"]


def test_pickle_roundtrip_drops_launch_metadata(tmp_path):
    c = load_case("deepfm_small")
    m = build_model(c["cfg"], "cpu")
    path = tmp_path / "model.h5"
    torch.save(m, path)
    m2 = torch.load(path, weights_only=False)
    assert m2._plan is None
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())
