"""Pin the oracle (oracle/ctr_oracle.py) against golden vectors recorded from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import ctr_oracle as O
from helpers import MODEL_CASES, load_case, load_layers, rel_err

TOL = 2e-6   # fp32 round-off between two orderings of the same fp32 arithmetic


@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_forward_backward_matches_reference(name):
    c = load_case(name)
    logit, y_pred, loss, grads = O.loss_and_grads(c["cfg"], c["state"], c["X"], c["y"])
    assert rel_err(logit, c["logit"]) < TOL
    assert rel_err(y_pred, c["y_pred"]) < TOL
    assert abs(float(loss) - c["loss"]) / abs(c["loss"]) < TOL
    assert set(grads) == set(c["grad"])
    for k, g in c["grad"].items():
        assert grads[k].shape == g.shape
        assert rel_err(grads[k], g) < 2e-5, k
    reg = float(O.regularization_loss(c["cfg"], c["state"]).sum())
    assert abs(reg - c["reg_loss"]) <= 1e-6 * max(1.0, abs(c["reg_loss"]))


def test_gather_is_bit_exact():
    c = load_case("deepfm_criteo_shape")
    cfg, P, X = c["cfg"], c["state"], c["X"]
    findex = O.feature_index(cfg)
    rows = O.embedding_rows(P, "embedding_dict.", X, cfg["dnn_columns"], findex)
    for col, r in zip([k for k in cfg["dnn_columns"] if k["type"] == "sparse"], rows):
        ids = X[:, findex[col["name"]][0]].long()
        assert torch.equal(r, P["embedding_dict." + col["embedding_name"] + ".weight"][ids])


def _params(layer, prefix=""):
    return {prefix + k[len("param/"):]: v for k, v in layer.items() if k.startswith("param/")}


LAYER_FUNCS = {
    "fm": lambda L, x: O.fm(x),
    "fm_small": lambda L, x: O.fm(x),
    "senet": lambda L, x: O.senet(L, "", x),
    "senet_f2": lambda L, x: O.senet(L, "", x),
    "bilinear_all": lambda L, x: O.bilinear(L, "", x, "all"),
    "bilinear_each": lambda L, x: O.bilinear(L, "", x, "each"),
    "bilinear_interaction": lambda L, x: O.bilinear(L, "", x, "interaction"),
    "cin_split": lambda L, x: O.cin(L, "", x, (32, 16), True, "relu"),
    "cin_nosplit_linear": lambda L, x: O.cin(L, "", x, (6, 4, 3), False, "linear"),
    "cin_one": lambda L, x: O.cin(L, "", x, (7,), True, "relu"),
    "cross_vector": lambda L, x: O.crossnet(L, "", x, "vector"),
    "cross_matrix": lambda L, x: O.crossnet(L, "", x, "matrix"),
    "cross_mix": lambda L, x: O.crossnet_mix(L, "", x),
    "dnn_relu": lambda L, x: O.dnn(L, "", x, "relu"),
    "dnn_sigmoid": lambda L, x: O.dnn(L, "", x, "sigmoid"),
    "dnn_linear_act": lambda L, x: O.dnn(L, "", x, "linear"),
}


LAYER_FUNCS_R2 = {
    "bipool": lambda L, x: O.bi_interaction(x),
    "afm": lambda L, x: O.afm_layer(L, "", [x[:, i] for i in range(x.shape[1])]),
    "afm_small": lambda L, x: O.afm_layer(L, "", [x[:, i] for i in range(x.shape[1])]),
    "interacting": lambda L, x: O.interacting_layer(L, "", x, 4, True, True),
    "interacting_nores": lambda L, x: O.interacting_layer(L, "", x, 2, False, False),
}


@pytest.mark.parametrize("tag", sorted(LAYER_FUNCS) + sorted(LAYER_FUNCS_R2))
def test_layer_matches_reference(tag):
    if tag in LAYER_FUNCS_R2:
        layer = load_layers("layers_r2.npz")[tag]
        LAYER_FUNCS[tag] = LAYER_FUNCS_R2[tag]
    else:
        layer = load_layers()[tag]
    P = {k: v.clone().requires_grad_(True) for k, v in _params(layer).items()}
    x = layer["x"].clone().requires_grad_(True)
    y = LAYER_FUNCS[tag](P, x)
    assert y.shape == layer["y"].shape
    assert rel_err(y, layer["y"]) < TOL
    (y * layer["w"]).sum().backward()
    assert rel_err(x.grad, layer["dx"]) < 1e-5
    for k, p in P.items():
        ref = layer["dparam/" + k]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert rel_err(got, ref) < 1e-5, k


def test_fp64_noise_floor():
    """The fp32 reference itself sits 3e-7..3e-6 away from an fp64 evaluation."""
    c = load_case("deepfm_criteo_shape")
    P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in c["state"].items()}
    logit64 = O.model_logit(c["cfg"], P64, c["X"])
    assert rel_err(c["logit"], logit64) < 1e-5
