"""BASELINE config #1: the reference example pipeline (DeepFM on criteo_sample, Adam, 3 epochs)
re-run through deepctr_torch_b200's fit()/predict() on the GPU and compared with the history and
predictions the unmodified reference produced on CPU (tests/golden/fit_criteo_sample.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, build_model

pytestmark = pytest.mark.gpu


def test_fit_predict_matches_reference_run():
    z = np.load(os.path.join(GOLDEN_DIR, "fit_criteo_sample.npz"))
    cfg = json.loads(str(z["cfg"]))
    names = json.loads(str(z["names"]))
    X = z["X"]
    x = {n: X[:, i].copy() for i, n in enumerate(names)}
    y = z["y"]
    m = build_model(cfg, "cuda:0")
    # same starting point as the reference run (bit-identical CPU initialisation is checked in
    # tests/test_host_logic.py; on a CUDA device `linear_model.weight` is drawn by the CUDA RNG in
    # the reference as well)
    m.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init/")})
    m.compile("adam", "binary_crossentropy", metrics=["binary_crossentropy"])
    hist = m.fit(x, y, batch_size=64, epochs=3, verbose=0, validation_split=0.2, shuffle=False)
    ref_hist = json.loads(str(z["history"]))
    for k, vals in ref_hist.items():
        assert np.allclose(hist.history[k], vals, rtol=2e-4, atol=1e-6), (k, hist.history[k], vals)
    pred = m.predict(x, batch_size=64)
    assert pred.shape == z["pred"].shape and pred.dtype == np.float64
    assert np.abs(pred - z["pred"]).max() <= 2e-4
    for k, v in m.state_dict().items():
        ref = z["final/" + k]
        assert np.abs(v.cpu().numpy() - ref).max() <= 5e-4 * max(1e-3, np.abs(ref).max()), k


def test_fit_with_callbacks_and_rowwise_sgd(tmp_path):
    from deepctr_torch_b200.callbacks import EarlyStopping, ModelCheckpoint
    z = np.load(os.path.join(GOLDEN_DIR, "fit_criteo_sample.npz"))
    cfg = json.loads(str(z["cfg"]))
    names = json.loads(str(z["names"]))
    x = {n: z["X"][:, i].copy() for i, n in enumerate(names)}
    m = build_model(cfg, "cuda:0", table_grad="rowwise", l2_reg_linear=0, l2_reg_embedding=0)
    m.compile("adagrad", "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    ck = ModelCheckpoint(str(tmp_path / "w.ckpt"), monitor="val_binary_crossentropy", save_best_only=True,
                         save_weights_only=True)
    es = EarlyStopping(monitor="val_binary_crossentropy", patience=0, mode="min")
    hist = m.fit(x, z["y"], batch_size=100, epochs=2, verbose=0, validation_split=0.5, callbacks=[es, ck])
    assert "val_auc" in hist.history and len(hist.history["loss"]) >= 1
    state = torch.load(tmp_path / "w.ckpt")
    m.load_state_dict(state)


def test_graphed_step_matches_eager():
    """The CUDA-graph replay of one step produces the same loss and gradients as the eager step."""
    from helpers import load_case
    c = load_case("deepfm_criteo_shape")
    results = []
    for graphed in (False, True):
        m = build_model(c["cfg"], "cuda:0", table_grad="rowwise")
        m.load_state_dict(c["state"])
        m.train()
        X, y = c["X"].cuda(), c["y"].cuda()
        if graphed:
            step = m.make_graphed_step(X.shape[0])
            loss = step(X, y)
            loss = step(X, y)            # replaying twice must not accumulate
        else:
            y_pred = m(X)
            loss = torch.nn.functional.binary_cross_entropy(y_pred.squeeze(1), y, reduction="sum")
            loss.backward()
        m.check_ids()
        grads = {k: (p.grad.to_dense() if p.grad.is_sparse else p.grad).detach().cpu().clone()
                 for k, p in m.named_parameters()}
        results.append((float(loss), grads))
    assert abs(results[0][0] - results[1][0]) <= 1e-5 * abs(results[0][0])
    for k in results[0][1]:
        a, b = results[0][1][k], results[1][1][k]
        assert (a - b).abs().max() <= 1e-5 * max(1e-6, a.abs().max()), k
