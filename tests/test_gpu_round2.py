"""Round-2 GPU parity: BASELINE-size logits (ReLU towers, uniform and Zipf ids) against the CPU oracle,
the adjacent models' layers against reference goldens, the fused row-wise optimizer against the
reference's own fit() runs, int32 id cells, the cuda:1 device guard."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, build_model, capture_logit, load_layers, rel_err
from oracle import ctr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOGIT_TOL, GRAD_TOL = 1e-5, 1e-4


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs #2 / #3 / #4 at full size: vocab 1M per table, batch 65 536 (32 768 for FiBiNET),
# ReLU everywhere.  Samples are independent, so the oracle evaluates a slice of the batch.
# ------------------------------------------------------------------------------------------------
def _baseline_cfg(model, D, **kw):
    cols = [O.sparse_col("C%d" % (i + 1), 1000000, D) for i in range(26)] + [O.dense_col("I%d" % (i + 1)) for i in range(13)]
    return O.make_cfg(model, cols, cols, init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0, **kw)


@pytest.mark.parametrize("model,D,B,n_check,kw", [
    ("DeepFM", 16, 65536, 65536, dict(dnn_hidden_units=[256, 128])),
    ("xDeepFM", 16, 65536, 4096, dict(dnn_hidden_units=[256, 256], cin_layer_size=[128, 128], cin_split_half=True)),
    ("FiBiNET", 32, 32768, 2048, dict(bilinear_type="interaction", dnn_hidden_units=[128, 128])),
    ("DCN", 16, 65536, 65536, dict(cross_num=2, cross_parameterization="vector", dnn_hidden_units=[128, 128])),
])
@pytest.mark.parametrize("zipf", [None, 1.05])
def test_logits_at_baseline_size(model, D, B, n_check, kw, zipf):
    cfg = _baseline_cfg(model, D, **kw)
    m = build_model(cfg, DEV, table_grad="rowwise")
    gen = torch.Generator(device=DEV).manual_seed(11)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=gen, device=DEV) * 0.05)
    X, y = O.synthetic_batch(cfg, B, seed=21, zipf_alpha=zipf)
    m.train()
    Xd = X.to(DEV)
    y_pred, logit = capture_logit(m, Xd)
    loss = torch.nn.functional.binary_cross_entropy(y_pred.squeeze(1), y.to(DEV), reduction="sum")
    loss.backward()                                   # the backward must run at this size too
    m.check_ids()
    # gather: bit-exact at full size
    E, dnn_input, lin, fm, blk = m.embed(Xd, want_fm=False)
    f = 7
    ids = Xd[:, f].long()
    table = m.embedding_dict["C%d" % (f + 1)].weight
    assert torch.equal(E[:, f, :], table[ids])
    state = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sl = slice(0, n_check)
    with torch.no_grad():
        ref_logit = O.model_logit(cfg, state, X[sl])
    err = rel_err(logit[sl].cpu(), ref_logit)
    assert err <= LOGIT_TOL, (model, zipf, err)


# ------------------------------------------------------------------------------------------------
# adjacent models: layers vs goldens recorded from the reference (layers_r2.npz)
# ------------------------------------------------------------------------------------------------
def _layer_check(tag, fn):
    layer = load_layers("layers_r2.npz")[tag]
    x = layer["x"].to(DEV).requires_grad_(True)
    P = {k[len("param/"):]: v.to(DEV).requires_grad_(True) for k, v in layer.items() if k.startswith("param/")}
    y = fn(x, P)
    assert tuple(y.shape) == tuple(layer["y"].shape), tag
    assert rel_err(y.detach().cpu(), layer["y"]) <= LOGIT_TOL, tag
    (y * layer["w"].to(DEV)).sum().backward()
    assert rel_err(x.grad.cpu(), layer["dx"]) <= GRAD_TOL, tag
    for k, p in P.items():
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(p).cpu()
        assert rel_err(got, layer["dparam/" + k]) <= GRAD_TOL, (tag, k)


def test_adjacent_layers_match_reference_golden():
    from deepctr_torch_b200 import ops

    def afm(x, P):
        att = ops.afm_attention(x, P["attention_W"], P["attention_b"], P["projection_h"])
        return ops.rowdot(att, P["projection_p"]).unsqueeze(1)

    def interacting(heads, res, scaling):
        def f(x, P):
            B, F, D = x.shape
            x2 = x.reshape(B * F, D)
            q = ops.dnn_layer(x2, P["W_Query"], None, "linear", w_kn=True).view(B, F, D)
            k = ops.dnn_layer(x2, P["W_key"], None, "linear", w_kn=True).view(B, F, D)
            v = ops.dnn_layer(x2, P["W_Value"], None, "linear", w_kn=True).view(B, F, D)
            r = ops.dnn_layer(x2, P["W_Res"], None, "linear", w_kn=True).view(B, F, D) if res else torch.zeros_like(q)
            return ops.field_attention(q, k, v, r, heads, (D // heads) ** -0.5 if scaling else 1.0)
        return f

    _layer_check("bipool", lambda x, P: ops.bi_interaction_pooling(x))
    _layer_check("afm", afm)
    _layer_check("afm_small", afm)
    _layer_check("interacting", interacting(4, True, True))
    _layer_check("interacting_nores", interacting(2, False, False))


# ------------------------------------------------------------------------------------------------
# f2: fused row-wise optimizer
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_rowwise_fit_matches_reference_fit(opt):
    """table_grad='rowwise' + the fused optimizer kernels reproduce the reference's own fit() with the dense
    torch optimizer (l2 = 0: rows outside the batch have zero gradient and do not move under sgd / adagrad)."""
    z = np.load(os.path.join(GOLDEN_DIR, "fit_%s_l2zero.npz" % opt))
    cfg = json.loads(str(z["cfg"]))
    names = json.loads(str(z["names"]))
    x = {n: z["X"][:, i].copy() for i, n in enumerate(names)}
    for int_ids in (False, True):
        m = build_model(cfg, DEV, table_grad="rowwise")
        m.use_int_ids(int_ids)
        m.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init/")})
        m.compile(opt, "binary_crossentropy", metrics=["binary_crossentropy"])
        from deepctr_torch_b200.optim import RowwiseOptimizer
        assert isinstance(m.optim, RowwiseOptimizer)
        hist = m.fit(x, z["y"], batch_size=64, epochs=3, verbose=0, validation_split=0.2, shuffle=False)
        ref_hist = json.loads(str(z["history"]))
        for k, vals in ref_hist.items():
            assert np.allclose(hist.history[k], vals, rtol=2e-4, atol=1e-6), (opt, int_ids, k, hist.history[k], vals)
        pred = m.predict(x, batch_size=64)
        assert np.abs(pred - z["pred"]).max() <= 2e-4
        for k, v in m.state_dict().items():
            ref = z["final/" + k]
            assert np.abs(v.cpu().numpy() - ref).max() <= 5e-4 * max(1e-3, np.abs(ref).max()), (opt, k)


@pytest.mark.parametrize("opt", ["adam", "rmsprop", "adagrad"])
def test_rowwise_optimizer_rule(opt):
    """Three steps on different batches with l2 > 0 against the row-wise rule evaluated on the CPU with the
    oracle's gradients (touched rows: g + 2*l2*w, state advanced only for touched rows)."""
    cols = [O.sparse_col("C%d" % i, 300 + 11 * i, 16) for i in range(6)] + [O.dense_col("I%d" % i) for i in range(3)]
    cfg = O.make_cfg("DeepFM", cols, cols, dnn_hidden_units=[32, 16], init_std=0.05, l2_reg_linear=1e-3,
                     l2_reg_embedding=1e-3)
    m = build_model(cfg, DEV, table_grad="rowwise")
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_((torch.randn(p.shape, generator=gen) * 0.05).to(DEV))
    m.compile(opt, "binary_crossentropy")
    m.train()
    cur = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    st = {}
    lr = {"adam": 1e-3, "rmsprop": 1e-2, "adagrad": 1e-2}[opt]
    table_keys = [k for k in cur if "embedding_dict" in k]
    for step in range(1, 4):
        X, y = O.synthetic_batch(cfg, 512, seed=50 + step)
        _, _, _, grads = O.loss_and_grads(cfg, cur, X, y)
        m.optim.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy(m(X.to(DEV)).squeeze(1), y.to(DEV), reduction="sum")
        (loss + m.get_regularization_loss().sum()).backward()
        m.optim.step()
        findex = O.feature_index(cfg)
        for k in table_keys:
            name = k.split(".")[-2]
            idx = X[:, findex[name][0]].long().unique()
            w, g = cur[k], grads[k]
            gg = g[idx] + 2e-3 * w[idx]
            s = st.setdefault(k, {"a": torch.zeros_like(w), "b": torch.zeros_like(w)})
            if opt == "adagrad":
                s["a"][idx] += gg * gg
                w[idx] -= lr * gg / (s["a"][idx].sqrt() + 1e-10)
            elif opt == "rmsprop":
                s["a"][idx] = 0.99 * s["a"][idx] + 0.01 * gg * gg
                w[idx] -= lr * gg / (s["a"][idx].sqrt() + 1e-8)
            else:
                s["a"][idx] = 0.9 * s["a"][idx] + 0.1 * gg
                s["b"][idx] = 0.999 * s["b"][idx] + 0.001 * gg * gg
                bc1, bc2 = 1 - 0.9 ** step, 1 - 0.999 ** step
                w[idx] -= (lr / bc1) * s["a"][idx] / (s["b"][idx].sqrt() / bc2 ** 0.5 + 1e-8)
        now = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        for k in cur:                        # dense parameters follow the torch optimizer on the GPU
            if k not in table_keys:
                cur[k] = now[k].clone()
        # rmsprop's first steps are lr / sqrt(1 - alpha) = 0.1 per coordinate whatever the gradient: fp32 round-off in
        # the gradients is amplified from step to step, the rule itself is identical
        tol = 1e-3 if opt == "rmsprop" else 2e-5
        for k in table_keys:
            assert rel_err(now[k], cur[k]) <= tol, (opt, step, k)


def test_two_forwards_before_backward_keep_their_own_plan():
    """the unique plan belongs to the autograd node that built it."""
    cols = [O.sparse_col("C%d" % i, 500, 16) for i in range(4)] + [O.dense_col("I0")]
    cfg = O.make_cfg("DeepFM", cols, cols, dnn_hidden_units=[16], init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0)
    m = build_model(cfg, DEV, table_grad="rowwise")
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, device=DEV) * 0.05)
    m.train()
    Xa, ya = O.synthetic_batch(cfg, 300, seed=1)
    Xb, yb = O.synthetic_batch(cfg, 200, seed=2)          # a different batch size on purpose
    bce = torch.nn.functional.binary_cross_entropy
    pa = m(Xa.to(DEV))
    pb = m(Xb.to(DEV))                                     # second forward before the first backward
    (bce(pa.squeeze(1), ya.to(DEV), reduction="sum") + bce(pb.squeeze(1), yb.to(DEV), reduction="sum")).backward()
    m.check_ids()
    state = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    _, _, _, ga = O.loss_and_grads(cfg, state, Xa, ya)
    _, _, _, gb = O.loss_and_grads(cfg, state, Xb, yb)
    for k, p in m.named_parameters():
        g = p.grad.to_dense() if p.grad.is_sparse else p.grad
        assert rel_err(g.cpu(), ga[k] + gb[k]) <= GRAD_TOL, k


# ------------------------------------------------------------------------------------------------
# f3: int32 id cells
# ------------------------------------------------------------------------------------------------
def test_int_ids_beyond_2_pow_24():
    """fp32-encoded ids are exact only below 2^24 (reference basemodel.py:242); int32 cells are not limited."""
    V = (1 << 24) + 1000
    cols = [O.sparse_col("big", V, 4), O.sparse_col("small", 50, 4), O.dense_col("I0")]
    cfg = O.make_cfg("DeepFM", cols, cols, dnn_hidden_units=[8], init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0)
    m = build_model(cfg, DEV, table_grad="rowwise").use_int_ids(True)
    ids = np.array([(1 << 24) + 1, (1 << 24) + 3, (1 << 24) + 999, 5, 16777217], dtype=np.int64)
    x = {"big": ids, "small": np.arange(5) % 50, "I0": np.linspace(0, 1, 5).astype("float32")}
    X = torch.from_numpy(m.pack_inputs(x)).to(DEV)
    E = m.embed(X)[0]
    assert torch.equal(E[:, 0, :], m.embedding_dict["big"].weight[torch.from_numpy(ids).to(DEV)])
    m.check_ids()
    # the same ids through fp32 cells collapse onto even neighbours: that is the limit being removed
    assert float(np.float32(ids[0])) != float(ids[0])
    # out-of-range int ids are still reported
    x_bad = dict(x, big=np.array([V, 1, 2, 3, 4], dtype=np.int64))
    m(torch.from_numpy(m.pack_inputs(x_bad)).to(DEV))
    with pytest.raises(IndexError):
        m.check_ids()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_cuda1_runs_on_cuda1():
    """launches follow the tensors' device, not the process-wide current device."""
    from helpers import load_case
    c = load_case("deepfm_small")
    assert torch.cuda.current_device() == 0
    m = build_model(c["cfg"], "cuda:1", table_grad="rowwise")
    m.load_state_dict(c["state"])
    m.train()
    y_pred, logit = capture_logit(m, c["X"].to("cuda:1"))
    torch.nn.functional.binary_cross_entropy(y_pred.squeeze(), c["y"].to("cuda:1"), reduction="sum").backward()
    m.check_ids()
    assert logit.device.index == 1 and rel_err(logit.cpu(), c["logit"]) <= LOGIT_TOL
