"""Locate where the Zipf-id gradient error of the tensor-core tower comes from: run the DeepFM step
under each GEMM engine, capture every dnn_layer input/output and its gradient, and compare against
an fp64 torch replay of the same tower (same inputs): element error, mask flips, coherent (column
sum) error."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle import ctr_oracle as O
from helpers import build_model
from deepctr_torch_b200 import ops

cols = [O.sparse_col("C%d" % i, 20000, 16) for i in range(26)] + [O.dense_col("I%d" % i) for i in range(13)]
cfg = O.make_cfg("DeepFM", cols, cols, init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0, dnn_hidden_units=[256, 128])
zipf = float(sys.argv[1]) if len(sys.argv) > 1 else 1.05
X, y = O.synthetic_batch(cfg, 4096, seed=11, zipf_alpha=zipf if zipf > 0 else None)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


for engine in ("simt", "pk"):
    os.environ["CTR_GEMM"] = engine
    m = build_model(cfg, "cuda:0", table_grad="rowwise")
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.device))
    rec = []
    orig = ops.dnn_layer

    def spy(x, W, b, act="relu", w_kn=False):
        x = x.detach().requires_grad_(True) if not x.requires_grad else x
        x.retain_grad()
        out = orig(x, W, b, act, w_kn)
        out.retain_grad()
        rec.append((x, W, b, out))
        return out

    ops.dnn_layer = spy
    m.train()
    yp = m(X.cuda())
    torch.nn.functional.binary_cross_entropy(yp.squeeze(1), y.cuda(), reduction="sum").backward()
    ops.dnn_layer = orig
    print("== engine %s  zipf %s" % (engine, zipf))
    for li, (x, W, b, out) in enumerate(rec):
        x64 = x.detach().double().requires_grad_(True)
        W64 = W.detach().double().requires_grad_(True)
        b64 = b.detach().double().requires_grad_(True)
        z64 = x64[:, :W64.shape[1]] @ W64.t() + b64
        h64 = torch.relu(z64)
        gout = out.grad.detach().double()
        (h64 * gout).sum().backward()
        flips = int(((out.detach() > 0) != (h64.detach() > 0)).sum())
        dz = gout * (out.detach() > 0)
        dz64 = gout * (h64.detach() > 0)
        dx = x.grad.detach().double()[:, :W64.shape[1]]
        dxe = (dx - x64.grad[:, :W64.shape[1]])
        print(" layer %d: fwd rel err %.2e  mask flips %d  |  dX rel err %.2e  colsum(dX err)/max|colsum dX| %.2e"
              "  dW rel err %.2e  db rel err %.2e  (db with own mask vs fp64 mask: %.2e)" % (
                  li, rel(out.detach(), h64.detach()), flips, rel(dx, x64.grad[:, :W64.shape[1]]),
                  float(dxe.sum(0).abs().max() / x64.grad.sum(0).abs().max()),
                  rel(W.grad[:, :W64.shape[1]], W64.grad), rel(b.grad, b64.grad), rel(dz.sum(0), dz64.sum(0))))
