"""Where does the gradient error of hot (Zipf) rows come from?  Compare SIMT vs tensor-core towers."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle import ctr_oracle as O
from helpers import build_model, rel_err

cols = [O.sparse_col("C%d" % i, 20000, 16) for i in range(26)] + [O.dense_col("I%d" % i) for i in range(13)]
cfg = O.make_cfg("DeepFM", cols, cols, init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0, dnn_hidden_units=[256, 128])
X, y = O.synthetic_batch(cfg, 4096, seed=11, zipf_alpha=1.05)
ref = None
for engine in ("simt", "tc1", "pk"):
    for mode in ("rowwise", "dense"):
        os.environ["CTR_GEMM"] = engine
        m = build_model(cfg, "cuda:0", table_grad=mode)
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.device))
        if ref is None:
            state64 = {k: v.detach().cpu().double() for k, v in m.state_dict().items()}
            _, _, _, ref = O.loss_and_grads(cfg, state64, X, y)
        m.train()
        yp = m(X.cuda())
        torch.nn.functional.binary_cross_entropy(yp.squeeze(1), y.cuda(), reduction="sum").backward()
        errs = {}
        for k, p in m.named_parameters():
            gg = p.grad.to_dense() if p.grad.is_sparse else p.grad
            errs[k] = rel_err(gg.cpu(), ref[k])
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(engine, mode, " ".join("%s=%.2e" % (k.replace("embedding_dict.", "E.").replace(".weight", ""), v) for k, v in worst))
