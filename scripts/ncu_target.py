"""Light-weight profiling target: DeepFM / xDeepFM-shaped eager steps with small tables so that the
set-up costs nothing under ncu.  Kernel behaviour per launch matches the bench (same batch, same D);
only the table height differs (rows are still far apart: 26 tables x 200k rows x 64 B = 333 MB)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import ctr_oracle as O
from helpers import build_model
from deepctr_torch_b200 import ops

model = sys.argv[1] if len(sys.argv) > 1 else "DeepFM"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, V = 65536, 200000
kw = {"DeepFM": dict(dnn_hidden_units=[256, 128]),
      "xDeepFM": dict(dnn_hidden_units=[256, 256], cin_layer_size=[128, 128]),
      "DCN": dict(cross_num=2, dnn_hidden_units=[128, 128], l2_reg_cross=0)}[model]
cols = [O.sparse_col("C%d" % i, V, 16) for i in range(26)] + [O.dense_col("I%d" % i) for i in range(13)]
cfg = O.make_cfg(model, cols, cols, init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0, **kw)
m = build_model(cfg, "cuda:0", table_grad="rowwise")
m.train()
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.cat([torch.randint(0, V, (B, 26), device="cuda", generator=g).float(),
               torch.rand(B, 13, device="cuda", generator=g)], 1)
y = (torch.rand(B, device="cuda", generator=g) < 0.25).float()
for _ in range(steps):
    m.zero_grad(set_to_none=True)
    loss = ops.binary_cross_entropy(m(X).squeeze(1), y, reduction="sum")
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
