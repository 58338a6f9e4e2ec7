"""Target for `ncu -k regex:gemm_(tsw|pk)_kernel`: the DeepFM tower's first-layer weight gradient at batch 65 536,
three launches through the TSW engine and three through the SS engine + column sum (CTR_GEMM_TSW=0)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepctr_torch_b200 import _lib, ops

shape = sys.argv[1] if len(sys.argv) > 1 else "dw1"
B, K, N = {"dw1": (65536, 432, 256), "dw2": (65536, 256, 128)}[shape]
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.randn(B, K, device="cuda", generator=g)
dY = torch.randn(B, N, device="cuda", generator=g)
W = torch.randn(N, K, device="cuda", generator=g)
dW = torch.empty(N, K, device="cuda")
db = torch.empty(N, device="cuda")
ops.ensure_gemm_scratch(torch.device("cuda:0"), B, K, N)
for tsw in ("1", "0"):
    os.environ["CTR_GEMM_TSW"] = tsw
    for _ in range(3):
        _lib.call("ctr_dnn_layer_bwd_chain", ops._ptr(X), K, ops._ptr(W), K, 1, None, 0, ops._ptr(dY), N, None, 0,
                  ops._ptr(dW), K, 1, ops._ptr(db), B, K, N, 0, 1, 0, ops._stream())
    torch.cuda.synchronize()
