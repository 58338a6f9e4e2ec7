"""Target for `ncu -k regex:gemm_(ts|pk)_kernel`: the DeepFM tower's first-layer forward GEMM at batch 65 536,
three launches through the TS engine and three through the SS engine (CTR_GEMM_TS=0)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepctr_torch_b200 import _lib, ops

shape = sys.argv[1] if len(sys.argv) > 1 else "fwd1"
M, N, K = {"fwd1": (65536, 256, 432), "fwd2": (65536, 128, 256), "dx1": (65536, 432, 256)}[shape]
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(M, K, device="cuda", generator=g)
Bm = torch.randn(N, K, device="cuda", generator=g)
C = torch.empty(M, N, device="cuda")
ops.ensure_gemm_scratch(torch.device("cuda:0"), M, K, N)
for ts in ("1", "0"):
    os.environ["CTR_GEMM_TS"] = ts
    for _ in range(3):
        _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(Bm), K, 1, ops._ptr(C), N, 0, ops._stream())
    torch.cuda.synchronize()
