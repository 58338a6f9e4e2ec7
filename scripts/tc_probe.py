"""Quick health probe of the tcgen05 GEMM (run under `timeout`): exits 0 when results are correct."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTR_GEMM"] = "tc"
from deepctr_torch_b200 import _lib, ops

worst = 0.0
for (M, N, K) in [(128, 32, 32), (128, 256, 64), (1000, 256, 429), (256, 429, 3000), (77, 33, 19)]:
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    C = torch.full((M, N), float("nan"), device="cuda")
    _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(B), K, 1, ops._ptr(C), N, 0, ops._stream())
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    print("tc probe M=%d N=%d K=%d rel err %.3e" % (M, N, K, err), flush=True)
    worst = max(worst, err if err == err else 1e9)
sys.exit(0 if worst < 5e-6 else 1)
