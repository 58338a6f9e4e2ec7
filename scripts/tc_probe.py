"""Health probes of the tcgen05 kernels (run under `timeout`): exit 0 when results are correct.
usage: tc_probe.py kvec | mnvec | scalar | cin"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["CTR_GEMM"] = "tc"
what = sys.argv[1] if len(sys.argv) > 1 else "kvec"
from deepctr_torch_b200 import _lib, ops


def gemm(A, sam, sak, Bm, sbn, sbk, M, N, K):
    C = torch.full((M, N), float("nan"), device="cuda")
    _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), sam, sak, ops._ptr(Bm), sbn, sbk, ops._ptr(C), N, 0, ops._stream())
    torch.cuda.synchronize()
    return C


worst = 0.0
g = torch.Generator(device="cuda").manual_seed(1)
if what in ("kvec", "scalar", "mnvec"):
    if what == "scalar":
        os.environ["CTR_TC_LOAD"] = "s"
    for (M, N, K) in [(128, 32, 32), (128, 256, 64), (1000, 256, 432), (300, 448, 96), (77, 36, 20)]:
        if what == "mnvec":      # A stored [K, M], B stored [K, N]: both MN-contiguous
            A = torch.randn(K, M if M % 4 == 0 else M + 4 - M % 4, device="cuda", generator=g)
            Bm = torch.randn(K, N, device="cuda", generator=g)
            C = gemm(A, 1, A.shape[1], Bm, 1, N, M, N, K)
            ref = A[:, :M].double().t() @ Bm.double()
        else:
            A = torch.randn(M, K, device="cuda", generator=g)
            Bm = torch.randn(N, K, device="cuda", generator=g)
            C = gemm(A, K, 1, Bm, K, 1, M, N, K)
            ref = A.double() @ Bm.double().t()
        err = float((C.double() - ref).abs().max() / ref.abs().max())
        print("tc probe %s M=%d N=%d K=%d rel err %.3e" % (what, M, N, K, err), flush=True)
        worst = max(worst, err if err == err else 1e9)
else:   # CIN forward + backward on the tensor cores vs fp64 torch
    from oracle import ctr_oracle as O
    B, M, D = 300, 26, 16
    sizes = (32, 16)
    E = (torch.randn(B, M, D, device="cuda", generator=g) * 0.5).requires_grad_(True)
    P = {}
    H = M
    params = []
    for k, n in enumerate(sizes):
        W = (torch.randn(n, H * M, 1, device="cuda", generator=g) * 0.1).requires_grad_(True)
        b = (torch.randn(n, device="cuda", generator=g) * 0.1).requires_grad_(True)
        params += [W, b]
        P["conv1ds.%d.weight" % k], P["conv1ds.%d.bias" % k] = W, b
        H = n // 2
    out = ops.cin(E, sizes, True, "relu", params)
    w = torch.randn(out.shape, device="cuda", generator=g)
    (out * w).sum().backward()
    torch.cuda.synchronize()
    P64 = {k: v.detach().double().cpu().requires_grad_(True) for k, v in P.items()}
    E64 = E.detach().double().cpu().requires_grad_(True)
    ref = O.cin(P64, "", E64, sizes, True, "relu")
    (ref * w.double().cpu()).sum().backward()

    def rel(a, b):
        return float((a.double().cpu() - b).abs().max() / b.abs().max())
    errs = {"out": rel(out.detach(), ref.detach()), "dE": rel(E.grad, E64.grad)}
    for k in P:
        errs["d" + k] = rel(P[k].grad, P64[k].grad)
    for k, v in errs.items():
        print("cin probe %-24s rel err %.3e" % (k, v), flush=True)
    worst = max(v if v == v else 1e9 for v in errs.values())
sys.exit(0 if worst < 1e-5 else 1)
