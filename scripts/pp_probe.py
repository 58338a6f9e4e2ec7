"""Health probe + timing of the PP GEMM engine (persistent, tile-pipelined SS engine, csrc/gemm_pk.cu gemm_pp_kernel).

    pp_probe.py check   correctness vs fp64 on shapes with M / N / K tails, bias+activation and act-mask paths (PP=1 and PP=0)
    pp_probe.py bench   the DeepFM tower forward / dgrad GEMMs at batch 65 536: PP vs the one-tile-per-CTA SS engine
Run under `timeout`: a protocol bug in a tcgen05 kernel traps (bounded mbarrier spins)."""





import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepctr_torch_b200 import _lib, ops

what = sys.argv[1] if len(sys.argv) > 1 else "check"
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)


def gemm(A, sam, sak, Bm, sbn, sbk, M, N, K, C=None):
    if C is None:
        C = torch.full((M, N), float("nan"), device="cuda")
    ops.ensure_gemm_scratch(DEV, M, K, N)
    ops.ensure_scratch_bytes(DEV, int(os.environ.get("CTR_TS_BREP", "1")) * (2 << 20))
    _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), sam, sak, ops._ptr(Bm), sbn, sbk, ops._ptr(C), N, 0, ops._stream())
    return C


if what == "check":
    worst = 0.0
    for (M, N, K) in [(65536, 256, 432), (65536, 128, 256), (65536, 432, 256), (70001, 130, 100), (66000, 256, 128),
                      (40000, 300, 64), (50000, 20, 600), (8192, 256, 432)]:
        K4 = (K + 3) // 4 * 4
        A = torch.randn(M, K4, device="cuda", generator=g)
        Bm = torch.randn(N, K4, device="cuda", generator=g)
        for ts in ("1", "0"):
            os.environ["CTR_GEMM_PP"] = ts
            C = gemm(A, K4, 1, Bm, K4, 1, M, N, K)
            torch.cuda.synchronize()
            ref = A[:, :K].double() @ Bm[:, :K].double().t()
            e = float((C.double() - ref).abs().max() / ref.abs().max())
            print("M=%d N=%d K=%d  PP=%s rel err %.3e" % (M, N, K, ts, e), flush=True)
            worst = max(worst, e if e == e else 1e9)
        # weights stored [K, N] (dgrad form)
        Bt = torch.randn(K, N, device="cuda", generator=g)
        os.environ["CTR_GEMM_PP"] = "1"
        C = gemm(A, K4, 1, Bt, 1, N, M, N, K)
        torch.cuda.synchronize()
        ref = A[:, :K].double() @ Bt.double()
        e = float((C.double() - ref).abs().max() / ref.abs().max())
        print("M=%d N=%d K=%d  PP=1 (B as [K,N]) rel err %.3e" % (M, N, K, e), flush=True)
        worst = max(worst, e if e == e else 1e9)
    # fused layer paths: bias + relu forward, chained backward with the act' mask on A
    os.environ["CTR_GEMM_PP"] = "1"
    B, K, N = 50000, 432, 256
    x = torch.randn(B, K, device="cuda", generator=g)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = (torch.randn(N, device="cuda", generator=g) * 0.05).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = ops.dnn_layer(xr, W, b, "tanh")
    w = torch.randn(B, N, device="cuda", generator=g)
    (y * w).sum().backward()
    x64 = x.double().requires_grad_(True)
    W64, b64 = W.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    y64 = torch.tanh(x64 @ W64.t() + b64)
    (y64 * w.double()).sum().backward()
    for name, got, ref in (("y", y, y64), ("dx", xr.grad, x64.grad), ("dW", W.grad, W64.grad), ("db", b.grad, b64.grad)):
        e = float((got.double() - ref).abs().max() / ref.abs().max())
        print("dnn_layer tanh B=%d K=%d N=%d  %s rel err %.3e" % (B, K, N, name, e), flush=True)
        worst = max(worst, e * 0.2)      # abs error of z ~ 4e-6 * max|z| goes straight into y = tanh(z); gradients: bar 1e-4
    sys.exit(0 if worst < 5e-6 else 1)

B = 65536
cases = [("fwd L1  C[B,256]   = X[B,432] W1^T", "nt", B, 256, 432),
         ("fwd L2  C[B,128]   = H[B,256] W2^T", "nt", B, 128, 256),
         ("dgrad L2 C[B,256]  = dZ[B,128] W2", "nn", B, 256, 128),
         ("dgrad L1 C[B,432]  = dZ[B,256] W1", "nn", B, 432, 256)]
for name, kind, M, N, K in cases:
    A = torch.randn(M, K, device="cuda", generator=g)
    if kind == "nt":
        Bm = torch.randn(N, K, device="cuda", generator=g)
        args = (A, K, 1, Bm, K, 1)
    else:
        Bm = torch.randn(K, N, device="cuda", generator=g)
        args = (A, K, 1, Bm, 1, N)
    C = torch.empty(M, N, device="cuda")
    line = "%-36s" % name
    for ts in ("1", "0"):
        os.environ["CTR_GEMM_PP"] = ts
        for _ in range(3):
            gemm(*args, M, N, K, C=C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gemm(*args, M, N, K, C=C)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
        line += "  PP=%s %7.1f us %6.1f TF/s (%.0f%% of the 3xTF32 ceiling)" % (ts, us, tf, 100 * tf / (1358.6 / 6))
    print(line, flush=True)
