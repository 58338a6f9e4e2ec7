"""Health probe + timing of the TSW weight-gradient engine (dZ^T through tensor memory, csrc/gemm_pk.cu).

    tsw_probe.py check   dW / db vs fp64 through ctr_dnn_layer_bwd_chain on tower shapes and ragged shapes
    tsw_probe.py bench   the DeepFM tower's two weight gradients at batch 65 536: TSW vs (SS engine + colsum)
Run under `timeout`: a protocol bug in a tcgen05 kernel traps (bounded mbarrier spins)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepctr_torch_b200 import _lib, ops

what = sys.argv[1] if len(sys.argv) > 1 else "check"
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
ACT_LINEAR, ACT_RELU = 0, 1


def wgrad(X, ldx, Y, dY, B, K, N, act, dy_is_dz, W):
    dW = torch.full((N, K), float("nan"), device="cuda")
    db = torch.full((N,), float("nan"), device="cuda")
    ops.ensure_gemm_scratch(DEV, B, K, N)
    _lib.call("ctr_dnn_layer_bwd_chain", ops._ptr(X), ldx, ops._ptr(W), K, 1, ops._ptr(Y) if Y is not None else None,
              Y.stride(0) if Y is not None else 0, ops._ptr(dY), dY.stride(0), None, 0, ops._ptr(dW), K, 1, ops._ptr(db),
              B, K, N, act, dy_is_dz, ACT_LINEAR, ops._stream())
    return dW, db


if what == "check":
    worst = 0.0
    for (B, K, N, masked) in [(65536, 432, 256, True), (65536, 256, 128, False), (5000, 432, 256, True), (4099, 100, 130, True),
                              (70001, 24, 40, False), (8192, 600, 20, True), (16384, 256, 384, False)]:
        K4 = (K + 3) // 4 * 4
        N4 = (N + 3) // 4 * 4
        X = torch.randn(B, K4, device="cuda", generator=g)
        Y = torch.relu(torch.randn(B, N4, device="cuda", generator=g))
        dY = torch.randn(B, N4, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g)
        dz64 = dY[:, :N].double() * ((Y[:, :N] > 0).double() if masked else 1.0)
        ref_w = dz64.t() @ X[:, :K].double()
        ref_b = dz64.sum(0)
        for tsw in ("1", "0"):
            os.environ["CTR_GEMM_TSW"] = tsw
            dW, db = wgrad(X, K4, Y if masked else None, dY, B, K, N, ACT_RELU if masked else ACT_LINEAR, 0 if masked else 1, W)
            torch.cuda.synchronize()
            ew = float((dW.double() - ref_w).abs().max() / ref_w.abs().max())
            eb = float((db.double() - ref_b).abs().max() / ref_b.abs().max())
            print("B=%d K=%d N=%d mask=%d  TSW=%s  dW rel err %.3e  db rel err %.3e" % (B, K, N, masked, tsw, ew, eb), flush=True)
            for e in (ew, eb):
                worst = max(worst, e if e == e else 1e9)
    print("worst", worst)
    sys.exit(0 if worst < 2e-5 else 1)      # K = 65 536 accumulation chain, truncating fp32 adds

B = 65536
for name, K, N, masked in [("dW1 [256,432] = dZ1^T X  (+db1)", 432, 256, False), ("dW2 [128,256] = (dY.relu'(Y))^T H1 (+db2)", 256, 128, True)]:
    X = torch.randn(B, K, device="cuda", generator=g)
    Y = torch.relu(torch.randn(B, N, device="cuda", generator=g))
    dY = torch.randn(B, N, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    for tsw in ("0", "1"):
        os.environ["CTR_GEMM_TSW"] = tsw
        args = (X, K, Y if masked else None, dY, B, K, N, ACT_RELU if masked else ACT_LINEAR, 0 if masked else 1, W)
        for _ in range(3):
            wgrad(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            wgrad(*args)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print("%-44s TSW=%s  %.1f us  (%.1f fp32-equivalent TFLOP/s)" % (name, tsw, us, 2.0 * B * K * N / us / 1e6), flush=True)
