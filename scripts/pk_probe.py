"""Health probe + micro-benchmark of the packed-operand tcgen05 GEMM (csrc/gemm_pk.cu).

    pk_probe.py check    correctness on small / odd shapes in the three operand layouts (exit 0 = OK)
    pk_probe.py bench    the DeepFM tower's GEMMs at batch 65 536: us per call and fp32-equivalent
                         TFLOP/s for every engine (pack kernels included in the pk time)
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
    _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), sam, sak, ops._ptr(Bm), sbn, sbk, ops._ptr(C), N, 0, ops._stream())
    return C


if what == "check":
    os.environ["CTR_GEMM"] = "pk"
    worst = 0.0
    for (M, N, K) in [(128, 32, 32), (128, 256, 64), (1000, 256, 432), (300, 448, 96), (77, 36, 20), (513, 96, 1664),
                      (256, 429, 3000), (8192, 256, 432), (4096, 448, 100), (300, 200, 20000)]:
        K4, M4 = (K + 3) // 4 * 4, (M + 3) // 4 * 4
        A = torch.randn(M, K4, device="cuda", generator=g)
        Bm = torch.randn(N, K4, device="cuda", generator=g)
        C = gemm(A, K4, 1, Bm, K4, 1, M, N, K)
        torch.cuda.synchronize()
        ref = A[:, :K].double() @ Bm[:, :K].double().t()
        e1 = float((C.double() - ref).abs().max() / ref.abs().max())
        At = torch.randn(K, M4, device="cuda", generator=g)
        Bt = torch.randn(K, N, device="cuda", generator=g)
        C = gemm(At, 1, M4, Bt, 1, N, M, N, K)
        torch.cuda.synchronize()
        ref = At[:, :M].double().t() @ Bt.double()
        e2 = float((C.double() - ref).abs().max() / ref.abs().max())
        print("pk probe M=%d N=%d K=%d  NT rel err %.3e  TN rel err %.3e" % (M, N, K, e1, e2), flush=True)
        for e in (e1, e2):
            worst = max(worst, e if e == e else 1e9)
    sys.exit(0 if worst < 5e-6 else 1)

if what == "bias":
    # Is the tensor core's fp32 accumulation rounding symmetric?  C(+A) + C(-A) is exactly 0 for
    # round-to-nearest / toward-zero, and about -2 x bias for a floor-like (toward -inf) truncation.
    for engine in ("pk", "simt"):
        os.environ["CTR_GEMM"] = engine
        for (M, N, K, pos) in [(256, 256, 512, True), (256, 256, 512, False), (1024, 256, 128, False)]:
            A = torch.rand(M, K, device="cuda", generator=g) if pos else torch.randn(M, K, device="cuda", generator=g)
            Bm = torch.rand(N, K, device="cuda", generator=g) if pos else torch.randn(N, K, device="cuda", generator=g)
            C1 = gemm(A, K, 1, Bm, K, 1, M, N, K).clone()
            An = (-A).contiguous()
            C2 = gemm(An, K, 1, Bm, K, 1, M, N, K).clone()
            torch.cuda.synchronize()
            ref = A.double() @ Bm.double().t()
            scale = float(ref.abs().mean())
            e1 = (C1.double() - ref) / scale
            e2 = (C2.double() + ref) / scale
            print("%s M=%d N=%d K=%d %s: mean signed err C(+A) %+.3e  C(-A) %+.3e  mean(C(+A)+C(-A)) %+.3e  rms err %.3e"
                  % (engine, M, N, K, "positive" if pos else "gaussian", float(e1.mean()), float(e2.mean()),
                     float(((C1 + C2).double() / scale).mean()), float(e1.pow(2).mean().sqrt())), flush=True)
    sys.exit(0)

B = 65536
cases = [("fwd L1  C[B,256]   = X[B,432] W1^T", "nt", B, 256, 432),
         ("fwd L2  C[B,128]   = H[B,256] W2^T", "nt", B, 128, 256),
         ("dgrad L2 C[B,256]  = dZ[B,128] W2", "nn", B, 256, 128),
         ("dgrad L1 C[B,432]  = dZ[B,256] W1", "nn", B, 432, 256),
         ("wgrad L2 C[128,256] = dZ^T H", "tn", 128, 256, B),
         ("wgrad L1 C[256,432] = dZ^T X", "tn", 256, 432, B)]
for name, kind, M, N, K in cases:
    if kind == "nt":
        A = torch.randn(M, K, device="cuda", generator=g); Bm = torch.randn(N, K, device="cuda", generator=g)
        args = (A, K, 1, Bm, K, 1)
    elif kind == "nn":
        A = torch.randn(M, K, device="cuda", generator=g); Bm = torch.randn(K, N, device="cuda", generator=g)
        args = (A, K, 1, Bm, 1, N)
    else:
        A = torch.randn(K, M, device="cuda", generator=g); Bm = torch.randn(K, N, device="cuda", generator=g)
        args = (A, 1, M, Bm, 1, N)
    C = torch.empty(M, N, device="cuda")
    line = "%-36s" % name
    for engine in ("pk", "tc1", "simt"):
        os.environ["CTR_GEMM"] = engine
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
        line += "  %s %7.1f us %6.1f TF/s" % (engine, us, 2.0 * M * N * K / (us * 1e-6) / 1e12)
    print(line, flush=True)
# split of the pk time between the pack passes and the MMA kernel (CUDA events around one call each)
os.environ["CTR_GEMM"] = "pk"
print("pk engine = pack(A) + pack(B) + gemm_pk_kernel; see the ncu launch list for the split")
