"""clock64 timeline of CTA (0,0,0) of the SS GEMM engine (gemm_pk_kernel; ctr_debug_set_buffer): per k stage, when the MMA
thread saw its A stage / its weight stage / finished issuing, and when converter thread 0 had its raw pieces / the free
A stage / had published it."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepctr_torch_b200 import _lib, ops

shape = sys.argv[1] if len(sys.argv) > 1 else "fwd1"
M, N, K, kind = {"fwd1": (65536, 256, 432, "nt"), "fwd2": (65536, 128, 256, "nt"), "dx1": (65536, 432, 256, "nn"),
                 "dx2": (65536, 256, 128, "nn")}[shape]
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(M, K, device="cuda", generator=g)
Bm = torch.randn(N, K, device="cuda", generator=g) if kind == "nt" else torch.randn(K, N, device="cuda", generator=g)
C = torch.empty(M, N, device="cuda")
ops.ensure_gemm_scratch(torch.device("cuda:0"), M, K, N)
dbg = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2:
        _lib.call("ctr_debug_set_buffer", ops._ptr(dbg))
    if kind == "nt":
        _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(Bm), K, 1, ops._ptr(C), N, 0, ops._stream())
    else:
        _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(Bm), 1, N, ops._ptr(C), N, 0, ops._stream())
torch.cuda.synchronize()
_lib.call("ctr_debug_set_buffer", None)
d = dbg.cpu().view(8, 64)
t0 = int(d[7, 2])
nkb = (K + 15) // 16
print("shape %s M=%d N=%d K=%d: cycles since kernel entry of CTA (0,0,0)" % (shape, M, N, K))
print("stage | mma: a_full b_full issued | conv t0: raw_ready stage_free published")
for i in range(nkb):
    row = [int(d[e, i]) - t0 if int(d[e, i]) else -1 for e in (0, 1, 2, 3, 4, 5)]
    print("%5d | %7d %7d %7d | %7d %7d %7d" % ((i,) + tuple(row)))
print("epilogue: accum ready %d, done %d" % (int(d[7, 0]) - t0, int(d[7, 1]) - t0))
