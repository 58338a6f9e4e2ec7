"""clock64 timeline of CTA (0,0) of the TS GEMM engine (ctr_debug_set_buffer): per k stage, when the MMA thread had its
weight stage / its A slots / finished issuing, when converter group 0 had its raw tile / a free slot / published the slot,
when the TMA producer issued the stage."""
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
dbg = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2:
        _lib.call("ctr_debug_set_buffer", ops._ptr(dbg))
    _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(Bm), K, 1, ops._ptr(C), N, 0, ops._stream())
torch.cuda.synchronize()
_lib.call("ctr_debug_set_buffer", None)
d = dbg.cpu().view(8, 64)
t0 = int(d[7, 2])
nkb = (K + 15) // 16
print("shape %s M=%d N=%d K=%d: cycles since kernel entry of CTA (0,0)" % (shape, M, N, K))
print("stage |  tma_issue  b_full  a_full  mma_issued | conv0: raw_ready slot_free published")
for i in range(nkb):
    row = [int(d[e, i]) - t0 if int(d[e, i]) else -1 for e in (6, 0, 1, 2, 3, 4, 5)]
    print("%5d | %10d %7d %7d %10d | %10d %9d %9d" % ((i,) + tuple(row)))
print("epilogue: accum ready %d, done %d" % (int(d[7, 0]) - t0, int(d[7, 1]) - t0))
