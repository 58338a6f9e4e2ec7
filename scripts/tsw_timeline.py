"""clock64 timeline of CTA (0,0,0) of the TSW weight-gradient engine (ctr_debug_set_buffer)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTR_GEMM_TSW"] = "1"
from deepctr_torch_b200 import _lib, ops

shape = sys.argv[1] if len(sys.argv) > 1 else "dw1"
B, K, N, masked = {"dw1": (65536, 432, 256, False), "dw2": (65536, 256, 128, True)}[shape]
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.randn(B, K, device="cuda", generator=g)
Y = torch.relu(torch.randn(B, N, device="cuda", generator=g))
dY = torch.randn(B, N, device="cuda", generator=g)
W = torch.randn(N, K, device="cuda", generator=g)
dW = torch.empty(N, K, device="cuda")
db = torch.empty(N, device="cuda")
dbg = torch.zeros(10 * 64, dtype=torch.int64, device="cuda")
ops.ensure_gemm_scratch(torch.device("cuda:0"), B, K, N)
for it in range(3):
    if it == 2:
        _lib.call("ctr_debug_set_buffer", ops._ptr(dbg))
    _lib.call("ctr_dnn_layer_bwd_chain", ops._ptr(X), K, ops._ptr(W), K, 1, ops._ptr(Y) if masked else None, N if masked else 0,
              ops._ptr(dY), N, None, 0, ops._ptr(dW), K, 1, ops._ptr(db), B, K, N, 1 if masked else 0, 0 if masked else 1, 0,
              ops._stream())
torch.cuda.synchronize()
_lib.call("ctr_debug_set_buffer", None)
d = dbg.cpu().view(10, 64)
t0 = int(d[7, 2])
print("shape %s B=%d K=%d N=%d: cycles since kernel entry of CTA (0,0,0)" % (shape, B, K, N))
print("stage | mma: b_full a_full issued | A conv: raw_ready slot_free published | B conv: raw_ready stage_free published")
for i in range(40):
    row = [int(d[e, i]) - t0 if int(d[e, i]) else -1 for e in (0, 1, 2, 3, 4, 5, 6, 8, 9)]
    print("%5d | %7d %7d %7d | %7d %7d %7d | %7d %7d %7d" % ((i,) + tuple(row)))
print("epilogue: accum ready %d, done %d" % (int(d[7, 0]) - t0, int(d[7, 1]) - t0))
