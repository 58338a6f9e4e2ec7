"""Per-stage clock64 timeline of the tcgen05 GEMM for the DeepFM tower shapes (CTA (0,0,0))."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTR_GEMM"] = "tc"
from deepctr_torch_b200 import _lib, ops

buf = torch.zeros(512, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
B = 65536
cases = {"fwd L1 (M=65536,N=256,K=432) KVEC/KVEC": lambda: ("nt", B, 256, 432),
         "fwd L2 (M=65536,N=128,K=256)": lambda: ("nt", B, 128, 256),
         "dW L1 (M=256,N=432,K=65536) TRANS/TRANS": lambda: ("tn", 256, 432, B)}
for name, f in cases.items():
    kind, M, N, K = f()
    if kind == "nt":
        A = torch.randn(M, K, device="cuda", generator=g); Bm = torch.randn(N, K, device="cuda", generator=g)
        args = (A, K, 1, Bm, K, 1)
    else:
        A = torch.randn(K, M, device="cuda", generator=g); Bm = torch.randn(K, N, device="cuda", generator=g)
        args = (A, 1, M, Bm, 1, N)
    C = torch.empty(M, N, device="cuda")
    def run():
        _lib.call("ctr_sgemm", M, N, K, ops._ptr(args[0]), args[1], args[2], ops._ptr(args[3]), args[4], args[5],
                  ops._ptr(C), N, 0, ops._stream())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    print("== %s: %.1f us per call, %.1f TFLOP/s (fp32-equivalent)" % (name, e0.elapsed_time(e1) * 100, 2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-4) / 1e12))
    buf.zero_()
    _lib.call("ctr_debug_set_buffer", ops._ptr(buf))
    run(); torch.cuda.synchronize()
    _lib.call("ctr_debug_set_buffer", None)
    t = buf.cpu().tolist()
    t0 = t[0]
    print("   prologue->first stage %d cyc; producers done at %d; accum ready %d; epilogue end %d" % (t[8] - t0, t[1] - t0, t[2] - t0, t[3] - t0))
    print("   kb: start  wait_empty  store(consume regs)  arrive  issue_next_loads | mma: full_at  issue   (cycles)")
    for kb in range(8):
        r = t[8 + kb * 8: 16 + kb * 8]
        if r[0] == 0:
            break
        print("   %2d: start %7d  +wait %6d  +store %5d  +arr %4d  +nextld %5d | full %7d  issue +%5d" %
              (kb, r[0] - t0, r[2] - r[0], r[3] - r[2], r[4] - r[3], r[1] - r[4], r[5] - t0, r[6] - r[5]))
