"""The GEMM engines behind every dense op (tcgen05 3xTF32 over packed operands / tcgen05 3xTF32 with
in-kernel split / FP32 FFMA tiles) against
an fp64 torch reference, for the operand layouts the tower uses (NT forward, NN dgrad, TN wgrad
with split-K), odd sizes, strided leading dimensions and the fused prologue/epilogues."""
import os

import pytest
import torch

from deepctr_torch_b200 import _lib, ops
from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sgemm(A, sam, sak, Bm, sbn, sbk, M, N, K, accumulate=False, C=None):
    if C is None:
        C = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.ensure_gemm_scratch(torch.device(DEV), M, K, N)
    _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), sam, sak, ops._ptr(Bm), sbn, sbk, ops._ptr(C), C.stride(0),
              1 if accumulate else 0, ops._stream())
    return C


@pytest.fixture(params=["pk", "pk-packed", "tc1", "simt"])
def engine(request):
    """pk: large operands are converted inside the GEMM (cp.async -> split), small ones packed;
    pk-packed: every operand goes through the pack kernels + TMA (CTR_PK_STREAM=0)."""
    old = {k: os.environ.get(k) for k in ("CTR_GEMM", "CTR_PK_STREAM")}
    os.environ["CTR_GEMM"] = request.param.split("-")[0]
    if request.param == "pk-packed":
        os.environ["CTR_PK_STREAM"] = "0"
    else:
        os.environ.pop("CTR_PK_STREAM", None)
    ops._scratch_need.clear()
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    ops._scratch_need.clear()


SHAPES = [(128, 256, 32), (1000, 256, 429), (4096, 128, 256), (77, 33, 19), (256, 429, 3000), (130, 520, 64),
          (5, 7, 3), (513, 96, 1664), (40000, 256, 432)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_nt_nn_tn_layouts(engine, M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    # NT: A [M,K] row-major (ld padded), B [N,K] row-major
    lda = (K + 3) // 4 * 4 + 4
    Abuf = torch.randn(M, lda, device=DEV, generator=g)
    Bm = torch.randn(N, K, device=DEV, generator=g)
    C = _sgemm(Abuf, lda, 1, Bm, K, 1, M, N, K)
    ref = Abuf[:, :K].double() @ Bm.double().t()
    assert rel_err(C.cpu(), ref.cpu()) <= 4e-6
    # NN (dgrad): B stored [K, N] row-major
    Bkn = torch.randn(K, N, device=DEV, generator=g)
    C = _sgemm(Abuf, lda, 1, Bkn, 1, N, M, N, K)
    ref = Abuf[:, :K].double() @ Bkn.double()
    assert rel_err(C.cpu(), ref.cpu()) <= 4e-6
    # TN (wgrad): A stored [K, M], B stored [K, N]; split-K path when K is large
    Akm = torch.randn(K, M, device=DEV, generator=g)
    C = _sgemm(Akm, 1, M, Bkn, 1, N, M, N, K)
    ref = Akm.double().t() @ Bkn.double()
    assert rel_err(C.cpu(), ref.cpu()) <= 4e-6
    # accumulate into C
    C0 = torch.randn(M, N, device=DEV, generator=g)
    C = _sgemm(Akm, 1, M, Bkn, 1, N, M, N, K, accumulate=True, C=C0.clone())
    assert rel_err(C.cpu(), (ref + C0.double()).cpu()) <= 4e-6


def test_wgrad_shape_split_k(engine):
    """dW[256,429] = dZ^T[256,B] X[B,429] with B = 65536 (the BASELINE tower's first layer)."""
    g = torch.Generator(device=DEV).manual_seed(1)
    B, N, K = 65536, 256, 429
    dZ = torch.randn(B, N, device=DEV, generator=g) * 0.01
    X = torch.randn(B, 432, device=DEV, generator=g)
    C = _sgemm(dZ, 1, N, X, 1, 432, N, K, B)
    ref = dZ.double().t() @ X[:, :K].double()
    # 65 536-term fp32 accumulations: tensor-core accumulators round toward zero, so the error grows
    # with the number of accumulation steps (1.2e-5 measured); this is a weight GRADIENT (bar 1e-4)
    assert rel_err(C.cpu(), ref.cpu()) <= 3e-5


@pytest.mark.parametrize("act", ["relu", "sigmoid", "tanh", "linear"])
def test_dnn_layer_fwd_bwd(engine, act):
    g = torch.Generator(device=DEV).manual_seed(3)
    B, K, N = 3000, 429, 256
    x = torch.randn(B, 432, device=DEV, generator=g)[:, :K].requires_grad_(True)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.1).requires_grad_(True)
    b = (torch.randn(N, device=DEV, generator=g) * 0.1).requires_grad_(True)
    y = ops.dnn_layer(x, W, b, act)
    w = torch.randn(B, N, device=DEV, generator=g)
    (y * w).sum().backward()
    xd, Wd, bd = x.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True), \
        b.detach().double().requires_grad_(True)
    z = xd @ Wd.t() + bd
    yr = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "linear": lambda t: t}[act](z)
    (yr * w.double()).sum().backward()
    assert rel_err(y.detach().cpu(), yr.detach().cpu()) <= 2e-5      # K = 429 fp32 accumulation + activation
    assert rel_err(x.grad.cpu(), xd.grad.cpu()) <= 5e-6
    assert rel_err(W.grad.cpu(), Wd.grad.cpu()) <= 5e-6
    assert rel_err(b.grad.cpu(), bd.grad.cpu()) <= 5e-6


@pytest.mark.parametrize("hidden", [(256, 256), (256, 128)])
def test_dnn_tower_large_batch_streaming(hidden):
    """The fused tower at a batch large enough that every activation operand is converted inside the
    GEMM (streamed, masked on the top layer, both operands streamed in the weight gradients)."""
    g = torch.Generator(device=DEV).manual_seed(11)
    B, K = 40000, 429
    x = torch.randn(B, 432, device=DEV, generator=g)[:, :K].requires_grad_(True)
    Ws, bs, dims = [], [], [K] + list(hidden)
    for i in range(len(hidden)):
        Ws.append((torch.randn(dims[i + 1], dims[i], device=DEV, generator=g) * 0.05).requires_grad_(True))
        bs.append((torch.randn(dims[i + 1], device=DEV, generator=g) * 0.05).requires_grad_(True))
    # tanh on purpose: same code paths as ReLU (act'(Y) prologue on the top layer, act'(X) epilogues below),
    # but smooth — at 10^7 units a handful of ReLU masks flip between two correctly rounded
    # implementations and each flip moves one sample's gradient row by ~10 % of the maximum (run 27)
    y = ops.dnn_tower(x, "tanh", Ws, bs)
    w = torch.randn(B, hidden[-1], device=DEV, generator=g)
    (y * w).sum().backward()
    xd = x.detach().double().requires_grad_(True)
    Wd = [W.detach().double().requires_grad_(True) for W in Ws]
    bd = [b.detach().double().requires_grad_(True) for b in bs]
    h = xd
    for W, b in zip(Wd, bd):
        h = torch.tanh(h @ W.t() + b)
    (h * w.double()).sum().backward()
    assert rel_err(y.detach().cpu(), h.detach().cpu()) <= 1e-5
    assert rel_err(x.grad.cpu(), xd.grad.cpu()) <= 2e-5
    for W, Wr in zip(Ws, Wd):
        assert rel_err(W.grad.cpu(), Wr.grad.cpu()) <= 5e-5      # K = 40 000 truncating accumulations
    for b, br in zip(bs, bd):
        assert rel_err(b.grad.cpu(), br.grad.cpu()) <= 5e-5


@pytest.mark.parametrize("M,N,K", [(8192, 256, 432), (65536, 128, 256), (5000, 432, 256), (66000, 256, 128)])
@pytest.mark.parametrize("variant", ["", "raw", "cluster2"])
def test_ts_engine_matches_fp64(M, N, K, variant, monkeypatch):
    """The opt-in TS engine (A operand through tensor memory, csrc/gemm_pk.cu): plain, with raw fp32 weight stages
    (lo derived in the CTA) and with the weight stages multicast across a 2-CTA cluster.  Run in a subprocess: the
    engine reads its switches once per process."""
    import subprocess
    import sys
    code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from deepctr_torch_b200 import _lib, ops
M, N, K = %d, %d, %d
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(M, K, device="cuda", generator=g); Bm = torch.randn(N, K, device="cuda", generator=g)
C = torch.full((M, N), float("nan"), device="cuda")
ops.ensure_gemm_scratch(torch.device("cuda:0"), M, K, N)
n0 = _lib.launch_count()
_lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(Bm), K, 1, ops._ptr(C), N, 0, ops._stream())
torch.cuda.synchronize()
ref = A.double() @ Bm.double().t()
print("ERR", float((C.double() - ref).abs().max() / ref.abs().max()))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), M, N, K)
    env = dict(os.environ, CTR_GEMM_TS="1", CTR_TS_BRAW="1" if variant == "raw" else "0",
               CTR_TS_CLUSTER="2" if variant == "cluster2" else "1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    err = float([l for l in r.stdout.splitlines() if l.startswith("ERR")][-1].split()[1])
    assert err <= 5e-6, (variant, err)


@pytest.mark.parametrize("B,K,N,masked", [(65536, 432, 256, True), (65536, 256, 128, False), (5000, 432, 256, True),
                                          (4099, 100, 132, True), (4099, 101, 130, True), (70001, 24, 40, False), (8192, 600, 20, True),
                                          (16384, 256, 384, False), (4097, 432, 256, False)])
def test_wgrad_engine_matches_fp64(B, K, N, masked):
    """The weight-gradient engine (csrc/gemm_pk.cu gemm_tsw_kernel: dZ^T through tensor memory, raw row segments by
    bulk copies, fused bias gradient) through ctr_dnn_layer_bwd_chain: dW and db against fp64, with and without the
    act'(Y) mask, with M / N tails inside a tile and a batch tail inside the last 16-sample stage.  Reference:
    autograd of nn.Linear inside DNN (layers/core.py:120-134).  Tolerance: the batch is the contraction, 65 536 samples
    are accumulated by truncating fp32 adds in chains of ~100 stages: 2e-5 of max|dW|."""
    g = torch.Generator(device="cuda").manual_seed(3)
    K4, N4 = (K + 3) // 4 * 4, (N + 3) // 4 * 4
    X = torch.randn(B, K4, device="cuda", generator=g)
    Y = torch.relu(torch.randn(B, N4, device="cuda", generator=g))
    dY = torch.randn(B, N4, device="cuda", generator=g)
    W = torch.randn(N, K4, device="cuda", generator=g)
    dW = torch.full((N, K4), float("nan"), device="cuda")
    db = torch.full((N,), float("nan"), device="cuda")
    ops.ensure_gemm_scratch(torch.device("cuda:0"), B, K, N)
    _lib.call("ctr_dnn_layer_bwd_chain", ops._ptr(X), K4, ops._ptr(W), K4, 1, ops._ptr(Y) if masked else None,
              N4 if masked else 0, ops._ptr(dY), N4, None, 0, ops._ptr(dW), K4, 1, ops._ptr(db), B, K, N,
              1 if masked else 0, 0 if masked else 1, 0, ops._stream())
    torch.cuda.synchronize()
    dz = dY[:, :N].double() * ((Y[:, :N] > 0).double() if masked else 1.0)
    ref_w = dz.t() @ X[:, :K].double()
    ref_b = dz.sum(0)
    assert float((dW[:, :K].double() - ref_w).abs().max() / ref_w.abs().max()) <= 2e-5
    assert float((db.double() - ref_b).abs().max() / ref_b.abs().max()) <= 5e-6


def test_fast_mode_is_single_pass_and_restorable():
    """ctr_set_gemm_passes(1) = the labelled NON-PARITY mode (single-pass TF32): ~1e-3 relative error, and the default
    3xTF32 mode comes back bit-for-bit when the switch is restored (bench.py reports the fast mode as a secondary line)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 32768, 256, 432
    A = torch.randn(M, K, device="cuda", generator=g)
    Bm = torch.randn(N, K, device="cuda", generator=g)
    ref = A.double() @ Bm.double().t()
    ops.ensure_gemm_scratch(torch.device("cuda:0"), M, K, N)

    def run():
        C = torch.full((M, N), float("nan"), device="cuda")
        _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K, 1, ops._ptr(Bm), K, 1, ops._ptr(C), N, 0, ops._stream())
        torch.cuda.synchronize()
        return C

    exact = run()
    prev = _lib.load().ctr_set_gemm_passes(1)
    try:
        assert prev == 3
        fast = run()
    finally:
        assert _lib.load().ctr_set_gemm_passes(prev) == 1
    again = run()
    e_exact = float((exact.double() - ref).abs().max() / ref.abs().max())
    e_fast = float((fast.double() - ref).abs().max() / ref.abs().max())
    assert e_exact <= 5e-6
    assert 1e-5 < e_fast < 5e-3, e_fast
    assert torch.equal(exact, again)


@pytest.mark.parametrize("M,N,K,kind", [(65536, 432, 256, "nn"), (65536, 256, 128, "nn"), (66000, 128, 256, "nt"),
                                        (40000, 300, 64, "nt"), (70001, 130, 100, "nt"), (65536, 256, 432, "nt")])
def test_pp_engine_matches_ss_engine(M, N, K, kind, monkeypatch):
    """The persistent tile-pipelined engine (gemm_pp_kernel: double-buffered accumulators, epilogue of tile t under the
    main loop of tile t+1) against the one-tile-per-CTA SS engine and fp64: same arithmetic order, so bit-identical.
    CTR_GEMM_PP is read at every launch: 1 forces the engine (any K), 0 disables it."""
    g = torch.Generator(device="cuda").manual_seed(9)
    K4 = (K + 3) // 4 * 4
    A = torch.randn(M, K4, device="cuda", generator=g)
    Bm = torch.randn(N, K4, device="cuda", generator=g) if kind == "nt" else torch.randn(K, N, device="cuda", generator=g)
    ops.ensure_gemm_scratch(torch.device("cuda:0"), M, K, N)
    outs = {}
    for pp in ("1", "0"):
        monkeypatch.setenv("CTR_GEMM_PP", pp)
        C = torch.full((M, N), float("nan"), device="cuda")
        if kind == "nt":
            _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K4, 1, ops._ptr(Bm), K4, 1, ops._ptr(C), N, 0, ops._stream())
        else:
            _lib.call("ctr_sgemm", M, N, K, ops._ptr(A), K4, 1, ops._ptr(Bm), 1, N, ops._ptr(C), N, 0, ops._stream())
        torch.cuda.synchronize()
        outs[pp] = C
    ref = A[:, :K].double() @ (Bm[:, :K].double().t() if kind == "nt" else Bm.double())
    assert float((outs["1"].double() - ref).abs().max() / ref.abs().max()) <= 5e-6
    assert torch.equal(outs["1"], outs["0"])
