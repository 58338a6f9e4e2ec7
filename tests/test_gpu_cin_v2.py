"""Parity of the second-generation CIN forward / weight gradient (csrc/cin_v2.cu, the default since round 2)
against the fp64 oracle, and against the round-1 kernels (CTR_CIN_V2=0)."""
import os

import pytest
import torch

from deepctr_torch_b200 import ops
from helpers import rel_err
from oracle import ctr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,M,D,sizes,split", [(300, 26, 16, (32, 16), True), (1000, 26, 16, (128, 128), True),
                                               (77, 7, 8, (6, 4, 3), False), (4096, 26, 16, (200,), True)])
def test_cin_v2_forward_matches_fp64(B, M, D, sizes, split):
    g = torch.Generator(device=DEV).manual_seed(B + M)
    E = torch.randn(B, M, D, device=DEV, generator=g) * 0.5
    params, P, H = [], {}, M
    for k, n in enumerate(sizes):
        W = torch.randn(n, H * M, 1, device=DEV, generator=g) * 0.1
        b = torch.randn(n, device=DEV, generator=g) * 0.1
        params += [W, b]
        P["conv1ds.%d.weight" % k], P["conv1ds.%d.bias" % k] = W, b
        H = n // 2 if (split and k != len(sizes) - 1) else n
    old = os.environ.get("CTR_CIN_V2")
    os.environ["CTR_CIN_V2"] = "1"
    try:
        out = ops.cin(E, sizes, split, "relu", params)
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("CTR_CIN_V2", None)
        else:
            os.environ["CTR_CIN_V2"] = old
    P64 = {k: v.double().cpu() for k, v in P.items()}
    ref = O.cin(P64, "", E.double().cpu(), sizes, split, "relu")
    assert rel_err(out.cpu(), ref) <= 1e-5


@pytest.mark.parametrize("B,M,sizes", [(300, 26, (32, 16)), (2048, 26, (128, 128))])
def test_cin_v2_weight_gradient_matches_fp64(B, M, sizes):
    """dW through cin_v2_dw_kernel (D = 16) — everything else of the backward stays on the round-1 kernels."""
    D, split = 16, True
    g = torch.Generator(device=DEV).manual_seed(B)
    E = (torch.randn(B, M, D, device=DEV, generator=g) * 0.5).requires_grad_(True)
    params, P, H = [], {}, M
    for k, n in enumerate(sizes):
        W = (torch.randn(n, H * M, 1, device=DEV, generator=g) * 0.1).requires_grad_(True)
        b = (torch.randn(n, device=DEV, generator=g) * 0.1).requires_grad_(True)
        params += [W, b]
        P["conv1ds.%d.weight" % k], P["conv1ds.%d.bias" % k] = W, b
        H = n // 2 if k != len(sizes) - 1 else n
    old = os.environ.get("CTR_CIN_V2")
    os.environ["CTR_CIN_V2"] = "1"
    try:
        out = ops.cin(E, sizes, split, "tanh", params)
        w = torch.randn(out.shape, device=DEV, generator=g)
        (out * w).sum().backward()
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("CTR_CIN_V2", None)
        else:
            os.environ["CTR_CIN_V2"] = old
    P64 = {k: v.detach().double().cpu().requires_grad_(True) for k, v in P.items()}
    E64 = E.detach().double().cpu().requires_grad_(True)
    ref = O.cin(P64, "", E64, sizes, split, "tanh")
    (ref * w.double().cpu()).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) <= 1e-5
    for k in P:
        assert rel_err(P[k].grad.cpu(), P64[k].grad) <= 1e-4, k
    assert rel_err(E.grad.cpu(), E64.grad) <= 1e-4
