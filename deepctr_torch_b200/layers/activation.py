"""Activation modules that carry parameters (reference ``deepctr_torch/layers/activation.py:6-84``).

The parameter-free activations (relu / sigmoid / tanh / linear) are fused into the GEMM epilogues of
``libctr_b200.so``; ``prelu`` has its own kernels (``ops.prelu``).  ``Dice`` normalises over the batch, so it sits
between two fused linears as torch CUDA ops, exactly like ``dnn_use_bn`` — it is not part of any BASELINE config.
Parameter names (``bn.*``, ``alpha``) are the reference's, so ``state_dict``s are interchangeable."""
from __future__ import annotations

import torch
import torch.nn as nn


class Dice(nn.Module):
    """``p = sigmoid(BN(x));  y = p * x + alpha * (1 - p) * x`` with a learned per-feature ``alpha`` (zeros at start).
    ``dim == 2``: x is ``[batch, features]``; ``dim == 3``: x is ``[batch, n, features]`` and the statistics run over
    batch and n per feature."""

    def __init__(self, emb_size, dim=2, epsilon=1e-8, device="cpu"):
        super().__init__()
        if dim not in (2, 3):
            raise ValueError("Dice: dim must be 2 or 3")
        self.dim = dim
        self.bn = nn.BatchNorm1d(emb_size, eps=epsilon)
        self.alpha = nn.Parameter(torch.zeros((emb_size,) if dim == 2 else (emb_size, 1)))
        self.to(device)

    def forward(self, x):
        if x.dim() != self.dim:
            raise ValueError("Dice(dim=%d) got a %d-d input" % (self.dim, x.dim()))
        if self.dim == 3:
            xt = x.transpose(1, 2)                       # BatchNorm1d wants the feature axis second
            gate = torch.sigmoid(self.bn(xt))
            return (gate * xt + self.alpha * (1.0 - gate) * xt).transpose(1, 2)
        gate = torch.sigmoid(self.bn(x))
        return gate * x + self.alpha * (1.0 - gate) * x


class Identity(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, inputs):
        return inputs


def activation_layer(act_name, hidden_size=None, dice_dim=2):
    """Module for an activation name (or an ``nn.Module`` subclass), reference ``activation.py:57-84``."""
    if isinstance(act_name, str):
        name = act_name.lower()
        table = {"sigmoid": nn.Sigmoid, "linear": Identity, "relu": nn.ReLU, "tanh": nn.Tanh, "prelu": nn.PReLU}
        if name == "dice":
            if not dice_dim:
                raise ValueError("dice needs dice_dim")
            return Dice(hidden_size, dice_dim)
        if name in table:
            return table[name]()
        raise NotImplementedError("activation %r" % (act_name,))
    if isinstance(act_name, type) and issubclass(act_name, nn.Module):
        return act_name()
    raise NotImplementedError("activation %r" % (act_name,))
