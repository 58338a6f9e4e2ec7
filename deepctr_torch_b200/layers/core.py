"""DNN tower and PredictionLayer with the reference's constructor arguments and parameter names
(reference ``deepctr_torch/layers/core.py:67-160``); the arithmetic runs in libctr_b200.so."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class DNN(nn.Module):
    """``Linear(+bias) -> [BN] -> activation -> dropout`` stack (reference core.py:67-134).

    ``linears`` holds ``nn.Linear`` modules purely as parameter containers, so ``state_dict`` keys
    (``linears.<i>.weight [out,in]``, ``linears.<i>.bias``) and the default initialisation
    (weights ``N(0, init_std)``, biases torch-default) are the reference's.  Without batch-norm
    each layer is ONE fused kernel call (GEMM + bias + activation).  ``use_bn`` / ``dropout_rate``
    are outside the benchmarked hot path and use torch's CUDA ops between the fused linears.
    """

    def __init__(self, inputs_dim, hidden_units, activation="relu", l2_reg=0, dropout_rate=0,
                 use_bn=False, init_std=0.0001, dice_dim=3, seed=1024, device="cpu"):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.seed = seed
        self.l2_reg = l2_reg
        self.use_bn = use_bn
        if len(hidden_units) == 0:
            raise ValueError("hidden_units is empty!!")
        if not isinstance(activation, str) or activation.lower() not in tuple(ops.ACT_CODES) + ("prelu", "dice"):
            raise NotImplementedError("DNN activation %r is not implemented by the CUDA tower" % (activation,))
        self.activation = activation.lower()
        units = [inputs_dim] + list(hidden_units)
        self.linears = nn.ModuleList([nn.Linear(units[i], units[i + 1]) for i in range(len(units) - 1)])
        if self.use_bn:
            self.bn = nn.ModuleList([nn.BatchNorm1d(units[i + 1]) for i in range(len(units) - 1)])
        if self.activation == "prelu":
            # same container and key names as the reference (`activation_layers.<i>.weight`, core.py:110-111)
            self.activation_layers = nn.ModuleList([nn.PReLU() for _ in range(len(units) - 1)])
        if self.activation == "dice":
            # batch-normalising activation (reference activation.py:6-45): torch CUDA ops between the fused linears,
            # keys `activation_layers.<i>.bn.*` / `.alpha` as in the reference
            from .activation import Dice
            self.activation_layers = nn.ModuleList([Dice(units[i + 1], dice_dim) for i in range(len(units) - 1)])
        for name, tensor in self.linears.named_parameters():
            if "weight" in name:
                nn.init.normal_(tensor, mean=0, std=init_std)
        self.to(device)

    def forward(self, inputs):
        x = inputs
        if self.activation == "dice":
            for i, lin in enumerate(self.linears):
                x = ops.dnn_layer(x, lin.weight, lin.bias, "linear")
                if self.use_bn:
                    x = self.bn[i](x)
                x = self.activation_layers[i](x)
                if self.dropout_rate > 0:
                    x = F.dropout(x, self.dropout_rate, self.training)
            return x
        if self.activation == "prelu":
            for i, lin in enumerate(self.linears):
                x = ops.dnn_layer(x, lin.weight, lin.bias, "linear")
                if self.use_bn:
                    x = self.bn[i](x)
                x = ops.prelu(x, self.activation_layers[i].weight)
                if self.dropout_rate > 0:
                    x = F.dropout(x, self.dropout_rate, self.training)
            return x
        if not self.use_bn and not (self.dropout_rate > 0 and self.training):
            # the whole stack as one autograd node: the backward hands dZ from layer to layer
            return ops.dnn_tower(x, self.activation, [lin.weight for lin in self.linears],
                                 [lin.bias for lin in self.linears])
        for i, lin in enumerate(self.linears):
            if self.use_bn:
                x = ops.dnn_layer(x, lin.weight, lin.bias, "linear")
                x = self.bn[i](x)
                x = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
                     "linear": lambda t: t}[self.activation](x)
            else:
                x = ops.dnn_layer(x, lin.weight, lin.bias, self.activation)
            if self.dropout_rate > 0:
                x = F.dropout(x, self.dropout_rate, self.training)
        return x


class PredictionLayer(nn.Module):
    """``sigmoid(sum of branch logits + bias)`` (reference core.py:137-160).

    Called with a single tensor it behaves like the reference module; the models call
    ``forward_terms`` so that the branch sum, the bias and the sigmoid are one kernel."""

    def __init__(self, task="binary", use_bias=True, **kwargs):
        if task not in ["binary", "multiclass", "regression"]:
            raise ValueError("task must be binary,multiclass or regression")
        super().__init__()
        self.use_bias = use_bias
        self.task = task
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros((1,)))

    def forward_terms(self, terms):
        return ops.predict(list(terms), self.bias if self.use_bias else None, self.task == "binary")

    def forward(self, X):
        return self.forward_terms([X])
