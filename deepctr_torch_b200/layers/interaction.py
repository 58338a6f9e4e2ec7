"""Interaction layers with the reference's constructors and parameter names
(reference ``deepctr_torch/layers/interaction.py``); every forward/backward is a CUDA kernel
of libctr_b200.so reached through ``deepctr_torch_b200.ops``."""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn

from .. import ops


class FM(nn.Module):
    """``0.5 * sum_d((sum_f e)^2 - sum_f e^2)`` -> ``[B,1]`` (reference interaction.py:12-34).
    Inside the models the FM term is produced by the fused gather kernel; this module is the
    stand-alone layer for callers that hold an assembled ``[B,F,D]`` block."""

    def forward(self, inputs):
        if inputs.dim() != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % inputs.dim())
        return ops.fm(inputs)


class SENETLayer(nn.Module):
    """reference interaction.py:64-101; ``excitation.0.weight [R,F]``, ``excitation.2.weight [F,R]``."""

    def __init__(self, filed_size, reduction_ratio=3, seed=1024, device="cpu"):
        super().__init__()
        self.seed = seed
        self.filed_size = filed_size
        self.reduction_size = max(1, filed_size // reduction_ratio)
        self.excitation = nn.Sequential(
            nn.Linear(self.filed_size, self.reduction_size, bias=False), nn.ReLU(),
            nn.Linear(self.reduction_size, self.filed_size, bias=False), nn.ReLU())
        self.to(device)

    def forward(self, inputs):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        return ops.senet(inputs, self.excitation[0].weight, self.excitation[2].weight)


class BilinearInteraction(nn.Module):
    """reference interaction.py:104-156; weights are ``nn.Linear(E,E,bias=False)`` containers in the
    reference's ModuleList layout (``bilinear.<p>.weight`` / ``bilinear.weight``)."""

    def __init__(self, filed_size, embedding_size, bilinear_type="interaction", seed=1024, device="cpu"):
        super().__init__()
        self.bilinear_type = bilinear_type
        self.seed = seed
        self.bilinear = nn.ModuleList()
        if self.bilinear_type == "all":
            self.bilinear = nn.Linear(embedding_size, embedding_size, bias=False)
        elif self.bilinear_type == "each":
            for _ in range(filed_size):
                self.bilinear.append(nn.Linear(embedding_size, embedding_size, bias=False))
        elif self.bilinear_type == "interaction":
            for _ in itertools.combinations(range(filed_size), 2):
                self.bilinear.append(nn.Linear(embedding_size, embedding_size, bias=False))
        else:
            raise NotImplementedError
        self.to(device)

    def stacked_weight(self):
        if self.bilinear_type == "all":
            return self.bilinear.weight.unsqueeze(0)
        return torch.stack([m.weight for m in self.bilinear], dim=0)

    def forward(self, inputs, weight=None):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        W = self.stacked_weight() if weight is None else weight
        return ops.bilinear(inputs, W, self.bilinear_type)


class CIN(nn.Module):
    """reference interaction.py:159-248; ``conv1ds.<k>`` are ``nn.Conv1d(H*M, size, 1)`` containers
    (default torch init, like the reference)."""

    def __init__(self, field_size, layer_size=(128, 128), activation="relu", split_half=True, l2_reg=1e-5,
                 seed=1024, device="cpu"):
        super().__init__()
        if len(layer_size) == 0:
            raise ValueError("layer_size must be a list(tuple) of length greater than 1")
        self.layer_size = layer_size
        self.field_nums = [field_size]
        self.split_half = split_half
        if not isinstance(activation, str) or activation.lower() not in ops.ACT_CODES:
            raise NotImplementedError("CIN activation %r is not implemented by the CUDA kernels" % (activation,))
        self.activation = activation.lower()
        self.l2_reg = l2_reg
        self.seed = seed
        self.conv1ds = nn.ModuleList()
        for i, size in enumerate(self.layer_size):
            self.conv1ds.append(nn.Conv1d(self.field_nums[-1] * self.field_nums[0], size, 1))
            if self.split_half:
                if i != len(self.layer_size) - 1 and size % 2 > 0:
                    raise ValueError("layer_size must be even number except for the last layer when split_half=True")
                self.field_nums.append(size // 2)
            else:
                self.field_nums.append(size)
        self.to(device)

    def forward(self, inputs):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        params = []
        for conv in self.conv1ds:
            params += [conv.weight, conv.bias]
        return ops.cin(inputs, self.layer_size, self.split_half, self.activation, params)


class CrossNet(nn.Module):
    """reference interaction.py:397-453; ``kernels [L,in,1|in]`` xavier-normal, ``bias [L,in,1]`` zeros."""

    def __init__(self, in_features, layer_num=2, parameterization="vector", seed=1024, device="cpu"):
        super().__init__()
        self.layer_num = layer_num
        self.parameterization = parameterization
        if self.parameterization == "vector":
            self.kernels = nn.Parameter(torch.Tensor(self.layer_num, in_features, 1))
        elif self.parameterization == "matrix":
            self.kernels = nn.Parameter(torch.Tensor(self.layer_num, in_features, in_features))
        else:
            raise ValueError("parameterization should be 'vector' or 'matrix'")
        self.bias = nn.Parameter(torch.Tensor(self.layer_num, in_features, 1))
        for i in range(self.kernels.shape[0]):
            nn.init.xavier_normal_(self.kernels[i])
        for i in range(self.bias.shape[0]):
            nn.init.zeros_(self.bias[i])
        self.to(device)

    def forward(self, inputs):
        return ops.crossnet(inputs, self.kernels, self.bias, self.parameterization)


class CrossNetMix(nn.Module):
    """reference interaction.py:456-534 (DCN-Mix): per layer and expert
    ``x0 (.) (U tanh(C tanh(V^T x_l)) + b)`` mixed with softmax gates + residual.

    The low-rank projections are fused GEMM(+tanh) kernel calls on the ``[in,r]`` / ``[r,r]``
    factors in place; the gate softmax, Hadamard product, mixture and residual are one kernel."""

    def __init__(self, in_features, low_rank=32, num_experts=4, layer_num=2, device="cpu"):
        super().__init__()
        self.layer_num = layer_num
        self.num_experts = num_experts
        self.U_list = nn.Parameter(torch.Tensor(self.layer_num, num_experts, in_features, low_rank))
        self.V_list = nn.Parameter(torch.Tensor(self.layer_num, num_experts, in_features, low_rank))
        self.C_list = nn.Parameter(torch.Tensor(self.layer_num, num_experts, low_rank, low_rank))
        self.gating = nn.ModuleList([nn.Linear(in_features, 1, bias=False) for _ in range(self.num_experts)])
        self.bias = nn.Parameter(torch.Tensor(self.layer_num, in_features, 1))
        for para in (self.U_list, self.V_list, self.C_list):
            for i in range(self.layer_num):
                nn.init.xavier_normal_(para[i])
        for i in range(len(self.bias)):
            nn.init.zeros_(self.bias[i])
        self.to(device)

    def forward(self, inputs):
        x0 = inputs
        xl = x0
        gate_w = torch.cat([g.weight for g in self.gating], dim=0)           # [E, in]
        for i in range(self.layer_num):
            gate = ops.dnn_layer(xl, gate_w, None, "linear")                 # [B, E]
            uvs = []
            for e in range(self.num_experts):
                v = ops.dnn_layer(xl, self.V_list[i, e], None, "tanh", w_kn=True)      # tanh(x V)   [B,r]
                v = ops.dnn_layer(v, self.C_list[i, e], None, "tanh")                  # tanh(v C^T) [B,r]
                uvs.append(ops.dnn_layer(v, self.U_list[i, e], None, "linear"))        # v U^T       [B,in]
            xl = ops.cross_mix_combine(x0, xl, torch.stack(uvs, dim=0), gate, self.bias[i])
        return xl


class BiInteractionPooling(nn.Module):
    """``0.5 ((sum_f e)^2 - sum_f e^2)`` keeping the embedding axis, ``[B,F,D] -> [B,1,D]``
    (reference interaction.py:37-61); one kernel."""

    def forward(self, inputs):
        if inputs.dim() != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % inputs.dim())
        return ops.bi_interaction_pooling(inputs)


class AFMLayer(nn.Module):
    """Attentional FM (reference interaction.py:250-331): parameters ``attention_W [D,A]``,
    ``attention_b [A]``, ``projection_h [A,1]``, ``projection_p [D,1]`` (xavier-normal / zeros like the
    reference).  Accepts the reference's list of ``[B,1,D]`` tensors or an assembled ``[B,F,D]`` block;
    pair products, attention net, softmax over the pairs and the weighted sum are one kernel."""

    def __init__(self, in_features, attention_factor=4, l2_reg_w=0, dropout_rate=0, seed=1024, device="cpu"):
        super().__init__()
        self.attention_factor = attention_factor
        self.l2_reg_w = l2_reg_w
        self.dropout_rate = dropout_rate
        self.seed = seed
        self.attention_W = nn.Parameter(torch.Tensor(in_features, attention_factor))
        self.attention_b = nn.Parameter(torch.Tensor(attention_factor))
        self.projection_h = nn.Parameter(torch.Tensor(attention_factor, 1))
        self.projection_p = nn.Parameter(torch.Tensor(in_features, 1))
        for tensor in [self.attention_W, self.projection_h, self.projection_p]:
            nn.init.xavier_normal_(tensor)
        nn.init.zeros_(self.attention_b)
        self.dropout = nn.Dropout(dropout_rate)
        self.to(device)

    def forward(self, inputs):
        E = torch.cat(list(inputs), dim=1) if isinstance(inputs, (list, tuple)) else inputs
        att = ops.afm_attention(E, self.attention_W, self.attention_b, self.projection_h)     # [B,D]
        if self.dropout_rate > 0:
            att = self.dropout(att)
        return ops.rowdot(att, self.projection_p).unsqueeze(1)


class InteractingLayer(nn.Module):
    """AutoInt's multi-head self-attention over the fields with residual + relu (reference
    interaction.py:334-394): ``W_Query / W_key / W_Value / W_Res [D,D]``, ``N(0, 0.05)``.  The four
    projections are GEMM kernel calls on the ``[B*F, D]`` view, the attention core is one kernel."""

    def __init__(self, embedding_size, head_num=2, use_res=True, scaling=False, seed=1024, device="cpu"):
        super().__init__()
        if head_num <= 0:
            raise ValueError("head_num must be a int > 0")
        if embedding_size % head_num != 0:
            raise ValueError("embedding_size is not an integer multiple of head_num!")
        self.att_embedding_size = embedding_size // head_num
        self.head_num = head_num
        self.use_res = use_res
        self.scaling = scaling
        self.seed = seed
        self.W_Query = nn.Parameter(torch.Tensor(embedding_size, embedding_size))
        self.W_key = nn.Parameter(torch.Tensor(embedding_size, embedding_size))
        self.W_Value = nn.Parameter(torch.Tensor(embedding_size, embedding_size))
        if self.use_res:
            self.W_Res = nn.Parameter(torch.Tensor(embedding_size, embedding_size))
        for tensor in self.parameters():
            nn.init.normal_(tensor, mean=0.0, std=0.05)
        self.to(device)

    def forward(self, inputs):
        if len(inputs.shape) != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % (len(inputs.shape)))
        B, F, D = inputs.shape
        x = inputs.reshape(B * F, D)
        q = ops.dnn_layer(x, self.W_Query, None, "linear", w_kn=True).view(B, F, D)
        k = ops.dnn_layer(x, self.W_key, None, "linear", w_kn=True).view(B, F, D)
        v = ops.dnn_layer(x, self.W_Value, None, "linear", w_kn=True).view(B, F, D)
        r = ops.dnn_layer(x, self.W_Res, None, "linear", w_kn=True).view(B, F, D) if self.use_res \
            else torch.zeros_like(q)
        scale = 1.0 / self.att_embedding_size ** 0.5 if self.scaling else 1.0
        return ops.field_attention(q, k, v, r, self.head_num, scale)
