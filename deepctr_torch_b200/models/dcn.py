"""DCN / DCN-M — same constructor and ``state_dict`` as reference ``deepctr_torch/models/dcn.py:44-96``."""
import torch
import torch.nn as nn

from .. import ops
from ..layers import DNN, CrossNet
from .basemodel import BaseModel


class _DeepCrossBase(BaseModel):
    """Shared wiring of DCN and DCNMix: ``dnn_linear(cat(cross_out, deep_out))`` is evaluated as two
    row-dots on the two halves of the weight, so the concatenation is never materialised."""

    def _head(self, X):
        E, dnn_input, lin, _, blk = self.embed(X)
        deep_in = blk if blk is not None else dnn_input
        terms = [lin]
        n_in = dnn_input.shape[1]
        w = self.dnn_linear.weight
        if len(self.dnn_hidden_units) > 0 and self.cross_num > 0:
            cross_out = self.crossnet(dnn_input)
            deep_out = self.dnn(deep_in)
            terms.append(ops.rowdot(cross_out, w[:, :n_in]))
            terms.append(ops.rowdot(deep_out, w[:, n_in:]))
        elif len(self.dnn_hidden_units) > 0:
            terms.append(ops.rowdot(self.dnn(deep_in), w))
        elif self.cross_num > 0:
            terms.append(ops.rowdot(self.crossnet(dnn_input), w))
        return self.out.forward_terms(terms)


class DCN(_DeepCrossBase):
    def __init__(self, linear_feature_columns, dnn_feature_columns, cross_num=2, cross_parameterization='vector',
                 dnn_hidden_units=(128, 128), l2_reg_linear=0.00001, l2_reg_embedding=0.00001,
                 l2_reg_cross=0.00001, l2_reg_dnn=0, init_std=0.0001, seed=1024, dnn_dropout=0,
                 dnn_activation='relu', dnn_use_bn=False, task='binary', device='cpu', gpus=None,
                 table_grad="dense"):
        super().__init__(linear_feature_columns=linear_feature_columns, dnn_feature_columns=dnn_feature_columns,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_grad=table_grad)
        self.dnn_hidden_units = dnn_hidden_units
        self.cross_num = cross_num
        self.dnn = DNN(self.compute_input_dim(dnn_feature_columns), dnn_hidden_units, activation=dnn_activation,
                       use_bn=dnn_use_bn, l2_reg=l2_reg_dnn, dropout_rate=dnn_dropout, init_std=init_std,
                       device=device)
        if len(self.dnn_hidden_units) > 0 and self.cross_num > 0:
            dnn_linear_in_feature = self.compute_input_dim(dnn_feature_columns) + dnn_hidden_units[-1]
        elif len(self.dnn_hidden_units) > 0:
            dnn_linear_in_feature = dnn_hidden_units[-1]
        elif self.cross_num > 0:
            dnn_linear_in_feature = self.compute_input_dim(dnn_feature_columns)
        self.dnn_linear = nn.Linear(dnn_linear_in_feature, 1, bias=False).to(device)
        self.crossnet = CrossNet(in_features=self.compute_input_dim(dnn_feature_columns), layer_num=cross_num,
                                 parameterization=cross_parameterization, device=device)
        self.add_regularization_weight(
            filter(lambda x: 'weight' in x[0] and 'bn' not in x[0], self.dnn.named_parameters()), l2=l2_reg_dnn)
        self.add_regularization_weight(self.dnn_linear.weight, l2=l2_reg_linear)
        self.add_regularization_weight(self.crossnet.kernels, l2=l2_reg_cross)
        self.to(device)

    def forward(self, X):
        return self._head(X)
