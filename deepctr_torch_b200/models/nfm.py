"""NFM — same constructor and ``state_dict`` as reference ``deepctr_torch/models/nfm.py:38-84``:
bi-interaction pooling of the embedding block (one kernel) feeding the DNN tower."""
import torch
import torch.nn as nn

from .. import ops
from ..layers import DNN, BiInteractionPooling
from .basemodel import BaseModel


class NFM(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(128, 128), l2_reg_embedding=1e-5,
                 l2_reg_linear=1e-5, l2_reg_dnn=0, init_std=0.0001, seed=1024, bi_dropout=0, dnn_dropout=0,
                 dnn_activation='relu', task='binary', device='cpu', gpus=None, table_grad="dense"):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_grad=table_grad)
        self.dnn = DNN(self.compute_input_dim(dnn_feature_columns, include_sparse=False) + self.embedding_size,
                       dnn_hidden_units, activation=dnn_activation, l2_reg=l2_reg_dnn, dropout_rate=dnn_dropout,
                       use_bn=False, init_std=init_std, device=device)
        self.dnn_linear = nn.Linear(dnn_hidden_units[-1], 1, bias=False).to(device)
        self.add_regularization_weight(
            filter(lambda x: 'weight' in x[0] and 'bn' not in x[0], self.dnn.named_parameters()), l2=l2_reg_dnn)
        self.add_regularization_weight(self.dnn_linear.weight, l2=l2_reg_dnn)
        self.bi_pooling = BiInteractionPooling()
        self.bi_dropout = bi_dropout
        if self.bi_dropout > 0:
            self.dropout = nn.Dropout(bi_dropout)
        self.to(device)

    def forward(self, X):
        E, dnn_input, lin, _, _ = self.embed(X)
        bi_out = self.bi_pooling(E)                     # [B,1,D]
        if self.bi_dropout:
            bi_out = self.dropout(bi_out)
        B, F, D = E.shape
        parts = [bi_out.reshape(B, D)]
        if dnn_input.shape[1] > F * D:                  # dense columns follow the embedding block
            parts.append(dnn_input[:, F * D:])
        x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        return self.out.forward_terms([lin, ops.rowdot(self.dnn(x), self.dnn_linear.weight)])
