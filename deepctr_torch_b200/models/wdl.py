"""WDL — same constructor and ``state_dict`` as reference ``deepctr_torch/models/wdl.py:38-80``:
the fused gather's linear term (wide) + the DNN tower on ``combined_dnn_input`` (deep)."""
import torch.nn as nn

from .. import ops
from ..layers import DNN
from .basemodel import BaseModel


class WDL(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 128), l2_reg_linear=1e-5,
                 l2_reg_embedding=1e-5, l2_reg_dnn=0, init_std=0.0001, seed=1024, dnn_dropout=0, dnn_activation='relu',
                 dnn_use_bn=False, task='binary', device='cpu', gpus=None, table_grad="dense"):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_grad=table_grad)
        self.use_dnn = len(dnn_feature_columns) > 0 and len(dnn_hidden_units) > 0
        if self.use_dnn:
            self.dnn = DNN(self.compute_input_dim(dnn_feature_columns), dnn_hidden_units,
                           activation=dnn_activation, l2_reg=l2_reg_dnn, dropout_rate=dnn_dropout, use_bn=dnn_use_bn,
                           init_std=init_std, device=device)
            self.dnn_linear = nn.Linear(dnn_hidden_units[-1], 1, bias=False).to(device)
            self.add_regularization_weight(
                filter(lambda x: 'weight' in x[0] and 'bn' not in x[0], self.dnn.named_parameters()), l2=l2_reg_dnn)
            self.add_regularization_weight(self.dnn_linear.weight, l2=l2_reg_dnn)
        self.to(device)

    def forward(self, X):
        E, dnn_input, lin, _, blk = self.embed(X, want_blk=self.use_dnn)
        terms = [lin]
        if self.use_dnn:
            terms.append(ops.rowdot(self.dnn(blk if blk is not None else dnn_input), self.dnn_linear.weight))
        return self.out.forward_terms(terms)
