"""IFM — same constructor and ``state_dict`` as reference ``deepctr_torch/models/ifm.py:38-93``:
the factor-estimating tower produces one input-aware weight per field; the re-weighting of the
embedding block, of the per-field linear weights and the softmax are ONE kernel (``ops.refine``)."""
import torch.nn as nn

from .. import ops
from ..inputs import SparseFeat, VarLenSparseFeat
from ..layers import DNN, FM
from .basemodel import BaseModel


class IFM(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 128), l2_reg_linear=0.00001,
                 l2_reg_embedding=0.00001, l2_reg_dnn=0, init_std=0.0001, seed=1024, dnn_dropout=0,
                 dnn_activation='relu', dnn_use_bn=False, task='binary', device='cpu', gpus=None, table_grad="dense"):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_grad=table_grad)
        if not len(dnn_hidden_units) > 0:
            raise ValueError("dnn_hidden_units is null!")
        self.fm = FM()
        self.factor_estimating_net = DNN(self.compute_input_dim(dnn_feature_columns, include_dense=False),
                                         dnn_hidden_units, activation=dnn_activation, l2_reg=l2_reg_dnn,
                                         dropout_rate=dnn_dropout, use_bn=dnn_use_bn, init_std=init_std, device=device)
        self.sparse_feat_num = len([c for c in dnn_feature_columns if isinstance(c, (SparseFeat, VarLenSparseFeat))])
        self.transform_weight_matrix_P = nn.Linear(dnn_hidden_units[-1], self.sparse_feat_num, bias=False).to(device)
        self.add_regularization_weight(
            filter(lambda x: 'weight' in x[0] and 'bn' not in x[0], self.factor_estimating_net.named_parameters()),
            l2=l2_reg_dnn)
        self.add_regularization_weight(self.transform_weight_matrix_P.weight, l2=l2_reg_dnn)
        self.to(device)

    def forward(self, X):
        E, _, _, _, _ = self.embed(X)
        if E is None or E.shape[1] == 0:
            raise ValueError("there are no sparse features")
        B, F, D = E.shape
        h = self.factor_estimating_net(E.reshape(B, F * D))
        P = ops.dnn_layer(h, self.transform_weight_matrix_P.weight, None, "linear")        # m'_x  [B,F]
        L, lin_dense = self.linear_field_terms(X)
        Er, lin_sparse = ops.refine(P, E, L, softmax=True)                                   # m = F*softmax(P)
        terms = [self.fm(Er).squeeze(1)]
        if lin_sparse is not None:
            terms.append(lin_sparse)
        if lin_dense is not None:
            terms.append(lin_dense)
        return self.out.forward_terms(terms)
