"""FiBiNET — same constructor and ``state_dict`` as reference ``deepctr_torch/models/fibinet.py:39-102``."""
import torch.nn as nn

from .. import ops
from ..inputs import DenseFeat, SparseFeat, VarLenSparseFeat
from ..layers import DNN, BilinearInteraction, SENETLayer
from .basemodel import BaseModel


class FiBiNET(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, bilinear_type='interaction',
                 reduction_ratio=3, dnn_hidden_units=(128, 128), l2_reg_linear=1e-5,
                 l2_reg_embedding=1e-5, l2_reg_dnn=0, init_std=0.0001, seed=1024, dnn_dropout=0,
                 dnn_activation='relu', task='binary', device='cpu', gpus=None, table_grad="dense"):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_grad=table_grad)
        self.linear_feature_columns = linear_feature_columns
        self.dnn_feature_columns = dnn_feature_columns
        self.field_size = len(self.embedding_dict)
        self.SE = SENETLayer(self.field_size, reduction_ratio, seed, device)
        self.Bilinear = BilinearInteraction(self.field_size, self.embedding_size, bilinear_type, seed, device)
        self.dnn = DNN(self.compute_input_dim(dnn_feature_columns), dnn_hidden_units,
                       activation=dnn_activation, l2_reg=l2_reg_dnn, dropout_rate=dnn_dropout, use_bn=False,
                       init_std=init_std, device=device)
        self.dnn_linear = nn.Linear(dnn_hidden_units[-1], 1, bias=False).to(device)

    def compute_input_dim(self, feature_columns, include_sparse=True, include_dense=True):
        cols = list(feature_columns) if feature_columns else []
        sparse = [c for c in cols if isinstance(c, (SparseFeat, VarLenSparseFeat))]
        dense = [c for c in cols if isinstance(c, DenseFeat)]
        field_size = len(sparse)
        dense_dim = sum(c.dimension for c in dense)
        sparse_dim = field_size * (field_size - 1) * sparse[0].embedding_dim
        return (sparse_dim if include_sparse else 0) + (dense_dim if include_dense else 0)

    def forward(self, X):
        E, dnn_input, lin, _, _ = self.embed(X)
        W = self.Bilinear.stacked_weight()           # shared by both passes (fibinet.py:82-83)
        senet_out = self.SE(E)
        n_emb = E.shape[1] * E.shape[2]
        dense = dnn_input[:, n_emb:] if dnn_input.shape[1] > n_emb else None
        # [bilinear(SENET(E)) | bilinear(E) | dense]: both passes write straight into the tower's input
        x = ops.fibinet_dnn_input(senet_out, E, W, self.Bilinear.bilinear_type, dense)
        dnn_logit = ops.rowdot(self.dnn(x), self.dnn_linear.weight)
        if len(self.linear_feature_columns) > 0 and len(self.dnn_feature_columns) > 0:
            terms = [lin, dnn_logit]
        elif len(self.linear_feature_columns) == 0:
            terms = [dnn_logit]
        elif len(self.dnn_feature_columns) == 0:
            terms = [lin]
        else:
            raise NotImplementedError
        return self.out.forward_terms(terms)
