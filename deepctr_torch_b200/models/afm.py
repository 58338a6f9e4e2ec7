"""AFM — same constructor and ``state_dict`` as reference ``deepctr_torch/models/afm.py:38-72``
(``fm.attention_W``, ``fm.attention_b``, ``fm.projection_h``, ``fm.projection_p``)."""
from ..inputs import split_columns
from ..layers import AFMLayer
from .basemodel import BaseModel


class AFM(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, use_attention=True, attention_factor=8,
                 l2_reg_linear=1e-5, l2_reg_embedding=1e-5, l2_reg_att=1e-5, afm_dropout=0, init_std=0.0001, seed=1024,
                 task='binary', device='cpu', gpus=None, table_grad="dense"):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_grad=table_grad)
        self.use_attention = use_attention
        if use_attention:
            self.fm = AFMLayer(self.embedding_size, attention_factor, l2_reg_att, afm_dropout, seed, device)
            self.add_regularization_weight(self.fm.attention_W, l2=l2_reg_att)
        else:
            from ..layers import FM
            self.fm = FM()
        self.to(device)

    def forward(self, X):
        _, dense, _ = split_columns(self.dnn_feature_columns)
        if len(dense) > 0:                              # support_dense=False (reference afm.py:57-58)
            raise ValueError("DenseFeat is not supported in dnn_feature_columns")
        has_emb = self._gather_plan(X.device).n_emb + len(self._plan.varlen) > 0
        E, _, lin, fm, _ = self.embed(X, want_fm=has_emb and not self.use_attention, want_blk=has_emb)
        terms = [lin]
        if has_emb:
            terms.append(self.fm(E).squeeze(1) if self.use_attention else fm)
        return self.out.forward_terms(terms)
