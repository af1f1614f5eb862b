"""Parameter containers mirroring prototype/model/image_encoder/base_transformer.py (identical copy in
text_encoder/): same class names, attribute names and nn.Module types, so state_dict keys and the
`isinstance`-based weight-decay grouping in prototype/utils/misc.py:296-381 are unchanged.

The arithmetic does NOT happen here: the owning tower runs all layers in one C-ABI call
(csrc/encoder.cu).  These modules only hold the fp32 master parameters.
"""
from collections import OrderedDict

import torch
from torch import nn


class LayerNorm(nn.LayerNorm):
    """base_transformer.py:10-18 (container only; eps 1e-5)."""


class QuickGELU(nn.Module):
    """base_transformer.py:24-26: x * sigmoid(1.702 x) — fused into the c_fc GEMM epilogue (csrc/gemm.cu)."""

    def forward(self, x):
        raise RuntimeError("declip_b200: QuickGELU is fused into the GEMM epilogue; run the owning tower instead")


class ResidualAttentionBlock(nn.Module):
    """base_transformer.py:29-53.  attn is a real nn.MultiheadAttention so `attn.in_proj_weight`,
    `attn.in_proj_bias`, `attn.out_proj.{weight,bias}` exist with the reference's shapes."""

    def __init__(self, d_model, n_head, attn_mask=None, dropout=0.):
        super().__init__()
        if dropout != 0.:
            raise NotImplementedError("declip_b200: attention dropout is 0 in every reference config")
        self.attn = nn.MultiheadAttention(d_model, n_head, dropout=dropout)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def forward(self, x):
        raise RuntimeError("declip_b200: blocks are executed by the owning tower (dc_vit_forward / dc_text_forward)")


class Transformer(nn.Module):
    """base_transformer.py:56-79."""

    def __init__(self, width, layers, heads, attn_mask=None, checkpoint=False, dropout=0., emb_dropout=0.):
        super().__init__()
        if emb_dropout != 0.:
            raise NotImplementedError("declip_b200: embedding dropout is 0 in every reference config")
        self.width = width
        self.layers = layers
        self.heads = heads
        self.checkpoint = checkpoint   # activations are saved in one workspace; flag kept for config compatibility
        self.dropout = nn.Dropout(emb_dropout)
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask, dropout=dropout)
                                         for _ in range(layers)])

    def forward(self, x):
        raise RuntimeError("declip_b200: Transformer is executed by the owning tower")
