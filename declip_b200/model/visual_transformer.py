"""VisualTransformer — mirror of prototype/model/image_encoder/visual_transformer.py (ViT-B/32, B/16).

Same constructor, parameter names/shapes, init and forward signature; the forward is one call into the
CUDA executor dc_vit_forward (csrc/encoder.cu) and the backward one call into dc_vit_backward.
"""
import torch
from torch import nn

from ..runtime import TowerRuntime, run_tower
from .base_transformer import LayerNorm, Transformer


class VisualTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, embed_dim, checkpoint, dropout=0,
                 emb_dropout=0):
        super().__init__()
        self.input_resolution = input_resolution
        self.patch_size = patch_size
        self.output_dim = embed_dim
        self.freeze_conv1 = True                                            # visual_transformer.py:12
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, checkpoint=checkpoint, dropout=dropout,
                                       emb_dropout=emb_dropout)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, embed_dim))
        self.initialize_parameters()
        self._rt = TowerRuntime("vit", self, layers, width, heads, (input_resolution // patch_size) ** 2 + 1, embed_dim,
                                res=input_resolution, patch=patch_size)

    def initialize_parameters(self):
        # visual_transformer.py:29-39
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)

    def train(self, mode=True):
        # visual_transformer.py:41-52: conv1 stays frozen / in eval mode
        self.training = mode
        for module in self.children():
            module.train(mode)
        if self.freeze_conv1:
            self.conv1.eval()
            for p in self.conv1.parameters():
                p.requires_grad = False
        return self

    def forward(self, x, return_dense=False, return_feature=False):
        """visual_transformer.py:55-82: x [, dense_feat] [, feature] — `feature` is ln_post(class token) before `proj`."""
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.input_resolution or x.shape[3] != self.input_resolution:
            raise ValueError("expected images [B,3,%d,%d], got %s" % (self.input_resolution, self.input_resolution,
                                                                      tuple(x.shape)))
        if not (return_dense or return_feature):
            return run_tower(self._rt, x)
        out = list(run_tower(self._rt, x, dense=return_dense, pre=return_feature))
        if return_dense:   # dense_feat = last-block patch tokens, bf16 [B, 49, width]
            out[1] = out[1].view(x.shape[0], self._rt.seq_len - 1, -1)
        return tuple(out)


def visual_transformer_B32(**kwargs):
    default_kwargs = {'layers': 12, 'heads': 12, 'input_resolution': 224, 'patch_size': 32, 'width': 768,
                      'checkpoint': False}
    default_kwargs.update(**kwargs)
    return VisualTransformer(**default_kwargs)


def visual_transformer_B16(**kwargs):
    raise NotImplementedError("declip_b200: ViT-B/16 has 197 tokens; the fused attention core supports L <= 80")
