"""FILIP wrapper — mirror of prototype/model/filip.py (token-wise late interaction, filip.py:28-142): same
constructor, parameter names (`image_mapping`, `text_mapping`, `logit_scale_dense`, `text_label_predictor`) and
output dict {'logits', 'dense_logits'}."""
import numpy as np
import torch
from torch import nn

from .. import functions as F_
from ..runtime import concurrent_towers
from .clip import CLIP
from .modified_resnet import modified_resnet_R50
from .text_transformer import text_transformers
from .visual_transformer import visual_transformer_B32

__all__ = ['filip_vitb32', 'filip_res50', 'FILIP']


def weighted_dense_logits(dense_feat_1, dense_feat_2, logit_scale_dense, top_k=16):
    """Token-wise late interaction with top-k token selection (filip.py:71-106, defilip.py:223-267): unit-norm tokens,
    the top_k tokens of each side by summed cross similarity are all-gathered, every local token takes its best match
    among the selected tokens of every sample, averaged over the local tokens.  dense_feat_* fp32 [B, n, dim]."""
    B, n1, dim = dense_feat_1.shape
    n2 = dense_feat_2.shape[1]
    d1 = F_.L2Normalize.apply(dense_feat_1.reshape(B * n1, dim), 0.0)
    d2 = F_.L2Normalize.apply(dense_feat_2.reshape(B * n2, dim), 0.0)
    s1, s2 = F_.token_scores(d1.detach(), d2.detach(), B, n1, n2)
    id1 = torch.topk(s1, k=top_k, dim=1).indices + torch.arange(B, device=s1.device)[:, None] * n1
    id2 = torch.topk(s2, k=top_k, dim=1).indices + torch.arange(B, device=s1.device)[:, None] * n2
    sel1 = F_.GatherRowsF32.apply(d1, id1.reshape(-1).to(torch.int32))                     # [B*k, dim]
    sel2 = F_.GatherRowsF32.apply(d2, id2.reshape(-1).to(torch.int32))
    sel1 = F_.AllGatherRows.apply(sel1)
    sel2 = F_.AllGatherRows.apply(sel2)
    l1 = F_.FilipLate.apply(d1, sel2, logit_scale_dense, n1, top_k)
    l2 = F_.FilipLate.apply(d2, sel1, logit_scale_dense, n2, top_k)
    return l1, l2


class FILIP(CLIP):
    def __init__(self, image_encode, text_encode, use_allgather, nn_size=2 ** 16, nn_topk=1, return_dense=False,
                 return_caption=False, return_nn_bank=False, text_mask_type=None, EDA=True, feature_dim=1024,
                 embed_dim=768, forward_type='split', dense_mapping_image=2048, dense_mapping_language=512,
                 dense_embed_dim=256, mask_rate=0.75, patch_number=14, text_mae_feature=False, return_simsiam=False,
                 two_view=False, sparse=False, select_topk=False):
        super().__init__(image_encode, text_encode, use_allgather)
        if return_caption:
            raise NotImplementedError("declip_b200: the captioning head is not on the hot path")
        self.return_dense = return_dense
        self.return_caption = return_caption
        self.text_mask_type = text_mask_type
        self.select_topk = select_topk
        if self.return_dense:
            self.image_mapping = nn.Linear(dense_mapping_image, dense_embed_dim)
            self.text_mapping = nn.Linear(dense_mapping_language, dense_embed_dim)
        self.logit_scale_dense = nn.Parameter(torch.ones([]))
        nn.init.constant_(self.logit_scale_dense, np.log(1 / 0.07))
        if text_mask_type is not None:
            enc_dim = self.encode_text.text_projection.weight.shape[-1]
            self.text_label_predictor = nn.Linear(enc_dim, self.encode_text.vocab_size)

    def encode_image(self, image, return_all=False):
        return self.visual(image, return_dense=return_all)

    def get_weighted_dense_logits(self, dense_feat_1, dense_feat_2, top_k=16):
        """filip.py:71-106.  dense_feat_* fp32 [B, n, 256]."""
        if not self.select_topk:
            raise NotImplementedError("declip_b200: FILIP without select_topk is a latent bug in the reference "
                                      "(selected_feat undefined, filip.py:90-94)")
        return weighted_dense_logits(dense_feat_1, dense_feat_2, self.logit_scale_dense, top_k)

    def forward(self, input, return_dict=False):
        if not return_dict:
            raise NotImplementedError()
        if not (self.training and self.use_allgather):
            raise NotImplementedError("declip_b200: FILIP.forward defines its logits only for training with "
                                      "use_allgather (filip.py:123-129)")
        images_1 = input['images'][:, :3]                                                      # filip.py:112
        if input.get('token_ids') is not None:
            text_in = input.get('mlm') if input.get('mlm') is not None else input['token_ids']
        else:
            text_in = self.sample_captions(input['captions'])
        with concurrent_towers():
            if self.text_mask_type is not None:
                text_features, word_features, text_labels = self.encode_text(text_in, mask_type=self.text_mask_type)
            else:
                text_features, word_features = self.encode_text(text_in, return_dense=True)
            image_features_1, image_features_d = self.encode_image(images_1, return_all=True)  # :119
        li, lt = F_.ClipLogits.apply(image_features_1, text_features, self.logit_scale, True, False)   # :121-129 (no clamp)
        ret = {'logits': (li, lt)}
        if self.return_dense:
            B = images_1.shape[0]
            d1 = F_.LinearBF16In.apply(image_features_d.reshape(-1, image_features_d.shape[-1]),
                                       self.image_mapping.weight, self.image_mapping.bias)     # :133
            d2 = F_.LinearBF16In.apply(word_features.reshape(-1, word_features.shape[-1]), self.text_mapping.weight,
                                       self.text_mapping.bias)                                 # :134
            ret['dense_logits'] = self.get_weighted_dense_logits(d1.view(B, -1, d1.shape[-1]), d2.view(B, -1, d2.shape[-1]))
        return ret


def filip_vitb32(**kwargs):
    """filip.py:156-163."""
    image_encode = visual_transformer_B32(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return FILIP(image_encode, text_encode, **kwargs['clip'], dense_mapping_image=768)


def filip_res50(**kwargs):
    """filip.py:146-153."""
    image_encode = modified_resnet_R50(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return FILIP(image_encode, text_encode, **kwargs['clip'], dense_mapping_image=2048)
