"""CLIP wrapper — mirror of prototype/model/clip.py: same constructor, attributes the solver pokes
(`logit_scale`, `encode_image`, `encode_text`, `visual`, `text_parameters`, ...), same forward signature
and return values (clip.py:118-146), backed by the CUDA towers and the fused contrastive head.
"""
import numpy as np
import torch
from torch import nn

from .. import functions as F_
from ..runtime import concurrent_towers
from .modified_resnet import modified_resnet_R50
from .text_transformer import text_transformers
from .visual_transformer import visual_transformer_B32

__all__ = ['clip_vitb32', 'clip_res50', 'CLIP']


class CLIP(nn.Module):
    def __init__(self, image_encode, text_encode, use_allgather, fused_head=False):
        super().__init__()
        self.use_allgather = use_allgather
        # fused_head=True: forward returns two HANDLES instead of logit strips — the strips, ClipInfoCELoss, accuracy and
        # the whole backward run inside csrc/head.cu and no [b, N] tensor exists; only declip_b200's ClipInfoCELoss
        # understands them.  False (default): real [b, N] strips for callers that read logits (clip_solver.py:417-422).
        self.fused_head = fused_head
        self.visual = image_encode
        self.encode_text = text_encode
        self.logit_scale = nn.Parameter(torch.ones([1]))
        nn.init.constant_(self.logit_scale, np.log(1 / 0.07))                     # clip.py:57-59

    # --- attributes used by the solvers' AdamW_SGD param grouping (clip_solver.py:259-281)
    def text_parameters(self):
        return [self.logit_scale, self.encode_text.positional_embedding]

    def text_modules(self):
        return [self.encode_text.transformer, self.encode_text.text_projection, self.encode_text.token_embedding,
                self.encode_text.ln_final]

    def visual_parameters(self):
        return []

    def visual_modules(self):
        return [self.visual]

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image)                                                  # clip.py:107-108

    def sample_captions(self, texts):
        return [text[0] for text in texts]                                         # clip.py:110-111

    def _texts(self, input):
        """`captions` (List[List[str]], the reference API) or pre-tokenised `token_ids` LongTensor [B,77]."""
        if 'token_ids' in input and input['token_ids'] is not None:
            return input['token_ids']
        return self.sample_captions(input['captions'])

    def forward(self, input, all_gather=False):
        images = input['images']
        texts = self._texts(input)
        # The text tower runs first so that autograd (latest node first) runs the IMAGE tower's backward first: its
        # gradient all-reduce — the larger bucket — is then hidden behind the text tower's backward (dist.py).
        # The reference encodes the image first (clip.py:126-127); the two towers are independent, the results identical.
        with concurrent_towers():        # text tower on a side stream, image tower on the current one (runtime.py)
            text_features = self.encode_text(texts)
            image_features = self.encode_image(images)
        gather = (self.training and self.use_allgather) or all_gather             # clip.py:136
        if self.fused_head and image_features.shape[1] % 256 == 0 and image_features.shape[1] <= 1024:
            return F_.fused_clip_head(image_features, text_features, self.logit_scale, gather, True)
        logits_per_image, logits_per_text = F_.ClipLogits.apply(image_features, text_features, self.logit_scale,
                                                                gather, True)
        return logits_per_image, logits_per_text


def clip_vitb32(**kwargs):
    """clip.py:158-165."""
    image_encode = visual_transformer_B32(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return CLIP(image_encode, text_encode, **kwargs['clip'])


def clip_res50(**kwargs):
    """clip.py:149-156."""
    image_encode = modified_resnet_R50(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return CLIP(image_encode, text_encode, **kwargs['clip'])
