"""SLIP wrapper — mirror of prototype/model/slip.py (CLIP on a base view + SimCLR on two augmented views through
`predictor_sim`, a 3-layer MLP on the image tower's PRE-projection feature): same constructor keywords, parameter names
and output dict {'logits', 'sim_features', 'features'} (slip.py:196-284).  `slip_res50` is not offered: the reference's
ModifiedResNet.forward has no `return_feature` argument (image_encoder/modified_resnet.py:192), so that factory cannot
run there either."""
import numpy as np
import torch
from torch import nn

from .. import functions as F_
from ..runtime import concurrent_towers
from .clip import CLIP
from .declip import _bn
from .text_transformer import text_transformers
from .visual_transformer import visual_transformer_B32

__all__ = ['slip_vitb32', 'SLIP']


class projection_MLP(nn.Module):
    """slip.py:49-108: Linear-BN-ReLU, Linear-BN-ReLU, Linear [-BN when out_bn]; BN is the shim's SyncBatchNorm2d, i.e.
    nn.BatchNorm1d with per-rank statistics (linklink/nn.py:4-5)."""

    def __init__(self, in_dim, hidden_dim=1024, out_dim=1024, num_layers=3, out_bn=True):
        super().__init__()
        self.num_layers = num_layers
        self.in_dim, self.hidden_dim, self.out_dim = in_dim, hidden_dim, out_dim
        self.linear1 = nn.Linear(in_dim, hidden_dim)
        self.bn1 = nn.BatchNorm1d(hidden_dim)
        self.relu1 = nn.ReLU(inplace=True)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.bn2 = nn.BatchNorm1d(hidden_dim)
        if self.num_layers == 3:
            self.relu2 = nn.ReLU(inplace=True)
            self.linear3 = nn.Linear(hidden_dim, out_dim)
            self.bn3 = nn.BatchNorm1d(hidden_dim)
            self.out_bn = out_bn

    def set_layers(self, num_layers):
        self.num_layers = num_layers

    def forward(self, x):
        x = F_.LinearF32.apply(x, self.linear1.weight, self.linear1.bias)
        x = _bn(self.bn1, x, True)
        x = F_.LinearF32.apply(x, self.linear2.weight, self.linear2.bias)
        x = _bn(self.bn2, x, self.num_layers == 3)
        if self.num_layers == 3:
            x = F_.LinearF32.apply(x, self.linear3.weight, self.linear3.bias)
            if self.out_bn:
                x = _bn(self.bn3, x, False)
        return x


class SLIP(CLIP):
    def __init__(self, image_encode, text_encode, use_allgather, EDA=True, feature_dim=1024, sim_dim=256,
                 forward_type='split', return_sim=False):
        nn.Module.__init__(self)
        # slip.py:111-120: this file's CLIP base registers the text tower as `text_encoder` (state_dict keys follow)
        self.use_allgather = use_allgather
        self.visual = image_encode
        self.text_encoder = text_encode
        self.logit_scale = nn.Parameter(torch.ones([1]))
        nn.init.constant_(self.logit_scale, np.log(1 / 0.07))
        self.return_sim = return_sim
        if self.return_sim:
            self.predictor_sim = projection_MLP(feature_dim, hidden_dim=4096, out_dim=sim_dim, out_bn=False)
        self.forward_type = forward_type

    def text_parameters(self):
        return [self.logit_scale, self.text_encoder.positional_embedding]                   # slip.py:123-132

    def text_modules(self):
        return [self.text_encoder.transformer, self.text_encoder.text_projection, self.text_encoder.token_embedding,
                self.text_encoder.ln_final]

    def visual_modules(self):
        return [self.visual, self.predictor_sim]

    def encode_text(self, text, text_mask_type=None, return_sim=False):                      # slip.py:239-245
        assert not return_sim
        return self.text_encoder(text, mask_type=text_mask_type) if text_mask_type else self.text_encoder(text)

    def encode_image(self, image, return_dense=False, return_sim=False):
        if not (return_dense or return_sim):
            return self.visual(image)
        out = self.visual(image, return_dense=return_dense, return_feature=return_sim)
        if return_sim:
            out = (*out[:-1], self.predictor_sim(out[-1]))                                   # slip.py:234-237
        return out

    def forward(self, input, return_dict=False):
        if not return_dict:
            raise NotImplementedError('Must Return A Dict')                                  # slip.py:284
        if not (self.training and self.use_allgather):
            raise NotImplementedError('2-View: Not Implemented')                             # slip.py:273-274
        if not self.return_sim:
            raise NotImplementedError("declip_b200: SLIP.forward reads predictor_sim, which only exists with return_sim")
        images = input['images']
        images_base, images_1, images_2 = images[:, :3], images[:, 3:6], images[:, 6:9]       # slip.py:241
        texts = self._texts(input)
        with concurrent_towers():
            text_features = self.encode_text(texts)
            image_features = self.encode_image(images_base)
            _, image_sim_1 = self.encode_image(images_1, return_sim=True)
            _, image_sim_2 = self.encode_image(images_2, return_sim=True)
        # exp(logit_scale) is NOT clamped here (slip.py:258), as in FILIP
        logits_per_image, logits_per_text = F_.ClipLogits.apply(image_features, text_features, self.logit_scale, True, False)
        image_features = F_.L2Normalize.apply(image_features, 0.0)
        text_features = F_.L2Normalize.apply(text_features, 1e-10)
        gathered_1 = F_.AllGatherRows.apply(image_sim_1)                                     # slip.py:264-265
        gathered_2 = F_.AllGatherRows.apply(image_sim_2)
        return {'logits': (logits_per_image, logits_per_text),
                'sim_features': (image_sim_1, gathered_1, image_sim_2, gathered_2),
                'features': (text_features, image_features)}


def slip_vitb32(**kwargs):
    """slip.py:299-306."""
    image_encode = visual_transformer_B32(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return SLIP(image_encode, text_encode, **kwargs['clip'])
