"""DeFILIP wrapper — mirror of prototype/model/defilip.py: DeCLIP (two image views, EDA-augmented caption, SimSiam heads,
nearest-neighbour text bank, MLM) plus FILIP's token-wise late-interaction logits on the patch / word tokens
(`return_filip`, `dense_aug`).  Same constructor keywords, parameter names (`image_mapping`, `text_mapping`,
`logit_scale_dense` on top of DECLIP's) and output dict keys ('filip', 'filip_aug').  Built entirely from kernels
the CLIP / DeCLIP / FILIP paths already use."""
import numpy as np
import torch
from torch import nn

from .. import functions as F_
from ..runtime import concurrent_towers
from .declip import DECLIP
from .filip import weighted_dense_logits
from .text_transformer import text_transformers
from .visual_transformer import visual_transformer_B32

__all__ = ['defilip_vitb32', 'DEFILIP']


class DEFILIP(DECLIP):
    def __init__(self, image_encode, text_encode, use_allgather, nn_size=2 ** 16, nn_topk=1, return_dense=False,
                 return_simsiam_text=False, return_simsiam_nn_text=False, return_caption=False, return_nn_bank=False,
                 text_mask_type=None, EDA=True, feature_dim=1024, forward_type='split', return_filip=False,
                 dense_embed_dim=256, dense_mapping_image=768, dense_mapping_language=512, dense_aug=False):
        super().__init__(image_encode, text_encode, use_allgather, nn_size=nn_size, nn_topk=nn_topk,
                         return_dense=return_dense, return_simsiam_text=return_simsiam_text,
                         return_simsiam_nn_text=return_simsiam_nn_text, return_caption=return_caption,
                         return_nn_bank=return_nn_bank, text_mask_type=text_mask_type, EDA=EDA, feature_dim=feature_dim,
                         forward_type=forward_type)
        self.return_filip = return_filip
        self.dense_aug = dense_aug
        if self.return_filip:                                                       # defilip.py:178-183
            if text_mask_type is None:
                raise NotImplementedError("declip_b200: return_filip needs text_mask_type='MLM' — without it the reference "
                                          "reads an undefined `word_features` (defilip.py:296-302,335)")
            self.select_topk = True
            self.logit_scale_dense = nn.Parameter(torch.ones([]))
            nn.init.constant_(self.logit_scale_dense, np.log(1 / 0.07))
            self.image_mapping = nn.Linear(dense_mapping_image, dense_embed_dim)
            self.text_mapping = nn.Linear(dense_mapping_language, dense_embed_dim)
        elif dense_aug:
            raise NotImplementedError("declip_b200: dense_aug without return_filip reads undefined logits in the reference")

    def get_weighted_dense_logits(self, dense_feat_1, dense_feat_2, top_k=16):
        return weighted_dense_logits(dense_feat_1, dense_feat_2, self.logit_scale_dense, top_k)   # defilip.py:223-267

    def _map(self, tokens, linear):
        B = tokens.shape[0]
        d = F_.LinearBF16In.apply(tokens.reshape(-1, tokens.shape[-1]), linear.weight, linear.bias)
        return d.view(B, -1, d.shape[-1])

    def forward(self, input, return_dict=False):
        if not self.return_filip:
            return super().forward(input, return_dict=return_dict)
        if not return_dict:
            raise NotImplementedError('Must Return A Dict')                                         # defilip.py:431
        if not (self.training and self.use_allgather):
            raise NotImplementedError('2-View: Not Implemented')                                    # defilip.py:393-394
        images = input['images']
        images_1, images_2 = images[:, :3], images[:, 3:]                                           # defilip.py:275
        ids, ids_aug, mlm = self._text_inputs(input)
        with concurrent_towers():
            text_features, word_features, text_labels = self.encode_text(mlm if mlm is not None else ids,
                                                                         mask_type=self.text_mask_type)   # :294
            text_features_aug, word_features_aug = self.encode_text(ids_aug, return_dense=True)      # :295
            image_features_1, image_d1 = self.encode_image(images_1, return_dense=True)              # :311-313
            image_features_2, image_d2 = self.encode_image(images_2, return_dense=True)
        z1 = self.projector(image_features_1)                                                        # :316-319
        z2 = self.projector(image_features_2)
        p1 = self.predictor(z1)
        p2 = self.predictor(z2)
        # token-wise late interaction                                                                 :331-342
        di1, di2 = self._map(image_d1, self.image_mapping), self._map(image_d2, self.image_mapping)
        dw1, dw2 = self._map(word_features, self.text_mapping), self._map(word_features_aug, self.text_mapping)
        filip = self.get_weighted_dense_logits(di1, dw1)
        filip_aug = None
        if self.dense_aug:
            filip_aug = (*self.get_weighted_dense_logits(di2, dw1), *self.get_weighted_dense_logits(di1, dw2),
                         *self.get_weighted_dense_logits(di2, dw2))
        image_features_1 = F_.L2Normalize.apply(image_features_1, 0.0)                                # :345-348
        image_features_2 = F_.L2Normalize.apply(image_features_2, 0.0)
        text_features = F_.L2Normalize.apply(text_features, 1e-10)
        text_features_aug = F_.L2Normalize.apply(text_features_aug, 1e-10)
        feats = [image_features_1, image_features_2, text_features, text_features_aug]
        I1, I2, T, TA = 0, 1, 2, 3
        pairs = [(I1, T), (I2, T), (I1, TA), (I2, TA), (T, I1), (T, I2), (TA, I1), (TA, I2)]        # :362-370
        if self.return_nn_bank:                                                                      # :372-392
            t_nn = self.nn_replacer_text(text_features.detach(), update=False)[0]
            t_nn = F_.L2Normalize.apply(t_nn, 1e-10)
            t_nn_aug = self.nn_replacer_text(text_features_aug.detach(), update=True)[0]
            t_nn_aug = F_.L2Normalize.apply(t_nn_aug, 1e-10)
            self.nn_replacer_text(text_features.detach(), update=True)
            feats += [t_nn, t_nn_aug]
            TN, TNA = 4, 5
            pairs += [(I1, TN), (I2, TN), (I1, TNA), (I2, TNA)]
        strips = F_.StripLogits.apply(self.logit_scale, 1.0, True, True, tuple(pairs), *feats)
        li1, li2, li1a, li2a, lt1, lt2, lt1a, lt2a = strips[:8]
        ret = {'logits': (li1, li2, lt1, lt2), 'logits_aug': (li1a, li2a, lt1a, lt2a),
               'simsiam_features': (p1, p2, z1, z2), 'features': (text_features, image_features_1, image_features_2),
               'filip': filip}                                                                        # :398-405
        if filip_aug is not None:
            ret['filip_aug'] = filip_aug
        if self.return_nn_bank:
            ret['nn_text_logits'] = tuple(strips[8:12])
        labels = text_labels.reshape(-1)                                                              # :420-428
        rows = torch.nonzero(labels != -100, as_tuple=False).reshape(-1)
        dev = word_features.device
        sel = labels[rows].to(dev)
        rows = rows.to(device=dev, dtype=torch.int32)
        ret['text_self_supervised'] = F_.MaskedLMHead.apply(word_features.reshape(-1, word_features.shape[-1]), rows, sel,
                                                            self.text_label_predictor.weight,
                                                            self.text_label_predictor.bias)
        return ret


def defilip_vitb32(**kwargs):
    """defilip.py:431-438."""
    image_encode = visual_transformer_B32(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return DEFILIP(image_encode, text_encode, **kwargs['clip'])
