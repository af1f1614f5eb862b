"""DECLIP wrapper — mirror of prototype/model/declip.py (DeCLIP: multi-view CLIP + SimSiam + nearest-neighbour
text supervision + masked-language-model loss).  Same constructor, sub-module / parameter names and output dict
(declip.py:132-336); every arithmetic op runs through the C ABI.
"""
import torch
from torch import nn

from .. import functions as F_
from ..runtime import concurrent_towers
from .clip import CLIP
from .modified_resnet import modified_resnet_R50
from .nn_memory_bank import NNMemoryBankModule
from .text_transformer import text_transformers
from .visual_transformer import visual_transformer_B32

__all__ = ['declip_vitb32', 'declip_res50', 'DECLIP']


def _bn(bn, x, relu):
    if bn.training and bn.track_running_stats:
        bn.num_batches_tracked += 1
    return F_.BatchNorm1dF.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, relu, bn.eps,
                                 bn.momentum)


class projection_MLP(nn.Module):
    """declip.py:33-90: Linear-BN-ReLU, Linear-BN-ReLU, Linear-BN."""

    def __init__(self, in_dim, hidden_dim=1024, out_dim=1024, num_layers=3):
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

    def set_layers(self, num_layers):
        self.num_layers = num_layers

    def forward(self, x):
        x = F_.LinearF32.apply(x, self.linear1.weight, self.linear1.bias)
        x = _bn(self.bn1, x, True)
        x = F_.LinearF32.apply(x, self.linear2.weight, self.linear2.bias)
        x = _bn(self.bn2, x, self.num_layers == 3)
        if self.num_layers == 3:
            x = F_.LinearF32.apply(x, self.linear3.weight, self.linear3.bias)
            x = _bn(self.bn3, x, False)
        return x


class prediction_MLP(nn.Module):
    """declip.py:92-130: Linear-BN-ReLU, Linear."""

    def __init__(self, in_dim, hidden_dim=512, out_dim=1024):
        super().__init__()
        self.in_dim, self.hidden_dim, self.out_dim = in_dim, hidden_dim, out_dim
        self.linear1 = nn.Linear(in_dim, hidden_dim)
        self.bn1 = nn.BatchNorm1d(hidden_dim)
        self.relu1 = nn.ReLU(inplace=True)
        self.layer2 = nn.Linear(hidden_dim, out_dim)

    def forward(self, x):
        x = F_.LinearF32.apply(x, self.linear1.weight, self.linear1.bias)
        x = _bn(self.bn1, x, True)
        return F_.LinearF32.apply(x, self.layer2.weight, self.layer2.bias)


class DECLIP(CLIP):
    def __init__(self, image_encode, text_encode, use_allgather, nn_size=2 ** 16, nn_topk=1, return_dense=False,
                 return_simsiam_text=False, return_simsiam_nn_text=False, return_caption=False, return_nn_bank=False,
                 text_mask_type=None, EDA=True, feature_dim=1024, forward_type='split', fused_head=False):
        super().__init__(image_encode, text_encode, use_allgather, fused_head=fused_head)
        self.projector = projection_MLP(feature_dim)
        self.predictor = prediction_MLP(1024)
        if return_dense:
            raise NotImplementedError('These are bugs in the model, Please Check The Codes!')      # declip.py:158
        if return_caption:
            raise NotImplementedError('Not Available')
        if return_simsiam_text or return_simsiam_nn_text:
            raise NotImplementedError("declip_b200: text-SimSiam heads are unused by the reference configs")
        self.return_dense = return_dense
        self.return_nn_bank = return_nn_bank
        self.return_caption = return_caption
        self.return_simsiam_text = return_simsiam_text
        self.return_simsiam_nn_text = return_simsiam_nn_text
        self.text_mask_type = text_mask_type
        self.EDA = EDA
        if forward_type not in ('split', 'image_concat'):
            raise NotImplementedError("declip_b200: forward_type %r (declip.py:225-232 knows 'split' / 'image_concat')" % (forward_type,))
        self.forward_type = forward_type
        if self.EDA:
            from ..eda import EDA as _EDA
            self.emd = _EDA()                                                                      # declip.py:154-155
        if text_mask_type is not None:
            enc_dim = self.encode_text.text_projection.weight.shape[-1]
            self.text_label_predictor = nn.Linear(enc_dim, self.encode_text.vocab_size)
        if self.return_nn_bank:
            self.nn_replacer_img = NNMemoryBankModule(size=nn_size, topk=nn_topk)
            self.nn_replacer_text = NNMemoryBankModule(size=nn_size, topk=nn_topk)

    def visual_modules(self):
        return [self.visual, self.predictor, self.projector]

    def text_modules(self):
        ret = super().text_modules()
        if self.text_mask_type is not None:
            ret.append(self.text_label_predictor)
        return ret

    def encode_image(self, image, return_dense=False):
        return self.visual(image, return_dense=return_dense) if return_dense else self.visual(image)

    def _text_inputs(self, input):
        """(ids, ids_aug, pre-masked (ids, labels) or None).  With strings the reference runs host EDA augmentation
        (declip.py:203-212, needs the `textaugment` package); pre-tokenised callers pass `token_ids_aug`."""
        if input.get('token_ids') is not None:
            ids = input['token_ids']
            return ids, input.get('token_ids_aug', ids), input.get('mlm')
        texts = self.sample_captions(input['captions'])
        ids = self.encode_text.tokenize(texts)
        if not self.EDA:
            if self.text_mask_type is not None:
                raise NotImplementedError('No EDA')                                                 # declip.py:212
            return ids, ids, None
        texts_aug = self.emd.augment_batch(texts)                                                   # declip.py:203-211
        return ids, self.encode_text.tokenize(texts_aug), None

    def _encode_two_views(self, images):
        """image_features of the two views stacked on the channel axis (declip.py:199,225-232).  The ViT tower has no
        cross-sample coupling, so both views always go through it as ONE 2B-sample pass — for a contiguous
        [B,6,H,W] batch that is a zero-copy [2B,3,H,W] view with the views interleaved — whatever `forward_type` says.
        The ResNet tower's BatchNorm statistics do depend on the grouping: 'image_concat' normalises over 2B samples,
        'split' over B twice (two passes, running statistics updated twice)."""
        B = images.shape[0]
        coupled = any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in self.visual.modules())
        if coupled and self.forward_type != 'image_concat':
            return self.encode_image(images[:, :3]), self.encode_image(images[:, 3:])
        if images.is_contiguous():
            both = self.encode_image(images.view(2 * B, 3, images.shape[2], images.shape[3]))
            return both[0::2], both[1::2]
        both = self.encode_image(torch.cat([images[:, :3], images[:, 3:]], dim=0))
        return both[:B], both[B:]

    def forward(self, input, return_dict=False):
        if not return_dict:
            raise NotImplementedError('Must Return A Dict')                                         # declip.py:336
        if not (self.training and self.use_allgather):
            raise NotImplementedError('2-View: Not Implemented')                                    # declip.py:301-302
        images = input['images']
        ids, ids_aug, mlm = self._text_inputs(input)
        with concurrent_towers():        # text passes on a side stream, the two-view image pass on the current one
            if self.text_mask_type is not None:
                # The masked caption and the EDA-view caption go through the text tower as ONE 2b-sample pass (no cross-
                # sample coupling in a transformer: identical to the reference's two calls, declip.py:215-217, with half
                # the launches and each weight read once); the dense word features are used for the first half only.
                if mlm is not None:
                    mlm_ids, text_labels = mlm
                else:
                    if self.text_mask_type != 'MLM':
                        raise NotImplementedError(self.text_mask_type)
                    from .text_utils import mask_tokens_batch
                    mlm_ids, text_labels = mask_tokens_batch(ids)                                    # text_transformer.py:154-162
                dev = self.encode_text.positional_embedding.device
                both = torch.cat([mlm_ids.to(dev, non_blocking=True), ids_aug.to(dev, non_blocking=True)], dim=0)
                feats_both, words_both = self.encode_text(both, return_dense=True)
                B = mlm_ids.shape[0]
                text_features, text_features_aug = feats_both[:B], feats_both[B:]
                word_features = words_both[:B]
            else:
                text_features = self.encode_text(ids)
                text_features_aug = self.encode_text(ids_aug) if self.EDA else text_features.detach()
            image_features_1, image_features_2 = self._encode_two_views(images)                     # declip.py:199,225-232
        # SimSiam heads                                                                              declip.py:238-241
        z1 = self.projector(image_features_1)
        z2 = self.projector(image_features_2)
        p1 = self.predictor(z1)
        p2 = self.predictor(z2)
        fused = self.fused_head and image_features_1.shape[1] % 256 == 0 and image_features_1.shape[1] <= 1024
        raw_i1, raw_i2, raw_t, raw_ta = image_features_1, image_features_2, text_features, text_features_aug
        # normalised features                                                                        declip.py:245-248
        image_features_1 = F_.L2Normalize.apply(image_features_1, 0.0)
        image_features_2 = F_.L2Normalize.apply(image_features_2, 0.0)
        text_features = F_.L2Normalize.apply(text_features, 1e-10)
        text_features_aug = F_.L2Normalize.apply(text_features_aug, 1e-10)
        feats = [image_features_1, image_features_2, text_features, text_features_aug]
        I1, I2, T, TA = 0, 1, 2, 3
        pairs = [(I1, T), (I2, T), (I1, TA), (I2, TA), (T, I1), (T, I2), (TA, I1), (TA, I2)]       # declip.py:271-279
        if self.return_nn_bank:                                                                     # declip.py:281-300
            raw_nn = self.nn_replacer_text(text_features.detach(), update=False)[0]
            t_nn = F_.L2Normalize.apply(raw_nn, 1e-10)
            raw_nna = self.nn_replacer_text(text_features_aug.detach(), update=True)[0]
            t_nn_aug = F_.L2Normalize.apply(raw_nna, 1e-10)
            self.nn_replacer_text(text_features.detach(), update=True)
            feats += [t_nn, t_nn_aug]
            TN, TNA = 4, 5
            pairs += [(I1, TN), (I2, TN), (I1, TNA), (I2, TNA)]
        ret = {'simsiam_features': (p1, p2, z1, z2), 'features': (text_features, image_features_1, image_features_2)}
        if fused:
            # fused head (csrc/head.cu): 4 symmetric pairs + 2 nearest-neighbour pairs over ONE gathered bf16 buffer; every
            # (logits_a, logits_b) below is a pair of HANDLES for ClipInfoCELoss — no [b, N] strip exists
            raws = [raw_i1, raw_i2, raw_t, raw_ta]
            eps = [0.0, 0.0, 1e-10, 1e-10]
            hp = [(I1, T, T, I1, 1), (I2, T, T, I2, 1), (I1, TA, TA, I1, 1), (I2, TA, TA, I2, 1)]
            if self.return_nn_bank:
                raws += [raw_nn.detach(), raw_nna.detach()]
                eps += [1e-10, 1e-10]
                hp += [(I1, TN, I1, TNA, 0), (I2, TN, I2, TNA, 0)]
            h = F_.fused_pair_heads(self.logit_scale, True, True, eps, hp, raws)
            ret['logits'] = (h[0][0], h[1][0], h[0][1], h[1][1])
            ret['logits_aug'] = (h[2][0], h[3][0], h[2][1], h[3][1])
            if self.return_nn_bank:
                ret['nn_text_logits'] = (h[4][0], h[5][0], h[4][1], h[5][1])
        else:
            strips = F_.StripLogits.apply(self.logit_scale, 1.0, True, True, tuple(pairs), *feats)
            li1, li2, li1a, li2a, lt1, lt2, lt1a, lt2a = strips[:8]
            ret['logits'] = (li1, li2, lt1, lt2)
            ret['logits_aug'] = (li1a, li2a, lt1a, lt2a)
            if self.return_nn_bank:
                ret['nn_text_logits'] = tuple(strips[8:12])
        if self.text_mask_type is not None:                                                         # declip.py:326-334
            labels = text_labels.reshape(-1)
            rows = torch.nonzero(labels != -100, as_tuple=False).reshape(-1)
            dev = word_features.device
            sel = labels[rows].to(dev)
            rows = rows.to(device=dev, dtype=torch.int32)
            ret['text_self_supervised'] = F_.MaskedLMHead.apply(word_features.reshape(-1, word_features.shape[-1]), rows,
                                                                sel, self.text_label_predictor.weight,
                                                                self.text_label_predictor.bias)
        return ret


def declip_vitb32(**kwargs):
    """declip.py:348-355."""
    image_encode = visual_transformer_B32(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return DECLIP(image_encode, text_encode, **kwargs['clip'])


def declip_res50(**kwargs):
    """declip.py:339-346."""
    image_encode = modified_resnet_R50(**kwargs['image_encode'])
    text_encode = text_transformers(**kwargs['text_encode'])
    return DECLIP(image_encode, text_encode, **kwargs['clip'])
