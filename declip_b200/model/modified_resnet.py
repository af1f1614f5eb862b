"""ModifiedResNet (CLIP-style RN50 / RN101) — mirror of prototype/model/image_encoder/modified_resnet.py: same
constructor, module tree and parameter names (Conv2d / BatchNorm2d / Linear containers so state_dicts and the
isinstance-based weight-decay grouping carry over), forward signature `forward(x, return_dense=False)`.

Execution: NHWC bf16 activations; 1x1 convs = tcgen05 GEMM, 3x3 convs = implicit GEMM over TMA boxes (csrc/conv_igemm.cu:
forward, input and weight gradient; no im2col matrix), the block's skip gradient added in conv1's dgrad epilogue, BatchNorm2d (+ReLU, +residual)
fused apply passes, AvgPool2d(2), AttentionPool2d through the same fused attention core as the transformers
(32 heads x 64, L = 50).  Only `use_sync_bn=False` is supported — the only mode that runs with the reference's own
linklink shim (`link.new_group` is missing, `SyncBatchNorm2d` aliases BatchNorm1d)."""
from collections import OrderedDict

import torch
from torch import nn

from .. import functions as F_
from .. import functions_conv as C_


def _bn(bn, x, res=None, relu=True):
    if not bn.training:     # model.eval(): running statistics (zero-shot evaluation, clip_solver.py:675-737)
        return C_.BatchNorm2dNHWC.apply(x, res, bn.weight, bn.bias, bn.running_mean, bn.running_var, relu, bn.eps,
                                        bn.momentum, False)
    if bn.track_running_stats:
        bn.num_batches_tracked += 1
    return C_.BatchNorm2dNHWC.apply(x, res, bn.weight, bn.bias, bn.running_mean, bn.running_var, relu, bn.eps, bn.momentum)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        self.stride = stride
        if stride > 1 or inplanes != planes * Bottleneck.expansion:
            self.downsample = nn.Sequential(OrderedDict([
                ("-1", nn.AvgPool2d(stride)),
                ("0", nn.Conv2d(inplanes, planes * self.expansion, 1, stride=1, bias=False)),
                ("1", nn.BatchNorm2d(planes * self.expansion))]))

    def run(self, x, B, H, W):
        """x: NHWC bf16 [B*H*W, C] -> (out, H', W')   (modified_resnet.py:40-56)."""
        out, x = C_.Conv1x1Skip.apply(x, self.conv1.weight)        # x: the skip branch, its gradient is added in conv1's dgrad
        out = _bn(self.bn1, out)
        out = _bn(self.bn2, C_.Conv3x3.apply(out, self.conv2.weight, B, H, W))
        h2, w2 = H, W
        if self.stride > 1:
            out = C_.AvgPool2.apply(out, B, H, W)
            h2, w2 = H // 2, W // 2
        out = C_.Conv1x1.apply(out, self.conv3.weight)
        identity = x
        if self.downsample is not None:
            idn = C_.AvgPool2.apply(x, B, H, W) if self.stride > 1 else x
            identity = _bn(self.downsample[2], C_.Conv1x1.apply(idn, self.downsample[1].weight), relu=False)
        out = _bn(self.bn3, out, res=identity, relu=True)          # relu(bn3(conv3) + identity)
        return out, h2, w2

    def forward(self, x):
        raise RuntimeError("declip_b200: Bottleneck runs inside ModifiedResNet.forward (NHWC executor)")


class AttentionPool2d(nn.Module):
    def __init__(self, spacial_dim, embed_dim, num_heads, output_dim=None):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.num_heads = num_heads
        if embed_dim != num_heads * 64:
            raise NotImplementedError("declip_b200: the fused attention core needs head_dim 64")

    def run(self, x, B, P):
        """x NHWC bf16 [B*P, C] -> fp32 [B, output_dim]   (modified_resnet.py:71-96: only token 0 is returned)."""
        C = x.shape[1]
        tok = C_.AttnPoolAssemble.apply(x, self.positional_embedding, B, P)                   # [B*(P+1), C]
        w = torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], dim=0)   # separate proj weights :78-85
        b = torch.cat([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias])
        qkv = C_.LinearBF16.apply(tok, w, b)
        att = C_.Attention.apply(qkv, B, P + 1, self.num_heads, False)
        first = att.view(B, P + 1, C)[:, 0].contiguous()                                      # x[0]
        return F_.LinearBF16In.apply(first, self.c_proj.weight, self.c_proj.bias)

    def forward(self, x):
        raise RuntimeError("declip_b200: AttentionPool2d runs inside ModifiedResNet.forward")


class ModifiedResNet(nn.Module):
    def __init__(self, layers, embed_dim, heads, input_resolution=224, width=64, bn_group_size=1, bn_var_mode=None,
                 bn_sync_stats=False, use_sync_bn=True):
        super().__init__()
        if use_sync_bn:
            raise NotImplementedError("declip_b200: use_sync_bn=True does not run in the reference either (link.new_group is "
                                      "missing from its shim and its SyncBatchNorm2d aliases BatchNorm1d, SURVEY.md 2.2); set "
                                      "use_sync_bn: False as experiments/clip_experiments/yfcc15m/yfcc15m_r50_clip does "
                                      "(yfcc15m_r50_declip and yfcc15m_r50_filip ask for it and need that one-line change)")
        self.output_dim = embed_dim
        self.input_resolution = input_resolution
        self.conv1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width // 2)
        self.conv2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width // 2)
        self.conv3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(width)
        self.avgpool = nn.AvgPool2d(2)
        self.relu = nn.ReLU(inplace=True)
        self._inplanes = width
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        feat_dim = width * 32
        self.attnpool = AttentionPool2d(input_resolution // 32, feat_dim, heads, embed_dim)
        self.adaptivepool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(2048, embed_dim)
        std = self.attnpool.c_proj.in_features ** -0.5                                        # modified_resnet.py:171-180
        for m in (self.attnpool.q_proj, self.attnpool.k_proj, self.attnpool.v_proj, self.attnpool.c_proj):
            nn.init.normal_(m.weight, std=std)
        for block in [self.layer1, self.layer2, self.layer3, self.layer4]:
            for name, param in block.named_parameters():
                if name.endswith("bn3.weight"):
                    nn.init.zeros_(param)

    def _make_layer(self, planes, blocks, stride=1):
        layers = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        for _ in range(1, blocks):
            layers.append(Bottleneck(self._inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x, return_dense=False):
        B, _, H, W = x.shape
        if H != self.input_resolution or W != self.input_resolution:
            raise ValueError("expected %dx%d images" % (self.input_resolution, self.input_resolution))
        if x.dtype != torch.float32 or x.stride(3) != 1 or x.stride(2) != W or x.stride(1) != H * W:
            x = x.float().contiguous()
        # stem                                                                                  modified_resnet.py:193-198
        h, w = H // 2, W // 2
        y = _bn(self.bn1, C_.StemConv.apply(x, self.conv1.weight))
        y = _bn(self.bn2, C_.Conv3x3.apply(y, self.conv2.weight, B, h, w))
        y = _bn(self.bn3, C_.Conv3x3.apply(y, self.conv3.weight, B, h, w))
        y = C_.AvgPool2.apply(y, B, h, w)
        h, w = h // 2, w // 2
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for block in layer:
                y, h, w = block.run(y, B, h, w)
        if h != 7:
            raise NotImplementedError("declip_b200: only the 7x7 AttentionPool2d head (224x224 input) is built")
        out = self.attnpool.run(y, B, h * w)
        if return_dense:
            return out, y.view(B, h * w, -1)                                                  # [B, 49, 2048] (bf16)
        return out


def modified_resnet_R50(**kwargs):
    default_kwargs = {'layers': (3, 4, 6, 3), 'heads': 32, 'input_resolution': 224, 'width': 64}
    default_kwargs.update(**kwargs)
    return ModifiedResNet(**default_kwargs)


def modified_resnet_R101(**kwargs):
    default_kwargs = {'layers': (3, 4, 23, 3), 'heads': 32, 'input_resolution': 224, 'width': 64}
    default_kwargs.update(**kwargs)
    return ModifiedResNet(**default_kwargs)
