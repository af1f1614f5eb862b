"""Layer-by-layer comparison of the NHWC bf16 ModifiedResNet executor with the fp32 restatement (oracle/resnet_ref.py run
on the GPU in fp32) — finds the first block whose output departs.  python tools/resnet_debug.py [batch] [layers...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from declip_b200 import functions_conv as C_  # noqa: E402
from declip_b200.model.modified_resnet import _bn, modified_resnet_R50  # noqa: E402
from oracle import resnet_ref, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
layers = tuple(int(a) for a in sys.argv[2:6]) if len(sys.argv) >= 6 else (3, 4, 6, 3)
dev = torch.device("cuda:0")
sd = synth.resnet_state_dict(seed=16, layers=layers, embed_dim=1024, prefix="")
m = modified_resnet_R50(embed_dim=1024, use_sync_bn=False, bn_group_size=1, layers=layers)
m.load_state_dict(sd, strict=True)
m = m.to(dev).train()
sdd = {k: v.to(dev) for k, v in sd.items()}
stats = {k: v.clone() for k, v in sdd.items() if "running_" in k}
x = synth.synth_images(B, seed=16).to(dev)


def nchw(y, h, w):
    return y.float().view(B, h, w, -1).permute(0, 3, 1, 2)


def cmp(name, mine, ref):
    a, b = mine.reshape(-1).float(), ref.reshape(-1).float()
    cos = (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()
    rel = ((a - b).norm() / (b.norm() + 1e-20)).item()
    print("%-22s cos %.6f rel %.4f  |ref| mean %.3f std %.3f" % (name, cos, rel, b.mean().item(), b.std().item()), flush=True)


with torch.no_grad():
    h = w = 112
    y = _bn(m.bn1, C_.StemConv.apply(x, m.conv1.weight))
    r = F.relu(resnet_ref._bn(F.conv2d(x, sdd["conv1.weight"], stride=2, padding=1), sdd, "bn1", stats))
    cmp("stem conv1", nchw(y, h, w), r)
    y = _bn(m.bn2, C_.Conv3x3.apply(y, m.conv2.weight, B, h, w))
    r = F.relu(resnet_ref._bn(F.conv2d(r, sdd["conv2.weight"], padding=1), sdd, "bn2", stats))
    cmp("stem conv2", nchw(y, h, w), r)
    y = _bn(m.bn3, C_.Conv3x3.apply(y, m.conv3.weight, B, h, w))
    r = F.relu(resnet_ref._bn(F.conv2d(r, sdd["conv3.weight"], padding=1), sdd, "bn3", stats))
    cmp("stem conv3", nchw(y, h, w), r)
    y = C_.AvgPool2.apply(y, B, h, w)
    r = F.avg_pool2d(r, 2)
    h = w = 56
    for li, layer in enumerate((m.layer1, m.layer2, m.layer3, m.layer4), 1):
        for bi, block in enumerate(layer):
            # feed BOTH paths the reference activation so errors do not compound: isolates the faulty block
            yin = r.permute(0, 2, 3, 1).reshape(-1, r.shape[1]).contiguous().bfloat16()
            y1, h2, w2 = block.run(yin, B, h, w)
            y, _, _ = block.run(y, B, h, w)
            r = resnet_ref.bottleneck(r, sdd, stats, "layer%d.%d." % (li, bi), 2 if (bi == 0 and li > 1) else 1)
            h, w = h2, w2
            cmp("layer%d.%d (isolated)" % (li, bi), nchw(y1, h, w), r)
            cmp("layer%d.%d (chained)" % (li, bi), nchw(y, h, w), r)
    out = m.attnpool.run(y, B, h * w)
    ro = resnet_ref.attention_pool(r, sdd, "attnpool.", 32)
    cmp("attnpool (chained)", out, ro)
