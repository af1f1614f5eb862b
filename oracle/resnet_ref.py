"""TEST INFRASTRUCTURE — CPU/fp32 restatement (oracle) of ModifiedResNet.forward
(prototype/model/image_encoder/modified_resnet.py:14-96,192-214), BatchNorm2d in training mode (per-rank statistics,
`use_sync_bn: False`).  Pinned against golden vectors generated from the reference's own clip_res50."""
import torch
import torch.nn.functional as F

from . import clip_ref


def _bn(x, sd, p, stats):
    return F.batch_norm(x, stats[p + ".running_mean"], stats[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        True, 0.1, 1e-5)


def bottleneck(x, sd, stats, p, stride):
    # modified_resnet.py:40-56
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1", stats))
    out = F.relu(_bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2", stats))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(F.conv2d(out, sd[p + "conv3.weight"]), sd, p + "bn3", stats)
    identity = x
    if (p + "downsample.0.weight") in sd:
        identity = F.avg_pool2d(x, stride) if stride > 1 else x
        identity = _bn(F.conv2d(identity, sd[p + "downsample.0.weight"]), sd, p + "downsample.1", stats)
    return F.relu(out + identity)


def attention_pool(x, sd, p, heads):
    # modified_resnet.py:71-96
    B, C, H, W = x.shape
    t = x.reshape(B, C, H * W).permute(0, 2, 1)                                   # [B, HW, C]
    t = torch.cat([t.mean(dim=1, keepdim=True), t], dim=1) + sd[p + "positional_embedding"][None]
    q = F.linear(t, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(t, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(t, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    L, hd = t.shape[1], C // heads
    q, k, v = (z.view(B, L, heads, hd).transpose(1, 2) for z in (q, k, v))
    a = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, L, C)
    return F.linear(o[:, 0], sd[p + "c_proj.weight"], sd[p + "c_proj.bias"])


def encode_image_resnet(images, sd, stats, prefix="visual.", heads=32):
    x = images
    for conv, bn, s in (("conv1", "bn1", 2), ("conv2", "bn2", 1), ("conv3", "bn3", 1)):       # :193-196
        x = F.relu(_bn(F.conv2d(x, sd[prefix + conv + ".weight"], stride=s, padding=1), sd, prefix + bn, stats))
    x = F.avg_pool2d(x, 2)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        bi = 0
        while (prefix + "layer%d.%d.conv1.weight" % (li, bi)) in sd:
            x = bottleneck(x, sd, stats, prefix + "layer%d.%d." % (li, bi), stride if bi == 0 else 1)
            bi += 1
    return attention_pool(x, sd, prefix + "attnpool.", heads)


def clip_res_step(sd, images, ids):
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running_" not in k) for k, v in sd.items()}
    stats = {k: v.detach().clone() for k, v in sd.items() if "running_" in k}
    fi = encode_image_resnet(images, params, stats)
    ft = clip_ref.encode_text(ids, params)
    li, lt, _, _ = clip_ref.clip_logits(fi, ft, params["logit_scale"])
    loss, labels = clip_ref.clip_info_ce(li, lt)
    loss.backward()
    return {"loss": loss.detach(), "logits_per_image": li.detach(), "image_features": fi.detach(),
            "text_features": ft.detach(), "grads": {k: p.grad for k, p in params.items() if p.grad is not None},
            "stats": stats}
