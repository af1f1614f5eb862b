"""TEST INFRASTRUCTURE — CPU/fp32 restatement (oracle) of the reference's dual-encoder hot path.

Plain PyTorch fp32 functions over a state_dict, each citing the reference lines it follows
(paths relative to /root/reference).  Pinned against the reference's own modules executed in the
build container: `tools/make_golden.py` runs prototype.model.clip_vitb32 + ClipInfoCELoss on the
synthetic weights/inputs of `oracle/synth.py` and stores the outputs in tests/golden/;
tests/test_oracle.py checks this file against those vectors.  The reference itself ships no golden
vectors or tests, so that pin is the only one available.

This is the CHECKER, never the product: nothing in declip_b200/ imports it.
"""
import math

import torch
import torch.nn.functional as F


def layer_norm(x, w, b):
    # base_transformer.py:10-18 — nn.LayerNorm, eps 1e-5, in the tensor's own dtype
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def quick_gelu(x):
    # base_transformer.py:24-26
    return x * torch.sigmoid(1.702 * x)


def attention(x, sd, p, heads, mask):
    """nn.MultiheadAttention(d, h)(x, x, x, attn_mask=mask)[0] on NLD input.
    base_transformer.py:33,44-48 (the reference runs it in LND; the math is layout independent)."""
    B, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, L, heads, hd).transpose(1, 2)
    k = k.view(B, L, heads, hd).transpose(1, 2)
    v = v.view(B, L, heads, hd).transpose(1, 2)
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, L, D)
    return F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])


def resblock(x, sd, p, heads, mask):
    # base_transformer.py:50-53
    x = x + attention(layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]), sd, p, heads, mask)
    h = F.linear(layer_norm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]), sd[p + "mlp.c_fc.weight"],
                 sd[p + "mlp.c_fc.bias"])
    return x + F.linear(quick_gelu(h), sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])


def _n_layers(sd, prefix):
    n = 0
    while (prefix + "%d.ln_1.weight" % n) in sd:
        n += 1
    return n


def encode_image(images, sd, prefix="visual.", heads=None, return_dense=False, return_feature=False):
    """VisualTransformer.forward — visual_transformer.py:55-82 (ret = [x] + [dense_feat] + [feature], :74-82)."""
    w = sd[prefix + "conv1.weight"]
    width, patch = w.shape[0], w.shape[-1]
    heads = heads or width // 64
    x = F.conv2d(images, w, stride=patch)                              # :56
    x = x.reshape(x.shape[0], width, -1).permute(0, 2, 1)              # :58-59
    cls = sd[prefix + "class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, width, dtype=x.dtype)
    x = torch.cat([cls, x], dim=1)                                     # :60-61
    x = x + sd[prefix + "positional_embedding"]                        # :62
    x = layer_norm(x, sd[prefix + "ln_pre.weight"], sd[prefix + "ln_pre.bias"])   # :63
    for i in range(_n_layers(sd, prefix + "transformer.resblocks.")):
        x = resblock(x, sd, prefix + "transformer.resblocks.%d." % i, heads, None)
    dense = x[:, 1:, :]                                                # :68
    x = layer_norm(x[:, 0, :], sd[prefix + "ln_post.weight"], sd[prefix + "ln_post.bias"])   # :69
    feature = x                                                        # :70
    x = x @ sd[prefix + "proj"]                                        # :72-73
    ret = [x] + ([dense] if return_dense else []) + ([feature] if return_feature else [])
    return ret[0] if len(ret) == 1 else tuple(ret)


def causal_mask(ctx):
    # text_transformer.py:136-142
    m = torch.empty(ctx, ctx)
    m.fill_(float("-inf"))
    m.triu_(1)
    return m


def encode_text(ids, sd, prefix="encode_text.", heads=None, return_dense=False):
    """TextTransformer.forward, 'Transformer' branch, ids already tokenised — text_transformer.py:183-204."""
    emb = sd[prefix + "token_embedding.weight"]
    width = emb.shape[1]
    heads = heads or width // 64
    x = F.embedding(ids, emb)                                          # :188
    x = x + sd[prefix + "positional_embedding"]                        # :190
    mask = causal_mask(ids.shape[1]).to(x.dtype)
    for i in range(_n_layers(sd, prefix + "transformer.resblocks.")):
        x = resblock(x, sd, prefix + "transformer.resblocks.%d." % i, heads, mask)
    x = layer_norm(x, sd[prefix + "ln_final.weight"], sd[prefix + "ln_final.bias"])   # :194
    words = x
    x = x[torch.arange(x.shape[0]), ids.argmax(dim=-1)]                # :203 (EOT = highest id)
    x = F.linear(x, sd[prefix + "text_projection.weight"], sd[prefix + "text_projection.bias"])
    return (x, words) if return_dense else x


def clip_logits(image_features, text_features, logit_scale, gathered_image=None, gathered_text=None):
    """CLIP.forward after the encoders — clip.py:129-141.  Returns (logits_per_image, logits_per_text)
    and the normalised features.  `gathered_*` (already normalised, [N,E]) model the all-gather path."""
    i = image_features / image_features.norm(dim=-1, keepdim=True)             # :129
    t = text_features / (text_features.norm(dim=-1, keepdim=True) + 1e-10)     # :130
    s = logit_scale.exp()                                                      # :133
    s.data = torch.clamp(s.data, max=100)                                      # :134 (grad flows as if unclamped)
    gi = i if gathered_image is None else gathered_image
    gt = t if gathered_text is None else gathered_text
    return s * i @ gt.t(), s * t @ gi.t(), i, t                                # :140-141


def clip_info_ce(logits_per_image, logits_per_text, rank=0):
    """ClipInfoCELoss.forward — loss_functions/loss.py:40-50."""
    bs, n = logits_per_image.shape
    labels = torch.arange(bs) if n == bs else rank * bs + torch.arange(bs)
    loss = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
    return loss, labels


def accuracy(output, target, topk=(1, 5)):
    """prototype/utils/misc.py:415-428."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * (100.0 / target.size(0)) for k in topk]


def clip_step(sd, images, ids, requires_grad=True):
    """One reference training step's forward+loss+backward (clip_solver.py:413-418,561) at world size 1.
    Returns dict(loss, logits_per_image, image_features, text_features, grads{key: tensor})."""
    params = {k: v.detach().clone().requires_grad_(requires_grad and k != "visual.conv1.weight") for k, v in sd.items()}
    fi = encode_image(images, params)
    ft = encode_text(ids, params)
    li, lt, _, _ = clip_logits(fi, ft, params["logit_scale"])
    loss, labels = clip_info_ce(li, lt)
    out = {"loss": loss.detach(), "logits_per_image": li.detach(), "logits_per_text": lt.detach(),
           "image_features": fi.detach(), "text_features": ft.detach(), "labels": labels}
    if requires_grad:
        loss.backward()
        out["grads"] = {k: p.grad for k, p in params.items() if p.grad is not None}
    return out
