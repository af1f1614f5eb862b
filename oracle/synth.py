"""TEST INFRASTRUCTURE (oracle side) — deterministic synthetic weights and inputs.

Both the reference modules (imported from /root/reference in the build container by
`oracle/ref_harness.py`) and the CUDA path load the SAME state_dict produced here, so parity never
depends on reproducing the reference's RNG call order.  Keys/shapes follow the reference's
state_dict (prototype/model/clip.py:51-60, image_encoder/visual_transformer.py:6-27,
text_encoder/text_transformer.py:27-44, image_encoder/base_transformer.py:29-42).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may
import this package.
"""
import math
import zlib

import torch

VOCAB = 49409          # simple_tokenizer.py:66-75 : 49408 BPE/byte entries + <|mask|>
SOT, EOT, MASK = 49407, 49408, 49406


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _randn(key, seed, shape, std):
    return torch.randn(*shape, generator=_gen(key, seed), dtype=torch.float32) * std


def _block_keys(prefix, width):
    return [
        (prefix + "attn.in_proj_weight", (3 * width, width), width ** -0.5),
        (prefix + "attn.in_proj_bias", (3 * width,), 0.02),
        (prefix + "attn.out_proj.weight", (width, width), width ** -0.5),
        (prefix + "attn.out_proj.bias", (width,), 0.02),
        (prefix + "ln_1.weight", (width,), None),
        (prefix + "ln_1.bias", (width,), 0.02),
        (prefix + "mlp.c_fc.weight", (4 * width, width), (2 * width) ** -0.5),
        (prefix + "mlp.c_fc.bias", (4 * width,), 0.02),
        (prefix + "mlp.c_proj.weight", (width, 4 * width), (2 * width) ** -0.5 * 0.5),
        (prefix + "mlp.c_proj.bias", (width,), 0.02),
        (prefix + "ln_2.weight", (width,), None),
        (prefix + "ln_2.bias", (width,), 0.02),
    ]


def clip_vit_state_dict(seed=0, embed_dim=512, v_layers=12, t_layers=12, v_width=768, t_width=512, res=224, patch=32,
                        ctx=77, vocab=VOCAB):
    """state_dict for `clip_vitb32` (clip.py:158-165) with overridable depth for small test models."""
    spec = [("logit_scale", (1,), "logit")]
    g2 = (res // patch) ** 2
    spec += [
        ("visual.class_embedding", (v_width,), v_width ** -0.5),
        ("visual.positional_embedding", (g2 + 1, v_width), 0.01 * 4),
        ("visual.proj", (v_width, embed_dim), v_width ** -0.5),
        ("visual.conv1.weight", (v_width, 3, patch, patch), (3 * patch * patch) ** -0.5),
        ("visual.ln_pre.weight", (v_width,), None), ("visual.ln_pre.bias", (v_width,), 0.02),
        ("visual.ln_post.weight", (v_width,), None), ("visual.ln_post.bias", (v_width,), 0.02),
    ]
    for i in range(v_layers):
        spec += _block_keys("visual.transformer.resblocks.%d." % i, v_width)
    spec += [
        ("encode_text.positional_embedding", (ctx, t_width), 0.01 * 4),
        ("encode_text.token_embedding.weight", (vocab, t_width), 0.02 * 4),
        ("encode_text.ln_final.weight", (t_width,), None), ("encode_text.ln_final.bias", (t_width,), 0.02),
        ("encode_text.text_projection.weight", (embed_dim, t_width), t_width ** -0.5),
        ("encode_text.text_projection.bias", (embed_dim,), 0.02),
    ]
    for i in range(t_layers):
        spec += _block_keys("encode_text.transformer.resblocks.%d." % i, t_width)
    sd = {}
    for key, shape, std in spec:
        if std == "logit":
            sd[key] = torch.full(shape, math.log(1 / 0.07), dtype=torch.float32)   # clip.py:59
        elif std is None:   # LayerNorm gain: 1 + noise so the affine path is exercised
            sd[key] = 1.0 + _randn(key, seed, shape, 0.1)
        else:
            sd[key] = _randn(key, seed, shape, std)
    return sd


def synth_images(batch, seed=0, channels=3, res=224):
    return torch.randn(batch, channels, res, res, generator=_gen("images", seed), dtype=torch.float32)


def synth_token_ids(batch, seed=0, ctx=77, vocab=VOCAB):
    """int64 [B,ctx]: SOT, random body of length U[8,ctx-2], EOT (= max id so argmax finds it;
    text_transformer.py:203), zero padding — the layout `tokenize` produces (text_transformer.py:144-180)."""
    g = _gen("ids", seed)
    ids = torch.zeros(batch, ctx, dtype=torch.int64)
    lens = torch.randint(8, ctx - 1, (batch,), generator=g)
    body = torch.randint(1, 49000, (batch, ctx), generator=g)
    for b in range(batch):
        n = int(lens[b])
        ids[b, 0] = SOT
        ids[b, 1:1 + n] = body[b, :n]
        ids[b, 1 + n] = EOT
    return ids


def declip_extra_state_dict(seed=0, feature_dim=512, t_width=512, vocab=VOCAB):
    """Extra DECLIP keys (declip.py:48-60,107-112,171): SimSiam projector / predictor (+ BatchNorm buffers) and the
    MLM head `text_label_predictor`."""
    sd = {}

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = _randn(name + ".weight", seed, (out_f, in_f), in_f ** -0.5)
        sd[name + ".bias"] = _randn(name + ".bias", seed, (out_f,), 0.02)

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + _randn(name + ".weight", seed, (c,), 0.1)
        sd[name + ".bias"] = _randn(name + ".bias", seed, (c,), 0.02)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    lin("projector.linear1", 1024, feature_dim); bn("projector.bn1", 1024)
    lin("projector.linear2", 1024, 1024); bn("projector.bn2", 1024)
    lin("projector.linear3", 1024, 1024); bn("projector.bn3", 1024)
    lin("predictor.linear1", 512, 1024); bn("predictor.bn1", 512)
    lin("predictor.layer2", 1024, 512)
    lin("text_label_predictor", vocab, t_width)
    return sd


def synth_bank(dim=512, size=1024, seed=0):
    """NN memory bank in the reference layout [dim, size], unit-norm columns (memory_bank.py:66-67)."""
    b = torch.randn(dim, size, generator=_gen("bank", seed))
    return torch.nn.functional.normalize(b, dim=0)


def synth_mlm(ids, seed=0):
    """Deterministic BERT-style masking of synthetic ids (mask_tokens.py:5-29 distribution): returns (masked_ids, labels)."""
    g = _gen("mlm", seed)
    ids = ids.clone()
    B, L = ids.shape
    lens = ids.argmax(1) + 1
    pos = torch.arange(L).unsqueeze(0)
    valid = (pos < lens.unsqueeze(1)) & (ids != SOT) & (ids != EOT) & (ids != MASK)
    masked = (torch.rand(B, L, generator=g) < 0.15) & valid
    labels = torch.where(masked, ids, torch.full_like(ids, -100))
    r = torch.rand(B, L, generator=g)
    ids[masked & (r < 0.8)] = MASK
    rnd = masked & (r >= 0.8) & (r < 0.9)
    ids[rnd] = torch.randint(0, VOCAB, (B, L), generator=g)[rnd]
    return ids, labels


def slip_state_dict(seed=0, embed_dim=512, v_layers=12, t_layers=12, feature_dim=768, sim_dim=256, hidden=4096):
    """state_dict of the reference SLIP (slip.py:111-120,196-204): the CLIP keys with the text tower under `text_encoder.`
    plus `predictor_sim` (Linear-BN x2, Linear, and the unused bn3)."""
    sd = {}
    for k, v in clip_vit_state_dict(seed=seed, embed_dim=embed_dim, v_layers=v_layers, t_layers=t_layers).items():
        sd[("text_encoder." + k[len("encode_text."):]) if k.startswith("encode_text.") else k] = v

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = _randn(name + ".weight", seed, (out_f, in_f), in_f ** -0.5)
        sd[name + ".bias"] = _randn(name + ".bias", seed, (out_f,), 0.02)

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + _randn(name + ".weight", seed, (c,), 0.1)
        sd[name + ".bias"] = _randn(name + ".bias", seed, (c,), 0.02)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    lin("predictor_sim.linear1", hidden, feature_dim); bn("predictor_sim.bn1", hidden)
    lin("predictor_sim.linear2", hidden, hidden); bn("predictor_sim.bn2", hidden)
    lin("predictor_sim.linear3", sim_dim, hidden); bn("predictor_sim.bn3", hidden)
    return sd


def filip_extra_state_dict(seed=0, v_width=768, t_width=512, dense_dim=256, vocab=VOCAB):
    """Extra FILIP keys (filip.py:40-55): token mappings, dense logit scale, (unused) MLM head."""
    sd = {"logit_scale_dense": torch.tensor(math.log(1 / 0.07), dtype=torch.float32)}
    for name, o, i in (("image_mapping", dense_dim, v_width), ("text_mapping", dense_dim, t_width),
                       ("text_label_predictor", vocab, t_width)):
        sd[name + ".weight"] = _randn(name + ".weight", seed, (o, i), i ** -0.5)
        sd[name + ".bias"] = _randn(name + ".bias", seed, (o,), 0.02)
    return sd


def resnet_state_dict(seed=0, layers=(3, 4, 6, 3), width=64, embed_dim=1024, res=224, prefix="visual.", bn3_scale=1.0):
    """state_dict of ModifiedResNet (modified_resnet.py:109-190) under `prefix`, BatchNorm buffers included.
    `bn3_scale` multiplies the gain of every block's last BatchNorm: the reference zero-initialises it
    (modified_resnet.py:177-180) so that a freshly built 16-block network is a near-identity map; with gains ~1 a RANDOM
    network of that depth amplifies any perturbation ~1.25x per block (measured: tools/resnet_debug.py) and no reduced-
    precision implementation can follow its fp32 gradients — full-depth parity cases use a small non-zero gain."""
    sd = {}

    def conv(name, cout, cin, k):
        sd[prefix + name + ".weight"] = _randn(prefix + name, seed, (cout, cin, k, k), (cin * k * k) ** -0.5)

    def bn(name, c):
        sd[prefix + name + ".weight"] = 1.0 + _randn(prefix + name + ".w", seed, (c,), 0.1)
        sd[prefix + name + ".bias"] = _randn(prefix + name + ".b", seed, (c,), 0.05)
        sd[prefix + name + ".running_mean"] = torch.zeros(c)
        sd[prefix + name + ".running_var"] = torch.ones(c)
        sd[prefix + name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", width // 2, 3, 3); bn("bn1", width // 2)
    conv("conv2", width // 2, width // 2, 3); bn("bn2", width // 2)
    conv("conv3", width, width // 2, 3); bn("bn3", width)
    inplanes = width
    for li, (planes, blocks, stride) in enumerate(zip((width, width * 2, width * 4, width * 8), layers, (1, 2, 2, 2)), 1):
        for bi in range(blocks):
            p = "layer%d.%d." % (li, bi)
            s = stride if bi == 0 else 1
            conv(p + "conv1", planes, inplanes, 1); bn(p + "bn1", planes)
            conv(p + "conv2", planes, planes, 3); bn(p + "bn2", planes)
            conv(p + "conv3", planes * 4, planes, 1); bn(p + "bn3", planes * 4)
            sd[prefix + p + "bn3.weight"] = sd[prefix + p + "bn3.weight"] * bn3_scale
            if s > 1 or inplanes != planes * 4:
                conv(p + "downsample.0", planes * 4, inplanes, 1); bn(p + "downsample.1", planes * 4)
            inplanes = planes * 4
    feat = width * 32
    sd[prefix + "attnpool.positional_embedding"] = _randn(prefix + "attnpool.pos", seed, ((res // 32) ** 2 + 1, feat),
                                                          feat ** -0.5)
    for n in ("k_proj", "q_proj", "v_proj"):
        sd[prefix + "attnpool.%s.weight" % n] = _randn(prefix + "attnpool." + n, seed, (feat, feat), feat ** -0.5)
        sd[prefix + "attnpool.%s.bias" % n] = _randn(prefix + "attnpool." + n + ".b", seed, (feat,), 0.02)
    sd[prefix + "attnpool.c_proj.weight"] = _randn(prefix + "attnpool.c_proj", seed, (embed_dim, feat), feat ** -0.5)
    sd[prefix + "attnpool.c_proj.bias"] = _randn(prefix + "attnpool.c_proj.b", seed, (embed_dim,), 0.02)
    sd[prefix + "fc.weight"] = _randn(prefix + "fc", seed, (embed_dim, 2048), 2048 ** -0.5)
    sd[prefix + "fc.bias"] = _randn(prefix + "fc.b", seed, (embed_dim,), 0.02)
    return sd


def clip_res_state_dict(seed=0, embed_dim=1024, layers=(3, 4, 6, 3), t_layers=12, bn3_scale=1.0):
    """clip_res50 (clip.py:149-156): ModifiedResNet image tower + the text transformer."""
    full = clip_vit_state_dict(seed=seed, embed_dim=embed_dim, v_layers=0, t_layers=t_layers)
    sd = {k: v for k, v in full.items() if not k.startswith("visual.")}
    sd.update(resnet_state_dict(seed=seed, layers=layers, embed_dim=embed_dim, bn3_scale=bn3_scale))
    return sd
