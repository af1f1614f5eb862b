"""TEST INFRASTRUCTURE — golden-vector format shared by tools/make_golden.py and the tests."""
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SAMPLE = 1024

CASES = {
    # name: dict(batch, v_layers, t_layers, embed_dim, seed)
    "clip_vitb32_l2_b8": dict(batch=8, v_layers=2, t_layers=2, embed_dim=512, seed=1),
    "clip_vitb32_l12_b32": dict(batch=32, v_layers=12, t_layers=12, embed_dim=512, seed=0),   # BASELINE configs[0]
    "clip_vitb32_l12_b512": dict(batch=512, v_layers=12, t_layers=12, embed_dim=512, seed=3),  # configs[1] per-GPU shape
}
DECLIP_CASES = {
    "declip_vitb32_l2_b8": dict(batch=8, v_layers=2, t_layers=2, embed_dim=512, seed=2, nn_size=1024),
    "declip_vitb32_l12_b64": dict(batch=64, v_layers=12, t_layers=12, embed_dim=512, seed=12, nn_size=4096),  # configs[2] depth
}


SLIP_CASES = {
    "slip_vitb32_l2_b8": dict(batch=8, v_layers=2, t_layers=2, embed_dim=512, seed=9),
}


def slip_inputs(c):
    from . import synth
    sd = synth.slip_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"], t_layers=c["t_layers"])
    images = torch.cat([synth.synth_images(c["batch"], seed=c["seed"], channels=6),
                        synth.synth_images(c["batch"], seed=c["seed"] + 50, channels=3)], dim=1)
    return sd, images, synth.synth_token_ids(c["batch"], seed=c["seed"])


DEFILIP_CASES = {
    # DeCLIP batch + token-wise late interaction on all four (view, caption) combinations (return_filip, dense_aug)
    "defilip_vitb32_l2_b8": dict(batch=8, v_layers=2, t_layers=2, embed_dim=512, seed=7, nn_size=1024),
}


def defilip_inputs(c):
    from . import synth
    sd, images, mlm_ids, mlm_labels, ids_aug, bank = declip_inputs(c)
    extra = synth.filip_extra_state_dict(seed=c["seed"])
    for k in ("logit_scale_dense", "image_mapping.weight", "image_mapping.bias", "text_mapping.weight", "text_mapping.bias"):
        sd[k] = extra[k]
    return sd, images, mlm_ids, mlm_labels, ids_aug, bank


FILIP_CASES = {
    "filip_vitb32_l2_b8": dict(batch=8, v_layers=2, t_layers=2, embed_dim=768, seed=4),
    "filip_vitb32_l12_b64": dict(batch=64, v_layers=12, t_layers=12, embed_dim=768, seed=14),   # configs[4] depth
}


def filip_inputs(c):
    from . import synth
    sd = synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"], t_layers=c["t_layers"])
    sd.update(synth.filip_extra_state_dict(seed=c["seed"]))
    images = synth.synth_images(c["batch"], seed=c["seed"], channels=6)
    ids = synth.synth_token_ids(c["batch"], seed=c["seed"])
    mlm_ids, mlm_labels = synth.synth_mlm(ids, seed=c["seed"])
    return sd, images, mlm_ids, mlm_labels


RES_CASES = {
    "clip_res50_l1111_b4": dict(batch=4, layers=(1, 1, 1, 1), t_layers=2, embed_dim=1024, seed=6),
    # configs[3] depth; last-BatchNorm gains x0.25 (the reference zero-initialises them): see synth.resnet_state_dict
    "clip_res50_l3463_b32": dict(batch=32, layers=(3, 4, 6, 3), t_layers=12, embed_dim=1024, seed=16, bn3_scale=0.25),
}


def res_inputs(c):
    from . import synth
    sd = synth.clip_res_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], layers=c["layers"], t_layers=c["t_layers"],
                                   bn3_scale=c.get("bn3_scale", 1.0))
    return sd, synth.synth_images(c["batch"], seed=c["seed"]), synth.synth_token_ids(c["batch"], seed=c["seed"])


def declip_inputs(c):
    """Synthetic DeCLIP batch for a case: (state_dict, images[B,6,H,W], mlm_ids, mlm_labels, ids_aug, bank[dim,size])."""
    from . import synth
    sd = synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"], t_layers=c["t_layers"])
    sd.update(synth.declip_extra_state_dict(seed=c["seed"], feature_dim=c["embed_dim"]))
    images = synth.synth_images(c["batch"], seed=c["seed"], channels=6)
    ids = synth.synth_token_ids(c["batch"], seed=c["seed"])
    mlm_ids, mlm_labels = synth.synth_mlm(ids, seed=c["seed"])
    ids_aug = synth.synth_token_ids(c["batch"], seed=c["seed"] + 100)
    bank = synth.synth_bank(c["embed_dim"], c["nn_size"], seed=c["seed"])
    return sd, images, mlm_ids, mlm_labels, ids_aug, bank


def sample_index(numel):
    step = max(1, numel // SAMPLE)
    return torch.arange(0, numel, step)[:SAMPLE]


def summarise_grads(grads):
    """Per-parameter: L2 norm + a deterministic strided sample (full tensor when small)."""
    out = {}
    for k, g in grads.items():
        g = g.detach().float().reshape(-1)
        out[k] = {"norm": g.norm().item(), "sample": g[sample_index(g.numel())].clone(), "numel": g.numel()}
    return out


def path(name):
    return os.path.join(GOLDEN_DIR, name + ".pt")


def load(name):
    return torch.load(path(name), map_location="cpu", weights_only=False)
