"""Checkpoint compatibility on the CPU (no GPU needed): every model family accepts — `strict=True` — the
state dict the reference's OWN module accepted when the goldens were produced (tools/make_golden.py loads the synthetic
weights of oracle/synth.py into the unmodified reference modules), every parameter the reference produced a gradient for
exists under the same name and size, and a `declip_b200.dist.DistModule` wrapper keeps the `module.` prefix under which the
reference's solver saves and restores checkpoints (clip_solver.py saves `self.model.state_dict()` of the wrapped model;
prototype/utils/misc.py:441-453 `load_state_model` -> `model.load_state_dict(state, strict=False)`)."""
import pytest
import torch

from declip_b200.model import model_entry
from oracle import golden, synth


def _text(c):
    return dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                embed_dim=c["embed_dim"], transformer_layers=c["t_layers"])


def _vit(c):
    return dict(embed_dim=c["embed_dim"], layers=c["v_layers"])


def _build(name):
    """(model, state dict with the reference's keys, golden) for one golden case — the configs of the GPU parity tests."""
    g = golden.load(name)
    c = g["case"]
    if name.startswith("clip_vitb32"):
        cfg = dict(type='clip_vitb32', kwargs=dict(image_encode=_vit(c), text_encode=_text(c), clip=dict(use_allgather=False)))
        sd = synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"], t_layers=c["t_layers"])
    elif name.startswith("clip_res50"):
        cfg = dict(type='clip_res50', kwargs=dict(
            image_encode=dict(embed_dim=c["embed_dim"], use_sync_bn=False, bn_group_size=1, layers=tuple(c["layers"])),
            text_encode=_text(c), clip=dict(use_allgather=False)))
        sd = golden.res_inputs(c)[0]
    elif name.startswith("declip"):
        cfg = dict(type='declip_vitb32', kwargs=dict(
            image_encode=_vit(c), text_encode=_text(c),
            clip=dict(use_allgather=True, text_mask_type='MLM', return_nn_bank=True, feature_dim=c["embed_dim"], nn_size=c["nn_size"])))
        sd = golden.declip_inputs(c)[0]
    elif name.startswith("filip"):
        cfg = dict(type='filip_vitb32', kwargs=dict(
            image_encode=_vit(c), text_encode=_text(c),
            clip=dict(use_allgather=True, text_mask_type='MLM', return_dense=True, select_topk=True, feature_dim=c["embed_dim"],
                      mask_rate=0.5, patch_number=14)))
        sd = golden.filip_inputs(c)[0]
    elif name.startswith("slip"):
        cfg = dict(type='slip_vitb32', kwargs=dict(
            image_encode=_vit(c), text_encode=_text(c), clip=dict(use_allgather=True, return_sim=True, feature_dim=768, sim_dim=256)))
        sd = golden.slip_inputs(c)[0]
    else:
        cfg = dict(type='defilip_vitb32', kwargs=dict(
            image_encode=_vit(c), text_encode=_text(c),
            clip=dict(use_allgather=True, text_mask_type='MLM', return_nn_bank=True, feature_dim=c["embed_dim"],
                      nn_size=c["nn_size"], return_filip=True, dense_aug=True)))
        sd = golden.defilip_inputs(c)[0]
    return model_entry(cfg), sd, g


@pytest.mark.parametrize("name", ["clip_vitb32_l2_b8", "clip_res50_l1111_b4", "declip_vitb32_l2_b8", "filip_vitb32_l2_b8",
                                  "slip_vitb32_l2_b8", "defilip_vitb32_l2_b8"])
def test_reference_state_dict_loads_strictly_and_gradient_names_match(name):
    model, sd, g = _build(name)
    model.load_state_dict(sd, strict=True)                     # same keys, same shapes as the reference's module
    params = dict(model.named_parameters())
    for k, ref in g["grads"].items():                          # what the reference's backward produced a gradient for
        assert k in params, k
        if "numel" in ref:
            assert params[k].numel() == ref["numel"], k
    if "param_names" in g:                                     # registration order too (optimizer state_dicts index by position)
        assert [k for k, _ in model.named_parameters()] == list(g["param_names"])


def test_dist_module_checkpoints_carry_the_module_prefix_and_round_trip():
    from declip_b200 import dist as ddist
    model, sd, _ = _build("clip_vitb32_l2_b8")
    model.load_state_dict(sd, strict=True)
    wrapped = ddist.DistModule(model)                           # no process group: world size 1, nothing is broadcast
    state = wrapped.state_dict()
    assert list(state) == ["module." + k for k in model.state_dict()]
    # a checkpoint as the reference's solver writes it: {'model': DistModule.state_dict(), ...}; restore it into a fresh model
    ckpt = {"model": {k: v.clone() for k, v in state.items()}, "last_iter": 7}
    fresh = ddist.DistModule(_build("clip_vitb32_l2_b8")[0])
    result = fresh.load_state_dict(ckpt["model"], strict=False)             # load_state_model's call (misc.py:446)
    assert not result.missing_keys and not result.unexpected_keys
    for k, v in fresh.module.state_dict().items():
        assert torch.equal(v, sd[k] if v.dtype == sd[k].dtype else sd[k].to(v.dtype)), k
    # a bare model consumes the same checkpoint after stripping the prefix
    bare = _build("clip_vitb32_l2_b8")[0]
    bare.load_state_dict({k[len("module."):]: v for k, v in ckpt["model"].items()}, strict=True)
