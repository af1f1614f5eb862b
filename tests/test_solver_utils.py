"""declip_b200.solver_utils against values produced by the reference's own classes (tools/make_golden_solver.py ->
tests/golden/solver_utils.json): parameter groups by module type, cosine/warm-up learning rates, logit-scale clamps."""
import json
import os

import pytest
import torch

from declip_b200 import solver_utils
from declip_b200.model import model_entry

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "solver_utils.json")))


def _groups(model):
    id2name = {id(p): n for n, p in model.named_parameters()}
    pg, type2num = solver_utils.param_group_all(model, G["pconfig"], G["default"])
    return [{"names": [id2name[id(p)] for p in g["params"]], "weight_decay": g["weight_decay"], "lr": g["lr"]} for g in pg], dict(type2num)


@pytest.mark.parametrize("case", ["clip_vitb32_l2", "clip_res50_l1111"])
def test_param_groups_match_reference(case):
    if case == "clip_vitb32_l2":
        cfg = dict(type="clip_vitb32", kwargs=dict(image_encode=dict(embed_dim=512, layers=2),
                                                   text_encode=dict(embed_dim=512, transformer_layers=2, bpe_path=None,
                                                                    text_encode_type="Transformer"),
                                                   clip=dict(use_allgather=False)))
    else:
        cfg = dict(type="clip_res50", kwargs=dict(image_encode=dict(embed_dim=1024, layers=(1, 1, 1, 1), use_sync_bn=False, bn_group_size=1),
                                                  text_encode=dict(embed_dim=1024, transformer_layers=1, bpe_path=None,
                                                                   text_encode_type="Transformer"),
                                                  clip=dict(use_allgather=False)))
    model = model_entry(cfg)
    groups, type2num = _groups(model)
    ref = G[case]
    assert len(groups) == len(ref["groups"])
    for mine, theirs in zip(groups, ref["groups"]):
        assert mine["names"] == theirs["names"]
        assert mine["weight_decay"] == theirs["weight_decay"] and mine["lr"] == theirs["lr"]
    assert type2num == ref["type2num"]
    # every parameter lands in exactly one group
    assert sorted(n for g in groups for n in g["names"]) == sorted(n for n, _ in model.named_parameters())


def test_cosine_schedule_matches_reference():
    w = [torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.AdamW([{"params": [w[0]], "lr": 1e-4}, {"params": [w[1]], "lr": 3e-4}])
    sch = solver_utils.scheduler_entry({"type": "Cosine", "kwargs": dict(G["sched"], optimizer=opt)})
    for it, want in zip(G["iters"], G["lrs"]):
        sch.step(it)
        got = sch.get_lr()
        assert got == pytest.approx(want, rel=1e-12, abs=1e-18), (it, got, want)


def test_logit_scale_clip_types():
    p = torch.nn.Parameter(torch.tensor([7.5]))
    c = solver_utils.LogitScaleClip(p, "logit_scale_param_value", 3, 6)   # yfcc15m configs (config.yaml:20-23)
    c.before()
    assert p.item() == 6.0
    p.data.fill_(1.0)
    c.after()
    assert p.item() == 3.0
    c = solver_utils.LogitScaleClip(p, "logit_scale_param_abs_min", 4.0)
    c.before()
    assert p.item() == 4.0
    c = solver_utils.LogitScaleClip(p, "logit_scale_param", 0.25)
    c.before()
    p.data.add_(1.0)
    c.after()
    assert p.item() == pytest.approx(4.25)
    c.before()
    p.data.sub_(0.1)
    c.after()
    assert p.item() == pytest.approx(4.15)


def test_factory_guards():
    """Factories that cannot work raise at construction: DeFILIP's dense path without MLM reads an undefined
    `word_features` in the reference (defilip.py:296-302,335); slip_res50 calls ModifiedResNet.forward with a
    `return_feature` argument it does not have (modified_resnet.py:192)."""
    kw = dict(image_encode=dict(embed_dim=512, layers=1),
              text_encode=dict(bpe_path=None, text_encode_type='Transformer', embed_dim=512, transformer_layers=1))
    with pytest.raises(NotImplementedError):
        model_entry(dict(type='defilip_vitb32', kwargs=dict(kw, clip=dict(use_allgather=True, return_filip=True, feature_dim=512))))
    with pytest.raises(NotImplementedError):
        model_entry(dict(type='slip_res50', kwargs=dict(kw, clip=dict(use_allgather=True))))
    with pytest.raises(KeyError):
        model_entry(dict(type='no_such_model', kwargs={}))
    m = model_entry(dict(type='slip_vitb32', kwargs=dict(kw, clip=dict(use_allgather=True, return_sim=True, feature_dim=768))))
    assert any(k.startswith('text_encoder.') for k in m.state_dict()) and not any(k.startswith('encode_text.') for k in m.state_dict())
