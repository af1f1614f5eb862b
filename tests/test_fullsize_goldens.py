"""CPU checks of the full-depth golden vectors (generated from the reference's own modules by tools/make_golden.py):
present, well-formed, and consistent with the oracle restatement where that is cheap."""
import os

import pytest
import torch

from oracle import golden

NEW = ["clip_vitb32_l12_b512", "declip_vitb32_l12_b64", "filip_vitb32_l12_b64", "clip_res50_l3463_b32"]


@pytest.mark.parametrize("name", NEW)
def test_golden_present_and_wellformed(name):
    assert os.path.exists(golden.path(name)), "run tools/make_golden.py %s in the build container" % name
    g = golden.load(name)
    assert "reference" in g["generator"]
    assert len(g["grads"]) > 300 and all(torch.isfinite(v["sample"]).all() for v in g["grads"].values())
    c = g["case"]
    depth = c.get("v_layers", None)
    if depth is not None:
        assert depth == 12 and c["t_layers"] == 12
    else:
        assert tuple(c["layers"]) == (3, 4, 6, 3) and c["t_layers"] == 12


def test_b512_golden_loss_is_near_ln_n():
    import math
    g = golden.load("clip_vitb32_l12_b512")
    assert g["logits_per_image"].shape == (512, 512)
    assert abs(g["loss"] - math.log(512)) < 0.5           # random init: close to the uniform-softmax loss
    assert torch.allclose(g["logits_per_image"], g["logits_per_text"].t(), atol=1e-4)
