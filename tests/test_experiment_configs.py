"""The `model:` section of every config the reference ships under experiments/ goes through `model_entry` unchanged
("experiments/*/config.yaml stay unchanged").  Reads /root/reference, so it runs in the build container
only (skipped on the GPU box).  Three shipped configs cannot run in the public reference itself; they must fail HERE with
the documented reason, not silently build something else:
  * yfcc15m_r50_declip (use_sync_bn defaults to True) and yfcc15m_r50_filip (use_sync_bn: True): the reference's linklink
    shim has no `new_group` and aliases SyncBatchNorm2d to BatchNorm1d;
  * yfcc15m_r50_slip: `ModifiedResNet.forward` has no `return_feature` argument (modified_resnet.py:192)."""
import glob
import os

import pytest
import yaml

from declip_b200.model import model_entry

REF = "/root/reference/experiments"
CONFIGS = sorted(glob.glob(os.path.join(REF, "*", "*", "*", "config.yaml")))
CANNOT_RUN_UPSTREAM = {"yfcc15m_r50_declip": "use_sync_bn", "yfcc15m_r50_filip": "use_sync_bn", "yfcc15m_r50_slip": "slip_res50"}


class _Cfg(dict):
    """attribute access like the reference's EasyDict"""
    __getattr__ = dict.get


def _attr(d):
    if isinstance(d, dict):
        return _Cfg({k: _attr(v) for k, v in d.items()})
    return [_attr(x) for x in d] if isinstance(d, list) else d


@pytest.mark.skipif(not CONFIGS, reason="the reference tree is not present on this machine")
@pytest.mark.parametrize("path", CONFIGS, ids=[p.split(os.sep)[-2] for p in CONFIGS])
def test_shipped_experiment_config_builds(path):
    name = path.split(os.sep)[-2]
    model_cfg = _attr(yaml.safe_load(open(path))["model"])
    if name in CANNOT_RUN_UPSTREAM:
        with pytest.raises(NotImplementedError, match=CANNOT_RUN_UPSTREAM[name]):
            model_entry(model_cfg)
        return
    model = model_entry(model_cfg)
    assert sum(p.numel() for p in model.parameters()) > 100e6
    # what the solver reads to build its per-tower parameter groups and to clamp the temperature (clip.py:64-92:
    # visual_parameters() is an empty list in the reference too — the whole image tower is in visual_modules())
    assert hasattr(model, "logit_scale") and list(model.text_parameters()) and list(model.text_modules())
    assert list(model.visual_modules()) and list(model.visual_parameters()) == []


@pytest.mark.skipif(not CONFIGS, reason="the reference tree is not present on this machine")
def test_all_eleven_shipped_configs_are_covered():
    assert len(CONFIGS) == 11 and set(CANNOT_RUN_UPSTREAM) <= {p.split(os.sep)[-2] for p in CONFIGS}
