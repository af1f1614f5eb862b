"""Generates tests/golden/solver_utils.json from the UNMODIFIED reference classes (build container only):
prototype.utils.misc.param_group_all on the reference clip_vitb32 / clip_res50 module trees with the pconfig of
experiments/clip_experiments/yfcc15m/yfcc15m_vit_clip/config.yaml:36-49, and prototype.lr_scheduler Cosine with the
kwargs of the same file (:52-59).
    python tools/make_golden_solver.py"""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import golden, ref_harness  # noqa: E402

PCONFIG = {"bn_w": {"weight_decay": 0}, "bn_b": {"weight_decay": 0}, "ln_w": {"weight_decay": 0},
           "ln_b": {"weight_decay": 0}, "bias": {"weight_decay": 0}, "logit_scale": {"weight_decay": 0}}
DEFAULT = {"lr": 1e-4, "weight_decay": 0.1, "betas": [0.9, 0.98], "eps": 1e-8}
SCHED = {"base_lr": 1e-4, "warmup_lr": 1e-3, "min_lr": 0.0, "warmup_steps": 2500, "max_iter": 128001}
ITERS = [1, 2, 3, 100, 1250, 2499, 2500, 2501, 10000, 64000, 100000, 128000, 128001]


def groups_of(model):
    if "easydict" not in sys.modules:      # imported at module scope by prototype/utils/misc.py:15, unused on this path
        ed = types.ModuleType("easydict")
        ed.EasyDict = dict
        sys.modules["easydict"] = ed
    from prototype.utils.misc import param_group_all
    id2name = {id(p): n for n, p in model.named_parameters()}
    pg, type2num = param_group_all(model, PCONFIG, DEFAULT)
    return ([{"names": [id2name[id(p)] for p in g["params"]], "weight_decay": g["weight_decay"], "lr": g["lr"]} for g in pg],
            dict(type2num))


def main():
    ref_harness.setup()
    os.environ.setdefault("SLURM_PROCID", "0")
    os.environ.setdefault("SLURM_NTASKS", "1")
    out = {"generator": "tools/make_golden_solver.py (reference param_group_all + CosineLRScheduler)", "pconfig": PCONFIG,
           "default": DEFAULT, "sched": SCHED, "iters": ITERS}
    vit = ref_harness.build_clip_vitb32(512, {"layers": 2}, {"transformer_layers": 2})
    out["clip_vitb32_l2"] = dict(zip(("groups", "type2num"), groups_of(vit)))
    from prototype.model import model_entry
    res = model_entry(dict(type="clip_res50", kwargs=dict(
        image_encode=dict(embed_dim=1024, use_sync_bn=False, bn_group_size=1, layers=(1, 1, 1, 1)),
        text_encode=dict(bpe_path=ref_harness._fake_bpe(), text_encode_type="Transformer",
                         text_model_utils=dict(random=False, freeze=False), embed_dim=1024, transformer_layers=1),
        clip=dict(use_allgather=False))))
    out["clip_res50_l1111"] = dict(zip(("groups", "type2num"), groups_of(res)))
    from prototype.lr_scheduler import scheduler_entry
    w = [torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.AdamW([{"params": [w[0]], "lr": 1e-4}, {"params": [w[1]], "lr": 3e-4}])
    cfg = types.SimpleNamespace(type="Cosine", kwargs=dict(SCHED, optimizer=opt))
    sch = scheduler_entry(cfg)
    lrs = []
    for it in ITERS:
        sch.step(it)
        lrs.append(sch.get_lr())
    out["lrs"] = lrs
    path = os.path.join(golden.GOLDEN_DIR, "solver_utils.json")
    json.dump(out, open(path, "w"), indent=0)
    print(path, os.path.getsize(path), "bytes;", [len(g["names"]) for g in out["clip_vitb32_l2"]["groups"]], lrs[:3], lrs[-2:])


if __name__ == "__main__":
    main()
