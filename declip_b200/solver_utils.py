"""Host-side pieces of the training step that sit right after backward, mirrored so that a
solver written against the reference keeps its behaviour when the model comes from `declip_b200.model`:

* `param_group_all`  — weight-decay / lr groups by module type            (prototype/utils/misc.py:266-412)
* `CosineLRScheduler`, `scheduler_entry` — linear warm-up + cosine decay    (prototype/lr_scheduler/scheduler.py:7-84,200-246,
                                                                            prototype/lr_scheduler/__init__.py:4-27)
* `LogitScaleClip`   — the `grad_clip.type: logit_scale_param*` parameter clamps around `optimizer.step()`
                                                                           (prototype/solver/clip_solver.py:500-522)

Nothing here touches the GPU hot path; it exists so `FusedAdamW(param_group_all(...)[0], ...)` + `scheduler_entry(...)`
reproduce the reference's optimiser state evolution (pinned by tests/test_solver_utils.py against values produced by the
reference's own classes, tests/golden/solver_utils.json)."""
import copy
import math
from collections import OrderedDict, defaultdict

import torch

_GROUP_KEYS = ("bn_w", "bn_b", "conv_b", "linear_b", "ln_w", "ln_b")
_OPTIONAL_KEYS = ("conv_dw_w", "conv_dw_b", "conv_dense_w", "conv_dense_b", "linear_w", "logit_scale", "bias")


def _pname(prefix, suffix):
    return prefix + "." + suffix if prefix else suffix


def param_group_all(model, config, default_config=None):
    """Returns (param_groups, type2num) like the reference: group 0 holds every parameter no rule claimed, then one
    group per rule key in the reference's order (the six fixed keys, then the optional ones present in `config`).
    A key listed in `config` gets `default_config` updated by `config[key]`, the others get `default_config`.
    Also returns, as attribute `param_group_all.last_names`, the parameter names of each group (the reference logs them)."""
    default_config = dict(default_config or {})
    groups = OrderedDict((k, []) for k in _GROUP_KEYS)
    for k in _OPTIONAL_KEYS:
        if k in config:
            groups[k] = []
    names = OrderedDict((k, []) for k in groups)
    claimed = set()
    type2num = defaultdict(int)

    def put(key, mod_name, mod, attr, tag=""):
        n = _pname(mod_name, attr)
        groups[key].append(getattr(mod, attr))
        names[key].append(n)
        claimed.add(n)
        type2num[mod.__class__.__name__ + "." + attr + tag] += 1

    bn_types = (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)
    for name, m in model.named_modules():
        if isinstance(m, torch.nn.Conv2d):
            depthwise = m.groups == m.in_channels
            if m.bias is not None:
                if "bias" in groups:
                    put("bias", name, m, "bias")
                elif "conv_dw_b" in groups and depthwise:
                    put("conv_dw_b", name, m, "bias", "(dw)")
                elif "conv_dense_b" in groups and m.groups == 1:
                    put("conv_dense_b", name, m, "bias", "(dense)")
                else:
                    put("conv_b", name, m, "bias")
            if "conv_dw_w" in groups and depthwise:
                put("conv_dw_w", name, m, "weight", "(dw)")
            elif "conv_dense_w" in groups and m.groups == 1:
                put("conv_dense_w", name, m, "weight", "(dense)")
        elif isinstance(m, torch.nn.Linear):
            if m.bias is not None:
                put("bias" if "bias" in groups else "linear_b", name, m, "bias")
            if "linear_w" in groups:
                put("linear_w", name, m, "weight")
        elif isinstance(m, bn_types):
            if m.weight is not None:
                put("bn_w", name, m, "weight")
            if m.bias is not None:
                put("bias" if "bias" in groups else "bn_b", name, m, "bias")
        elif isinstance(m, torch.nn.LayerNorm):
            if m.weight is not None:
                put("ln_w", name, m, "weight")
            if m.bias is not None:
                put("bias" if "bias" in groups else "ln_b", name, m, "bias")
    normal, normal_names = [], []
    for name, p in model.named_parameters():
        if "logit_scale" in groups and "logit_scale" in name:
            groups["logit_scale"].append(p)
            names["logit_scale"].append(name)
            claimed.add(name)
        if name not in claimed:
            normal.append(p)
            normal_names.append(name)
    param_groups = [{"params": normal, **default_config}]
    for key, plist in groups.items():
        cfg = copy.deepcopy(default_config)
        if key in config:
            cfg.update(config[key])
        param_groups.append({"params": plist, **cfg})
    param_group_all.last_names = OrderedDict([("normal", normal_names)] + list(names.items()))
    return param_groups, type2num


class CosineLRScheduler:
    """Linear warm-up from `base_lr` to `warmup_lr` over `warmup_steps` iterations, then cosine decay to `min_lr` at
    `max_iter`; every param group keeps its own `initial_lr` scaled by the same factor."""

    def __init__(self, optimizer, max_iter, min_lr, base_lr, warmup_lr, warmup_steps, last_iter=0):
        if not isinstance(optimizer, torch.optim.Optimizer):
            raise TypeError("%s is not an Optimizer" % type(optimizer).__name__)
        assert warmup_steps >= 2 or warmup_steps == 0
        if warmup_steps == 0:
            assert base_lr == warmup_lr
        self.optimizer = optimizer
        self.max_iter, self.min_lr = max_iter, min_lr
        self.base_lr, self.warmup_lr, self.warmup_steps = base_lr, warmup_lr, warmup_steps
        if last_iter == 0:
            for g in optimizer.param_groups:
                g.setdefault("initial_lr", g["lr"])
        else:
            for i, g in enumerate(optimizer.param_groups):
                if "initial_lr" not in g:
                    raise KeyError("param 'initial_lr' is not specified in param_groups[%d] when resuming an optimizer" % i)
        self.base_lrs = [g["initial_lr"] for g in optimizer.param_groups]
        self.last_iter = last_iter

    def _scale(self):
        it = self.last_iter
        if self.warmup_steps >= 2 and it < self.warmup_steps:
            target = (self.warmup_lr - self.base_lr) / (self.warmup_steps - 1) * (it - 1) + self.base_lr
        else:
            ratio = (it - self.warmup_steps) / (self.max_iter - self.warmup_steps)
            target = self.min_lr + (self.warmup_lr - self.min_lr) * (1 + math.cos(math.pi * ratio)) / 2
        return target / self.base_lr

    def get_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def step(self, this_iter=None):
        self.last_iter = self.last_iter + 1 if this_iter is None else this_iter
        s = self._scale()
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = s * base


Cosine = CosineLRScheduler


def scheduler_entry(config):
    """config: {'type': 'Cosine' | 'CosineEpoch', 'kwargs': {...}} (attribute or item access)."""
    typ = config["type"] if isinstance(config, dict) else config.type
    kwargs = dict(config["kwargs"] if isinstance(config, dict) else config.kwargs)
    if typ == "CosineEpoch":
        ratio = kwargs["max_iter"] / kwargs.pop("max_epoch")
        if "warmup_epoch" in kwargs:
            kwargs["warmup_steps"] = max(round(kwargs.pop("warmup_epoch") * ratio), 2)
        typ = "Cosine"
    if typ != "Cosine":
        raise NotImplementedError("lr scheduler %r (the shipped experiments use Cosine only)" % typ)
    return CosineLRScheduler(**kwargs)


class LogitScaleClip:
    """`grad_clip` of the CLIP solvers for the parameter-clamp types: call `before()` ahead of backward / step and
    `after()` behind `optimizer.step()`.
        logit_scale_param_value   : clamp to [value, max_value] on both sides of the step (yfcc15m configs: [3, 6])
        logit_scale_param_abs_min : clamp to >= value on both sides
        logit_scale_param         : limit the change made by one step to +-value"""

    def __init__(self, logit_scale, type, value, max_value=None):   # noqa: A002 (the config key is called `type`)
        if type not in ("logit_scale_param_value", "logit_scale_param_abs_min", "logit_scale_param"):
            raise NotImplementedError("grad_clip.type %r" % type)
        self.p, self.type, self.value, self.max_value = logit_scale, type, value, max_value
        self._before = None

    def _clamp(self):
        if self.type == "logit_scale_param_value":
            self.p.data.clamp_(min=self.value, max=self.max_value)
        elif self.type == "logit_scale_param_abs_min":
            self.p.data.clamp_(min=self.value)

    def before(self):
        if self.type == "logit_scale_param":
            self._before = self.p.data.clone()
        else:
            self._clamp()

    def after(self):
        if self.type == "logit_scale_param":
            self.p.data.copy_(torch.minimum(torch.maximum(self.p.data, self._before - self.value), self._before + self.value))
        else:
            self._clamp()
