"""FusedAdamW — drop-in for `torch.optim.AdamW` (the optimiser of every reference experiment config, built by
prototype/optimizer/__init__.py:18-26 `optim_entry`) backed by ONE multi-tensor CUDA launch per step
(`dc_adamw_multi`).  Same constructor arguments, param-group semantics (per-group lr / weight_decay as produced by
prototype/utils/misc.py:267-412 `param_group_all`), `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq`)."""
import ctypes

import torch

from . import _lib
from ._lib import AdamWEntry


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("declip_b200: amsgrad is not used by the reference configs")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._table = None
        self._sig = None
        self._gbuf = {}      # small parameters whose autograd gradient is a fresh tensor every step (logit_scale)
        self._last_ptr = {}

    def _build(self, plist):
        n = len(plist)
        entries = (AdamWEntry * n)()
        mx = 0
        for i, (p, group) in enumerate(plist):
            st = self.state[p]
            entries[i].param = p.data_ptr()
            entries[i].grad = self._grad_ptr(p)
            entries[i].exp_avg = st["exp_avg"].data_ptr()
            entries[i].exp_avg_sq = st["exp_avg_sq"].data_ptr()
            entries[i].numel = p.numel()
            entries[i].lr = float(group["lr"])
            entries[i].weight_decay = float(group["weight_decay"])
            mx = max(mx, p.numel())
        dev = plist[0][0].device
        self._table = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8).to(dev)
        self._n, self._max = n, mx

    def _grad_ptr(self, p):
        g = self._gbuf.get(p)
        return g.data_ptr() if g is not None else p.grad.data_ptr()

    def _stabilise_small_grads(self, plist):
        """Gradients produced by autograd (not the towers' flat buffers) live in a new tensor each step; copying the
        small ones into persistent buffers keeps the device pointer table valid from step to step."""
        for p, _ in plist:
            cur = p.grad.data_ptr()
            if p in self._gbuf:
                self._gbuf[p].copy_(p.grad)                      # known to move every step (e.g. logit_scale)
            elif p.numel() <= 4096 and self._last_ptr.get(p, cur) != cur:
                self._gbuf[p] = p.grad.detach().clone()          # moved since the last step: give it a stable home
            self._last_ptr[p] = cur

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        by_cfg = {}
        for group in self.param_groups:
            key = (tuple(group["betas"]), float(group["eps"]))
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_cfg.setdefault(key, []).append((p, group))
        if len(by_cfg) > 1:
            raise NotImplementedError("FusedAdamW: all param groups must share betas / eps (they do in every config)")
        for (betas, eps), plist in by_cfg.items():
            steps = {self.state[p]["step"] for p, _ in plist}
            if len(steps) != 1:
                raise RuntimeError("FusedAdamW: parameters with different step counts")
            self._stabilise_small_grads(plist)
            sig = tuple((p.data_ptr(), self._grad_ptr(p), self.state[p]["exp_avg"].data_ptr(), float(g["lr"]),
                         float(g["weight_decay"])) for p, g in plist)
            if sig != self._sig:          # pointers, lr (scheduler) or wd changed -> rebuild the device table
                self._build(plist)
                self._sig = sig
            dev = plist[0][0].device
            lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
            _lib.check(lib.dc_adamw_multi(ctypes.c_void_p(self._table.data_ptr()), self._n, self._max, betas[0], betas[1],
                                          eps, steps.pop(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "dc_adamw_multi")
        return loss
