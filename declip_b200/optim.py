"""FusedAdamW — drop-in for `torch.optim.AdamW` (the optimiser of every reference experiment config, built by
prototype/optimizer/__init__.py:18-26 `optim_entry`) backed by ONE multi-tensor CUDA launch per step
(`dc_adamw_multi`).  Same constructor arguments, param-group semantics (per-group lr / weight_decay as produced by
prototype/utils/misc.py:267-412 `param_group_all`), `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq`; `step`
may be an int or — as torch >= 1.12 checkpoints store it — a tensor).

The kernel also rewrites the bf16 shadow (`runtime.register_shadow`) of every GEMM weight from the updated fp32
master in the same pass, and the step bumps each parameter's autograd version counter exactly as an in-place
`torch.optim` update would, so nothing downstream can keep reading a stale copy."""
import ctypes

import torch

from . import _lib
from ._lib import AdamWEntry

MAX_GROUPS = 32


def _bump_versions(params):
    """In-place updates through raw pointers are invisible to autograd's version counters; bump them by hand."""
    torch._C._autograd._unsafe_set_version_counter(params, [p._version + 1 for p in params])


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("declip_b200: amsgrad is not used by the reference configs")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > MAX_GROUPS:
            raise NotImplementedError("FusedAdamW: at most %d param groups" % MAX_GROUPS)
        self._tables = {}    # step value -> (signature, device table, n, max_numel)
        self._gbuf = {}      # small parameters whose autograd gradient is a fresh tensor every step (logit_scale)
        self._last_ptr = {}
        self._staging = None

    # ------------------------------------------------------------------ device pointer table
    @staticmethod
    def _shadow_ptr(p):
        sh = getattr(p, "_dc_shadow", None)
        return sh.data_ptr() if sh is not None else 0

    def _grad_ptr(self, p):
        g = self._gbuf.get(p)
        return g.data_ptr() if g is not None else p.grad.data_ptr()

    def _build(self, plist):
        n = len(plist)
        entries = (AdamWEntry * n)()
        mx = 0
        for i, (p, gi) in enumerate(plist):
            st = self.state[p]
            entries[i].param = p.data_ptr()
            entries[i].grad = self._grad_ptr(p)
            entries[i].exp_avg = st["exp_avg"].data_ptr()
            entries[i].exp_avg_sq = st["exp_avg_sq"].data_ptr()
            entries[i].shadow = self._shadow_ptr(p) or None
            entries[i].numel = p.numel()
            entries[i].group = gi
            mx = max(mx, p.numel())
        dev = plist[0][0].device
        raw = bytes(entries)
        # staged through pinned memory: the (rare) rebuild is an async copy on the current stream, not a pageable one
        if self._staging is None or self._staging.numel() < len(raw):
            self._staging = torch.empty(max(len(raw), 1 << 16), dtype=torch.uint8).pin_memory()
        self._staging[:len(raw)].copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        table = torch.empty(len(raw), dtype=torch.uint8, device=dev)
        table.copy_(self._staging[:len(raw)], non_blocking=True)
        torch.cuda.current_stream().synchronize()      # staging buffer is reused by the next rebuild
        return table, n, mx

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

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        cfgs = {(tuple(g["betas"]), float(g["eps"])) for g in self.param_groups}
        if len(cfgs) > 1:
            raise NotImplementedError("FusedAdamW: all param groups must share betas / eps (they do in every config)")
        (betas, eps), = cfgs
        by_step = {}          # parameters that skipped iterations (grad None) carry their own step count
        for gi, group in enumerate(self.param_groups):
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
                step = int(st["step"]) + 1            # int(): torch.optim.AdamW checkpoints hold a tensor step
                st["step"] = step
                by_step.setdefault(step, []).append((p, gi))
        if not by_step:
            return loss
        lr = (ctypes.c_float * MAX_GROUPS)(*[float(g["lr"]) for g in self.param_groups])
        wd = (ctypes.c_float * MAX_GROUPS)(*[float(g["weight_decay"]) for g in self.param_groups])
        for key in [k for k in self._tables if k not in by_step and k + 1 not in by_step]:
            del self._tables[key]
        for step, plist in by_step.items():
            self._stabilise_small_grads(plist)
            sig = tuple((p.data_ptr(), self._grad_ptr(p), self.state[p]["exp_avg"].data_ptr(), self._shadow_ptr(p), gi)
                        for p, gi in plist)
            cached = self._tables.pop(step - 1, None) or self._tables.get(step)
            if cached is None or cached[0] != sig:      # pointers changed -> rebuild the device table (lr / wd are not in it)
                cached = (sig,) + self._build(plist)
            self._tables[step] = cached
            _, table, n, mx = cached
            dev = plist[0][0].device
            lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
            _lib.check(lib.dc_adamw_multi(ctypes.c_void_p(table.data_ptr()), n, mx, lr, wd, len(self.param_groups),
                                          betas[0], betas[1], eps, step,
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dc_adamw_multi")
            plain = [p for p, _ in plist]
            _bump_versions(plain)
            for p in plain:
                if getattr(p, "_dc_shadow", None) is not None:
                    p._dc_shadow_version = p._version      # the kernel just rewrote it from the new master
        return loss
