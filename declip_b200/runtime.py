"""Host-side runtime of one encoder tower: bf16 weight shadows, flat fp32 gradient buffer, pointer
tables and workspace for the C++ executors (csrc/encoder.cu), plus the autograd bridge.

PyTorch's role here is plumbing only — it owns device memory (caching allocator), the current
stream and autograd bookkeeping.  All arithmetic is in declip_b200/_C.so.
"""
import ctypes
import os
import threading

import torch

from . import _lib
from ._lib import CastEntry, TowerCfg

_PTR = ctypes.c_void_p


def _stream():
    return _PTR(torch.cuda.current_stream().cuda_stream)


def _ptr_array(ptrs):
    arr = (_PTR * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


LAYER_BF16 = ("attn.in_proj_weight", "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")
LAYER_F32 = ("ln_1.weight", "ln_1.bias", "attn.in_proj_bias", "attn.out_proj.bias", "ln_2.weight", "ln_2.bias",
             "mlp.c_fc.bias", "mlp.c_proj.bias")
LAYER_GRADS = ("attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "ln_1.weight",
               "ln_1.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias", "ln_2.weight",
               "ln_2.bias")
VIT_BF16 = ("conv1.weight", "proj")
VIT_F32 = ("class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias", "ln_post.weight", "ln_post.bias")
VIT_GRADS = ("class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias", "ln_post.weight", "ln_post.bias",
             "proj")
TEXT_BF16 = ("text_projection.weight",)
TEXT_F32 = ("token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection.bias")
TEXT_GRADS = ("token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias",
              "text_projection.weight", "text_projection.bias")


def register_shadow(p, shadow):
    """Attach a bf16 mirror (a tensor whose first p.numel() elements follow p's layout) to a parameter.  Whoever
    updates the fp32 master either bumps p._version (torch.optim, copy_, load_state_dict) — the owner then re-casts
    — or rewrites the mirror itself and records the version it corresponds to (FusedAdamW)."""
    p._dc_shadow = shadow
    p._dc_shadow_version = -1


def shadow_current(p, ptr=None):
    sh = getattr(p, "_dc_shadow", None)
    return sh is not None and (ptr is None or sh.data_ptr() == ptr) and p._dc_shadow_version == p._version


def weight_shadow(p, pad_rows=0):
    """bf16 mirror of a head / conv weight outside the two towers, cast only when the master changed (the towers
    keep theirs in one flat buffer, TowerRuntime._prepare).  pad_rows > 0: the mirror has that many leading rows
    (zero tail) for GEMM operands that need a 16-byte-aligned or padded extent."""
    sh = getattr(p, "_dc_shadow", None)
    rows = max(pad_rows, p.shape[0]) if p.dim() > 1 else p.shape[0]
    shape = (rows,) + tuple(p.shape[1:])
    if sh is None or sh.device != p.device or tuple(sh.shape) != shape:
        sh = torch.zeros(shape, device=p.device, dtype=torch.bfloat16)
        register_shadow(p, sh)
    if p._dc_shadow_version != p._version:
        lib = _lib.init(p.device.index if p.device.index is not None else torch.cuda.current_device())
        src = p.detach()
        if not src.is_contiguous() or src.dtype != torch.float32:
            src = src.float().contiguous()
        _lib.check(lib.dc_cast_f32_bf16(_PTR(src.data_ptr()), _PTR(sh.data_ptr()), src.numel(), _stream()),
                   "dc_cast_f32_bf16")
        p._dc_shadow_version = p._version
    return sh


# ---------------------------------------------------------------------------------------------------------------------
# Two-stream tower execution.  The image and the text tower are independent in the forward and in the backward; each
# is a chain of persistent kernels (one CTA per SM) whose last wave leaves most SMs idle (300 cluster tiles on 74
# clusters = 4.05 waves for every N = 768 output).  Issued on two streams, the CTAs of one tower's next kernel occupy
# the SMs the other tower's tail wave leaves free — late-starting CTAs are the high-numbered ones, which own one tile
# less, so the static tile schedules balance by themselves.
#   Inside `with concurrent_towers():` the FIRST tower runtime that runs is applied under a side stream (it forks
#   before the main stream has any tower work queued); the region's exit joins.  Autograd runs a node's backward on
#   the stream its forward was applied on and orders it after the event recorded when its incoming gradient was
#   produced (the head's backward) — so the same tower runs on the side stream in the backward too, concurrently with
#   the other tower on the main stream, and its multi-GB workspace is allocated, used and freed on ONE stream (a
#   cross-stream `record_stream` on it made the caching allocator unable to reuse the block in time: intermittent
#   40 ms steps while it grew the pool with cudaMalloc).  Parameter gradients are written by raw kernels, not by
#   autograd, so an end-of-backward callback joins the side stream into the stream `backward()` was called on.
TOWER_STREAMS = os.environ.get("DECLIP_B200_TOWER_STREAMS", "1") != "0"
_tls = threading.local()
_side_streams = {}
_side_lock = threading.Lock()


def _side_stream(dev):
    dev = torch.device(dev)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    with _side_lock:
        if key not in _side_streams:
            _side_streams[key] = torch.cuda.Stream(device=dev)
        return _side_streams[key]


class concurrent_towers:
    """Context manager around the encoder calls of one model forward (see above)."""

    def __enter__(self):
        self.prev = getattr(_tls, "region", None)
        _tls.region = self if TOWER_STREAMS else None
        self.side_rt, self.outputs, self.stream, self.dev, self.main = None, [], None, None, None
        return self

    def __exit__(self, *exc):
        _tls.region = self.prev
        if self.stream is not None:
            self.main.wait_stream(self.stream)
            for t in self.outputs:              # small tensors allocated on the side stream's pool, consumed on the main stream
                t.record_stream(self.main)
        return False

    def stream_for(self, rt, dev):
        """side stream for the first runtime seen in this region (and its later passes), None (= current) otherwise."""
        if self.side_rt is None:
            self.side_rt, self.dev = rt, dev
            self.stream = _side_stream(dev)
            self.main = torch.cuda.current_stream(dev)
            self.stream.wait_stream(self.main)      # inputs produced on the main stream
        return self.stream if rt is self.side_rt else None


def _join_side_after_backward(side, main):
    """Queue the join of the side stream into `main` (and into whatever stream is current when the engine finishes).
    One callback per tower node: a handful per step, and a repeated wait on an already-joined stream is free."""
    def finish():
        main.wait_stream(side)
        cur = torch.cuda.current_stream(side.device)
        if cur != main and cur != side:
            cur.wait_stream(side)
    torch.autograd.Variable._execution_engine.queue_callback(finish)


class TowerRuntime:
    """Per-module state for dc_{vit,text}_{forward,backward}.  `kind` is 'vit' or 'text'."""

    def __init__(self, kind, module, layers, width, heads, seq_len, embed_dim, res=0, patch=0, vocab=0):
        self.kind = kind
        self.module = module
        self.layers, self.width, self.heads, self.seq_len, self.embed_dim = layers, width, heads, seq_len, embed_dim
        self.res, self.patch, self.vocab = res, patch, vocab
        pre = "transformer.resblocks.%d."
        extra_bf16, extra_f32, extra_grads = (VIT_BF16, VIT_F32, VIT_GRADS) if kind == "vit" else (TEXT_BF16, TEXT_F32,
                                                                                                    TEXT_GRADS)
        self.bf16_names = [pre % l + n for l in range(layers) for n in LAYER_BF16] + list(extra_bf16)
        self.f32_names = [pre % l + n for l in range(layers) for n in LAYER_F32] + list(extra_f32)
        self.grad_names = [pre % l + n for l in range(layers) for n in LAYER_GRADS] + list(extra_grads)
        self._key = None          # (device, tuple of param data_ptrs)
        self._versions = None
        self.grad_flat = None
        self._outstanding = 0     # forwards (under autograd) whose backward has not run yet
        self.grad_ready_hook = None   # called (rt) when every outstanding backward of this tower has been enqueued
        self.progress_hook = None     # (ctypes callback, every): dc_set_backward_progress_cb during the LAST outstanding backward

    # ------------------------------------------------------------------ parameter plumbing
    def _params(self):
        return dict(self.module.named_parameters())

    def _prepare(self, params):
        dev = params[self.f32_names[0]].device
        key = (dev, tuple(params[n].data_ptr() for n in self.bf16_names + self.f32_names))
        if key == self._key:
            return
        for n in self.bf16_names + self.f32_names:
            p = params[n]
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
                raise RuntimeError("declip_b200: parameter %s must be a contiguous fp32 CUDA tensor (master weights "
                                   "stay fp32; bf16 shadows are internal)" % n)
        self.lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
        # bf16 shadows of the GEMM weights, one flat buffer
        offs, total = [], 0
        for n in self.bf16_names:
            offs.append(total)
            total += (params[n].numel() + 127) // 128 * 128
        self.shadow = torch.empty(total, device=dev, dtype=torch.bfloat16)
        entries = (CastEntry * len(self.bf16_names))()
        max_numel = 0
        for i, n in enumerate(self.bf16_names):
            entries[i].src = params[n].data_ptr()
            entries[i].dst = self.shadow.data_ptr() + 2 * offs[i]
            entries[i].numel = params[n].numel()
            max_numel = max(max_numel, params[n].numel())
        raw = bytes(entries)
        self.cast_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self.cast_n, self.cast_max = len(self.bf16_names), max_numel
        self.w_bf16 = _ptr_array([self.shadow.data_ptr() + 2 * o for o in offs])
        self._shadow_ptrs = [self.shadow.data_ptr() + 2 * o for o in offs]
        for n, o in zip(self.bf16_names, offs):     # FusedAdamW rewrites these slices in its own pass (optim.py)
            register_shadow(params[n], self.shadow[o:o + params[n].numel()])
        self.w_f32 = _ptr_array([params[n].data_ptr() for n in self.f32_names])
        # flat fp32 gradient buffer in the executor's table order
        goffs, gtotal = [], 0
        for n in self.grad_names:
            goffs.append(gtotal)
            gtotal += (params[n].numel() + 63) // 64 * 64
        self.grad_flat = torch.zeros(gtotal, device=dev, dtype=torch.float32)
        self.grad_offs = goffs
        self.layer_grad_end = goffs[12 * self.layers]       # the per-layer gradients come first, then the tower's own
        self.grad_ptrs = _ptr_array([self.grad_flat.data_ptr() + 4 * o for o in goffs])
        self._key = key
        self._versions = None

    def refresh_shadows(self, params):
        """fp32 master -> bf16 shadow for every GEMM weight (one launch); skipped when nothing changed."""
        ps = [params[n] for n in self.bf16_names]
        if all(shadow_current(p, ptr) for p, ptr in zip(ps, self._shadow_ptrs)):
            return
        _lib.check(self.lib.dc_multi_cast_f32_bf16(_PTR(self.cast_table.data_ptr()), self.cast_n, self.cast_max, _stream()),
                   "dc_multi_cast_f32_bf16")
        for p, ptr, n in zip(ps, self._shadow_ptrs, self.bf16_names):
            if getattr(p, "_dc_shadow", None) is None or p._dc_shadow.data_ptr() != ptr:
                off = (ptr - self.shadow.data_ptr()) // 2
                register_shadow(p, self.shadow[off:off + p.numel()])
            p._dc_shadow_version = p._version

    def cfg(self, batch):
        c = TowerCfg()
        c.layers, c.width, c.heads, c.seq_len = self.layers, self.width, self.heads, self.seq_len
        c.causal = 1 if self.kind == "text" else 0
        c.batch, c.embed_dim, c.res, c.patch, c.vocab = batch, self.embed_dim, self.res, self.patch, self.vocab
        return c

    def workspace(self, cfg, dev):
        """A fresh workspace per forward (torch's caching allocator makes this cheap): a module may be run
        several times before backward (DeCLIP encodes two image views, declip.py:231-232)."""
        nbytes = self.lib.dc_tower_workspace_bytes(ctypes.byref(cfg))
        if nbytes == 0:
            raise RuntimeError("declip_b200: bad tower config: " + _lib.last_error())
        return torch.empty(nbytes, device=dev, dtype=torch.uint8)

    # ------------------------------------------------------------------ executors
    def pre_features(self, cfg, ws):
        """bf16 [batch, width]: the pre-projection feature of the forward that filled `ws` (dc_tower_pre_features)."""
        out = torch.empty(cfg.batch, self.width, device=ws.device, dtype=torch.bfloat16)
        _lib.check(self.lib.dc_tower_pre_features(ctypes.byref(cfg), _PTR(ws.data_ptr()), _PTR(out.data_ptr()), _stream()),
                   "dc_tower_pre_features")
        return out

    def forward(self, inp, params, dense=False):
        self._prepare(params)
        self.refresh_shadows(params)
        batch = inp.shape[0]
        cfg = self.cfg(batch)
        ws = self.workspace(cfg, inp.device)
        feats = torch.empty(batch, self.embed_dim, device=inp.device, dtype=torch.float32)
        if self.kind == "vit":
            if inp.dtype != torch.float32 or inp.stride(3) != 1 or inp.stride(2) != self.res or inp.stride(1) != self.res ** 2:
                inp = inp.float().contiguous()
            dense_out = None
            if dense:   # patch tokens of the last block, x[:, 1:, :] (visual_transformer.py:68)
                dense_out = torch.empty(batch * (self.seq_len - 1), self.width, device=inp.device, dtype=torch.bfloat16)
            _lib.check(self.lib.dc_vit_forward(ctypes.byref(cfg), _PTR(inp.data_ptr()), inp.stride(0), self.w_bf16,
                                               self.w_f32, _PTR(ws.data_ptr()), _PTR(feats.data_ptr()),
                                               _PTR(dense_out.data_ptr()) if dense else None, _stream()),
                       "dc_vit_forward")
            return feats, inp, cfg, ws, dense_out
        else:
            if inp.dtype != torch.int64 or not inp.is_contiguous():
                inp = inp.long().contiguous()
            words = None
            if dense:   # ln_final of every token (words_feat, text_transformer.py:194-201)
                words = torch.empty(batch * self.seq_len, self.width, device=inp.device, dtype=torch.bfloat16)
            _lib.check(self.lib.dc_text_forward(ctypes.byref(cfg), _PTR(inp.data_ptr()), self.w_bf16, self.w_f32,
                                                _PTR(ws.data_ptr()), _PTR(feats.data_ptr()),
                                                _PTR(words.data_ptr()) if dense else None, _stream()),
                       "dc_text_forward")
            return feats, inp, cfg, ws, words
        return feats, inp, cfg, ws, None

    def backward(self, cfg, inp, ws, dfeats, params, dense=False, dwords=None, dpre=None):
        """Accumulates parameter gradients straight into `p.grad` (views of one flat fp32 buffer per tower, the
        DDP/"main_grad" pattern), so the gradient all-reduce is a single NCCL call per tower and a module that is
        run several times per step accumulates correctly.  Ownership rules per parameter:
          p.grad is None            -> zero its slice, attach the slice as p.grad
          p.grad is already a slice -> accumulate in place (also covers zero_grad(set_to_none=False))
          p.grad is a foreign tensor-> compute into a private buffer and add."""
        views = getattr(self, "_views", None)
        if views is None or self._views_key is not self.grad_flat:
            views = [self.grad_flat[o:o + params[n].numel()].view_as(params[n])
                     for n, o in zip(self.grad_names, self.grad_offs)]
            self._views, self._views_key = views, self.grad_flat
        foreign = []
        all_fresh = True
        for n, v in zip(self.grad_names, views):
            p = params[n]
            if not p.requires_grad:
                continue
            g = p.grad
            if g is None:
                continue
            all_fresh = False
            if g.data_ptr() != v.data_ptr() or g.dtype != torch.float32:
                foreign.append(n)
        ptrs = self.grad_ptrs
        tmp = None
        if foreign:
            tmp = torch.zeros_like(self.grad_flat)
            ptrs = _ptr_array([tmp.data_ptr() + 4 * o for o in self.grad_offs])
        elif all_fresh:
            self.grad_flat.zero_()
        else:
            for n, v in zip(self.grad_names, views):
                if params[n].requires_grad and params[n].grad is None:
                    v.zero_()
        if dfeats is None:
            dfeats = torch.zeros(cfg.batch, self.embed_dim, device=ws.device, dtype=torch.float32)
        dfeats = dfeats.float().contiguous()
        if dwords is not None:
            dwords = dwords.to(torch.bfloat16).contiguous()
        self.writing_flat = tmp is None         # this pass accumulates straight into grad_flat (the slices are / become p.grad)
        hook = self.progress_hook if (self.progress_hook is not None and self._outstanding == 1 and not foreign) else None
        if hook is not None:     # layer-group progress reports -> gradient buckets go on the wire during the backward
            self.lib.dc_set_backward_progress_cb(ctypes.cast(hook[0], _PTR), None, hook[1])
        try:
            self._run_backward(cfg, inp, dfeats, dense, dwords, dpre, ptrs, ws)
        finally:
            if hook is not None:
                self.lib.dc_set_backward_progress_cb(None, None, 1)
        self._finish_backward(params, views, tmp)

    def _run_backward(self, cfg, inp, dfeats, dense, dwords, dpre, ptrs, ws):
        if self.kind == "vit":
            if dpre is not None:
                dpre = dpre.to(torch.bfloat16).contiguous()
            _lib.check(self.lib.dc_vit_backward_pre(ctypes.byref(cfg), _PTR(dfeats.data_ptr()),
                                                    _PTR(dpre.data_ptr()) if dpre is not None else None,
                                                    _PTR(dwords.data_ptr()) if dwords is not None else None, self.w_bf16,
                                                    self.w_f32, ptrs, _PTR(ws.data_ptr()), _stream()), "dc_vit_backward")
        else:
            _lib.check(self.lib.dc_text_backward(ctypes.byref(cfg), _PTR(inp.data_ptr()), _PTR(dfeats.data_ptr()),
                                                 int(dense), _PTR(dwords.data_ptr()) if dwords is not None else None,
                                                 self.w_bf16, self.w_f32, ptrs, _PTR(ws.data_ptr()), _stream()),
                       "dc_text_backward")

    def _finish_backward(self, params, views, tmp):
        if tmp is not None:
            for n, o in zip(self.grad_names, self.grad_offs):
                p = params[n]
                if not p.requires_grad:
                    continue
                t = tmp[o:o + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = t.clone()
                else:
                    p.grad.add_(t)
        else:
            for n, v in zip(self.grad_names, views):
                p = params[n]
                if p.requires_grad and p.grad is None:
                    p.grad = v


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _forward_on_stream(rt, inp, params, dense, pre):
    """rt.forward on the stream the enclosing concurrent_towers() region assigns (the current stream outside one)."""
    region = getattr(_tls, "region", None)
    stream = region.stream_for(rt, inp.device) if region is not None and inp.is_cuda else None
    with (torch.cuda.stream(stream) if stream is not None else _NullCtx()):
        if stream is not None:
            inp.record_stream(stream)
        feats, inp_used, cfg, ws, words = rt.forward(inp, params, dense)
        outs = [feats]
        if dense:
            outs.append(words)
        if pre:
            outs.append(rt.pre_features(cfg, ws).float())
    if stream is not None:
        region.outputs.extend(outs)
    return outs, (inp_used, cfg, ws)


class _TowerFunction(torch.autograd.Function):
    """features [, words] = tower(inp; params).  One C-ABI call forward, one backward.  Parameter gradients are
    written into p.grad by the runtime (see TowerRuntime.backward); autograd only carries d(features), d(words)."""

    @staticmethod
    def forward(ctx, rt, inp, anchor, dense, pre=False):
        params = rt._params()
        feats, inp_used, cfg, ws, words = rt.forward(inp, params, dense)      # on the stream run_tower applied us under
        outs = [feats]
        if dense:
            outs.append(words)
        if pre:
            outs.append(rt.pre_features(cfg, ws).float())
        rt._outstanding += 1
        ctx.rt, ctx.cfg, ctx.inp, ctx.ws, ctx.params, ctx.dense, ctx.pre = rt, cfg, inp_used, ws, params, dense, pre
        ctx.streams = getattr(_tls, "apply_streams", None)        # (side, main) when applied under the side stream
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, dfeats, *rest):
        rt = ctx.rt
        rest = list(rest)
        dwords = rest.pop(0) if ctx.dense else None
        dpre = rest.pop(0) if ctx.pre else None
        if ctx.streams is not None:     # the engine runs this node on the side stream it was applied under
            side, main = ctx.streams
            for t in (dfeats, dwords, dpre):
                if t is not None:
                    t.record_stream(side)           # small gradients produced on the main stream
            _join_side_after_backward(side, main)
        rt.backward(ctx.cfg, ctx.inp, ctx.ws, dfeats, ctx.params, ctx.dense, dwords, dpre)
        ctx.ws = None
        rt._outstanding = max(0, rt._outstanding - 1)
        if rt._outstanding == 0 and rt.grad_ready_hook is not None:
            rt.grad_ready_hook(rt)      # e.g. DistModule: start this tower's gradient all-reduce now, overlapped
        return None, None, None, None, None


def run_tower(rt, inp, dense=False, pre=False):
    """Run the tower through autograd (training) or directly (no_grad / eval).  `anchor` is any trainable
    parameter: it makes the output require grad so backward is invoked.  dense=True also returns ln_final of every
    token as bf16 [B*L, D]; pre=True also returns the pre-projection feature fp32 [B, width] (last)."""
    params = rt._params()
    if torch.is_grad_enabled():
        anchor = next((p for p in params.values() if p.requires_grad), None)
        if anchor is not None:
            region = getattr(_tls, "region", None)
            stream = region.stream_for(rt, inp.device) if region is not None and inp.is_cuda else None
            if stream is None:
                _tls.apply_streams = None
                return _TowerFunction.apply(rt, inp, anchor, dense, pre)
            inp.record_stream(stream)
            _tls.apply_streams = (stream, region.main)
            try:
                with torch.cuda.stream(stream):
                    out = _TowerFunction.apply(rt, inp, anchor, dense, pre)
            finally:
                _tls.apply_streams = None
            region.outputs.extend(out if isinstance(out, tuple) else (out,))
            return out
    outs, _ = _forward_on_stream(rt, inp, params, dense, pre)
    return outs[0] if len(outs) == 1 else tuple(outs)
