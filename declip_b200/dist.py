"""Data-parallel wrapper — replaces prototype/utils/dist.py (DistModule) and the linklink shim
(linklink/__init__.py) with torch.distributed + NCCL.

Gradient sync.  The towers accumulate their parameter gradients into ONE flat fp32 buffer each
(runtime.TowerRuntime.grad_flat, laid out layer by layer).  The reference all-reduces every parameter
separately from a grad-accumulator hook (dist.py:63-74: 302 small fp32 collectives, loss pre-divided by world
size so SUM == mean, clip_solver.py:418).  Here the C++ backward executor reports progress every
`bucket_layers` transformer layers (dc_set_backward_progress_cb); each report turns the finished slice of the
flat buffer into one bucket that is, on a side stream,
    cast fp32 -> bf16 (dc_cast_f32_bf16)  ->  NCCL all-reduce (SUM)  ->  cast back into the fp32 .grad slice,
overlapping the backward of the remaining layers and of the other tower.  bf16 on the wire halves the bytes
(302 MB instead of 605 MB per step for CLIP ViT-B/32); the fp32 master gradients stay local.  Only the last
bucket of the tower that finishes last (its first layers + embeddings) is exposed.

Optional (`nccl_ctas` > 0, off by default): a dedicated communicator capped to that many CTAs plus
`dc_set_sm_reserve(nccl_ctas)`, so that the persistent GEMM / attention kernels leave exactly the SMs NCCL occupies.
Measured on 2 x B200 (profiles/r02_scaling_variants_n2.md): it LOSES — reserving 8 SMs for the whole backward costs
more (30.7 ms/step) than the occasional extra wave it avoids (29.5 without, same buckets); fewer, larger buckets win
(6 layers per bucket: 28.6 ms; 3 layers: 29.5; one per tower: 28.8; N = 1: 27.4-28.0).
"""
import ctypes
import os

import torch
import torch.distributed as dist
from torch.nn import Module


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def barrier():
    if get_world_size() > 1:
        dist.barrier()


_PROGRESS_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int)


class DistModule(Module):
    def __init__(self, module, sync=False, overlap=True, bucket_layers=None, grad_dtype=None, nccl_ctas=None):
        super().__init__()
        self.module = module
        self.sync = sync
        self.bucket_layers = int(os.environ.get("DECLIP_B200_BUCKET_LAYERS", "6")) if bucket_layers is None else bucket_layers
        grad_dtype = grad_dtype or os.environ.get("DECLIP_B200_GRAD_DTYPE", "bf16")
        if grad_dtype not in ("bf16", "fp32"):
            raise ValueError("grad_dtype must be 'bf16' or 'fp32'")
        self.grad_bf16 = grad_dtype == "bf16"
        self.nccl_ctas = int(os.environ.get("DECLIP_B200_NCCL_CTAS", "0")) if nccl_ctas is None else nccl_ctas
        self._pending = []        # (rt id, covered parameter ids, side-stream event) of buckets launched during backward
        self._sent = {}           # id(rt) -> lowest layer whose gradients are already on the wire
        self._side = None
        self._group = None
        self._stage = {}          # id(rt) -> bf16 staging buffer (one per tower, same layout as grad_flat)
        self._reserved = False
        self.broadcast_params()
        overlap = overlap and os.environ.get("DECLIP_B200_OVERLAP", "1") != "0"      # A/B switch (tools/gpu_round2x_n2.sh)
        if overlap and get_world_size() > 1:
            for rt in self._runtimes():
                rt.grad_ready_hook = self._on_tower_grads_ready
                if self.bucket_layers > 0:
                    rt.progress_hook = (self._make_progress_cb(rt), self.bucket_layers)

    # ------------------------------------------------------------------ plumbing
    def _grad_group(self):
        """A communicator of its own for gradient buckets, with NCCL's CTA count capped (one CTA per channel): the
        persistent compute kernels give up exactly that many SMs while buckets are in flight."""
        if self._group is None:
            self._group = dist.group.WORLD
            if dist.get_backend() == "nccl" and self.nccl_ctas > 0:
                try:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = self.nccl_ctas
                    opts.config.min_ctas = min(self.nccl_ctas, 4)
                    self._group = dist.new_group(backend="nccl", pg_options=opts)
                except Exception:     # noqa: BLE001 - older torch without NCCL config options: use the world group
                    self._group = dist.group.WORLD
        return self._group

    def _reserve_sms(self, on):
        if on == self._reserved or self.nccl_ctas <= 0:
            return
        params = next(self.module.parameters(), None)
        if params is None or not params.is_cuda:
            return
        from . import _lib
        _lib.load().dc_set_sm_reserve(self.nccl_ctas if on else 0)
        self._reserved = on

    def _reduce_slice(self, rt, pieces):
        """pieces: list of (lo, hi) element ranges of rt.grad_flat that are final in stream order.  Runs on the side
        stream: [fp32 -> bf16] -> all-reduce -> [bf16 -> fp32]; returns the completion event."""
        from . import _lib
        flat = rt.grad_flat
        if self._side is None:
            self._side = torch.cuda.Stream(device=flat.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(flat.device))
        group = self._grad_group()
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            if not self.grad_bf16:
                for lo, hi in pieces:
                    dist.all_reduce(flat[lo:hi], group=group)
            else:
                lib = rt.lib
                stage = self._stage.get(id(rt))
                if stage is None or stage.numel() != flat.numel():
                    stage = self._stage[id(rt)] = torch.empty(flat.numel(), device=flat.device, dtype=torch.bfloat16)
                st = ctypes.c_void_p(self._side.cuda_stream)
                # pack the pieces back to back so that one collective carries them all
                off, spans = 0, []
                for lo, hi in pieces:
                    n = hi - lo
                    _lib.check(lib.dc_cast_f32_bf16(ctypes.c_void_p(flat.data_ptr() + 4 * lo), ctypes.c_void_p(stage.data_ptr() + 2 * off),
                                                    n, st), "dc_cast_f32_bf16")
                    spans.append((lo, off, n))
                    off += (n + 7) // 8 * 8
                dist.all_reduce(stage[:off], group=group)
                for lo, so, n in spans:
                    _lib.check(lib.dc_cast_bf16_f32(ctypes.c_void_p(stage.data_ptr() + 2 * so), ctypes.c_void_p(flat.data_ptr() + 4 * lo),
                                                    n, 1.0, 0, st), "dc_cast_bf16_f32")
            done = torch.cuda.Event()
            done.record(self._side)
        return done

    def _make_progress_cb(self, rt):
        def cb(_user, layer):
            try:
                self._on_layers_done(rt, layer)
            except Exception as e:      # noqa: BLE001 - never unwind through the C frame
                self._cb_error = e
        return _PROGRESS_CB(cb)

    def _on_layers_done(self, rt, layer):
        """Called (on the launching thread, between kernel launches) when the gradients of transformer layers >= `layer`
        are complete in stream order."""
        if rt.grad_flat is None or not getattr(rt, "writing_flat", False):
            return
        hi_layer = self._sent.get(id(rt), rt.layers)
        if layer >= hi_layer:
            return
        lo, hi = rt.grad_offs[12 * layer], rt.grad_offs[12 * hi_layer]
        self._reserve_sms(True)
        done = self._reduce_slice(rt, [(lo, hi)])
        self._sent[id(rt)] = layer
        self._pending.append((id(rt), done))

    def _on_tower_grads_ready(self, rt):
        """Called from the tower's backward as soon as its last kernel is enqueued: reduce what has not been sent yet —
        the first layers and the embeddings / projection — overlapping the backward of the other tower (the reference
        overlaps per parameter through grad-accumulator hooks, dist.py:63-74)."""
        if rt.grad_flat is None or not self._aliased(rt):
            return
        hi_layer = self._sent.pop(id(rt), rt.layers)
        pieces = []
        if hi_layer > 0:
            pieces.append((0, rt.grad_offs[12 * hi_layer]))
        if rt.grad_flat.numel() > rt.layer_grad_end:
            pieces.append((rt.layer_grad_end, rt.grad_flat.numel()))
        if len(pieces) == 2 and pieces[0][1] == pieces[1][0]:
            pieces = [(pieces[0][0], pieces[1][1])]
        self._reserve_sms(True)
        done = self._reduce_slice(rt, pieces)
        self._pending.append((id(rt), done))
        self._covered_rts = getattr(self, "_covered_rts", set()) | {id(rt)}

    @staticmethod
    def _aliased(rt):
        params = rt._params()
        return all((not params[n].requires_grad) or (params[n].grad is not None and
                   params[n].grad.data_ptr() == rt.grad_flat.data_ptr() + 4 * o)
                   for n, o in zip(rt.grad_names, rt.grad_offs))

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def _runtimes(self):
        return [m._rt for m in self.module.modules() if hasattr(m, "_rt")]

    def sync_gradients(self):
        """dist.py:76-83 equivalent: SUM-all-reduce every gradient, then make the result visible to the
        optimizer stream (no device-wide synchronize)."""
        if get_world_size() == 1:
            return
        if os.environ.get("DECLIP_B200_SKIP_GRAD_SYNC") == "1":      # measurement only: the step WITHOUT its all-reduce
            return
        err = getattr(self, "_cb_error", None)
        if err is not None:
            self._cb_error = None
            raise err
        covered_rts = getattr(self, "_covered_rts", set())
        self._covered_rts = set()
        covered = set()
        cur = torch.cuda.current_stream() if torch.cuda.is_available() else None
        for rt in self._runtimes():
            if rt.grad_flat is None:
                continue
            params = rt._params()
            if id(rt) in covered_rts:                         # reduced in buckets on the side stream during backward
                covered.update(id(params[n]) for n in rt.grad_names)
            elif self._aliased(rt):
                self._sent.pop(id(rt), None)
                self._pending.append((id(rt), self._reduce_slice(rt, [(0, rt.grad_flat.numel())])))
                covered.update(id(params[n]) for n in rt.grad_names)
        for _, done in self._pending:
            cur.wait_event(done)
        self._pending = []
        self._sent = {}
        self._reserve_sms(False)
        rest = [p.grad for p in self.module.parameters() if p.grad is not None and id(p) not in covered]
        if rest:
            # heads outside the towers (DeCLIP projector / predictor / MLM head, FILIP mappings, logit_scale): large
            # gradients are reduced in place (no staging copy), the many small ones travel as one flat tensor
            big = [g for g in rest if g.numel() >= (1 << 20) and g.is_contiguous()]
            small = [g for g in rest if not (g.numel() >= (1 << 20) and g.is_contiguous())]
            for g in big:
                dist.all_reduce(g)
            if small:
                flat = torch.cat([g.reshape(-1).float() for g in small])
                dist.all_reduce(flat)
                off = 0
                for g in small:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()

    def broadcast_params(self):
        """dist.py:85-88."""
        if get_world_size() == 1:
            return
        for _, p in self.module.state_dict().items():
            dist.broadcast(p, 0)
