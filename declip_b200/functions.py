"""autograd bridges for the contrastive head (clip.py:129-141, loss.py:40-50) over the C ABI.

Data-parallel exchange (reference AllGather, clip.py:25-49) uses torch.distributed/NCCL:
forward all-gathers the bf16 L2-normalised features; backward reduce-scatters the gathered-side
feature gradients (the reference all-reduces the full [W,b,D] gradient and slices — same result,
W x fewer bytes kept).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops

_PTR = ctypes.c_void_p


def _stream():
    return _PTR(torch.cuda.current_stream().cuda_stream)


def dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def l2norm_fwd(x, eps, rows_pad=None):
    """y = x / (||x|| + eps) in bf16; optionally zero-padded to `rows_pad` rows (TMA needs 16-byte row strides
    on the gathered dimension, so odd batch sizes are padded internally)."""
    lib = ops.lib_for(x)
    n, d = x.shape
    if rows_pad is not None and rows_pad != n:
        y = torch.zeros(rows_pad, d, device=x.device, dtype=torch.bfloat16)
    else:
        y = torch.empty(n, d, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_l2norm_fwd(_PTR(x.data_ptr()), _PTR(y.data_ptr()), None, n, d, eps, _stream()), "dc_l2norm_fwd")
    return y


def l2norm_bwd(dy, x, eps):
    lib = ops.lib_for(x)
    n, d = x.shape
    dx = torch.empty_like(x)
    _lib.check(lib.dc_l2norm_bwd(_PTR(dy.data_ptr()), _PTR(x.data_ptr()), _PTR(dx.data_ptr()), n, d, eps, _stream()),
               "dc_l2norm_bwd")
    return dx


def cast_bf16(x):
    lib = ops.lib_for(x)
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_cast_f32_bf16(_PTR(x.data_ptr()), _PTR(y.data_ptr()), x.numel(), _stream()), "dc_cast_f32_bf16")
    return y


def dot_into(a, b, out):
    lib = ops.lib_for(a)
    _lib.check(lib.dc_dot_f32(_PTR(a.data_ptr()), _PTR(b.data_ptr()), a.numel(), _PTR(out.data_ptr()), _stream()),
               "dc_dot_f32")


class ClipLogits(torch.autograd.Function):
    """(logits_per_image, logits_per_text) = CLIP.forward after the encoders — clip.py:129-141.

    image/text features fp32 [b,E]; logit_scale the raw parameter [1].  s = min(exp(ls), 100) is used in the
    forward while d s / d ls = exp(ls) even when clamped (the reference clamps `.data`, clip.py:133-134)."""

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, gather, clamp):
        image_features = image_features.float().contiguous()
        text_features = text_features.float().contiguous()
        b, e = image_features.shape
        rank, world = dist_info()
        gather = bool(gather) and world > 1
        if gather and b % 8:
            raise RuntimeError("declip_b200: per-rank batch must be a multiple of 8 when features are all-gathered")
        bp = (b + 7) // 8 * 8
        i_n = l2norm_fwd(image_features, 0.0, bp)      # clip.py:129
        t_n = l2norm_fwd(text_features, 1e-10, bp)     # clip.py:130
        if gather:
            both = torch.cat([i_n, t_n], dim=1)                       # one collective for both towers
            allb = torch.empty(world * b, 2 * e, device=i_n.device, dtype=torch.bfloat16)
            dist.all_gather_into_tensor(allb, both)
            i_all, t_all = allb[:, :e], allb[:, e:]
        else:
            i_all, t_all = i_n, t_n
        n_pad = i_all.shape[0]
        n = world * b if gather else b
        s_raw = logit_scale.detach().float().exp().reshape(1)
        s_used = torch.clamp(s_raw, max=100.0) if clamp else s_raw           # clip.py:133-134
        # the scale stays on the device (alpha_dev): no host sync in the step
        li = ops.gemm(i_n[:b], t_all, epilogue=ops.EPI_F32, alpha_dev=s_used)   # [b,N] = s I_loc T_all^T  clip.py:140
        lt = ops.gemm(t_n[:b], i_all, epilogue=ops.EPI_F32, alpha_dev=s_used)   # [b,N] = s T_loc I_all^T  clip.py:141
        ctx.save_for_backward(image_features, text_features, i_n, t_n, i_all, t_all, li, lt, s_raw, s_used)
        ctx.gather, ctx.rank, ctx.world, ctx.n = gather, rank, world, n
        if n_pad != n:
            return li[:, :n], lt[:, :n]
        return li, lt

    @staticmethod
    def backward(ctx, dli, dlt):
        image_features, text_features, i_n, t_n, i_all, t_all, li, lt, s_raw, s_used = ctx.saved_tensors
        b, e = image_features.shape
        n, n_pad = ctx.n, li.shape[1]
        if n_pad != n:      # odd batch: zero-padded columns (rare path, plain copies)
            pad = torch.zeros(2, b, n_pad, device=li.device, dtype=torch.float32)
            pad[0, :, :n] = dli
            pad[1, :, :n] = dlt
            dli, dlt = pad[0], pad[1]
        else:
            dli = dli.contiguous().float()
            dlt = dlt.contiguous().float()
        dli16, dlt16 = cast_bf16(dli), cast_bf16(dlt)
        i_loc, t_loc = i_n[:b], t_n[:b]
        # local-side terms: dI_n = s * dli T_all ; dT_n = s * dlt I_all          (B read MN-major)
        di_n = ops.gemm(dli16, t_all, b_mn_major=True, epilogue=ops.EPI_F32, alpha_dev=s_used)
        dt_n = ops.gemm(dlt16, i_all, b_mn_major=True, epilogue=ops.EPI_F32, alpha_dev=s_used)
        # gathered-side terms: dT_all = s * dli^T I_loc ; dI_all = s * dlt^T T_loc   (A, B read MN-major)
        dt_all = ops.gemm(dli16, i_loc, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32, alpha_dev=s_used)
        di_all = ops.gemm(dlt16, t_loc, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32, alpha_dev=s_used)
        if ctx.gather:
            both = torch.cat([di_all, dt_all], dim=1)
            mine = torch.empty(b, 2 * e, device=both.device, dtype=torch.float32)
            dist.reduce_scatter_tensor(mine, both, op=dist.ReduceOp.SUM)       # == all_reduce + slice, clip.py:43-49
            di_n = di_n + mine[:, :e]
            dt_n = dt_n + mine[:, e:]
        else:
            di_n = di_n + di_all[:b]
            dt_n = dt_n + dt_all[:b]
        d_img = l2norm_bwd(di_n.contiguous(), image_features, 0.0)
        d_txt = l2norm_bwd(dt_n.contiguous(), text_features, 1e-10)
        # d logit_scale = exp(ls) * sum(dlogits * logits) / s_used
        acc = torch.zeros(1, device=li.device, dtype=torch.float32)
        dot_into(dli, li, acc)
        dot_into(dlt, lt, acc)
        dls = acc * (s_raw / s_used)
        return d_img, d_txt, dls.view(1), None, None


class ClipInfoCE(torch.autograd.Function):
    """ClipInfoCELoss.forward — loss.py:40-50 — fused with accuracy top-1/top-5 (misc.py:415-428)."""

    @staticmethod
    def forward(ctx, li, lt, label0, stats):
        lib = ops.lib_for(li)
        if li.stride(1) != 1:
            li = li.contiguous()
        if lt.stride(1) != 1:
            lt = lt.contiguous()
        b, n = li.shape
        acc = torch.zeros(2, device=li.device, dtype=torch.float32)
        cnt = torch.zeros(2, device=li.device, dtype=torch.int32)
        lse_i = torch.empty(b, device=li.device, dtype=torch.float32)
        lse_t = torch.empty(b, device=li.device, dtype=torch.float32)
        _lib.check(lib.dc_ce_strip_fwd(_PTR(li.data_ptr()), li.stride(0), b, n, label0, _PTR(acc.data_ptr()),
                                       _PTR(cnt.data_ptr()), _PTR(cnt.data_ptr() + 4), _PTR(lse_i.data_ptr()),
                                       _stream()), "dc_ce_strip_fwd")
        _lib.check(lib.dc_ce_strip_fwd(_PTR(lt.data_ptr()), lt.stride(0), b, n, label0, _PTR(acc.data_ptr() + 4), None,
                                       None, _PTR(lse_t.data_ptr()), _stream()), "dc_ce_strip_fwd")
        ctx.save_for_backward(li, lt, lse_i, lse_t)
        ctx.label0 = label0
        if stats is not None:
            stats["top1_count"], stats["top5_count"], stats["rows"] = cnt[0:1], cnt[1:2], b
        return (acc[0] + acc[1]) / (2.0 * b)

    @staticmethod
    def backward(ctx, g):
        li, lt, lse_i, lse_t = ctx.saved_tensors
        lib = ops.lib_for(li)
        b, n = li.shape
        g = g.contiguous().float()
        dli = torch.empty(b, n, device=li.device, dtype=torch.float32)
        dlt = torch.empty(b, n, device=li.device, dtype=torch.float32)
        for z, lse, d in ((li, lse_i, dli), (lt, lse_t, dlt)):
            _lib.check(lib.dc_ce_strip_bwd(_PTR(z.data_ptr()), z.stride(0), b, n, ctx.label0, _PTR(lse.data_ptr()),
                                           _PTR(g.data_ptr()), 1.0 / (2.0 * b), _PTR(d.data_ptr()), d.stride(0), 1,
                                           _stream()), "dc_ce_strip_bwd")
        return dli, dlt, None, None
