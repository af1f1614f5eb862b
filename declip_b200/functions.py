"""autograd bridges for the contrastive head (clip.py:129-141, loss.py:40-50) over the C ABI.

Data-parallel exchange (reference AllGather, clip.py:25-49) uses torch.distributed/NCCL:
forward all-gathers the bf16 L2-normalised features; backward reduce-scatters the gathered-side
feature gradients (the reference all-reduces the full [W,b,D] gradient and slices — same result,
W x fewer bytes kept).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib, ops
from .runtime import weight_shadow

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
    _lib.check(lib.dc_l2norm_fwd(_PTR(x.data_ptr()), _PTR(y.data_ptr()), None, None, n, d, eps, _stream()),
               "dc_l2norm_fwd")
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
        _lib.check(lib.dc_ce_strip_fwd(_PTR(li.data_ptr()), li.stride(0), b, n, label0, None, None, _PTR(acc.data_ptr()),
                                       _PTR(cnt.data_ptr()), _PTR(cnt.data_ptr() + 4), _PTR(lse_i.data_ptr()),
                                       _stream()), "dc_ce_strip_fwd")
        _lib.check(lib.dc_ce_strip_fwd(_PTR(lt.data_ptr()), lt.stride(0), b, n, label0, None, None, _PTR(acc.data_ptr() + 4), None,
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
            _lib.check(lib.dc_ce_strip_bwd(_PTR(z.data_ptr()), z.stride(0), b, n, ctx.label0, None, None, _PTR(lse.data_ptr()),
                                           _PTR(g.data_ptr()), 1.0 / (2.0 * b), _PTR(d.data_ptr()), d.stride(0), 1,
                                           _stream()), "dc_ce_strip_bwd")
        return dli, dlt, None, None


# =====================================================================================================================
# DeCLIP / FILIP building blocks (declip.py, loss.py, nnclr_modules) — each a thin autograd bridge over the C ABI.
# =====================================================================================================================
class L2Normalize(torch.autograd.Function):
    """y = x / (||x|| + eps) as a real fp32 tensor (clip.py:129-130; DeCLIP returns these in ret_dict['features'])."""

    @staticmethod
    def forward(ctx, x, eps):
        lib = ops.lib_for(x)
        x = x.float().contiguous()
        n, d = x.shape
        y = torch.empty_like(x)
        _lib.check(lib.dc_l2norm_fwd(_PTR(x.data_ptr()), None, _PTR(y.data_ptr()), None, n, d, float(eps), _stream()),
                   "dc_l2norm_fwd")
        ctx.save_for_backward(x)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return l2norm_bwd(dy.float().contiguous(), x, ctx.eps), None


def _to_bf16_rows(x, rows_pad):
    """fp32 [n,d] -> bf16 [rows_pad,d] (zero tail): operands of the strip GEMMs."""
    n, d = x.shape
    y = cast_bf16(x)
    if rows_pad != n:
        z = torch.zeros(rows_pad, d, device=x.device, dtype=torch.bfloat16)
        z[:n] = y
        return z
    return y


class StripLogits(torch.autograd.Function):
    """Generalised logit strips over ALREADY-NORMALISED features: for each (i, j) in `pairs`,
        strip = s * F_i(local rows) @ F_j(all ranks)^T          [b, N]
    with one all-gather for all features (declip.py:264-269) and s = min(exp(logit_scale), 100) when `clamp`
    (gradient flows as exp(logit_scale), clip.py:133-134) or the constant `scale_const` when logit_scale is None
    (NT-Xent temperature).  Covers CLIP (2 strips), DeCLIP (8 + 4 NN strips, declip.py:271-300), NTXentLoss."""

    @staticmethod
    def forward(ctx, logit_scale, scale_const, gather, clamp, pairs, *feats):
        nf = len(feats)
        feats = [f.float().contiguous() for f in feats]
        b, e = feats[0].shape
        rank, world = dist_info()
        gather = bool(gather) and world > 1
        if gather and b % 8:
            raise RuntimeError("declip_b200: per-rank batch must be a multiple of 8 when features are all-gathered")
        bp = (b + 7) // 8 * 8
        loc = torch.cat([_to_bf16_rows(f, bp) for f in feats], dim=1)          # [bp, nf*E] bf16
        if gather:
            allf = torch.empty(world * b, nf * e, device=loc.device, dtype=torch.bfloat16)
            dist.all_gather_into_tensor(allf, loc)
        else:
            allf = loc
        n = world * b if gather else b
        if logit_scale is not None:
            s_raw = logit_scale.detach().float().exp().reshape(1)
            s_used = torch.clamp(s_raw, max=100.0) if clamp else s_raw
        else:
            s_raw = s_used = torch.full((1,), float(scale_const), device=loc.device, dtype=torch.float32)
        strips = []
        for (i, j) in pairs:
            strips.append(ops.gemm(loc[:b, i * e:(i + 1) * e], allf[:, j * e:(j + 1) * e], epilogue=ops.EPI_F32,
                                   alpha_dev=s_used))
        ctx.save_for_backward(loc, allf, s_raw, s_used, *strips)
        ctx.meta = (nf, b, e, n, gather, pairs, logit_scale is not None, None)
        n_pad = allf.shape[0]
        return tuple(s[:, :n] if n_pad != n else s for s in strips)

    @staticmethod
    def backward(ctx, *dstrips):
        loc, allf, s_raw, s_used = ctx.saved_tensors[:4]
        strips = ctx.saved_tensors[4:]
        nf, b, e, n, gather, pairs, has_ls, _ = ctx.meta
        n_pad = allf.shape[0]
        need = ctx.needs_input_grad[5:]
        dloc = torch.zeros(nf, b, e, device=loc.device, dtype=torch.float32)
        dall = torch.zeros(nf, n_pad, e, device=loc.device, dtype=torch.float32)
        acc = torch.zeros(1, device=loc.device, dtype=torch.float32)
        for k, (i, j) in enumerate(pairs):
            ds = dstrips[k]
            if ds is None:
                continue
            if n_pad != n:
                pad = torch.zeros(b, n_pad, device=loc.device, dtype=torch.float32)
                pad[:, :n] = ds
                ds = pad
            else:
                ds = ds.contiguous().float()
            ds16 = cast_bf16(ds)
            if need[i]:    # dF_i(local) += s * dS F_j(all)                      (B read MN-major)
                ops.gemm(ds16, allf[:, j * e:(j + 1) * e], b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=dloc[i],
                         alpha_dev=s_used)
            if need[j]:    # dF_j(all) += s * dS^T F_i(local)                   (A, B read MN-major)
                ops.gemm(ds16, loc[:b, i * e:(i + 1) * e], a_mn_major=True, b_mn_major=True,
                         epilogue=ops.EPI_F32_ATOMIC, out=dall[j], alpha_dev=s_used)
            if has_ls:
                dot_into(ds, strips[k], acc)
        if gather:
            mine = torch.empty(nf, b, e, device=loc.device, dtype=torch.float32)
            # reduce-scatter over ranks of every feature's gathered-side gradient == all_reduce + slice (clip.py:43-49)
            send = dall.view(nf, -1, b, e).transpose(0, 1).contiguous()        # [W, nf, b, e]
            dist.reduce_scatter_tensor(mine.view(-1), send.view(-1), op=dist.ReduceOp.SUM)
            dloc = dloc + mine
        else:
            dloc = dloc + dall[:, :b]
        grads = [dloc[i] if need[i] else None for i in range(nf)]
        dls = (acc * (s_raw / s_used)).view(1) if has_ls else None
        return (dls, None, None, None, None) + tuple(grads)


def _split_bf16(x):
    """x (fp32) = hi + lo with hi, lo bf16: three bf16 tensor-core GEMMs (hi*hi + hi*lo + lo*hi) then reproduce an
    fp32 product to ~1e-5 relative — used for the tiny fp32 heads where BatchNorm over a small batch amplifies
    rounding noise."""
    x = x.float().contiguous()
    hi = cast_bf16(x)
    lo = cast_bf16(x - hi.float())
    return hi, lo


def _gemm3(a, b, **kw):
    """sum of the three significant partial products of split operands a = (hi, lo), b = (hi, lo)."""
    out = ops.gemm(a[0], b[0], epilogue=ops.EPI_F32, **kw)
    kw.pop("bias", None)
    ops.gemm(a[0], b[1], epilogue=ops.EPI_F32_ATOMIC, out=out, **kw)
    ops.gemm(a[1], b[0], epilogue=ops.EPI_F32_ATOMIC, out=out, **kw)
    return out


class LinearF32(torch.autograd.Function):
    """y = x W^T + b for the small fp32 heads (SimSiam projector/predictor, declip.py:48-60,107-112): split-bf16
    tensor-core GEMMs with fp32 I/O (near-fp32 accuracy).  in/out features must be multiples of 8."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xs = _split_bf16(x)
        ws = _split_bf16(weight)
        y = _gemm3(xs, ws, bias=bias)
        ctx.save_for_backward(*xs, *ws)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xh, xl, wh, wl = ctx.saved_tensors
        dy = dy.float().contiguous()
        dys = _split_bf16(dy)
        dx = _gemm3(dys, (wh, wl), b_mn_major=True) if ctx.needs_input_grad[0] else None
        dw = _gemm3(dys, (xh, xl), a_mn_major=True, b_mn_major=True)
        db = dy.sum(0) if ctx.has_bias else None
        return dx, dw, db


class BatchNorm1dF(torch.autograd.Function):
    """nn.BatchNorm1d (+ optional fused ReLU) — declip.py:49-60,108-110.  Batch statistics are PER RANK, as in the
    reference (nn.BatchNorm1d, not SyncBN)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, relu, eps, momentum):
        lib = ops.lib_for(x)
        x = x.float().contiguous()
        rows, c = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(c, device=x.device, dtype=torch.float32)
        rstd = torch.empty(c, device=x.device, dtype=torch.float32)
        _lib.check(lib.dc_batchnorm_fwd(_PTR(x.data_ptr()), _PTR(gamma.data_ptr()), _PTR(beta.data_ptr()),
                                        _PTR(y.data_ptr()), _PTR(mean.data_ptr()), _PTR(rstd.data_ptr()),
                                        _PTR(running_mean.data_ptr()), _PTR(running_var.data_ptr()), rows, c, float(eps),
                                        float(momentum), int(training), int(relu), _stream()), "dc_batchnorm_fwd")
        ctx.save_for_backward(x, y, gamma, mean, rstd)
        ctx.training, ctx.relu = bool(training), bool(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, rstd = ctx.saved_tensors
        lib = ops.lib_for(x)
        rows, c = x.shape
        dy = dy.float().contiguous()
        dx = torch.empty_like(x)
        dg = torch.zeros(c, device=x.device, dtype=torch.float32)
        db = torch.zeros(c, device=x.device, dtype=torch.float32)
        _lib.check(lib.dc_batchnorm_bwd(_PTR(dy.data_ptr()), _PTR(x.data_ptr()), _PTR(y.data_ptr()), _PTR(gamma.data_ptr()),
                                        _PTR(mean.data_ptr()), _PTR(rstd.data_ptr()), _PTR(dx.data_ptr()),
                                        _PTR(dg.data_ptr()), _PTR(db.data_ptr()), rows, c, int(ctx.training), int(ctx.relu),
                                        _stream()), "dc_batchnorm_bwd")
        return dx, dg, db, None, None, None, None, None, None


class CosineMean(torch.autograd.Function):
    """D(p, z) = mean_r cos(p_r, stopgrad(z_r)) — loss_functions/loss.py:52-58."""

    @staticmethod
    def forward(ctx, p, z):
        lib = ops.lib_for(p)
        p = p.float().contiguous()
        z = z.detach().float().contiguous()
        n, d = p.shape
        cosv = torch.empty(n, device=p.device, dtype=torch.float32)
        _lib.check(lib.dc_cosine_rows_fwd(_PTR(p.data_ptr()), _PTR(z.data_ptr()), _PTR(cosv.data_ptr()), n, d, _stream()),
                   "dc_cosine_rows_fwd")
        ctx.save_for_backward(p, z)
        return cosv.mean()

    @staticmethod
    def backward(ctx, g):
        p, z = ctx.saved_tensors
        lib = ops.lib_for(p)
        n, d = p.shape
        grow = (g.float() / n).expand(n).contiguous()
        dp = torch.empty_like(p)
        _lib.check(lib.dc_cosine_rows_bwd(_PTR(p.data_ptr()), _PTR(z.data_ptr()), _PTR(grow.data_ptr()), 1.0,
                                          _PTR(dp.data_ptr()), n, d, _stream()), "dc_cosine_rows_bwd")
        return dp, None


class RowCE(torch.autograd.Function):
    """mean_r CE(logits[r, :cols], labels[r]) with an int64 label array (MLM head, declip.py:330-333)."""

    @staticmethod
    def forward(ctx, logits, labels, cols):
        lib = ops.lib_for(logits)
        n = logits.shape[0]
        acc = torch.zeros(1, device=logits.device, dtype=torch.float32)
        lse = torch.empty(n, device=logits.device, dtype=torch.float32)
        _lib.check(lib.dc_ce_strip_fwd(_PTR(logits.data_ptr()), logits.stride(0), n, cols, 0, _PTR(labels.data_ptr()),
                                       None, _PTR(acc.data_ptr()), None, None, _PTR(lse.data_ptr()), _stream()),
                   "dc_ce_strip_fwd")
        ctx.save_for_backward(logits, labels, lse)
        ctx.cols = cols
        return acc[0] / n

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse = ctx.saved_tensors
        lib = ops.lib_for(logits)
        n = logits.shape[0]
        g = g.contiguous().float().reshape(1)
        d = torch.zeros_like(logits)
        _lib.check(lib.dc_ce_strip_bwd(_PTR(logits.data_ptr()), logits.stride(0), n, ctx.cols, 0, _PTR(labels.data_ptr()),
                                       None, _PTR(lse.data_ptr()), _PTR(g.data_ptr()), 1.0 / n, _PTR(d.data_ptr()), d.stride(0), 1,
                                       _stream()), "dc_ce_strip_bwd")
        return d, None, None


class MaskedLMHead(torch.autograd.Function):
    """text_label_predictor on the masked positions only, + cross-entropy — declip.py:326-334.  The reference runs the
    512 -> 49409 Linear on all B*77 tokens and then boolean-masks (7.8 GB of logits at b=512); gathering the masked
    rows first is mathematically identical.  words bf16 [B*L, D]; rows int32 [n] (masked token rows); labels int64 [n]."""

    @staticmethod
    def forward(ctx, words, rows, labels, weight, bias):
        lib = ops.lib_for(words)
        n = rows.numel()
        v, d = weight.shape
        vp = (v + 7) // 8 * 8
        if n == 0:
            raise RuntimeError("declip_b200: MLM head called with no masked token in the batch (declip.py:326-334 "
                               "would return nan)")
        w16 = weight_shadow(weight, pad_rows=vp)       # [vp, d] bf16, re-cast only when the master changed
        bpad = torch.zeros(vp, device=words.device, dtype=torch.float32)
        bpad[:v] = bias
        x = torch.empty(n, d, device=words.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_gather_rows(_PTR(words.data_ptr()), _PTR(rows.data_ptr()), _PTR(x.data_ptr()), n, d, _stream()),
                   "dc_gather_rows")
        logits = ops.gemm(x, w16, bias=bpad, epilogue=ops.EPI_F32)                       # [n, vp] fp32
        acc = torch.zeros(1, device=words.device, dtype=torch.float32)
        lse = torch.empty(n, device=words.device, dtype=torch.float32)
        _lib.check(lib.dc_ce_strip_fwd(_PTR(logits.data_ptr()), logits.stride(0), n, v, 0, _PTR(labels.data_ptr()),
                                       None, _PTR(acc.data_ptr()), None, None, _PTR(lse.data_ptr()), _stream()),
                   "dc_ce_strip_fwd")
        ctx.save_for_backward(x, rows, labels, w16, logits, lse)
        ctx.shape = (words.shape[0], v, vp, d)
        return acc[0] / n

    @staticmethod
    def backward(ctx, g):
        x, rows, labels, w16, logits, lse = ctx.saved_tensors
        lib = ops.lib_for(x)
        m, v, vp, d = ctx.shape
        n = x.shape[0]
        g = g.contiguous().float().reshape(1)
        dl = torch.zeros(n, vp, device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_ce_strip_bwd(_PTR(logits.data_ptr()), logits.stride(0), n, v, 0, _PTR(labels.data_ptr()),
                                       None, _PTR(lse.data_ptr()), _PTR(g.data_ptr()), 1.0 / n, _PTR(dl.data_ptr()), dl.stride(0),
                                       0, _stream()), "dc_ce_strip_bwd")
        dw = ops.gemm(dl, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC)[:v]     # [v, d]
        db = ops.colsum(dl)[:v]
        dx = ops.gemm(dl, w16, b_mn_major=True, epilogue=ops.EPI_BF16)                             # [n, d] bf16
        dwords = torch.zeros(m, d, device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_scatter_rows(_PTR(dx.data_ptr()), _PTR(rows.data_ptr()), _PTR(dwords.data_ptr()), n, d, _stream()),
                   "dc_scatter_rows")
        return dwords, None, None, dw, db


def nn_lookup(query, bank, bank16):
    """Top-1 nearest neighbour of each query row in the memory bank (cosine similarity) —
    nnclr_modules/nn_memory_bank.py:54-64.  query fp32 [b,d]; bank fp32 [size,d] (raw rows, returned un-normalised as in
    the reference); bank16 bf16 [size,d] = F.normalize(bank).  No gradient (the reference detaches)."""
    lib = ops.lib_for(query)
    query = query.detach().float().contiguous()
    b, d = query.shape
    q16 = torch.empty(b, d, device=query.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_l2norm_fwd(_PTR(query.data_ptr()), _PTR(q16.data_ptr()), None, None, b, d, 1e-12, _stream()),
               "dc_l2norm_fwd")
    sim = ops.gemm(q16, bank16, epilogue=ops.EPI_F32)                                    # [b, size]
    idx = torch.empty(b, device=query.device, dtype=torch.int32)
    _lib.check(lib.dc_argmax_rows(_PTR(sim.data_ptr()), sim.stride(0), b, sim.shape[1], _PTR(idx.data_ptr()), _stream()),
               "dc_argmax_rows")
    out = torch.empty(b, d, device=query.device, dtype=torch.float32)
    _lib.check(lib.dc_gather_rows_f32(_PTR(bank.data_ptr()), _PTR(idx.data_ptr()), _PTR(out.data_ptr()), b, d, _stream()),
               "dc_gather_rows_f32")
    return out, idx


def normalize_rows_bf16(x, eps=1e-12):
    """bf16 F.normalize(x, dim=1) of fp32 rows (memory-bank shadow)."""
    lib = ops.lib_for(x)
    x = x.float().contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_l2norm_fwd(_PTR(x.data_ptr()), _PTR(y.data_ptr()), None, None, x.shape[0], x.shape[1], eps, _stream()),
               "dc_l2norm_fwd")
    return y


# =====================================================================================================================
# FILIP (filip.py:71-142)
# =====================================================================================================================
class LinearBF16In(torch.autograd.Function):
    """y(fp32) = x(bf16) W^T + b on dense token features (image_mapping / text_mapping, filip.py:40-41,133-134)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w16 = weight_shadow(weight)
        x = x.contiguous()
        y = ops.gemm(x, w16, bias=bias, epilogue=ops.EPI_F32)
        ctx.save_for_backward(x, w16)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w16 = ctx.saved_tensors
        dy16 = cast_bf16(dy.float().contiguous())
        db = ops.colsum(dy16)
        dx = ops.gemm(dy16, w16, b_mn_major=True, epilogue=ops.EPI_BF16)
        dw = ops.gemm(dy16, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC)
        return dx, dw, db


class GatherRowsF32(torch.autograd.Function):
    """out[i] = x[idx[i]] (fp32 rows, distinct idx) — the top-k token selection (filip.py:83-88)."""

    @staticmethod
    def forward(ctx, x, idx):
        lib = ops.lib_for(x)
        x = x.contiguous()
        n, d = idx.numel(), x.shape[1]
        out = torch.empty(n, d, device=x.device, dtype=torch.float32)
        _lib.check(lib.dc_gather_rows_f32(_PTR(x.data_ptr()), _PTR(idx.data_ptr()), _PTR(out.data_ptr()), n, d, _stream()),
                   "dc_gather_rows_f32")
        ctx.save_for_backward(idx)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        lib = ops.lib_for(dout)
        dout = dout.float().contiguous()
        dx = torch.zeros(ctx.rows, dout.shape[1], device=dout.device, dtype=torch.float32)
        _lib.check(lib.dc_add_rows_f32(_PTR(dout.data_ptr()), _PTR(idx.data_ptr()), _PTR(dx.data_ptr()), idx.numel(),
                                       dout.shape[1], _stream()), "dc_add_rows_f32")
        return dx, None


class AllGatherRows(torch.autograd.Function):
    """CLIP.all_gather (clip.py:25-49,113-116) on torch.distributed/NCCL: forward all-gather along dim 0, backward
    reduce-scatter (== the reference's all-reduce + slice)."""

    @staticmethod
    def forward(ctx, x):
        rank, world = dist_info()
        ctx.world = world
        if world == 1:
            return x
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x)
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.world == 1:
            return g
        g = g.contiguous()
        out = torch.empty((g.shape[0] // ctx.world,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
        return out


class FilipLate(torch.autograd.Function):
    """logits[i, l] = mean_j max_m  s * <d[i, j, :], sel[l, m, :]>   (filip.py:93-104).
    d fp32 [B*n, dim] (normalised dense tokens of this rank), sel fp32 [N*group, dim] (gathered selected tokens),
    logit_scale_dense the raw parameter (s = exp, not clamped, filip.py:75)."""

    @staticmethod
    def forward(ctx, d, sel, logit_scale_dense, n, group):
        lib = ops.lib_for(d)
        d16 = cast_bf16(d.float().contiguous())
        sel16 = cast_bf16(sel.float().contiguous())
        batch = d.shape[0] // n
        ncand = sel.shape[0] // group
        s = logit_scale_dense.detach().float().exp().reshape(1)
        out = torch.empty(batch, ncand, device=d.device, dtype=torch.float32)
        if group == 16:
            # fused: the GEMM epilogue reduces every group of 16 score columns to (max, arg-max) straight out of TMEM —
            # the [B*n, N*16] score matrix (6.6 + 10.3 GB fp32 at N = 4096) is never written
            mx, arg = ops.gemm(d16, sel16, epilogue=ops.EPI_F32_GROUPMAX16, alpha_dev=s)      # [B*n, N] fp32 / uint8
            _lib.check(lib.dc_groupmax_mean_fwd(_PTR(mx.data_ptr()), mx.stride(0), batch, n, 1, ncand, _PTR(out.data_ptr()),
                                                out.stride(0), None, _stream()), "dc_groupmax_mean_fwd")
        else:
            G = ops.gemm(d16, sel16, epilogue=ops.EPI_F32, alpha_dev=s)                      # [B*n, N*group]
            arg = torch.empty(batch * n, ncand, device=d.device, dtype=torch.uint8)
            _lib.check(lib.dc_groupmax_mean_fwd(_PTR(G.data_ptr()), G.stride(0), batch, n, group, ncand, _PTR(out.data_ptr()),
                                                out.stride(0), _PTR(arg.data_ptr()), _stream()), "dc_groupmax_mean_fwd")
        ctx.save_for_backward(d16, sel16, arg, s, out)
        ctx.meta = (batch, n, group, ncand)
        return out

    @staticmethod
    def backward(ctx, dout):
        d16, sel16, arg, s, out = ctx.saved_tensors
        lib = ops.lib_for(d16)
        batch, n, group, ncand = ctx.meta
        dout = dout.float().contiguous()
        # The backward GEMMs contract over the one-hot operand dG[r, l * group + m] = [m == arg[r, l]] * dout[i(r), l] / n.
        # It is built per BLOCK of candidates (<= 512 MiB of bf16, DECLIP_B200_FILIP_CHUNK overrides the block size):
        # dd accumulates over the blocks, dsel's row blocks are independent — at N = 4096 gathered candidates the
        # full [B*n, N*group] operand would be 3.3 GB per rank, the block is 0.4 GB at every world size.
        rows, dim = batch * n, d16.shape[1]
        cb = int(os.environ.get("DECLIP_B200_FILIP_CHUNK", "0")) or max(64, (512 << 20) // (rows * group * 2) // 64 * 64)
        cb = min(ncand, max(8, cb // 8 * 8))
        dG = torch.empty(rows, cb * group, device=d16.device, dtype=torch.bfloat16)
        dd = None
        dsel = torch.empty(ncand * group, dim, device=d16.device, dtype=torch.float32)
        for c0 in range(0, ncand, cb):
            nc = min(cb, ncand - c0)
            blk = dG if nc == cb else dG.view(-1)[:rows * nc * group].view(rows, nc * group)
            _lib.check(lib.dc_groupmax_mean_bwd_ex(_PTR(dout.data_ptr() + 4 * c0), dout.stride(0), _PTR(arg.data_ptr() + c0),
                                                   arg.stride(0), batch, n, group, nc, _PTR(blk.data_ptr()), blk.stride(0),
                                                   _stream()), "dc_groupmax_mean_bwd")
            sel_blk = sel16[c0 * group:(c0 + nc) * group]
            part = ops.gemm(blk, sel_blk, b_mn_major=True, epilogue=ops.EPI_F32, alpha_dev=s)            # [B*n, dim]
            dd = part if dd is None else dd.add_(part)
            ops.gemm(blk, d16, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32, alpha_dev=s,
                     out=dsel[c0 * group:(c0 + nc) * group])                                          # [nc*group, dim]
        # d logit_scale_dense = sum(dout * out): out is linear in s = exp(ls), so d out / d ls = out
        acc = torch.zeros(1, device=d16.device, dtype=torch.float32)
        dot_into(dout, out, acc)
        return dd, dsel, acc.reshape(()), None, None


def token_scores(d1, d2, batch, n1, n2):
    lib = ops.lib_for(d1)
    dim = d1.shape[-1]
    s1 = torch.empty(batch, n1, device=d1.device, dtype=torch.float32)
    s2 = torch.empty(batch, n2, device=d1.device, dtype=torch.float32)
    _lib.check(lib.dc_token_scores(_PTR(d1.data_ptr()), _PTR(d2.data_ptr()), batch, n1, n2, dim, _PTR(s1.data_ptr()),
                                   _PTR(s2.data_ptr()), _stream()), "dc_token_scores")
    return s1, s2


# =====================================================================================================================
# NT-Xent family (loss_functions/nt_xent.py)
# =====================================================================================================================
class MatmulNT(torch.autograd.Function):
    """s * A B^T (fp32 [R,E] x fp32 [C,E] -> fp32 [R,C]) on the tensor-core GEMM, gradients to both operands."""

    @staticmethod
    def forward(ctx, a, b, scale):
        a16 = _to_bf16_rows(a.float().contiguous(), a.shape[0])
        cp = (b.shape[0] + 7) // 8 * 8
        b16 = _to_bf16_rows(b.float().contiguous(), cp)                 # columns padded to a multiple of 8
        out = ops.gemm(a16, b16, epilogue=ops.EPI_F32, alpha=float(scale))
        ctx.save_for_backward(a16, b16)
        ctx.scale, ctx.cols = float(scale), b.shape[0]
        return out[:, :b.shape[0]] if cp != b.shape[0] else out

    @staticmethod
    def backward(ctx, ds):
        a16, b16 = ctx.saved_tensors
        cp = b16.shape[0]
        if cp != ctx.cols:
            pad = torch.zeros(ds.shape[0], cp, device=ds.device, dtype=torch.float32)
            pad[:, :ctx.cols] = ds
            ds = pad
        ds16 = cast_bf16(ds.contiguous().float())
        da = ops.gemm(ds16, b16, b_mn_major=True, epilogue=ops.EPI_F32, alpha=ctx.scale)
        db = ops.gemm(ds16, a16, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32, alpha=ctx.scale)[:ctx.cols]
        return da, db, None


class MaskedRowCE(torch.autograd.Function):
    """sum_r CE(logits[r, all columns except skip[r]], labels[r]) / denom  — NT_Xent's [positive | negatives] CE."""

    @staticmethod
    def forward(ctx, logits, labels, skip, denom):
        lib = ops.lib_for(logits)
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        r, c = logits.shape
        acc = torch.zeros(1, device=logits.device, dtype=torch.float32)
        lse = torch.empty(r, device=logits.device, dtype=torch.float32)
        _lib.check(lib.dc_ce_strip_fwd(_PTR(logits.data_ptr()), logits.stride(0), r, c, 0, _PTR(labels.data_ptr()),
                                       _PTR(skip.data_ptr()), _PTR(acc.data_ptr()), None, None, _PTR(lse.data_ptr()),
                                       _stream()), "dc_ce_strip_fwd")
        ctx.save_for_backward(logits, labels, skip, lse)
        ctx.denom = float(denom)
        return acc[0] / denom

    @staticmethod
    def backward(ctx, g):
        logits, labels, skip, lse = ctx.saved_tensors
        lib = ops.lib_for(logits)
        r, c = logits.shape
        g = g.contiguous().float().reshape(1)
        d = torch.empty(r, c, device=logits.device, dtype=torch.float32)
        _lib.check(lib.dc_ce_strip_bwd(_PTR(logits.data_ptr()), logits.stride(0), r, c, 0, _PTR(labels.data_ptr()),
                                       _PTR(skip.data_ptr()), _PTR(lse.data_ptr()), _PTR(g.data_ptr()), 1.0 / ctx.denom,
                                       _PTR(d.data_ptr()), d.stride(0), 1, _stream()), "dc_ce_strip_bwd")
        return d, None, None, None


# =====================================================================================================================
# Fused distributed contrastive head (csrc/head.cu): normalise -> [gather] -> strips + ClipInfoCELoss + accuracy, and the
# whole backward, in three launches; no [b, N] strip in HBM.  clip.py:129-146, loss.py:40-50, misc.py:415-428.
# =====================================================================================================================
class HeadLayout:
    """Float offsets into the head workspace (dc_head_layout)."""

    _cache = {}

    def __init__(self, lib, b, e):
        arr = (ctypes.c_longlong * 8)()
        _lib.check(lib.dc_head_layout(b, e, arr), "dc_head_layout")
        self.out, self.lse, self.g, self.alab, self.ea, self.cnt, self.dxn, self.total = list(arr)

    @classmethod
    def get(cls, lib, b, e):
        key = (b, e)
        if key not in cls._cache:
            cls._cache[key] = cls(lib, b, e)
        return cls._cache[key]


def head_args(b, n, e, ld, row0, x_base, y_srcs, ws, x_off=(0, None), y_off=(None, 0), cross=True, strips=None):
    """dc_head_args for the symmetric pair: direction 0 = image rows x text columns, direction 1 = the transpose; the image
    feature sits at columns [0, e) of a row, the text feature at [e, 2e) unless offsets are given."""
    a = _lib.HeadArgs()
    a.b, a.n, a.e, a.ld, a.n_src, a.row0 = b, n, e, ld, len(y_srcs), row0
    xo = (x_off[0], e if x_off[1] is None else x_off[1])
    yo = (e if y_off[0] is None else y_off[0], y_off[1])
    a.x_off[0], a.x_off[1], a.y_off[0], a.y_off[1] = xo[0], xo[1], yo[0], yo[1]
    a.cross = 1 if cross else 0
    a.x_base = x_base
    for i, p in enumerate(y_srcs):
        a.y_src[i] = p
    a.ws = ws
    if strips is not None:
        a.strips[0], a.strips[1] = strips[0].data_ptr(), strips[1].data_ptr()
        a.ld_strip = strips[0].stride(0)
    return a


class HeadInfo:
    """What the fused head leaves behind for ClipInfoCELoss / accuracy: `parts` = (sum_i CE(logits_per_image[i]),
    sum_i CE(logits_per_text[i])) carrying the autograd graph, and the device-side top-1 / top-5 counts."""

    def __init__(self, parts, ws, b, n, layout):
        self.parts, self.ws, self.b, self.n, self.layout = parts, ws, b, n, layout

    @property
    def top1_count(self):
        return self.ws[self.layout.out + 4:self.layout.out + 5]

    @property
    def top5_count(self):
        return self.ws[self.layout.out + 5:self.layout.out + 6]


class _SymmFeatures:
    """Feature exchange of the fused head through SYMMETRIC memory (torch.distributed._symmetric_memory: every rank's
    buffer is mapped into every process) instead of an NCCL all-gather.  Two flavours:
      push : every rank owns a full gather buffer [N, F*E]; head_prep_kernel stores its normalised rows into ALL of them
             (a one-shot all-gather by peer stores over NVLink fused into the normalisation kernel), a device-side barrier
             follows, and the head kernels read their own buffer through the local L2 (dc_head_args.n_src = 1);
      pull : every rank owns only its rows [b, F*E]; the head kernels read the peers' rows through one TMA tensor map per
             rank (n_src = world).  No traffic before the kernels, but every row block re-reads the peer tiles over NVLink
             (peer reads bypass the local L2): measured +1.7 ms/step at 2 GPUs — kept as an option, not the default.
    Two buffers alternate: the backward of step i re-reads buffer i % 2, and no rank can overwrite it (prepare of step
    i + 2) before every rank has passed the barrier of step i + 1, i.e. has enqueued its step-i backward."""

    _cache = {}

    def __init__(self, rows, ld, dev):
        import torch.distributed._symmetric_memory as symm_mem
        self.bufs = [symm_mem.empty((rows, ld), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        self.hdls = [symm_mem.rendezvous(t, dist.group.WORLD) for t in self.bufs]
        self.ptrs = [[int(p) for p in h.buffer_ptrs] for h in self.hdls]
        self.i = 0

    @classmethod
    def get(cls, rows, ld, dev):
        key = (rows, ld, str(dev))
        if key not in cls._cache:
            cls._cache[key] = cls(rows, ld, dev)
        return cls._cache[key]

    def next(self):
        k = self.i & 1
        self.i += 1
        return self.bufs[k], self.hdls[k], self.ptrs[k]


def symm_head_mode(b):
    """DECLIP_B200_SYMM_HEAD = push | pull | 0 (default 0: NCCL all-gather).  pull needs b % 256 == 0."""
    import os
    mode = os.environ.get("DECLIP_B200_SYMM_HEAD", "0")
    if mode in ("1", "push"):
        return "push"
    if mode == "pull" and b % 256 == 0:
        return "pull"
    return None


class FusedClipHead(torch.autograd.Function):
    """parts[2], workspace = fused CLIP head on the raw tower outputs (image_features, text_features fp32 [b, E]).
    Exchange steps when gathering: one bf16 feature all-gather in the forward, one all-gather of 2b+2 floats per rank
    (row log-sum-exps + upstream gradients) in the backward — instead of the reference's all-reduce of two [N, E]
    gradients (clip.py:43-49)."""

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, gather, clamp):
        lib = ops.lib_for(image_features)
        img = image_features.float().contiguous()
        txt = text_features.float().contiguous()
        b, e = img.shape
        rank, world = dist_info()
        gather = bool(gather) and world > 1
        n, row0 = (world * b, rank * b) if gather else (b, 0)
        dev = img.device
        L = HeadLayout.get(lib, b, e)
        ws = torch.empty(L.total, device=dev, dtype=torch.float32)
        symm = symm_head_mode(b) if gather else None
        push = None
        if symm == "pull":
            local, hdl, srcs = _SymmFeatures.get(b, 2 * e, dev).next()
            allb = local
        elif symm == "push":
            allb, hdl, ptrs = _SymmFeatures.get(n, 2 * e, dev).next()
            local = allb[row0:row0 + b]
            srcs = [allb.data_ptr()]
            push = [p for r, p in enumerate(ptrs) if r != rank]
        else:
            allb = torch.empty(n, 2 * e, device=dev, dtype=torch.bfloat16)
            local = allb[row0:row0 + b]
            srcs = [allb.data_ptr()]
        ls = logit_scale.detach()
        if ls.dtype != torch.float32 or not ls.is_cuda:
            raise RuntimeError("declip_b200: logit_scale must be an fp32 CUDA parameter")
        feats = (_PTR * 2)(img.data_ptr(), txt.data_ptr())
        eps = (ctypes.c_float * 2)(0.0, 1e-10)                                         # clip.py:129-130
        smax = 100.0 if clamp else float("inf")
        if push is not None:
            peers = (_PTR * len(push))(*push)
            _lib.check(lib.dc_head_prepare_push(feats, eps, 2, b, e, _PTR(local.data_ptr()), peers, len(push), row0,
                                                _PTR(ws.data_ptr()), _PTR(ls.data_ptr()), smax, _stream()),
                       "dc_head_prepare_push")
        else:
            _lib.check(lib.dc_head_prepare(feats, eps, 2, b, e, _PTR(local.data_ptr()), _PTR(ws.data_ptr()),
                                           _PTR(ls.data_ptr()), smax, _stream()), "dc_head_prepare")
        if symm is not None:
            hdl.barrier(channel=0)                                                     # every rank's rows are in place
        elif gather:
            dist.all_gather_into_tensor(allb, local)                                  # in place: `local` is rank's slice
        args = head_args(b, n, e, 2 * e, row0, local.data_ptr(), srcs, ws.data_ptr())
        _lib.check(lib.dc_head_forward(ctypes.byref(args), _stream()), "dc_head_forward")
        ctx.save_for_backward(img, txt, ws, allb)
        ctx.meta = (b, n, e, row0, gather, world, L)
        ctx.srcs = srcs
        parts = ws[L.out:L.out + 2].clone()      # its own storage: an autograd output must not alias the workspace
        ctx.mark_non_differentiable(ws)
        return parts, ws

    @staticmethod
    def backward(ctx, gparts, _gws):
        img, txt, ws, allb = ctx.saved_tensors
        b, n, e, row0, gather, world, L = ctx.meta
        lib = ops.lib_for(img)
        g = gparts.contiguous().float()
        if gather:
            ws[L.g:L.g + 2].copy_(g)
            exch = torch.empty(world, 2 * b + 2, device=img.device, dtype=torch.float32)
            dist.all_gather_into_tensor(exch.view(-1), ws[L.lse:L.lse + 2 * b + 2])
        else:
            exch = ws[L.lse:L.lse + 2 * b + 2]
        d_img, d_txt = torch.empty_like(img), torch.empty_like(txt)
        local = allb if len(ctx.srcs) > 1 else allb[row0:row0 + b]
        args = head_args(b, n, e, 2 * e, row0, local.data_ptr(), ctx.srcs, ws.data_ptr())
        xraw = (_PTR * 2)(img.data_ptr(), txt.data_ptr())
        eps = (ctypes.c_float * 2)(0.0, 1e-10)
        dxo = (_PTR * 2)(d_img.data_ptr(), d_txt.data_ptr())
        _lib.check(lib.dc_head_backward(ctypes.byref(args), _PTR(g.data_ptr()), _PTR(exch.data_ptr()), xraw, eps, dxo,
                                        _stream()), "dc_head_backward")
        return d_img, d_txt, ws[L.out + 10:L.out + 11], None, None


class FusedPairHeads(torch.autograd.Function):
    """Several contrastive pairs over ONE gathered feature buffer, each through the fused head kernels — DeCLIP's four
    symmetric image/text pairs (declip.py:271-279) and its two nearest-neighbour pairs (declip.py:281-300, both strips
    image -> text, the bank rows carry no gradient).  feats: F raw feature tensors fp32 [b, E] (normalised inside, eps per
    feature); pairs: tuple of (x0, y0, x1, y1, cross) feature indices — direction d scores feature x_d (local rows) against
    feature y_d (all ranks); cross = 1 when direction 1 is the transpose of direction 0.  Returns parts [P, 2] (sum of row
    cross-entropies per pair and direction) and the per-pair workspaces."""

    @staticmethod
    def forward(ctx, logit_scale, gather, clamp, eps, pairs, *feats):
        lib = ops.lib_for(feats[0])
        feats = [f.float().contiguous() for f in feats]
        nf = len(feats)
        b, e = feats[0].shape
        rank, world = dist_info()
        gather = bool(gather) and world > 1
        n, row0 = (world * b, rank * b) if gather else (b, 0)
        dev = feats[0].device
        L = HeadLayout.get(lib, b, e)
        npair = len(pairs)
        ws = torch.empty(npair, L.total, device=dev, dtype=torch.float32)
        allb = torch.empty(n, nf * e, device=dev, dtype=torch.bfloat16)
        local = allb[row0:row0 + b]
        ls = logit_scale.detach()
        fp = (_PTR * nf)(*[f.data_ptr() for f in feats])
        ep = (ctypes.c_float * nf)(*[float(x) for x in eps])
        for k in range(npair):        # one prepare per pair: it also clears that pair's accumulators (normalising again is ~4 us)
            _lib.check(lib.dc_head_prepare(fp, ep, nf, b, e, _PTR(local.data_ptr()), _PTR(ws[k].data_ptr()), _PTR(ls.data_ptr()),
                                           100.0 if clamp else float("inf"), _stream()), "dc_head_prepare")
        if gather:
            dist.all_gather_into_tensor(allb, local)
        for k, (x0, y0, x1, y1, cross) in enumerate(pairs):
            args = head_args(b, n, e, nf * e, row0, local.data_ptr(), [allb.data_ptr()], ws[k].data_ptr(),
                             x_off=(x0 * e, x1 * e), y_off=(y0 * e, y1 * e), cross=bool(cross))
            _lib.check(lib.dc_head_forward(ctypes.byref(args), _stream()), "dc_head_forward")
        ctx.save_for_backward(ws, allb, *feats)
        ctx.meta = (b, n, e, row0, gather, world, L, nf, tuple(pairs), tuple(float(x) for x in eps))
        parts = ws[:, L.out:L.out + 2].clone()
        ctx.mark_non_differentiable(ws)
        return parts, ws

    @staticmethod
    def backward(ctx, gparts, _gws):
        ws, allb = ctx.saved_tensors[:2]
        feats = ctx.saved_tensors[2:]
        b, n, e, row0, gather, world, L, nf, pairs, eps = ctx.meta
        lib = ops.lib_for(allb)
        dev = allb.device
        npair = len(pairs)
        g = gparts.contiguous().float()                                   # [P, 2]
        xw = 2 * b + 2
        if gather:
            ws[:, L.g:L.g + 2].copy_(g)
            mine = ws[:, L.lse:L.lse + xw].contiguous()                  # [P, 2b+2]
            exch_all = torch.empty(world, npair, xw, device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(exch_all.view(-1), mine.view(-1))
            exch_all = exch_all.transpose(0, 1).contiguous()              # [P, W, 2b+2]
        local = allb[row0:row0 + b]
        grads = [None] * nf
        need = ctx.needs_input_grad[5:]
        dls = None
        for k, (x0, y0, x1, y1, cross) in enumerate(pairs):
            exch = exch_all[k] if gather else ws[k, L.lse:L.lse + xw]
            outs = [torch.empty(b, e, device=dev, dtype=torch.float32) if need[x] else None for x in (x0, x1)]
            args = head_args(b, n, e, nf * e, row0, local.data_ptr(), [allb.data_ptr()], ws[k].data_ptr(),
                             x_off=(x0 * e, x1 * e), y_off=(y0 * e, y1 * e), cross=bool(cross))
            xraw = (_PTR * 2)(feats[x0].data_ptr(), feats[x1].data_ptr())
            ep = (ctypes.c_float * 2)(eps[x0], eps[x1])
            dxo = (_PTR * 2)(outs[0].data_ptr() if outs[0] is not None else None,
                             outs[1].data_ptr() if outs[1] is not None else None)
            _lib.check(lib.dc_head_backward(ctypes.byref(args), _PTR(g[k].data_ptr()), _PTR(exch.data_ptr()), xraw, ep, dxo,
                                            _stream()), "dc_head_backward")
            for x, o in zip((x0, x1), outs):
                if o is not None:
                    grads[x] = o if grads[x] is None else grads[x] + o
            d = ws[k, L.out + 10:L.out + 11]
            dls = d if dls is None else dls + d
        return (dls, None, None, None, None) + tuple(grads)


def fused_pair_heads(logit_scale, gather, clamp, eps, pairs, feats):
    """-> list of (logits_a, logits_b) HANDLE pairs, one per entry of `pairs` (see FusedPairHeads / fused_clip_head)."""
    parts, ws = FusedPairHeads.apply(logit_scale, gather, clamp, tuple(eps), tuple(pairs), *feats)
    b, e = feats[0].shape
    rank, world = dist_info()
    n = world * b if (gather and world > 1) else b
    L = HeadLayout.get(ops.lib_for(feats[0]), b, e)
    out = []
    for k in range(len(pairs)):
        info = HeadInfo(parts[k], ws[k], b, n, L)
        la, lb = ws[k, 0:1].expand(b, n), ws[k, 1:2].expand(b, n)
        la._dc_head = lb._dc_head = info
        out.append((la, lb))
    return out


def fused_clip_head(image_features, text_features, logit_scale, gather, clamp=True):
    """-> (logits_per_image, logits_per_text) HANDLES: zero-stride [b, N] views that carry `._dc_head` (HeadInfo) for
    declip_b200.loss_functions.ClipInfoCELoss; they hold no logits (use the compat path when a caller reads them)."""
    parts, ws = FusedClipHead.apply(image_features, text_features, logit_scale, gather, clamp)
    b, e = image_features.shape
    rank, world = dist_info()
    n = world * b if (gather and world > 1) else b
    info = HeadInfo(parts, ws, b, n, HeadLayout.get(ops.lib_for(image_features), b, e))
    li = ws[0:1].expand(b, n)
    lt = ws[1:2].expand(b, n)
    li._dc_head = lt._dc_head = info
    return li, lt
