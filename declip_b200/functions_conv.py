"""autograd bridges for the ModifiedResNet image tower (prototype/model/image_encoder/modified_resnet.py) over the
C ABI.  Activations are NHWC bf16 matrices [B*H*W, C]; 1x1 convolutions are the tcgen05 GEMM, 3x3 convolutions the
implicit-GEMM kernels of csrc/conv_igemm.cu (im2col + GEMM only for shapes outside their domain and the 3-channel stem)."""
import ctypes

import os

import torch

from . import _lib, ops
from .functions import cast_bf16
from .runtime import weight_shadow

_PTR = ctypes.c_void_p


def _stream():
    return _PTR(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else _PTR(t.data_ptr())


class StemConv(torch.autograd.Function):
    """conv1 = Conv2d(3, 32, 3, stride 2, pad 1) on the fp32 NCHW image -> NHWC bf16 [B*H/2*W/2, 32]
    (modified_resnet.py:150,193-195).  The image needs no gradient."""

    @staticmethod
    def forward(ctx, images, weight):
        lib = ops.lib_for(images)
        B, C, H, W = images.shape
        assert C == 3 and images.stride(3) == 1 and images.stride(2) == W and images.stride(1) == H * W
        rows = B * (H // 2) * (W // 2)
        col = torch.empty(rows, 32, device=images.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_im2col_stem(_p(images), images.stride(0), _p(col), B, H, W, _stream()), "dc_im2col_stem")
        cout = weight.shape[0]
        wp = torch.zeros(cout, 32, device=images.device, dtype=torch.float32)
        wp[:, :27] = weight.permute(0, 2, 3, 1).reshape(cout, 27)          # (ky, kx, c) order of the im2col
        y = ops.gemm(col, cast_bf16(wp))
        ctx.save_for_backward(col)
        ctx.wshape = tuple(weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        cout = ctx.wshape[0]
        dw = ops.gemm(dy.contiguous(), col, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC)   # [cout, 32]
        dw = dw[:, :27].reshape(cout, 3, 3, 3).permute(0, 3, 1, 2).contiguous()
        return None, dw


class Conv1x1(torch.autograd.Function):
    """1x1 convolution on NHWC == GEMM (modified_resnet.py:20,27,34)."""

    @staticmethod
    def forward(ctx, x, weight):
        w16 = weight_shadow(weight).view(weight.shape[0], -1)      # [cout, cin] bf16, re-cast only when the master changed
        y = ops.gemm(x, w16)
        ctx.save_for_backward(x, w16)
        ctx.wshape = tuple(weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w16 = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.gemm(dy, w16, b_mn_major=True) if ctx.needs_input_grad[0] else None
        dw = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC).reshape(ctx.wshape)
        return dx, dw


class Conv1x1Skip(torch.autograd.Function):
    """conv1 of a Bottleneck together with the block's skip connection (modified_resnet.py:40-56: `identity = x` next to
    `self.conv1(x)`): returns (conv1(x), x).  Routing the identity branch through the second output hands its gradient to
    THIS backward, where the input-gradient GEMM adds it in its epilogue (fp32 accumulator + bf16 aux, one rounding) —
    instead of autograd materialising both branch gradients and summing them in a separate elementwise pass
    (16 launches, 2.6 ms of the res50 step, 6 B/element of HBM traffic)."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.set_materialize_grads(False)      # an unused output's gradient stays None instead of a zero tensor
        w16 = weight_shadow(weight).view(weight.shape[0], -1)
        y = ops.gemm(x, w16)
        ctx.save_for_backward(x, w16)
        ctx.wshape = tuple(weight.shape)
        return y, x.detach()          # same storage; as an output of this node it carries the skip branch's gradient back here

    @staticmethod
    def backward(ctx, dy, dskip):
        x, w16 = ctx.saved_tensors
        dx = None
        if dy is None:                # only the skip branch was used
            return dskip, None
        dy = dy.contiguous()
        if ctx.needs_input_grad[0]:
            if dskip is not None:
                dx = ops.gemm(dy, w16, b_mn_major=True, epilogue=ops.EPI_BF16_RESID, aux=dskip.contiguous())
            else:
                dx = ops.gemm(dy, w16, b_mn_major=True)
        dw = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC).reshape(ctx.wshape)
        return dx, dw


def _igemm_conv(lib, x, w_flat, B, H, W, C, cout):
    y = torch.empty(B * H * W, cout, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_conv3x3_igemm(_p(x), _p(w_flat), _p(y), B, H, W, C, cout, _stream()), "dc_conv3x3_igemm")
    return y


def _conv3x3_shadows(weight):
    """bf16 [Cout, 9*Cin] (ky, kx, c) forward operand and [Cin, 9*Cout] flipped / transposed dgrad operand of a 3x3 weight,
    re-derived only when the fp32 master changed (version counter; FusedAdamW bumps it)."""
    cache = getattr(weight, "_dc_conv3", None)
    if cache is not None and cache[0] == weight._version and cache[1].device == weight.device:
        return cache[1], cache[2]
    cout, cin = weight.shape[0], weight.shape[1]
    w = weight.detach()
    wf = cast_bf16(w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous())
    wb = cast_bf16(w.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, 9 * cout).contiguous())
    weight._dc_conv3 = (weight._version, wf, wb)
    return wf, wb


class Conv3x3(torch.autograd.Function):
    """3x3 / pad 1 / stride 1 convolution on NHWC (modified_resnet.py:23,151-154).  C, Cout multiples of 32 (every conv of
    the tower, the 32-channel stem included): implicit GEMM (csrc/conv_igemm.cu) for the forward, the input gradient (same
    kernel, flipped / transposed weights) and the weight gradient (contraction over pixel boxes) — no [rows, 9C] matrix is
    written in any direction.  Other shapes: im2col + GEMM, col2im for the dgrad."""

    @staticmethod
    def forward(ctx, x, weight, B, H, W):
        lib = ops.lib_for(x)
        C = x.shape[1]
        cout = weight.shape[0]
        igemm = bool(lib.dc_conv3x3_igemm_supported(H, W, C, cout)) and bool(lib.dc_conv3x3_igemm_supported(H, W, cout, C))
        if igemm:
            w16, wb16 = _conv3x3_shadows(weight)
            y = _igemm_conv(lib, x.contiguous(), w16, B, H, W, C, cout)
            ctx.save_for_backward(x, wb16)
        else:
            col = torch.empty(B * H * W, 9 * C, device=x.device, dtype=torch.bfloat16)
            _lib.check(lib.dc_im2col3x3(_p(x), _p(col), B, H, W, C, _stream()), "dc_im2col3x3")
            w16 = cast_bf16(weight.permute(0, 2, 3, 1).reshape(cout, 9 * C).contiguous())
            y = ops.gemm(col, w16)
            ctx.save_for_backward(x, w16)
        ctx.geom = (B, H, W, C, cout, igemm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wsaved = ctx.saved_tensors
        lib = ops.lib_for(x)
        B, H, W, C, cout, igemm = ctx.geom
        dy = dy.contiguous()
        if (igemm and lib.dc_conv3x3_wgrad_igemm_supported(H, W, C, cout)
                and os.environ.get("DECLIP_B200_CONV_WGRAD", "igemm") == "igemm"):
            # contraction over spatial TMA boxes of dy and (shifted) x: no im2col matrix for the weight gradient either
            dw = torch.zeros(cout, 9 * C, device=x.device, dtype=torch.float32)
            _lib.check(lib.dc_conv3x3_wgrad_igemm(_p(dy), _p(x), _p(dw), B, H, W, C, cout, _stream()), "dc_conv3x3_wgrad_igemm")
        else:
            col = torch.empty(B * H * W, 9 * C, device=x.device, dtype=torch.bfloat16)
            _lib.check(lib.dc_im2col3x3(_p(x), _p(col), B, H, W, C, _stream()), "dc_im2col3x3")
            dw = ops.gemm(dy, col, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC)     # [cout, 9C]
        dw = dw.reshape(cout, 3, 3, C).permute(0, 3, 1, 2).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            if igemm:      # dx = conv3x3(dy, flipped / transposed weights): the same kernel, no dcol / col2im
                dx = _igemm_conv(lib, dy, wsaved, B, H, W, cout, C)
            else:
                dcol = ops.gemm(dy, wsaved, b_mn_major=True)                                        # [rows, 9C]
                dx = torch.empty_like(x)
                _lib.check(lib.dc_col2im3x3(_p(dcol), _p(dx), B, H, W, C, _stream()), "dc_col2im3x3")
        return dx, dw, None, None, None


class BatchNorm2dNHWC(torch.autograd.Function):
    """nn.BatchNorm2d (training: per-rank batch statistics, `use_sync_bn: False` — the only mode that works with the
    reference's shim) fused with the optional residual add and ReLU (modified_resnet.py:40-56)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, running_mean, running_var, relu, eps, momentum, training=True):
        lib = ops.lib_for(x)
        rows, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(C, device=x.device, dtype=torch.float32)
        rstd = torch.empty(C, device=x.device, dtype=torch.float32)
        scratch = torch.empty(2 * C, device=x.device, dtype=torch.float32)
        _lib.check(lib.dc_bn2d_fwd(_p(x), _p(gamma), _p(beta), _p(res), _p(y), _p(mean), _p(rstd), _p(running_mean),
                                   _p(running_var), _p(scratch), rows, C, float(eps), float(momentum), int(bool(training)),
                                   int(relu), _stream()), "dc_bn2d_fwd")
        ctx.bn_training = bool(training)
        ctx.save_for_backward(x, y, gamma, mean, rstd)
        ctx.relu, ctx.has_res = bool(relu), res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.bn_training:
            raise NotImplementedError("declip_b200: backward through eval-mode BatchNorm2d (frozen statistics) is not built")
        x, y, gamma, mean, rstd = ctx.saved_tensors
        lib = ops.lib_for(x)
        rows, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if ctx.has_res else None
        dg = torch.zeros(C, device=x.device, dtype=torch.float32)
        db = torch.zeros(C, device=x.device, dtype=torch.float32)
        scratch = torch.empty(2 * C, device=x.device, dtype=torch.float32)
        _lib.check(lib.dc_bn2d_bwd(_p(dy), _p(x), _p(y), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dres), _p(dg), _p(db),
                                   _p(scratch), rows, C, int(ctx.relu), _stream()), "dc_bn2d_bwd")
        return dx, dres, dg, db, None, None, None, None, None, None


class AvgPool2(torch.autograd.Function):
    """nn.AvgPool2d(2) on NHWC (anti-aliased stride, modified_resnet.py:25,31-32,155)."""

    @staticmethod
    def forward(ctx, x, B, H, W):
        lib = ops.lib_for(x)
        C = x.shape[1]
        y = torch.empty(B * (H // 2) * (W // 2), C, device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_avgpool2(_p(x), _p(y), B, H, W, C, 0, _stream()), "dc_avgpool2")
        ctx.geom = (B, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.geom
        lib = ops.lib_for(dy)
        dy = dy.contiguous()
        dx = torch.empty(B * H * W, C, device=dy.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_avgpool2(_p(dy), _p(dx), B, H, W, C, 1, _stream()), "dc_avgpool2")
        return dx, None, None, None


class AttnPoolAssemble(torch.autograd.Function):
    """tokens = cat([mean token, x]) + positional_embedding (modified_resnet.py:72-74): x bf16 [B*P, C] -> [B*(P+1), C]."""

    @staticmethod
    def forward(ctx, x, pos, B, P):
        lib = ops.lib_for(x)
        C = x.shape[1]
        tok = torch.empty(B * (P + 1), C, device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_attnpool_assemble(_p(x), _p(pos), _p(tok), B, P, C, _stream()), "dc_attnpool_assemble")
        ctx.geom = (B, P, C)
        return tok

    @staticmethod
    def backward(ctx, dtok):
        B, P, C = ctx.geom
        lib = ops.lib_for(dtok)
        dtok = dtok.contiguous()
        dx = torch.empty(B * P, C, device=dtok.device, dtype=torch.bfloat16)
        _lib.check(lib.dc_attnpool_assemble_bwd(_p(dtok), _p(dx), B, P, C, _stream()), "dc_attnpool_assemble_bwd")
        dpos = ops.colsum(dtok.view(B, (P + 1) * C)).view(P + 1, C)
        return dx, dpos, None, None


class LinearBF16(torch.autograd.Function):
    """y(bf16) = x(bf16) W^T + b."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w16 = weight_shadow(weight)
        x = x.contiguous()
        y = ops.gemm(x, w16, bias=bias)
        ctx.save_for_backward(x, w16)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w16 = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.gemm(dy, w16, b_mn_major=True)
        dw = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC)
        return dx, dw, ops.colsum(dy)


class Attention(torch.autograd.Function):
    """softmax(q k^T / 8) v per (sample, head) on packed qkv bf16 [B*L, 3*D] (csrc/attention.cu)."""

    @staticmethod
    def forward(ctx, qkv, B, L, heads, causal):
        out, lse = ops.attention_fwd(qkv.contiguous(), B, L, heads, causal)
        ctx.save_for_backward(qkv, out, lse)
        ctx.geom = (B, L, heads, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        B, L, heads, causal = ctx.geom
        return ops.attention_bwd(qkv, out, dout.contiguous(), lse, B, L, heads, causal), None, None, None, None
