"""Thin tensor-level wrappers over the C ABI (include/declip_b200.h).

torch is used here only for device memory and streams: every function takes CUDA tensors,
passes raw pointers + sizes to the library and returns tensors it allocated through torch's
caching allocator.  No arithmetic happens in Python/PyTorch on this path.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs

EPI_BF16, EPI_BF16_GELU, EPI_BF16_RESID, EPI_BF16_DGELU, EPI_F32, EPI_F32_ATOMIC, EPI_F32_GROUPMAX16 = range(7)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def lib_for(t):
    if not t.is_cuda:
        raise RuntimeError("declip_b200 ops need CUDA tensors on an sm_100a device; there is no CPU path")
    return _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, epilogue=EPI_BF16, alpha=1.0, bias=None, aux=None, out=None,
         out2=None, splits=0, block_n=0, alpha_dev=None, colsum=None):
    """out[M,N] (op)= epilogue(alpha * sum_k A(m,k) B(n,k)).

    a: [M,K] (or [K,M] if a_mn_major), b: [N,K] (or [K,N] if b_mn_major); both bf16, last dim contiguous.
    """
    lib = lib_for(a)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and a.dim() == 2 and b.dim() == 2
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, "contraction mismatch %d vs %d" % (K, Kb)
    f32_out = epilogue in (EPI_F32, EPI_F32_ATOMIC, EPI_F32_GROUPMAX16)
    if epilogue == EPI_F32_GROUPMAX16:       # out [M, N/16] fp32 group maxima, out2 [M, N/16] uint8 arg-max
        assert N % 16 == 0
        if out is None:
            out = torch.empty(M, N // 16, device=a.device, dtype=torch.float32)
        if out2 is None:
            out2 = torch.empty(M, N // 16, device=a.device, dtype=torch.uint8)
        assert out2.dtype == torch.uint8 and out.shape == (M, N // 16) and out2.shape == (M, N // 16)
    if out is None:
        if epilogue == EPI_F32_ATOMIC:
            out = torch.zeros(M, N, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(M, N, device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    assert out.dtype == (torch.float32 if f32_out else torch.bfloat16) and out.stride(-1) == 1
    if epilogue == EPI_BF16_GELU and out2 is None:
        out2 = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    args = GemmArgs()
    args.A, args.lda, args.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn_major)
    args.B, args.ldb, args.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn_major)
    args.M, args.N, args.K = M, N, K
    args.epilogue, args.alpha = epilogue, float(alpha)
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.out2, args.ldo2 = (out2.data_ptr(), out2.stride(0)) if out2 is not None else (None, 0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        args.bias = bias.data_ptr()
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and aux.stride(-1) == 1
        args.aux, args.ldaux = aux.data_ptr(), aux.stride(0)
    args.splits, args.block_n = splits, block_n
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.numel() == N
        args.colsum = colsum.data_ptr()
    if alpha_dev is not None:
        assert alpha_dev.dtype == torch.float32 and alpha_dev.is_cuda
        args.alpha_dev = alpha_dev.data_ptr()
    _lib.check(lib.dc_gemm_bf16(ctypes.byref(args), _stream()), "dc_gemm_bf16")
    if epilogue in (EPI_BF16_GELU, EPI_F32_GROUPMAX16):
        return out, out2
    return out


# ------------------------------------------------------------------ op-level wrappers (tests, tools)
def layernorm_fwd(x, gamma, beta, eps=1e-5):
    lib = lib_for(x)
    rows, width = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.check(lib.dc_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), rows, width, eps,
                                    _stream()), "dc_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, with_colsum=False):
    lib = lib_for(x)
    rows, width = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.zeros(width, device=x.device, dtype=torch.float32)
    dbeta = torch.zeros(width, device=x.device, dtype=torch.float32)
    dcol = torch.zeros(width, device=x.device, dtype=torch.float32) if with_colsum else None
    _lib.check(lib.dc_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), _ptr(dx),
                                    _ptr(dgamma), _ptr(dbeta), _ptr(dcol), rows, width, _stream()), "dc_layernorm_bwd")
    if with_colsum:
        return dx, dgamma, dbeta, dcol
    return dx, dgamma, dbeta


def colsum(x, out=None):
    lib = lib_for(x)
    rows, cols = x.shape
    if out is None:
        out = torch.zeros(cols, device=x.device, dtype=torch.float32)
    _lib.check(lib.dc_colsum_bf16(_ptr(x), x.stride(0), _ptr(out), rows, cols, _stream()), "dc_colsum_bf16")
    return out


def attention_fwd(qkv, batch, L, heads, causal):
    lib = lib_for(qkv)
    D = heads * 64
    out = torch.empty(batch * L, D, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(batch * heads * L, device=qkv.device, dtype=torch.float32)
    _lib.check(lib.dc_attention_fwd(_ptr(qkv), _ptr(out), _ptr(lse), batch, L, heads, int(causal), _stream()),
               "dc_attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, lse, batch, L, heads, causal, dbias=None):
    lib = lib_for(qkv)
    dqkv = torch.empty_like(qkv)
    _lib.check(lib.dc_attention_bwd(_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(dbias), batch, L,
                                    heads, int(causal), _stream()), "dc_attention_bwd")
    return dqkv


def patchify(images, patch):
    lib = lib_for(images)
    B, C, R, _ = images.shape
    g = R // patch
    out = torch.empty(B * g * g, 3 * patch * patch, device=images.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_patchify(_ptr(images), images.stride(0), _ptr(out), B, R, patch, _stream()), "dc_patchify")
    return out


def text_embed(ids, table, pos):
    lib = lib_for(table)
    B, L = ids.shape
    W = table.shape[1]
    x = torch.empty(B * L, W, device=table.device, dtype=torch.bfloat16)
    _lib.check(lib.dc_text_embed(_ptr(ids), _ptr(table), _ptr(pos), _ptr(x), B, L, W, _stream()), "dc_text_embed")
    return x


def text_embed_bwd(ids, dx, vocab):
    lib = lib_for(dx)
    B, L = ids.shape
    W = dx.shape[1]
    dtable = torch.zeros(vocab, W, device=dx.device, dtype=torch.float32)
    dpos = torch.zeros(L, W, device=dx.device, dtype=torch.float32)
    _lib.check(lib.dc_text_embed_bwd(_ptr(ids), _ptr(dx), _ptr(dtable), _ptr(dpos), None, B, L, W, _stream()),
               "dc_text_embed_bwd")
    return dtable, dpos


def eot_index(ids):
    lib = lib_for(ids)
    B, L = ids.shape
    out = torch.empty(B, device=ids.device, dtype=torch.int32)
    _lib.check(lib.dc_eot_index(_ptr(ids), _ptr(out), B, L, _stream()), "dc_eot_index")
    return out
