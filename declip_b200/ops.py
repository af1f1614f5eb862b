"""Thin tensor-level wrappers over the C ABI (include/declip_b200.h).

torch is used here only for device memory and streams: every function takes CUDA tensors,
passes raw pointers + sizes to the library and returns tensors it allocated through torch's
caching allocator.  No arithmetic happens in Python/PyTorch on this path.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs

EPI_BF16, EPI_BF16_GELU, EPI_BF16_RESID, EPI_BF16_DGELU, EPI_F32, EPI_F32_ATOMIC = range(6)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def lib_for(t):
    if not t.is_cuda:
        raise RuntimeError("declip_b200 ops need CUDA tensors on an sm_100a device; there is no CPU path")
    return _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, epilogue=EPI_BF16, alpha=1.0, bias=None, aux=None, out=None,
         out2=None, splits=0, block_n=0):
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
    f32_out = epilogue in (EPI_F32, EPI_F32_ATOMIC)
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
    _lib.check(lib.dc_gemm_bf16(ctypes.byref(args), _stream()), "dc_gemm_bf16")
    if epilogue == EPI_BF16_GELU:
        return out, out2
    return out
