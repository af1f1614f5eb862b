"""Runs one GEMM shape repeatedly (for ncu): python tools/gemm_one.py <2cta:0|1> <M> <N> <K> [epi]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from declip_b200 import _lib, ops  # noqa: E402

two, M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
epi = sys.argv[5] if len(sys.argv) > 5 else "bias"
dev = torch.device("cuda:0")
a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
bias = torch.randn(N, device=dev)
o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
ops.lib_for(a)
_lib.set_gemm_2cta(bool(two))
for _ in range(8):
    if epi == "gelu":
        ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BF16_GELU, out=o1, out2=o2)
    else:
        ops.gemm(a, w, bias=bias, out=o1)
torch.cuda.synchronize()
