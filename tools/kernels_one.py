"""Runs the attention / LayerNorm kernels at the bench shapes a few times (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from declip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 512
for (L, H, causal) in ((50, 12, False), (77, 8, True)):
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16()
    dout = torch.randn(B * L, D, device=dev).bfloat16()
    dbias = torch.zeros(3 * D, device=dev)
    for _ in range(3):
        out, lse = ops.attention_fwd(qkv, B, L, H, causal)
        ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal, dbias=dbias)
    x = torch.randn(B * L, D, device=dev).bfloat16()
    g = torch.ones(D, device=dev)
    b = torch.zeros(D, device=dev)
    for _ in range(3):
        y, mean, rstd = ops.layernorm_fwd(x, g, b)
        ops.layernorm_bwd(dout, x, g, mean, rstd, dout, with_colsum=True)
torch.cuda.synchronize()
