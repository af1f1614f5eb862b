"""Launches the attention core a few times at the bench shapes (for ncu). Usage: python tools/attn_one.py [vit|text]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from declip_b200 import _lib, ops

_lib.init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "text"
B, L, H, causal = (512, 77, 8, 1) if which == "text" else (512, 50, 12, 0)
D = H * 64
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
dout = torch.randn(B * L, D, device="cuda").bfloat16()
dbias = torch.zeros(3 * D, device="cuda")
for _ in range(3):
    out, lse = ops.attention_fwd(qkv, B, L, H, causal)
    ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal, dbias=None if os.environ.get("NO_DBIAS") else dbias)
torch.cuda.synchronize()
print("done")
