"""Runs the fused contrastive head (csrc/head.cu: prep, forward, backward) a few times at the shape rank 0 of a W-rank job
sees — b local rows against N = W*b gathered columns, the other ranks' features synthesised — for ncu / timing.
    python tools/head_one.py [b=512] [world=8] [e=512] [iters=5]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from declip_b200 import _lib, functions as F_, ops  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
e = int(sys.argv[3]) if len(sys.argv) > 3 else 512
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
P = ctypes.c_void_p
n = b * world
torch.manual_seed(0)
img, txt = torch.randn(b, e, device=dev), torch.randn(b, e, device=dev)
ls = torch.tensor([2.659], device=dev)
lib = ops.lib_for(img)
L = F_.HeadLayout.get(lib, b, e)
allb = torch.nn.functional.normalize(torch.randn(n, 2, e, device=dev), dim=-1).view(n, 2 * e).bfloat16()   # peers' rows
ws = torch.empty(L.total, device=dev)
exch = torch.zeros(world, 2 * b + 2, device=dev)
exch[:, :2 * b] = 8.0                                      # plausible row LSEs of the peers
exch[:, 2 * b:] = 0.5 / b / world
g = torch.full((2,), 0.5 / b / world, device=dev)
d_img, d_txt = torch.empty_like(img), torch.empty_like(txt)
feats, eps = (P * 2)(img.data_ptr(), txt.data_ptr()), (ctypes.c_float * 2)(0.0, 1e-10)
xraw, dxo = (P * 2)(img.data_ptr(), txt.data_ptr()), (P * 2)(d_img.data_ptr(), d_txt.data_ptr())
st = P(torch.cuda.current_stream().cuda_stream)
args = F_.head_args(b, n, e, 2 * e, 0, allb.data_ptr(), [allb.data_ptr()], ws.data_ptr())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(iters):
    ev[0].record()
    _lib.check(lib.dc_head_prepare(feats, eps, 2, b, e, P(allb.data_ptr()), P(ws.data_ptr()), P(ls.data_ptr()), 100.0, st), "prep")
    ev[1].record()
    _lib.check(lib.dc_head_forward(ctypes.byref(args), st), "fwd")
    ev[2].record()
    exch[0, :2 * b] = ws[L.lse:L.lse + 2 * b]
    _lib.check(lib.dc_head_backward(ctypes.byref(args), P(g.data_ptr()), P(exch.data_ptr()), xraw, eps, dxo, st), "bwd")
    ev[3].record()
torch.cuda.synchronize()
print("head b=%d N=%d E=%d: prep %.1f us, fwd %.1f us, bwd (+exch copy) %.1f us; loss parts %s" %
      (b, n, e, 1e3 * ev[0].elapsed_time(ev[1]), 1e3 * ev[1].elapsed_time(ev[2]), 1e3 * ev[2].elapsed_time(ev[3]),
       ws[L.out:L.out + 2].tolist()))
