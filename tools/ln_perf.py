"""LayerNorm backward timing at the bench shapes (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from declip_b200 import _lib, ops

_lib.init(0)
for rows, W in ((25600, 768), (39424, 512)):
    x = torch.randn(rows, W, device="cuda").bfloat16()
    dy = torch.randn(rows, W, device="cuda").bfloat16()
    dres = torch.randn(rows, W, device="cuda").bfloat16()
    g = torch.randn(W, device="cuda"); b = torch.randn(W, device="cuda")
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(10):
        flush.sum()          # evict with CLEAN lines: a zero_() flush leaves ~126 MB of dirty L2 whose write-back is charged to the timed kernel
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dx, dg, db, dc = ops.layernorm_bwd(dy, x, g, mean, rstd, dres, with_colsum=True)
        e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    # fp32 check
    xf = x.float().requires_grad_(True)
    yf = torch.nn.functional.layer_norm(xf, (W,), g, b, 1e-5)
    yf.backward(dy.float())
    ref = xf.grad + dres.float()
    err = float((dx.float() - ref).abs().max()); 
    cg = float(torch.nn.functional.cosine_similarity(dg, (dy.float() * ((x.float() - x.float().mean(1, keepdim=True)) * rstd[:, None])).sum(0), dim=0))
    print(json.dumps({"rows": rows, "W": W, "us": round(sum(ts) / len(ts), 1), "floor_us": round(rows * W * 2 * 4 / 7.0e6, 1),
                      "dx_maxerr": round(err, 4), "dgamma_cos": round(cg, 6), "dbeta_err": float((db - dy.float().sum(0)).abs().max()),
                      "dcol_err": float((dc - ref.sum(0)).abs().max() / ref.sum(0).abs().max()),
                      "mode": "v1" if os.environ.get("DC_LN_BWD_V1") == "1" else ("group" if os.environ.get("DC_LN_BWD_GROUP") == "1" else "pipe")}))
