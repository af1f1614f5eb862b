"""LayerNorm backward timing at the bench shapes.  Device time per launch over 24 back-to-back launches on four rotating
operand sets (4 x 157 MB > the 126 MB L2, so every launch reads from HBM) with pre-allocated outputs and direct C-ABI
calls — the first version of this tool timed ONE launch through the Python wrapper (allocations + three fills) behind a
flush kernel and mostly measured the host's enqueue latency: three very different kernels all "took" 52-60 us."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from declip_b200 import _lib, ops

_lib.init(0)
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
mode = "v1" if os.environ.get("DC_LN_BWD_V1") == "1" else ("group" if os.environ.get("DC_LN_BWD_GROUP") == "1" else "pipe")
for rows, W in ((25600, 768), (39424, 512)):
    sets = []
    for i in range(4):
        x = torch.randn(rows, W, device="cuda").bfloat16()
        dy = torch.randn(rows, W, device="cuda").bfloat16()
        dres = torch.randn(rows, W, device="cuda").bfloat16()
        sets.append((x, dy, dres, torch.empty_like(x)))
    g = torch.randn(W, device="cuda"); b = torch.randn(W, device="cuda")
    stats = [ops.layernorm_fwd(s[0], g, b)[1:] for s in sets]
    dg = torch.zeros(W, device="cuda"); db = torch.zeros(W, device="cuda"); dc = torch.zeros(W, device="cuda")

    def launch(i):
        x, dy, dres, dx = sets[i % 4]
        mean, rstd = stats[i % 4]
        rc = lib.dc_layernorm_bwd(P(dy), P(x), P(g), P(mean), P(rstd), P(dres), P(dx), P(dg), P(db), P(dc), rows, W, st)
        assert rc == 0

    for i in range(8):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 24
    e0.record()
    for i in range(N):
        launch(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / N
    # fp32 check of one set
    x, dy, dres, dx = sets[0]
    dg.zero_(); db.zero_(); dc.zero_(); launch(0); torch.cuda.synchronize()
    xf = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xf, (W,), g, b, 1e-5).backward(dy.float())
    ref = xf.grad + dres.float()
    nbytes = rows * W * 2 * 4
    # forward, same protocol (reads x, writes y: 2 passes)
    ys = [torch.empty_like(s_[0]) for s_ in sets]

    def fwd(i):
        mean, rstd = stats[i % 4]
        rc = lib.dc_layernorm_fwd(P(sets[i % 4][0]), P(g), P(b), P(ys[i % 4]), P(mean), P(rstd), rows, W, ctypes.c_float(1e-5), st)
        assert rc == 0

    for i in range(8):
        fwd(i)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(N):
        fwd(i)
    f1.record(); torch.cuda.synchronize()
    fus = f0.elapsed_time(f1) * 1e3 / N
    print(json.dumps({"rows": rows, "W": W, "mode": "fwd", "us": round(fus, 1), "GBps": round(nbytes / 2 / fus / 1e3, 0),
                      "floor_us_at_6584GBps": round(nbytes / 2 / 6584.5e3, 1)}))
    print(json.dumps({"rows": rows, "W": W, "mode": mode, "us": round(us, 1), "GBps": round(nbytes / us / 1e3, 0),
                      "floor_us_at_6584GBps": round(nbytes / 6584.5e3, 1), "dx_maxerr": round(float((dx.float() - ref).abs().max()), 4),
                      "dbeta_err": float((db - dy.float().sum(0)).abs().max()),
                      "dcol_relerr": float((dc - ref.sum(0)).abs().max() / ref.sum(0).abs().max())}))
