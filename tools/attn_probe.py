"""tcgen05 attention core vs the mma.sync core vs an fp32 torch restatement; timing at the bench shapes.
Usage (GPU box): python tools/attn_probe.py [--perf]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from declip_b200 import _lib, ops


def torch_ref(qkv, dout, B, L, H, causal):
    D = H * 64
    x = qkv.float().view(B, L, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    s = q @ k.transpose(-1, -2) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=qkv.device).triu(1)
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B * L, D)
    o.backward(dout.float())
    lse = torch.logsumexp(s, -1).reshape(-1)
    return o.detach(), lse.detach(), x.grad.reshape(B * L, 3 * D)


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


def run(B, L, H, causal, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, device="cuda", generator=g) * 1.5).bfloat16()
    dout = torch.randn(B * L, D, device="cuda", generator=g).bfloat16()
    res = {}
    for name, tc in (("tc", True), ("mma", False)):
        if not tc and L > 80:
            continue
        _lib.set_attention_tc(tc)
        out, lse = ops.attention_fwd(qkv, B, L, H, causal)
        dbias = torch.zeros(3 * D, device="cuda")
        dqkv = ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal, dbias=dbias)
        torch.cuda.synchronize()
        res[name] = (out, lse, dqkv, dbias)
    _lib.set_attention_tc(True)
    ro, rl, rd = torch_ref(qkv, dout, B, L, H, causal)
    rb = rd.sum(0)
    line = {"B": B, "L": L, "H": H, "causal": causal}
    ok = True
    for name, (out, lse, dqkv, dbias) in res.items():
        e = {"out_cos": cos(out, ro), "out_maxerr": float((out.float() - ro).abs().max()),
             "lse_maxerr": float((lse - rl).abs().max()), "dqkv_cos": cos(dqkv, rd),
             "dqkv_maxerr": float((dqkv.float() - rd).abs().max()), "dbias_cos": cos(dbias, rb),
             "dbias_relerr": float((dbias - rb).abs().max() / (rb.abs().max() + 1e-30))}
        line[name] = {k: round(v, 6) for k, v in e.items()}
        if name == "tc":
            ok = e["out_cos"] > 0.9999 and e["dqkv_cos"] > 0.9995 and e["lse_maxerr"] < 2e-2 and e["dbias_cos"] > 0.999
            ok = ok and bool(torch.isfinite(dqkv.float()).all()) and bool(torch.isfinite(out.float()).all())
    line["ok"] = ok
    print(json.dumps(line), flush=True)
    return ok


def perf(B, L, H, causal):
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
    dout = torch.randn(B * L, D, device="cuda").bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    line = {"perf": True, "B": B, "L": L, "H": H, "causal": causal}
    for name, tc in (("tc", True), ("mma", False)):
        _lib.set_attention_tc(tc)
        out, lse = ops.attention_fwd(qkv, B, L, H, causal)
        dbias = torch.zeros(3 * D, device="cuda")
        tf, tb = [], []
        for i in range(8):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out, lse = ops.attention_fwd(qkv, B, L, H, causal)
            e[1].record()
            ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal, dbias=dbias)
            e[2].record()
            torch.cuda.synchronize()
            if i >= 3:
                tf.append(e[0].elapsed_time(e[1]) * 1e3)
                tb.append(e[1].elapsed_time(e[2]) * 1e3)
        line[name] = {"fwd_us": round(sum(tf) / len(tf), 1), "bwd_us": round(sum(tb) / len(tb), 1)}
    _lib.set_attention_tc(True)
    fb = B * L * D * 2 * 4
    line["hbm_floor_us"] = {"fwd": round(fb / 7.0e6, 1), "bwd": round(B * L * D * 2 * 7 / 7.0e6, 1)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    _lib.init(0)
    allok = True
    for cfg in [(2, 50, 12, 0), (64, 50, 12, 0), (2, 77, 8, 1), (64, 77, 8, 1), (5, 77, 8, 0), (4, 50, 32, 0),
                (2, 64, 4, 0), (2, 33, 2, 1), (3, 17, 2, 0), (2, 128, 2, 1), (1, 100, 3, 0), (300, 50, 12, 0)]:
        allok &= run(*cfg)
    print("ALL_OK" if allok else "MISMATCH", flush=True)
    if "--perf" in sys.argv:
        perf(512, 50, 12, 0)
        perf(512, 77, 8, 1)
