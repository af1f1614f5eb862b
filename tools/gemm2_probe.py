"""Bring-up probe of the 2-CTA (cta_group::2) GEMM: correctness vs torch for every major-ness / epilogue, then timing
against the 1-CTA kernel on the hot-path shapes.   python tools/gemm2_probe.py [check|perf]"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(mode):
    import torch
    from declip_b200 import _lib, ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ops.lib_for(torch.zeros(1, device=dev))

    def mk(r, c):
        return (torch.randn(r, c, device=dev) * 0.5).bfloat16()

    if mode == "check":
        bad = 0
        _lib.set_gemm_2cta(True)
        for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 1), (1, 0)):
            for (M, N, K) in ((256, 256, 64), (256, 256, 256), (512, 768, 768), (1000, 2304, 768), (25600, 768, 3072),
                              (3072, 768, 4096), (304, 520, 200)):
                a = mk(K, M) if a_mn else mk(M, K)
                b = mk(K, N) if b_mn else mk(N, K)
                want = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
                got = ops.gemm(a, b, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), epilogue=ops.EPI_F32)
                torch.cuda.synchronize()
                rel = ((got - want).norm() / want.norm()).item()
                ok = rel < 1e-5
                bad += (not ok)
                print(json.dumps({"a_mn": a_mn, "b_mn": b_mn, "M": M, "N": N, "K": K, "rel": rel, "ok": ok}), flush=True)
        M, N, K = 1000, 768, 512
        a, b = mk(M, K), mk(N, K)
        bias = torch.randn(N, device=dev)
        aux = mk(M, N)
        want = a.float() @ b.float().t()
        h, u = ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
        uu = want + bias
        r1 = ((u.float() - uu).norm() / uu.norm()).item()
        r2 = ((h.float() - uu * torch.sigmoid(1.702 * uu)).norm() / uu.norm()).item()
        o = ops.gemm(a, b, bias=bias, aux=aux, epilogue=ops.EPI_BF16_RESID)
        r3 = ((o.float() - (want + bias + aux.float())).norm() / want.norm()).item()
        acc = torch.ones(N, K, device=dev)
        dy, xx = mk(4096, N), mk(4096, K)
        ops.gemm(dy, xx, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=acc)
        r4 = ((acc - (dy.float().t() @ xx.float() + 1)).norm() / acc.norm()).item()
        print(json.dumps({"epi": [r1, r2, r3, r4]}), flush=True)
        bad += sum(1 for r in (r1, r2, r3) if r > 6e-3) + (r4 > 1e-5)
        print("CHECK bad=%d" % bad, flush=True)
    else:
        def bench(name, fn, flops, iters=20):
            for _ in range(3):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            return ms, flops / ms / 1e9
        M = 25600
        for (N, K, tag) in ((2304, 768, "qkv"), (3072, 768, "fc"), (768, 3072, "proj"), (768, 768, "out")):
            a, w = mk(M, K), mk(N, K)
            bias = torch.randn(N, device=dev)
            o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            dy = mk(M, N)
            dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
            dw = torch.zeros(N, K, device=dev)
            fl = 2.0 * M * N * K
            row = {"shape": tag}
            for mode2 in (0, 1):
                _lib.set_gemm_2cta(bool(mode2))
                row["bias_%dcta" % (mode2 + 1)] = round(bench("", lambda: ops.gemm(a, w, bias=bias, out=o1), fl)[1], 1)
                row["gelu_%dcta" % (mode2 + 1)] = round(bench("", lambda: ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BF16_GELU, out=o1, out2=o2), fl)[1], 1)
                row["dgrad_%dcta" % (mode2 + 1)] = round(bench("", lambda: ops.gemm(dy, w, b_mn_major=True, out=dx), fl)[1], 1)
                row["wgrad_%dcta" % (mode2 + 1)] = round(bench("", lambda: ops.gemm(dy, a, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=dw), fl)[1], 1)
            print(json.dumps(row), flush=True)
        a, w = mk(8192, 8192), mk(8192, 8192)
        o = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
        for mode2 in (0, 1):
            _lib.set_gemm_2cta(bool(mode2))
            print(json.dumps({"square8192_%dcta" % (mode2 + 1): round(bench("", lambda: ops.gemm(a, w, out=o), 2.0 * 8192 ** 3, 10)[1], 1)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker(sys.argv[1])
    else:
        os.makedirs("gpurun_out", exist_ok=True)
        for m in ("check", "perf"):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), m], capture_output=True, text=True, timeout=200)
                out = r.stdout + ("\nSTDERR:\n" + r.stderr[-2000:] if r.returncode else "")
            except subprocess.TimeoutExpired as ex:
                out = "TIMEOUT\n" + (ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else "TIMEOUT"
            open("gpurun_out/gemm2_%s.log" % m, "w").write(out)
            print("== %s\n%s" % (m, out[-3500:]), flush=True)
