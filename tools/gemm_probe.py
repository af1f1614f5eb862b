"""GPU bring-up probe for the tcgen05 GEMM: every operand major-ness x epilogue x ragged shape,
checked against torch fp32 matmul, plus a first timing of the hot-path shapes.
Run each group in its own process so one trapping kernel does not poison the others:
    python tools/gemm_probe.py            (driver: spawns the groups)
    python tools/gemm_probe.py GROUP      (worker)
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["kk", "kmn", "mnk", "mnmn", "epi", "perf"]


def worker(group):
    import torch
    from declip_b200 import ops
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    res = []

    def ref(a, b, a_mn, b_mn):
        A = a.float().t() if a_mn else a.float()
        B = b.float().t() if b_mn else b.float()
        return A @ B.t()

    def mk(M, K, mn, ones=False):
        shape = (K, M) if mn else (M, K)
        if ones:
            return torch.ones(shape, device=dev, dtype=torch.bfloat16)
        return (torch.randn(shape, device=dev) * 0.5).to(torch.bfloat16)

    def check(name, got, want, tol=2e-2):
        got = got.float()
        err = (got - want).abs()
        scale = want.abs().max().item() + 1e-6
        bad = (err > tol * scale)
        info = {"name": name, "max_err": err.max().item(), "scale": scale, "bad_frac": bad.float().mean().item()}
        if bad.any():
            idx = bad.nonzero()[:6].tolist()
            info["first_bad"] = [(i, j, got[i, j].item(), want[i, j].item()) for i, j in idx]
            rows_bad = bad.any(1).nonzero().flatten()
            cols_bad = bad.any(0).nonzero().flatten()
            info["bad_rows"] = (rows_bad.min().item(), rows_bad.max().item(), rows_bad.numel())
            info["bad_cols"] = (cols_bad.min().item(), cols_bad.max().item(), cols_bad.numel())
        info["ok"] = not bad.any().item()
        print(json.dumps(info), flush=True)
        res.append(info)

    def run_major(a_mn, b_mn):
        shapes = [(128, 128, 64), (128, 256, 64), (128, 256, 128), (256, 512, 256), (384, 768, 768),
                  (200, 264, 200), (50, 512, 768), (1000, 2304, 768), (512, 4096, 512)]
        for bn in (256, 128):
            for (M, N, K) in shapes:
                for ones in (True, False):
                    a = mk(M, K, a_mn, ones)
                    b = mk(N, K, b_mn, ones)
                    out = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, epilogue=ops.EPI_F32, block_n=bn)
                    torch.cuda.synchronize()
                    check("maj a_mn=%d b_mn=%d bn=%d M=%d N=%d K=%d ones=%d" % (a_mn, b_mn, bn, M, N, K, ones), out,
                          ref(a, b, a_mn, b_mn))

    if group == "kk":
        run_major(False, False)
    elif group == "kmn":
        run_major(False, True)
    elif group == "mnk":
        run_major(True, False)
    elif group == "mnmn":
        run_major(True, True)
    elif group == "epi":
        M, N, K = 1000, 768, 512
        a = mk(M, K, False)
        b = mk(N, K, False)
        bias = torch.randn(N, device=dev)
        aux = (torch.randn(M, N, device=dev)).to(torch.bfloat16)
        want = ref(a, b, False, False)
        o = ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BF16)
        check("epi bf16+bias", o, want + bias)
        o = ops.gemm(a, b, epilogue=ops.EPI_BF16, alpha=0.5)
        check("epi bf16 alpha", o, 0.5 * want)
        h, u = ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
        uu = want + bias
        check("epi gelu u", u, uu)
        check("epi gelu h", h, uu * torch.sigmoid(1.702 * uu))
        o = ops.gemm(a, b, bias=bias, aux=aux, epilogue=ops.EPI_BF16_RESID)
        check("epi resid", o, want + bias + aux.float())
        o = ops.gemm(a, b, aux=aux, epilogue=ops.EPI_BF16_DGELU)
        x = aux.float()
        s = torch.sigmoid(1.702 * x)
        check("epi dgelu", o, want * (s * (1 + 1.702 * x * (1 - s))))
        # split-K atomic accumulate (wgrad shape): dW[N,K] = dY^T X, contraction over M tokens
        Mt, No, Ki = 4096, 768, 512
        dy = mk(No, Mt, True)   # stored [Mt, No]
        xx = mk(Ki, Mt, True)   # stored [Mt, Ki]
        acc = torch.ones(No, Ki, device=dev)
        for sp in (0, 1, 3, 7):
            acc.fill_(1.0)
            ops.gemm(dy, xx, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=acc, splits=sp)
            check("epi atomic splits=%d" % sp, acc, dy.float().t() @ xx.float() + 1.0)
    elif group == "perf":
        def bench(name, M, N, K, a_mn, b_mn, epi, bn=0, splits=0, iters=20):
            a = mk(M, K, a_mn)
            b = mk(N, K, b_mn)
            kw = {}
            if epi == ops.EPI_F32_ATOMIC:
                kw["out"] = torch.zeros(M, N, device=dev)
            for _ in range(3):
                ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, epilogue=epi, block_n=bn, splits=splits, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(iters):
                ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, epilogue=epi, block_n=bn, splits=splits, **kw)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            print(json.dumps({"perf": name, "M": M, "N": N, "K": K, "bn": bn, "ms": ms, "tflops": tf}), flush=True)
            # torch (cuBLAS) reference timing for the same math
            A = a.t() if a_mn else a
            B = b if b_mn else b.t()
            for _ in range(3):
                torch.matmul(A, B)
            torch.cuda.synchronize()
            s.record()
            for _ in range(iters):
                torch.matmul(A, B)
            e.record()
            torch.cuda.synchronize()
            ms2 = s.elapsed_time(e) / iters
            print(json.dumps({"perf_cublas": name, "ms": ms2, "tflops": 2.0 * M * N * K / ms2 / 1e9}), flush=True)

        for bn in (256, 128):
            bench("fwd qkv vis", 25600, 2304, 768, False, False, ops.EPI_BF16, bn)
            bench("fwd fc vis", 25600, 3072, 768, False, False, ops.EPI_BF16, bn)
            bench("fwd proj vis", 25600, 768, 3072, False, False, ops.EPI_BF16, bn)
            bench("fwd qkv txt", 39424, 1536, 512, False, False, ops.EPI_BF16, bn)
            bench("dgrad fc vis", 25600, 768, 3072, False, True, ops.EPI_BF16, bn)
            bench("wgrad fc vis", 3072, 768, 25600, True, True, ops.EPI_F32_ATOMIC, bn)
            bench("wgrad qkv txt", 1536, 512, 39424, True, True, ops.EPI_F32_ATOMIC, bn)
        bench("square 8192", 8192, 8192, 8192, False, False, ops.EPI_BF16, 256)
    nbad = sum(1 for r in res if not r["ok"])
    print("GROUP %s: %d checks, %d bad" % (group, len(res), nbad), flush=True)


def main():
    if len(sys.argv) > 1:
        worker(sys.argv[1])
        return
    os.makedirs("gpurun_out", exist_ok=True)
    for g in GROUPS:
        t0 = time.time()
        with open("gpurun_out/gemm_probe_%s.log" % g, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), g], stdout=f, stderr=subprocess.STDOUT,
                                   timeout=240)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        tail = open("gpurun_out/gemm_probe_%s.log" % g).read().strip().splitlines()[-3:]
        print("== %s rc=%s %.1fs\n   %s" % (g, rc, time.time() - t0, "\n   ".join(tail)), flush=True)


if __name__ == "__main__":
    main()
