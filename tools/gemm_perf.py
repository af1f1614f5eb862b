"""Times the GEMM at the hot-path shapes with every epilogue (CUDA events, inputs > L2 or rotated).
    python tools/gemm_perf.py [only_name_substring]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from declip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
only = sys.argv[1] if len(sys.argv) > 1 else ""


def mk(r, c):
    return (torch.randn(r, c, device=dev) * 0.5).bfloat16()


def bench(name, fn, flops, iters=20):
    if only and only not in name:
        return
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
    print(json.dumps({"name": name, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}), flush=True)


M = 25600
for (N, K, tag) in ((2304, 768, "qkv"), (768, 768, "out"), (3072, 768, "fc"), (768, 3072, "proj")):
    a, w = mk(M, K), mk(N, K)
    bias = torch.randn(N, device=dev)
    aux = mk(M, N)
    o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    bench("vis %s plain" % tag, lambda: ops.gemm(a, w, out=o1), fl)
    bench("vis %s bias" % tag, lambda: ops.gemm(a, w, bias=bias, out=o1), fl)
    bench("vis %s resid" % tag, lambda: ops.gemm(a, w, bias=bias, aux=aux, epilogue=ops.EPI_BF16_RESID, out=o1), fl)
    bench("vis %s gelu" % tag, lambda: ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BF16_GELU, out=o1, out2=o2), fl)
    # dgrad: dx[M,K] = dy[M,N] W[N,K]
    dy = mk(M, N)
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    auxk = mk(M, K)
    bench("vis %s dgrad" % tag, lambda: ops.gemm(dy, w, b_mn_major=True, out=dx), fl)
    bench("vis %s dgrad_dgelu" % tag, lambda: ops.gemm(dy, w, b_mn_major=True, aux=auxk, epilogue=ops.EPI_BF16_DGELU, out=dx), fl)
    dw = torch.zeros(N, K, device=dev)
    bench("vis %s wgrad" % tag, lambda: ops.gemm(dy, a, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=dw), fl)
    for bn in (128, 256):
        bench("vis %s wgrad bn%d" % (tag, bn), lambda: ops.gemm(dy, a, a_mn_major=True, b_mn_major=True,
                                                               epilogue=ops.EPI_F32_ATOMIC, out=dw, block_n=bn), fl)
