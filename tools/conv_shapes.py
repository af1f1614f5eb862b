"""Per-shape device time of the Python-composed GEMM / conv / BatchNorm calls of the ResNet tower (CUDA events around
every call, towers serialised): python tools/conv_shapes.py [--batch 512] -> table sorted by total time with the
achieved TFLOP/s and GB/s of each launch shape."""
import argparse
import collections
import os
import sys

os.environ["DECLIP_B200_TOWER_STREAMS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from declip_b200 import _lib, ops  # noqa: E402
from declip_b200.optim import FusedAdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--top", type=int, default=60)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, run, host_inputs = bench.build_workload("res50", dev, args.batch, 1, "fused")
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
inp = {k: v.to(dev) for k, v in host_inputs(torch.Generator().manual_seed(0)).items()}
records = []
recording = [False]
_gemm = ops.gemm


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, **kw):
    if not recording[0]:
        return _gemm(a, b, a_mn_major=a_mn_major, b_mn_major=b_mn_major, **kw)
    K, M = a.shape if a_mn_major else a.shape[::-1]
    N = b.shape[1] if b_mn_major else b.shape[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = _gemm(a, b, a_mn_major=a_mn_major, b_mn_major=b_mn_major, **kw)
    e1.record()
    osz = out.element_size()
    records.append(("gemm M=%d N=%d K=%d %s%s epi=%s" % (M, N, K, "T" if a_mn_major else "N", "T" if b_mn_major else "N",
                                                         kw.get("epilogue", 0)), e0, e1, 2.0 * M * N * K,
                    2.0 * (M * K + N * K) + osz * M * N))
    return out


ops.gemm = gemm
lib = _lib.load()


class Timed:
    def __init__(self, name, fn, desc):
        self.name, self.fn, self.desc = name, fn, desc

    def __call__(self, *a):
        if not recording[0]:
            return self.fn(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = self.fn(*a)
        e1.record()
        label, flops, nbytes = self.desc(a)
        records.append((self.name + " " + label, e0, e1, flops, nbytes))
        return rc


class LibProxy:
    def __init__(self, lib):
        self._lib = lib
        self._w = {
            "dc_conv3x3_igemm": lambda a: ("B=%d H=%d W=%d C=%d Cout=%d" % a[3:8], 2.0 * a[3] * a[4] * a[5] * 9 * a[6] * a[7],
                                          2.0 * a[3] * a[4] * a[5] * (a[6] + a[7])),
            "dc_conv3x3_wgrad_igemm": lambda a: ("B=%d H=%d W=%d C=%d Cout=%d" % a[3:8], 2.0 * a[3] * a[4] * a[5] * 9 * a[6] * a[7],
                                                2.0 * a[3] * a[4] * a[5] * (a[6] + a[7])),
            "dc_bn2d_fwd": lambda a: ("rows=%d C=%d res=%d" % (a[10], a[11], a[3] is not None), 0.0,
                                      2.0 * a[10] * a[11] * (3 + (a[3] is not None))),
            "dc_bn2d_bwd": lambda a: ("rows=%d C=%d relu=%d dres=%d" % (a[11], a[12], a[13], a[7] is not None), 0.0,
                                      2.0 * a[11] * a[12] * (2 + 2 * a[13] + 2 + 1 + (a[7] is not None))),
            "dc_im2col3x3": lambda a: ("B=%d H=%d W=%d C=%d" % a[2:6], 0.0, 2.0 * a[2] * a[3] * a[4] * a[5] * 10),
            "dc_col2im3x3": lambda a: ("B=%d H=%d W=%d C=%d" % a[2:6], 0.0, 2.0 * a[2] * a[3] * a[4] * a[5] * 10),
            "dc_avgpool2": lambda a: ("B=%d H=%d W=%d C=%d bwd=%d" % a[2:7], 0.0, 2.0 * a[2] * a[3] * a[4] * a[5] * 1.25),
        }

    def __getattr__(self, k):
        f = getattr(self._lib, k)
        if k in self._w:
            return Timed(k, f, self._w[k])
        return f


proxy = LibProxy(lib)
_lib_for = ops.lib_for
ops.lib_for = lambda t: (_lib_for(t), proxy)[1]


def step():
    loss = run(model, inp)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
recording[0] = True
N = 2
for _ in range(N):
    step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, e0, e1, fl, nb in records:
    t = e0.elapsed_time(e1)
    a = agg.setdefault(name, [0, 0.0, fl, nb])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values()) / N
print("# res50 tower, per-shape device time (events around each call, towers serialised), batch %d: %.2f ms/step in %d calls"
      % (args.batch, tot, len(records) // N))
print("| call | calls/step | ms/step | avg us | TFLOP/s | GB/s (algorithmic) |\n|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:args.top]:
    avg = a[1] / a[0]
    print("| %s | %d | %.3f | %.1f | %.0f | %.0f |" % (k, a[0] // N, a[1] / N, avg * 1e3, a[2] / avg / 1e9, a[3] / avg / 1e6))
