"""Per-step device time of the resident CLIP step (CUDA events around every step) with the Python garbage collector
enabled and disabled, and the collector's own pauses (gc.callbacks): is the 1-in-8 slow step a gen-2 collection on the
launching thread?   python tools/step_jitter.py [--steps 48]"""
import argparse
import gc
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from declip_b200.optim import FusedAdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=48)
ap.add_argument("--config", default="clip")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, run, host_inputs = bench.build_workload(args.config, dev, 512, 1, "fused")
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
inps = [{k: v.to(dev) for k, v in host_inputs(torch.Generator().manual_seed(i)).items()} for i in range(2)]
pauses = []
_t0 = [0.0]


def cb(phase, info):
    if phase == "start":
        _t0[0] = time.perf_counter()
    else:
        pauses.append((info["generation"], (time.perf_counter() - _t0[0]) * 1e3))


gc.callbacks.append(cb)


def step(i):
    loss = run(model, inps[i & 1])
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


def measure(label):
    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    pauses.clear()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    host = []
    evs[0].record()
    for i in range(args.steps):
        t = time.perf_counter()
        step(i)
        evs[i + 1].record()
        host.append((time.perf_counter() - t) * 1e3)
    torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    srt = sorted(ms)
    print("%s: mean %.2f median %.2f min %.2f max %.2f ms; steps > median + 1 ms: %s" % (
        label, sum(ms) / len(ms), srt[len(ms) // 2], srt[0], srt[-1],
        [(i, round(m, 1)) for i, m in enumerate(ms) if m > srt[len(ms) // 2] + 1.0]))
    print("   host enqueue ms/step: mean %.2f max %.2f; gc pauses (generation, ms): %s" % (
        sum(host) / len(host), max(host), [(g, round(p, 2)) for g, p in pauses if p > 0.2]))


measure("gc enabled ")
gc.collect()
gc.disable()
measure("gc disabled")
gc.enable()
gc.collect()
gc.freeze()
measure("gc frozen  ")
