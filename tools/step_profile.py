"""Per-kernel device time of the training step measured in situ (torch.profiler / CUPTI, no replay, no cache
flush): python tools/step_profile.py [--config clip|declip|filip|res50] [--batch 512] -> table sorted by total time."""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from declip_b200.optim import FusedAdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="clip")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--top", type=int, default=32)
ap.add_argument("--head", default="fused")
args = ap.parse_args()
b = args.batch
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, run, host_inputs = bench.build_workload(args.config, dev, b, 1, args.head)
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
inp = {k: v.to(dev) for k, v in host_inputs(torch.Generator().manual_seed(0)).items()}


def step():
    loss = run(model, inp)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    name = re.sub(r"\(.*", "", ev.name)[:110]
    tot[name] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    cnt[name] += 1
T = sum(tot.values())
print("# in-situ kernel time per step, config %s (torch.profiler, %d steps, batch %d): total %.2f ms/step, %d launches/step"
      % (args.config, N, b, T / N / 1e3, sum(cnt.values()) / N))
print("| kernel | launches/step | ms/step | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:args.top]:
    print("| `%s` | %.0f | %.3f | %.1f%% | %.1f |" % (k, cnt[k] / N, v / N / 1e3, 100 * v / T, v / cnt[k]))
