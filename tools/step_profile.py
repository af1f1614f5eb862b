"""Per-kernel device time of the training step measured in situ (torch.profiler / CUPTI, no replay, no cache
flush): python tools/step_profile.py [batch] -> table sorted by total time."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from declip_b200.loss_functions import ClipInfoCELoss  # noqa: E402
from declip_b200.model import model_entry  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
cfg = dict(type='clip_vitb32', kwargs=dict(
    image_encode=dict(embed_dim=512),
    text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                     embed_dim=512), clip=dict(use_allgather=True)))
torch.manual_seed(0)
model = model_entry(cfg).to(dev).train()
crit = ClipInfoCELoss()
from declip_b200.optim import FusedAdamW  # noqa: E402
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
images = torch.randn(b, 3, 224, 224, device=dev)
ids = torch.zeros(b, 77, dtype=torch.long, device=dev)
ids[:, 0] = 49407
ids[:, 1:20] = torch.randint(1, 49000, (b, 19), device=dev)
ids[:, 20] = 49408


def step():
    li, lt = model({"images": images, "captions": None, "token_ids": ids})
    loss, _ = crit(li, lt)
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
    name = re.sub(r"\(.*", "", ev.name)[:100]
    tot[name] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    cnt[name] += 1
T = sum(tot.values())
print("# in-situ kernel time per step (torch.profiler, %d steps, batch %d): total %.2f ms/step" % (N, b, T / N / 1e3))
print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:28]:
    print("| `%s` | %.0f | %.3f | %.1f%% |" % (k, cnt[k] / N, v / N / 1e3, 100 * v / T))
