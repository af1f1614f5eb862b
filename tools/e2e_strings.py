"""End-to-end step rate when the captions arrive as STRINGS, i.e. through the reference's own call signature
`model({'images': ..., 'captions': [[str], ...]})` (clip.py:118-127): host C++ BPE tokenisation inside
TextTransformer.forward, H2D copies of the images from pinned memory, forward + ClipInfoCELoss + backward + FusedAdamW.
The real merges table is a download the reference does not ship, so the 277-merge test table and captions over its toy
vocabulary are used — the kernels see the same shapes.  Three legs: (a) strings tokenised inside forward (pinned ids,
asynchronous H2D), (b) the same batches through declip_b200.tokenizer.CaptionPipeline (worker thread tokenises and uploads
step i+1 while the device runs step i), (c) pre-tokenised ids resident on the device (the bench's `value` path).
Usage (GPU box): python tools/e2e_strings.py [steps]"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from declip_b200.loss_functions import ClipInfoCELoss  # noqa: E402
from declip_b200.model import model_entry  # noqa: E402
from declip_b200.optim import FusedAdamW  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b = 512
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
merges = os.path.join(ROOT, "tests", "golden", "bpe_small_merges.txt")
model = model_entry(dict(type='clip_vitb32', kwargs=dict(
    image_encode=dict(embed_dim=512),
    text_encode=dict(bpe_path=merges, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False), embed_dim=512),
    clip=dict(use_allgather=False)))).to(dev).train()
crit = ClipInfoCELoss()
opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
rng = random.Random(0)
words = ("a photo of the big red dog running in park two dogs are playing with ball on grass quick brown fox jumps over lazy "
         "man riding bicycle down street woman holding umbrella rain it's beautiful day coffee tea 2023 children flower "
         "mountains lakes rivers sunset airplane flying blue sky people walking beach ocean waves photography").split()
caps = [[[" ".join(rng.choice(words) for _ in range(rng.randint(6, 24))) + "."] for _ in range(b)] for _ in range(2)]
g = torch.Generator().manual_seed(1)
host = [torch.randn(b, 3, 224, 224, generator=g).pin_memory() for _ in range(2)]
stage = [torch.empty(b, 3, 224, 224, device=dev) for _ in range(2)]
loss_host = torch.zeros(steps + 4).pin_memory()


def step(i):
    s = i % 2
    stage[s].copy_(host[s], non_blocking=True)
    li, lt = model({"images": stage[s], "captions": caps[s]})       # strings -> C++ BPE -> ids -> H2D inside forward
    loss, _ = crit(li, lt)
    loss.backward()
    opt.step()
    model.logit_scale.data.clamp_(3.0, 6.0)
    opt.zero_grad(set_to_none=True)
    loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)


def step_inputs(inp, i):
    li, lt = model(inp)
    loss, _ = crit(li, lt)
    loss.backward()
    opt.step()
    model.logit_scale.data.clamp_(3.0, 6.0)
    opt.zero_grad(set_to_none=True)
    loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)


def timed(fn, n):
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s_.record()
    fn(n)
    e_.record()
    torch.cuda.synchronize()
    return s_.elapsed_time(e_) / n


for i in range(3):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
tok = model.encode_text.tokenizer
tok.tokenize([c[0] for c in caps[0]], 77)
t_tok = time.perf_counter() - t0
s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s_ev.record()
for i in range(steps):
    step(3 + i)
e_ev.record()
torch.cuda.synchronize()
ms = s_ev.elapsed_time(e_ev) / steps
# (b) CaptionPipeline: host text work + uploads of step i+1 overlap the device work of step i
from declip_b200.tokenizer import CaptionPipeline  # noqa: E402


def batches(n):
    for i in range(n):
        yield {"images": host[i % 2], "captions": caps[i % 2]}


def run_pipeline(n):
    for i, inp in enumerate(CaptionPipeline(batches(n), tok, dev)):
        step_inputs(inp, i)


run_pipeline(3)
ms_pipe = timed(run_pipeline, steps)
# (c) ids resident on the device
ids_dev = [tok.tokenize([c[0] for c in caps[s]], 77).to(dev) for s in range(2)]
imgs_dev = [h.to(dev) for h in host]


def run_ids(n):
    for i in range(n):
        step_inputs({"images": imgs_dev[i % 2], "captions": None, "token_ids": ids_dev[i % 2]}, i)


run_ids(3)
ms_ids = timed(run_ids, steps)
print(json.dumps({"e2e_from_strings_pairs_per_s": round(b / ms * 1e3), "ms_per_step": round(ms, 2), "steps": steps,
                  "pipeline_pairs_per_s": round(b / ms_pipe * 1e3), "pipeline_ms_per_step": round(ms_pipe, 2),
                  "resident_ids_pairs_per_s": round(b / ms_ids * 1e3), "resident_ids_ms_per_step": round(ms_ids, 2),
                  "strings_vs_ids": round(ms_ids / ms, 4), "pipeline_vs_ids": round(ms_ids / ms_pipe, 4),
                  "host_tokenize_ms_per_512_captions": round(t_tok * 1e3, 2), "h2d_bytes_per_step": b * 3 * 224 * 224 * 4 + b * 77 * 8,
                  "tokenizer": "C++ BPE (dc_bpe_tokenize), 277-merge test table", "last_loss": float(loss_host[steps + 2])}))
