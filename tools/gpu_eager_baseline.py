"""The honest GPU baseline: the reference's OWN PyTorch modules (staged copy oracle/_ref, unmodified) in
eager mode on the B200 — fp32 and under torch.autocast(bf16) — one CLIP ViT-B/32 training step (fwd + ClipInfoCELoss + bwd
+ torch.optim.AdamW), same synthetic shapes as bench.py.  This, not the CPU arm, is what a user of the reference would
otherwise run on this GPU.  TEST / MEASUREMENT INFRASTRUCTURE: nothing here is on the product path.
    python tools/gpu_eager_baseline.py [batch] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_harness, synth  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
ref_harness.setup()
model = ref_harness.build_clip_vitb32(512).to(dev).train()
crit = ref_harness.clip_loss_fn()
ids = synth.synth_token_ids(b, seed=0).to(dev)
ref_harness.set_token_ids(model, ids)
images = synth.synth_images(b, seed=0).to(dev)
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
res = {"batch": b, "gpu": torch.cuda.get_device_name(0), "reference_root": ref_harness.REF_ROOT}
for name, ctx in (("fp32", None), ("fp32_tf32", "tf32"), ("autocast_bf16", torch.bfloat16)):
    torch.backends.cuda.matmul.allow_tf32 = ctx == "tf32"
    torch.backends.cudnn.allow_tf32 = ctx == "tf32"

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ctx is torch.bfloat16):
            li, lt = model({"images": images, "captions": [["x"]] * b})
            loss, _ = crit(li.float(), lt.float())
        loss.backward()
        opt.step()
        return loss
    try:
        for _ in range(2):
            step()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(steps):
            loss = step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
        res[name] = {"ms_per_step": round(ms, 2), "pairs_per_s": round(b / ms * 1e3), "loss": round(loss.item(), 4),
                     "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    except Exception as ex:     # noqa: BLE001
        res[name] = {"error": repr(ex)[:300]}
print(json.dumps(res))
