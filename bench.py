#!/usr/bin/env python
"""bench.py — image-text pairs/s of one CLIP ViT-B/32 training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL)

native arm   : the CUDA path of this repo through its public API — CLIP.forward (dict in, two logit
               strips out) + ClipInfoCELoss + backward + flat-bucket gradient all-reduce + AdamW step —
               on synthetic 224x224 images / 77-token ids, random-init ViT-B/32 + 12-layer text tower,
               bf16 storage / fp32 accumulate, per-GPU batch 512 (BASELINE configs[1]: global 4096 on 8 GPUs).
               `value` = inputs resident in HBM; `e2e` = pinned-host inputs copied H2D every step
               (double-buffered on a copy stream) + D2H read of the loss, all inside the timed region.
reference arm: the reference's own CPU implementation of the same step (oracle restatement of its modules,
               oracle/clip_ref.py — the Python reference cannot travel to the GPU box), all host cores,
               a bounded sample (bs 32) per step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_PAIR = 43.9   # fwd 14.78 + 2x bwd, frozen conv1 bwd skipped
CPU_SAMPLE_BATCH = 32         # BASELINE configs[0]


_T0 = time.time()


def log(msg):
    if os.environ.get("RANK", "0") == "0":
        print("[bench %.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


# NCCL prints "NCCL version ..." on STDOUT when NCCL_DEBUG=VERSION (set in some images): keep stdout to the one JSON line
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def pick_cpu_threads():
    """The reference's CPU path is PyTorch eager; more threads is not monotonically faster on these GEMM sizes
    (128 threads measured 40x SLOWER than 8 on the GPU box).  Time the path's dominant op (the c_fc Linear of a
    bs-32 ViT step, [1600x768]x[768x3072]) at a few thread counts and use the fastest — the best the host can do."""
    import torch
    total = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, total) if c <= total})
    a, w = torch.randn(1600, 768), torch.randn(3072, 768)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.linear(a, w)
        t0 = time.perf_counter()
        for _ in range(5):
            torch.nn.functional.linear(a, w)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    log("cpu arm: picked %d of %d host threads" % (best, total))
    return best


def cpu_reference_steps(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """Times the oracle port of the reference step (fwd + ClipInfoCELoss + bwd, fp32) on the host cores.  The optimiser
    step is left out on purpose: at the sample batch of 32 a full AdamW over 151 M parameters (several seconds on a
    host CPU) would be amortised over 16x fewer pairs than in the native arm's b = 512 step and inflate the ratio."""
    import torch
    from oracle import clip_ref, synth
    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    sd = synth.clip_vit_state_dict(seed=0)
    images = synth.synth_images(batch, seed=0)
    ids = synth.synth_token_ids(batch, seed=0)
    log("cpu arm: %d threads, warm-up" % cores)
    for _ in range(warmup):
        tw = time.perf_counter()
        clip_ref.clip_step(sd, images, ids)
        tw = time.perf_counter() - tw
    steps = max(1, min(steps, int(25.0 / max(tw, 1e-3))))     # bound the sample to ~25 s of CPU work
    log("cpu arm: warm step %.2fs -> timing %d steps" % (tw, steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        clip_ref.clip_step(sd, images, ids)
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d steps of bs %d (fwd+loss+bwd, fp32, oracle/clip_ref.py) in %.1f s" % (steps, batch, dt),
            "ms_per_step": 1e3 * dt / steps}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    cb = cpu_reference_steps(steps, warm)
    steps = int(cb["sample"].split()[0])
    line = {
        "impl": "reference", "metric": "image-text pairs/sec", "value": cb["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CLIP ViT-B/32 + 12L text transformer, one training step (fwd+ClipInfoCELoss+bwd)",
                   "sample_batch": CPU_SAMPLE_BATCH, "seq_len": 77, "image": "3x224x224"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ native arm
def synthetic_token_ids(batch, gen, ctx=77):
    """int64 [B,77]: SOT 49407, random body of length U[8,75], EOT 49408 (= max id, found by argmax), zero pad —
    the layout TextTransformer.tokenize produces."""
    import torch
    ids = torch.zeros(batch, ctx, dtype=torch.int64)
    lens = torch.randint(8, ctx - 1, (batch,), generator=gen)
    body = torch.randint(1, 49000, (batch, ctx), generator=gen)
    pos = torch.arange(ctx).unsqueeze(0)
    ids = torch.where((pos >= 1) & (pos <= lens.unsqueeze(1)), body, ids)
    ids[:, 0] = 49407
    ids[torch.arange(batch), lens + 1] = 49408
    return ids


def run_native(args):
    import torch
    import torch.distributed as dist
    from declip_b200 import _lib, ops
    from declip_b200.dist import DistModule
    from declip_b200.loss_functions import ClipInfoCELoss
    from declip_b200.model import model_entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b = args.batch
    torch.manual_seed(1234)
    cfg = dict(type='clip_vitb32', kwargs=dict(
        image_encode=dict(embed_dim=512),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=512),
        clip=dict(use_allgather=True)))
    model = DistModule(model_entry(cfg).to(dev).train())
    crit = ClipInfoCELoss()
    from declip_b200.optim import FusedAdamW
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1)
    # synthetic inputs (pinned host copies for the e2e leg)
    g = torch.Generator().manual_seed(100 + rank)
    host_imgs = [torch.randn(b, 3, 224, 224, generator=g).pin_memory() for _ in range(2)]
    host_ids = [synthetic_token_ids(b, g).pin_memory() for _ in range(2)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    dev_ids = [h.to(dev) for h in host_ids]
    loss_host = torch.zeros(max(args.steps, 1), dtype=torch.float32).pin_memory()

    def step(images, ids):
        li, lt = model({"images": images, "captions": None, "token_ids": ids})
        loss, _ = crit(li, lt)
        (loss / world).backward()                      # clip_solver.py:418
        model.sync_gradients()                         # dist.py:76-83
        opt.step()
        model.module.logit_scale.data.clamp_(3.0, 6.0)  # grad_clip config: logit_scale param clamp, clip_solver.py:507-522
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(k)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def loop_resident(k):
        for i in range(k):
            step(dev_imgs[i % 2], dev_ids[i % 2])

    copy_stream = torch.cuda.Stream()
    stage_imgs = [torch.empty_like(dev_imgs[0]) for _ in range(2)]
    stage_ids = [torch.empty_like(dev_ids[0]) for _ in range(2)]

    def loop_e2e(k):
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream()

        def upload(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(done[s])
                stage_imgs[s].copy_(host_imgs[s], non_blocking=True)
                stage_ids[s].copy_(host_ids[s], non_blocking=True)
                ready[s].record(copy_stream)
        upload(0)
        for i in range(k):
            s = i % 2
            if i + 1 < k:
                upload(i + 1)
            main.wait_event(ready[s])
            loss = step(stage_imgs[s], stage_ids[s])
            done[s].record(main)
            loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)

    log("model + inputs ready; warm-up")
    for _ in range(max(args.warmup, 3)):
        step(dev_imgs[0], dev_ids[0])
    torch.cuda.synchronize()
    log("warm-up done; timing %d resident steps" % args.steps)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(loop_resident, args.steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    log("resident: %.2f ms/step" % (ms / args.steps))
    e2e = None
    if not args.no_e2e:
        loop_e2e(2)   # warm the copy path
        torch.cuda.synchronize()
        log("timing e2e")
        ms_e2e = timed(loop_e2e, args.steps)
        e2e = {"value": world * b * args.steps / (ms_e2e / 1e3), "unit": "pairs/s",
               "h2d_bytes_per_step": host_imgs[0].numel() * 4 + host_ids[0].numel() * 8, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / args.steps, "last_loss": float(loss_host[args.steps - 1])}

    # ---- roofline of the dominant kernel: the tcgen05 GEMM, at its largest-share launch shape (ViT c_fc forward)
    roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks.get("bf16_tflops"), "measured burst (MEASURED_PEAKS.json)") if peaks.get("bf16_tflops") else (
            1590.0, "fallback (B200_PROFILING.md)")
        M, N, K = b * 50, 3072, 768
        a = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        bias = torch.zeros(N, device=dev)
        o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BF16_GELU, out=o1, out2=o2)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        iters = 20
        s.record()
        for _ in range(iters):
            ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BF16_GELU, out=o1, out2=o2)
        e.record()
        torch.cuda.synchronize()
        kms = s.elapsed_time(e) / iters
        ach = 2.0 * M * N * K / (kms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                # DRAM bytes of this launch shape from the ncu --set full capture (profiles/r01_ncu_gemm_gelu_v7.md:
                # dram read 44.3 MB + write 260.3 MB at b=512; algorithmic 44.0 + 314.6 MB, part of the output is still
                # in L2 at kernel end) scaled to the batch in use
                "traffic": 304.6e6 * (b / 512.0),
                "kernel": "gemm2_bf16_kernel<K-major,K-major> (cta_group::2) c_fc+QuickGELU fwd M=%d N=%d K=%d" % (M, N, K),
                "kernel_ms": kms, "peak_source": which,
                "step_mfu": {"achieved_tflops_per_gpu": b * args.steps / (ms / 1e3) * TRAIN_GFLOP_PER_PAIR / 1e3,
                             "peak_sustained": peaks.get("bf16_tflops_sustained"),
                             "frac_of_sustained": (b * args.steps / (ms / 1e3) * TRAIN_GFLOP_PER_PAIR / 1e3) /
                             peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None}}

    log("roofline probe done")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # separate process: its thread pool / a slow host cannot stall or perturb the GPU arm
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4",
                                "--warmup", "1"], capture_output=True, text=True, timeout=240)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:   # reported, never silently dropped
            cpu = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "cpu baseline leg failed: %r" % (ex,)}
        log("cpu baseline done")

    if rank == 0:
        line = {
            "metric": "image-text pairs/sec", "value": world * b * args.steps / (ms / 1e3), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "CLIP ViT-B/32 + 12L text transformer training step: fwd + ClipInfoCELoss + bwd + "
                                   "grad all-reduce + AdamW (BASELINE configs[1], per-GPU batch %d)" % b,
                       "global_batch": world * b, "seq_len": 77, "image": "3x224x224 fp32", "parallelism": "dp%d" % world,
                       "optimizer": "declip_b200.optim.FusedAdamW (one multi-tensor launch)", "l2": "inputs+activations >> 126 MB L2 (no flush needed)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
