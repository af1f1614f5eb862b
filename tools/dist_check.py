"""Multi-GPU parity (run under torchrun, one rank per GPU, NCCL): W ranks x b pairs with the gathered contrastive
head + bucketed (bf16, per layer group) gradient all-reduce must reproduce the single-process global-batch step —
loss and every parameter gradient — checked on rank 0 against BOTH the same CUDA path run on the global batch and the
CPU fp32 oracle restatement of the reference (oracle/clip_ref.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port 29611 \
        tools/dist_check.py [--batch 8] [--layers 2] [--head strips|fused|both] [--oracle 1]
Prints DIST_CHECK_OK on rank 0 when everything agrees."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def cos(a, c):
    a, c = a.float().reshape(-1), c.float().reshape(-1)
    return (torch.dot(a, c) / (a.norm() * c.norm() + 1e-20)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)   # the fused head needs 32 | b when gathering
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--head", default="both", choices=["strips", "fused", "both"])
    ap.add_argument("--oracle", type=int, default=1)
    args = ap.parse_args()
    from declip_b200.dist import DistModule
    from declip_b200.loss_functions import ClipInfoCELoss
    from declip_b200.model import model_entry
    from oracle import synth   # test infrastructure: seeded weights / inputs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b, layers = args.batch, args.layers
    sd = synth.clip_vit_state_dict(seed=3, v_layers=layers, t_layers=layers)

    def build(use_allgather, fused):
        cfg = dict(type='clip_vitb32', kwargs=dict(
            image_encode=dict(embed_dim=512, layers=layers),
            text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                             embed_dim=512, transformer_layers=layers),
            clip=dict(use_allgather=use_allgather, fused_head=fused)))
        m = model_entry(cfg)
        m.load_state_dict(sd, strict=True)
        return m.to(dev).train()

    images = synth.synth_images(world * b, seed=3).to(dev)
    ids = synth.synth_token_ids(world * b, seed=3).to(dev)
    sl = slice(rank * b, (rank + 1) * b)
    ok, msgs = True, []
    ref_grads = ref_loss = None
    if rank == 0:
        # ---- single-process global batch on the same CUDA path (compat head)
        ref = build(False, False)
        gi, gt = ref({"images": images, "captions": None, "token_ids": ids})
        gl, _ = ClipInfoCELoss()(gi, gt)
        gl.backward()
        torch.cuda.synchronize()
        ref_loss = gl.item()
        ref_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}
        ref_strip = gi[:b].detach().clone()
        del ref
        if args.oracle:
            from oracle import clip_ref
            out = clip_ref.clip_step(sd, images.cpu(), ids.cpu())
            orc_loss, orc_grads = out["loss"].item(), out["grads"]
    for head in (["strips", "fused"] if args.head == "both" else [args.head]):
        # ---- distributed step: rank r owns pairs [r*b, (r+1)*b)
        model = DistModule(build(True, head == "fused"), bucket_layers=1)
        crit = ClipInfoCELoss()
        for it in range(2):                 # twice: the second pass runs with p.grad re-created after zero_grad(set_to_none)
            model.zero_grad(set_to_none=True)
            li, lt = model({"images": images[sl], "captions": None, "token_ids": ids[sl]})
            assert li.shape == (b, world * b), li.shape
            loss, labels = crit(li, lt)
            assert labels[0].item() == rank * b                                   # loss.py:45
            (loss / world).backward()                                             # clip_solver.py:418
            model.sync_gradients()
        loss_sum = (loss.detach() / world).clone()
        dist.all_reduce(loss_sum)
        torch.cuda.synchronize()
        if rank == 0:
            d = abs(ref_loss - loss_sum.item())
            msgs.append("[%s] W=%d b=%d L=%d: loss dist %.6f global %.6f |d| %.2e" % (head, world, b, layers, loss_sum.item(), ref_loss, d))
            ok &= d < 2e-3
            if head == "strips":
                dl = (li - ref_strip).abs().max().item()       # rank 0's strip is the first b rows of the global logits
                msgs.append("[%s] logit strip max |d| %.3e" % (head, dl))
                ok &= dl < 5e-2
            pd = dict(model.module.named_parameters())
            worst, worst_o = (1.0, ""), (1.0, "")
            for k, g in ref_grads.items():
                a = pd[k].grad
                cs, nr = cos(a, g), (a.float().norm() / (g.float().norm() + 1e-20)).item()
                if cs < worst[0]:
                    worst = (cs, k)
                if not (cs > 0.995 and 0.97 < nr < 1.03):       # bf16 buckets: 2^-9 relative per element
                    ok = False
                    msgs.append("[%s] GRAD MISMATCH %s cos %.5f norm ratio %.4f" % (head, k, cs, nr))
                if args.oracle:
                    co = cos(a.cpu(), orc_grads[k])
                    if co < worst_o[0]:
                        worst_o = (co, k)
            msgs.append("[%s] worst grad cosine vs global CUDA step %.6f (%s)" % (head, worst[0], worst[1]))
            if args.oracle:
                do = abs(orc_loss - loss_sum.item())
                msgs.append("[%s] vs CPU oracle: |d loss| %.2e, worst grad cosine %.5f (%s)" % (head, do, worst_o[0], worst_o[1]))
                ok &= do < 5e-3 and worst_o[0] > 0.98
        del model
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0:
        print("\n".join(msgs))
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
