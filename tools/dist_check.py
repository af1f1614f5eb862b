"""Multi-GPU parity (run under torchrun, one rank per GPU, NCCL): W ranks x b pairs with the gathered contrastive
head + flat-bucket gradient all-reduce must reproduce the single-process global-batch step (loss and gradients).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py
Prints DIST_CHECK_OK on rank 0 when everything agrees."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from declip_b200.dist import DistModule
    from declip_b200.loss_functions import ClipInfoCELoss
    from declip_b200.model import model_entry
    from oracle import synth   # test infrastructure: seeded weights / inputs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b = 8
    layers = 2

    def build(use_allgather):
        cfg = dict(type='clip_vitb32', kwargs=dict(
            image_encode=dict(embed_dim=512, layers=layers),
            text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                             embed_dim=512, transformer_layers=layers),
            clip=dict(use_allgather=use_allgather)))
        m = model_entry(cfg)
        m.load_state_dict(synth.clip_vit_state_dict(seed=3, v_layers=layers, t_layers=layers), strict=True)
        return m.to(dev).train()

    images = synth.synth_images(world * b, seed=3).to(dev)
    ids = synth.synth_token_ids(world * b, seed=3).to(dev)
    # ---- distributed step: rank r owns pairs [r*b, (r+1)*b)
    model = DistModule(build(True))
    crit = ClipInfoCELoss()
    sl = slice(rank * b, (rank + 1) * b)
    li, lt = model({"images": images[sl], "captions": None, "token_ids": ids[sl]})
    assert li.shape == (b, world * b), li.shape
    loss, labels = crit(li, lt)
    assert labels[0].item() == rank * b                                   # loss.py:45
    (loss / world).backward()                                             # clip_solver.py:418
    model.sync_gradients()
    loss_sum = (loss.detach() / world).clone()
    dist.all_reduce(loss_sum)
    torch.cuda.synchronize()
    ok = True
    msgs = []
    if rank == 0:
        # ---- single-process global batch on the same CUDA path
        ref = build(False)
        gi, gt = ref({"images": images, "captions": None, "token_ids": ids})
        gl, _ = ClipInfoCELoss()(gi, gt)
        gl.backward()
        torch.cuda.synchronize()
        d = abs(gl.item() - loss_sum.item())
        msgs.append("loss dist %.6f global %.6f |d| %.2e" % (loss_sum.item(), gl.item(), d))
        ok &= d < 2e-3
        # rank 0's strip is the first b rows of the global logits
        dl = (li - gi[:b]).abs().max().item()
        msgs.append("logit strip max |d| %.3e" % dl)
        ok &= dl < 5e-2
        pd, pr = dict(model.module.named_parameters()), dict(ref.named_parameters())
        worst = 1.0
        for k, p in pr.items():
            if p.grad is None:
                continue
            a, c = pd[k].grad.float().reshape(-1), p.grad.float().reshape(-1)
            cs = (torch.dot(a, c) / (a.norm() * c.norm() + 1e-20)).item()
            nr = (a.norm() / (c.norm() + 1e-20)).item()
            if cs < worst:
                worst = cs
            if not (cs > 0.995 and 0.97 < nr < 1.03):
                ok = False
                msgs.append("GRAD MISMATCH %s cos %.5f norm ratio %.4f" % (k, cs, nr))
        msgs.append("worst grad cosine %.6f" % worst)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0:
        print("\n".join(msgs))
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
