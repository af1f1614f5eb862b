"""CPU, gloo, world_size 2: the host-side logic of the N>1 path (declip_b200/dist.py and the rank/label
plumbing) — no CUDA kernels are involved."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from declip_b200 import dist as ddist
        from declip_b200 import functions as F_
        from declip_b200 import linklink_shim as link
        assert ddist.get_rank() == rank and ddist.get_world_size() == world
        assert link.get_rank() == rank and link.get_world_size() == world
        assert F_.dist_info() == (rank, world)
        torch.manual_seed(rank)                      # different init per rank -> broadcast must equalise
        net = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.LayerNorm(8))
        m = ddist.DistModule(net)
        w = net[0].weight.detach().clone()
        gathered = [torch.zeros_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        assert all(torch.equal(g, gathered[0]) for g in gathered)            # dist.py:85-88 semantics
        x = torch.full((4, 8), float(rank + 1))
        (m(x).sum() / world).backward()                                        # loss pre-divided: clip_solver.py:418
        local = net[0].bias.grad.clone()
        m.sync_gradients()
        tot = local.clone()
        dist.all_reduce(tot)
        assert torch.allclose(net[0].bias.grad, tot)                           # SUM == mean of per-rank grads
        # reference label rule (loss.py:42-45): rank * bs + arange(bs) when the strip is not square
        bs = 4
        labels = rank * bs + torch.arange(bs)
        assert labels[0].item() == rank * bs
        ddist.barrier()
        link.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_dist_module_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
