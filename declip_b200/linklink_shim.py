"""Replacement for the reference's `linklink` package (linklink/__init__.py:13-71) on torch.distributed + NCCL.

Rank / world size come from torch.distributed (the reference reads SLURM env vars, which breaks torchrun launches,
linklink/__init__.py:22-27); `initialize` uses the torchrun env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
import functools
import os

import torch
import torch.distributed as dist

allreduce = dist.all_reduce
allgather = dist.all_gather
broadcast = dist.broadcast
init_process_group = dist.init_process_group
allreduce_async = functools.partial(dist.all_reduce, async_op=True)


def synchronize():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_local_rank():
    return int(os.environ.get("LOCAL_RANK", get_rank() % max(1, torch.cuda.device_count() or 1)))


def barrier():
    if get_world_size() > 1:
        dist.barrier()


def initialize(backend='nccl'):
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12345")
    rank = int(os.environ.get("RANK", os.environ.get("SLURM_PROCID", "0")))
    world = int(os.environ.get("WORLD_SIZE", os.environ.get("SLURM_NTASKS", "1")))
    if backend == 'nccl' and torch.cuda.is_available():
        torch.cuda.set_device(get_local_rank() if "LOCAL_RANK" in os.environ else rank % torch.cuda.device_count())
    dist.init_process_group(backend=backend, rank=rank, world_size=world)


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
