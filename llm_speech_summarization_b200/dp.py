"""Data-parallel plumbing: one process per GPU, utterances sharded across ranks, no data-path collective in the
forward / loss path (the LLM is frozen and replicated). The only collective this path ever
needs is the SUM all-reduce of the encoder+projector gradients once per optimizer step (NCCL on the GPUs, gloo
in the CPU tests); forward-only configurations need none.

The reference has no distributed code at all (REF/README.md:29,86); its accumulation semantics are
"total_loss / grad_accum_interval, summed over grad_accum_interval utterances" (REF/trainer.py:372-384), which
sharding preserves when every rank scales by the GLOBAL window and gradients are summed.
"""
from __future__ import annotations

import os
from typing import Iterable, List

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            # The only collective of this path is the gradient all-reduce, hidden under the encoder backward: it needs
            # ~20 GB/s, not the whole chip. Few NCCL CTAs = few SMs taken from the persistent GEMMs it overlaps with
            # (EncoderTrainer.comm_sms reads the same variable). An explicit setting in the environment wins.
            os.environ.setdefault("NCCL_MAX_CTAS", "8")
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


_NATIVE_COMM = {"world": 1}


def ensure_native_comm() -> int:
    """Give the library's current handle its own NCCL communicator over the ranks of the default process group
    (include/b2s.h: b2s_comm_unique_id / b2s_comm_init): rank 0 mints the id, torch.distributed carries it to the others.
    The gradient all-reduce of the training step then runs through the C ABI (b2s_allreduce_grads) on the library's
    communication stream; torch.distributed stays the plumbing (rendezvous, barriers, scalar reductions).
    Returns the communicator's world size (1 = not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1
    if _NATIVE_COMM["world"] == dist.get_world_size():
        return _NATIVE_COMM["world"]
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
    if rank == 0:
        _lib.check(lib.b2s_comm_unique_id(buf), "comm_unique_id")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0)
    buf = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(box[0])
    _lib.check(lib.b2s_comm_init(buf, rank, world), "comm_init")
    _NATIVE_COMM["world"] = world
    return world


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Utterance i goes to rank i mod world (round-robin keeps per-rank work equal to within one utterance)."""
    return list(range(rank, n_items, world))


def local_accum_scale(grad_accum_interval: int) -> float:
    """Every rank divides its per-utterance loss by the GLOBAL accumulation window (REF/trainer.py:373); summing
    the gradients over ranks then reproduces the single-GPU accumulated gradient."""
    return 1.0 / float(grad_accum_interval)


def allreduce_sum_(tensors: Iterable[torch.Tensor], bucket_bytes: int = 64 << 20) -> None:
    """In-place SUM all-reduce of a list of gradient tensors, coalesced into ~bucket_bytes flat buckets so the
    launch count stays small (NVSwitch bandwidth is uniform; buckets are sized for latency, not link count)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        bucket, size = [], 0

    for t in tensors:
        if bucket and (bucket[0].dtype != t.dtype or size + t.numel() * t.element_size() > bucket_bytes):
            flush()
        bucket.append(t)
        size += t.numel() * t.element_size()
    flush()


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
