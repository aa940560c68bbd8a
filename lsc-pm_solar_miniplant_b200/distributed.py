"""Multi-GPU sharding of the hot path: one process per GPU, batches partitioned over ranks, and ONE
collective on the small integer tally block — an all-gather when ranks own whole batches (disjoint rows), an
all-reduce when they split photon ranges (NCCL over NVLink on GPUs, gloo in the CPU tests).  Photons are independent, so there is no data-path exchange; the reference's only parallelism
is the same photon-level split over CPU processes (``scene.simulate(num_rays, workers)``, reference
``miniplant/simulation_runner.py:62``).

Because every photon's random stream is keyed by (seed, batch id, photon index), the reduced tallies are
bit-identical for any world size.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Sequence

import numpy as np


def env_rank() -> tuple:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not distributed."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: Optional[str] = None):
    """Initialise ``torch.distributed`` from the torchrun environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE)."""
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    if world == 1 and "MASTER_ADDR" not in os.environ:
        return rank, local_rank, world
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {"device_id": torch.device(f"cuda:{local_rank}")} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def shard_batches(n_batches: int, rank: int, world: int) -> np.ndarray:
    """Round-robin ownership ``batch b -> rank b mod world`` (time points of a day are neighbours, so
    round-robin also balances the per-batch cost)."""
    return np.arange(rank, n_batches, world, dtype=np.int64)


def shard_photons(photons_per_batch: int, rank: int, world: int) -> tuple:
    """Contiguous photon sub-range of every batch for this rank: (begin, count)."""
    base, extra = divmod(int(photons_per_batch), world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def all_reduce_tallies(tallies):
    """Sum the int64 tally block over all ranks, in place (no-op when not distributed)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tallies, op=dist.ReduceOp.SUM)
    return tallies


def all_gather_rows(part, n_batches: int, rank: int, world: int):
    """Rows of a round-robin partition (``shard_batches``) -> the full ``[n_batches, stride]`` block on every rank.

    Every row has exactly one owner, so nothing needs adding: an all-gather moves half the bytes of the all-reduce it
    replaces and does no arithmetic.  Ranks pad their share to ``ceil(n_batches / world)`` rows; rank r's i-th row is
    batch ``r + i * world``."""
    import torch
    import torch.distributed as dist

    per = -(-n_batches // world)
    stride = part.shape[1]
    padded = part
    if part.shape[0] < per:
        padded = torch.zeros((per, stride), dtype=part.dtype, device=part.device)
        padded[:part.shape[0]] = part
    if not (dist.is_available() and dist.is_initialized() and world > 1):
        return padded[:n_batches].clone()
    gathered = torch.empty((world, per, stride), dtype=part.dtype, device=part.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(gathered.view(world * per, stride), padded.contiguous())
    else:   # gloo (CPU tests)
        dist.all_gather(list(gathered.unbind(0)), padded.contiguous())
    return gathered.permute(1, 0, 2).reshape(per * world, stride)[:n_batches].contiguous()


def trace_partitioned(n_batches: int, stride: int, photons_per_batch: int, trace_fn: Callable, rank: int, world: int,
                      by: str = "batch", device=None):
    """Run ``trace_fn(batch_indices, photon_begin, photon_count) -> int64 [len(batch_indices), stride]`` on
    this rank's share and return the full ``[n_batches, stride]`` block on every rank.

    ``by="batch"`` gives each rank whole batches (the drivers' thousands of time points): rows are disjoint, the
    collective is an all-gather;  ``by="photon"`` splits every batch's photon range (few, large batches): rows add,
    the collective is an all-reduce."""
    import torch

    if by == "batch":
        mine = shard_batches(n_batches, rank, world)
        part = torch.zeros((0, stride), dtype=torch.int64, device=device)
        if len(mine):
            part = torch.as_tensor(trace_fn(mine, 0, int(photons_per_batch)), device=device)
        return all_gather_rows(part, n_batches, rank, world)
    if by == "photon":
        full = torch.zeros((n_batches, stride), dtype=torch.int64, device=device)
        begin, count = shard_photons(photons_per_batch, rank, world)
        if count:
            part = trace_fn(np.arange(n_batches, dtype=np.int64), begin, count)
            full += torch.as_tensor(part, device=full.device)
        return all_reduce_tallies(full)
    raise ValueError("by must be 'batch' or 'photon'")


def trace_lights_distributed(arrays, lights, photons_per_batch: int, seed: int, by: str = "batch"):
    """CUDA tracing of ``lights`` (a sequence of ``LightArrays`` or a ``lights.BatchTable``) sharded over the ranks of
    the initialised process group; every rank returns the full ``[n_batches, n_bins + N_COUNTERS]`` int64 tensor (on
    its GPU).  ``seed`` must be the same on every rank (``miniplant.full_simulation.agreed_seed`` broadcasts one)."""
    import torch
    import torch.distributed as dist

    from . import backend
    from .lights import BatchTable

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = torch.cuda.current_device()
    scene = backend.scene_handle(arrays, device)
    table = isinstance(lights, BatchTable)

    def trace_fn(batch_indices, photon_begin, photon_count):
        with torch.cuda.device(device):
            if table:
                plan = backend.Plan(scene, lights.take(batch_indices))
            else:
                plan = backend.Plan(scene, [lights[int(b)] for b in batch_indices], [int(b) for b in batch_indices])
            tallies = plan.new_tallies()
            plan.trace_into(tallies, photon_count, seed, photon_begin)
            torch.cuda.current_stream().synchronize()   # the plan's buffers must outlive the launch
        return tallies

    return trace_partitioned(len(lights), scene.stride, photons_per_batch, trace_fn, rank, world, by=by,
                             device=f"cuda:{device}")
