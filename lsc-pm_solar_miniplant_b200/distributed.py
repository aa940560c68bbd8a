"""Multi-GPU sharding of the hot path: one process per GPU, batches partitioned over ranks, and ONE
collective — an all-reduce of the small integer tally block (NCCL over NVLink on GPUs, gloo in the CPU
tests).  Photons are independent, so there is no data-path exchange; the reference's only parallelism
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


def trace_partitioned(n_batches: int, stride: int, photons_per_batch: int, trace_fn: Callable, rank: int, world: int,
                      by: str = "batch", device=None):
    """Run ``trace_fn(batch_indices, photon_begin, photon_count) -> int64 [len(batch_indices), stride]`` on
    this rank's share and all-reduce into the full ``[n_batches, stride]`` block.

    ``by="batch"`` gives each rank whole batches (the drivers' thousands of time points);
    ``by="photon"`` splits every batch's photon range (few, large batches)."""
    import torch

    full = torch.zeros((n_batches, stride), dtype=torch.int64, device=device)
    if by == "batch":
        mine = shard_batches(n_batches, rank, world)
        if len(mine):
            part = trace_fn(mine, 0, int(photons_per_batch))
            full[torch.as_tensor(mine, device=full.device)] = torch.as_tensor(part, device=full.device)
    elif by == "photon":
        begin, count = shard_photons(photons_per_batch, rank, world)
        if count:
            part = trace_fn(np.arange(n_batches, dtype=np.int64), begin, count)
            full += torch.as_tensor(part, device=full.device)
    else:
        raise ValueError("by must be 'batch' or 'photon'")
    return all_reduce_tallies(full)


def trace_lights_distributed(arrays, lights: Sequence, photons_per_batch: int, seed: int, by: str = "batch"):
    """CUDA tracing of ``lights`` sharded over the ranks of the initialised process group; every rank
    returns the full reduced ``[n_batches, n_bins + N_COUNTERS]`` int64 tensor (on its GPU)."""
    import torch
    import torch.distributed as dist

    from . import backend

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = torch.cuda.current_device()
    scene = backend.scene_handle(arrays, device)

    def trace_fn(batch_indices, photon_begin, photon_count):
        plan = backend.Plan(scene, [lights[int(b)] for b in batch_indices], [int(b) for b in batch_indices])
        tallies = plan.new_tallies()
        plan.trace_into(tallies, photon_count, seed, photon_begin)
        return tallies

    return trace_partitioned(len(lights), scene.stride, photons_per_batch, trace_fn, rank, world, by=by,
                             device=f"cuda:{device}")
