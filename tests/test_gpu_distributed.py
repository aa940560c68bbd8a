"""The product's N>1 path on real GPUs: NCCL ranks, one per GPU, through
``distributed.trace_lights_distributed`` in both partitions, against the single-GPU result — BIT-identical.

Skipped on a one-GPU box (the round-end suite); run with ``gpurun --gpus 2|4|8 -- python -m pytest tests/test_gpu_distributed.py -m gpu``.
``bench.py`` repeats the comparison after every multi-GPU timed run (its ``verify`` key), so the SCALE record carries
the same evidence for 2, 4 and 8 ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_BATCHES, PHOTONS, SEED = 11, 30_011, 77      # odd on purpose: ragged shares in both partitions


def _gpu_count() -> int:
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _job():
    """A small year: direct and diffuse batches with different suns and spectra, as a BatchTable."""
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import lscpm_b200

    lscpm_b200.install()
    from lscpm_b200 import lights
    from lscpm_b200.miniplant.full_simulation import plate_geometry
    from lscpm_b200.workloads import AM15_LIKE_W, SPECTRL2_SLICE_NM

    arrays = plate_geometry(30.0, True)
    n_direct = 7
    elevation = np.linspace(15.0, 75.0, n_direct)
    azimuth = np.linspace(110.0, 250.0, n_direct)
    flux = AM15_LIKE_W * SPECTRL2_SLICE_NM * (1.0 + 0.1 * np.arange(N_BATCHES)[:, None] * np.linspace(0, 1, len(SPECTRL2_SLICE_NM)))
    table = lights.BatchTable.concatenate([
        lights.direct_batches(30.0, elevation, azimuth, flux=flux[:n_direct], x=SPECTRL2_SLICE_NM),
        lights.diffuse_batches(30.0, flux=flux[n_direct:], x=SPECTRL2_SLICE_NM)])
    assert len(table) == N_BATCHES
    return arrays, table


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import lscpm_b200  # noqa: F401
    from lscpm_b200 import distributed

    assert distributed.init_process_group(backend="nccl") == (rank, rank, world)
    arrays, table = _job()
    for by in ("batch", "photon"):
        full = distributed.trace_lights_distributed(arrays, table, PHOTONS, SEED, by=by)
        assert full.is_cuda and full.device.index == rank
        np.save(os.path.join(out_dir, f"{by}_{world}_{rank}.npy"), full.cpu().numpy())
    # the list-of-LightArrays form of the same job
    full = distributed.trace_lights_distributed(arrays, list(table), PHOTONS, SEED, by="batch")
    np.save(os.path.join(out_dir, f"list_{world}_{rank}.npy"), full.cpu().numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_rank_tallies_are_bit_identical_to_one_gpu(world, built_library, tmp_path):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpu_count()}")
    import torch.multiprocessing as mp
    from lscpm_b200 import backend

    arrays, table = _job()
    single = backend.trace_lights(arrays, table, PHOTONS, seed=SEED, device=0)
    n_bins = single.shape[1] - 2
    assert (single[:, :n_bins].sum(axis=1) == PHOTONS).all()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for kind in ("batch", "photon", "list"):
        for rank in range(world):
            got = np.load(tmp_path / f"{kind}_{world}_{rank}.npy")
            assert np.array_equal(got, single), f"{kind} partition, world {world}, rank {rank}: tallies differ from the 1-GPU run"


def test_one_gpu_photon_ranges_and_seeds_accumulate(built_library):
    """The additive contract the N-rank reduction rests on, on one device: tracing two photon sub-ranges into the same
    block equals tracing the whole range."""
    import torch
    from lscpm_b200 import backend

    arrays, table = _job()
    scene = backend.scene_handle(arrays, 0)
    plan = backend.Plan(scene, table)
    whole, parts = plan.new_tallies(), plan.new_tallies()
    plan.trace_into(whole, PHOTONS, SEED)
    cut = 12_345
    plan.trace_into(parts, cut, SEED, photon_begin=0)
    plan.trace_into(parts, PHOTONS - cut, SEED, photon_begin=cut)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    assert np.array_equal(whole.cpu().numpy(), backend.trace_lights(arrays, table, PHOTONS, seed=SEED, device=0))
