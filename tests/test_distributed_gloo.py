"""The N>1 path on CPU: world_size-2 gloo process group, batches / photon ranges partitioned over the ranks, the
integer tally block all-reduced.  The tracer is injected (the C oracle stands in for the CUDA launch, which needs a
GPU); what is under test is the partitioning, the Philox-style (seed, batch, photon) keying contract — the union
of the shards equals the single-process run bit for bit — and the collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_trace_fn(n_batches):
    import sys

    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import lscpm_b200

    lscpm_b200.install()
    import c_oracle
    from lscpm_b200 import _native as nat, backend, flatten
    from lscpm_b200.workloads import standard_scene

    scenes = [standard_scene(tilt=0, elevation=30 + 10 * b, wavelength=555.0) for b in range(n_batches)]
    arrays = flatten.flatten_scene(scenes[0])
    lights = [flatten.flatten_light(s) for s in scenes]
    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    batches, _ = backend.pack_batches(lights, list(range(n_batches)))
    stride = oracle.n_bins + 2

    def trace_fn(batch_indices, photon_begin, photon_count):
        out = np.zeros((len(batch_indices), stride), dtype=np.int64)
        for row, b in enumerate(batch_indices):
            out[row] = oracle.trace(batches[int(b)], photon_begin=photon_begin, photon_count=photon_count, seed=99, threads=1)
        return torch.from_numpy(out)

    return trace_fn, stride


def _worker(rank, world, port, n_batches, ppb, by, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys

    sys.path.insert(0, ROOT)
    import lscpm_b200  # noqa: F401
    from lscpm_b200 import distributed

    got = distributed.init_process_group(backend="gloo")
    assert got == (rank, rank, world)
    trace_fn, stride = _oracle_trace_fn(n_batches)
    full = distributed.trace_partitioned(n_batches, stride, ppb, trace_fn, rank, world, by=by, device="cpu")
    np.save(os.path.join(out_dir, f"{by}_{rank}.npy"), full.numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("by,n_batches", [("batch", 5), ("photon", 5), ("batch", 1)])   # 1 batch: rank 1 owns nothing
def test_world_size_2_gloo(by, n_batches, tmp_path):
    ppb, world = 601, 2
    mp.spawn(_worker, args=(world, _free_port(), n_batches, ppb, by, str(tmp_path)), nprocs=world, join=True)
    trace_fn, stride = _oracle_trace_fn(n_batches)
    whole = trace_fn(np.arange(n_batches), 0, ppb).numpy()
    for rank in range(world):
        got = np.load(tmp_path / f"{by}_{rank}.npy")
        assert got.shape == (n_batches, stride)
        assert np.array_equal(got, whole), f"rank {rank}: reduced tallies differ from the single-process run"
    assert np.all(whole[:, :stride - 2].sum(axis=1) == ppb)


def test_shard_helpers():
    from lscpm_b200 import distributed

    assert list(distributed.shard_batches(7, 1, 3)) == [1, 4]
    covered = np.concatenate([distributed.shard_batches(10, r, 4) for r in range(4)])
    assert sorted(covered) == list(range(10))
    spans = [distributed.shard_photons(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert distributed.env_rank() == (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
                                      int(os.environ.get("WORLD_SIZE", 1)))
