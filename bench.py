#!/usr/bin/env python
"""Benchmark of the LSC-PM photon-tracing hot path (contract: one JSON line on stdout from rank 0).

    python bench.py --gpus N --steps K --warmup W          # CUDA path (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W  # CPU arm: the oracle port on all host cores

Workload (BASELINE.json configs[1]): the standard LSC-PM scene (Lumogen F Red 305 plate + 16 methylene-blue
capillaries), tilt 0 deg / sun at zenith, wavelengths drawn from an AM1.5G-like 27-knot photon-flux table on
the SPECTRL2 slice the reference samples.  One step = one pass of the hot path over `--batches` batches of
`--photons` photons per GPU (default 64 x 1e6: the drivers launch many time points at once); weak scaling.

`value`   photons/s with batch descriptors and spectra already resident in HBM (plan created outside the
          timed region).  A job is K steps whose tallies accumulate in the owning rank's rows (tallies add); at
          N>1 ONE all-gather of those rows ends the job, inside the timed region.  After it, one more step is
          compared bit for bit with rank 0 tracing every rank's batches alone (`verify`).
`e2e`     the same metric through the public API with HOST inputs: descriptor lowering + H2D, launch,
          D2H of the tallies, every step.
`roofline`  FP32/INT issue roofline of the trace kernel: events/s x 224 lane-instr per
          event against n_SM x 128 lanes x the SM clock sampled during the run.  The path keeps photon state
          in registers, so the HBM view (`roofline_hbm`, 96 B/event if state round-tripped) is informational.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

I_EVT = 224            # lane-instructions per photon-event, frozen from SURVEY.md section 8(d)
BYTES_PER_EVENT = 96   # algorithmic bytes per event if photon state round-tripped HBM (it does not)
METRIC = "photons/sec traced to termination"


WORKLOAD = "config2"
_RESULT_FD = 1


def claim_stdout() -> None:
    """stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    version banner there when NCCL_DEBUG is set), so the real stdout is kept aside for the result and descriptor 1 is
    pointed at stderr for everything else."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


def workload_scene():
    import lscpm_b200

    lscpm_b200.install()
    from lscpm_b200 import flatten
    from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene

    extra = {}
    if WORKLOAD == "config5":     # scaled mini-plant: every batch is one plate of the array, dense channel grid
        extra = dict(capillary_count=46, capillary_pitch=0.01)
    scene = standard_scene(tilt=0, elevation=90, azimuth=180, spectrum=am15_like_photon_spectrum(), **extra)
    return flatten.flatten_scene(scene)


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, sm_max, reasons = [], [], set()
        for t, line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clock, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            inside = t0 <= t <= t1 + 0.2
            if inside:
                sm.append(clock)
                sm_max.append(cmax)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(sm_max), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            data = json.load(fh)
        data["_source"] = "measured"
        return data
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "_source": "fallback"}


def cpu_oracle_throughput(arrays, photons: int, threads: int) -> tuple:
    """photons/s of the C oracle (CPU port of the reference algorithm) on `threads` host threads."""
    import c_oracle
    from lscpm_b200 import _native as nat, backend

    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    batches, _ = backend.pack_batches([arrays.light])
    spectra = [arrays.light.spectrum] if arrays.light.spectrum is not None else None
    if photons <= 0:
        # bounded sample: calibrate, then size the run for about 10 s of CPU work
        t0 = time.perf_counter()
        oracle.trace(batches[0], spectra=spectra, photon_count=200_000, seed=1, threads=threads)
        rate = 200_000 / max(time.perf_counter() - t0, 1e-6)
        photons = int(min(max(rate * 10.0, 1_000_000), 200_000_000))
    else:
        oracle.trace(batches[0], spectra=spectra, photon_count=2000, seed=1, threads=threads)  # warm
    t0 = time.perf_counter()
    out = oracle.trace(batches[0], spectra=spectra, photon_count=photons, seed=2, threads=threads)
    dt = time.perf_counter() - t0
    return photons / dt, dt, photons


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arrays = workload_scene()
    cores = os.cpu_count() or 1
    # each step is a bounded sample of the workload; the whole run is sized to end within a few minutes
    _, _, calibrated = cpu_oracle_throughput(arrays, args.cpu_photons, cores)
    sample = args.cpu_photons
    if args.cpu_photons <= 0:   # ~10 s of CPU work per step, the whole run capped near one minute
        sample = max(int(calibrated * min(1.0, 6.0 / max(args.steps + args.warmup, 1))), 200_000)
    for _ in range(max(args.warmup, 0)):
        cpu_oracle_throughput(arrays, sample, cores)
    total, elapsed = 0, 0.0
    for _ in range(args.steps):
        rate, dt, _ = cpu_oracle_throughput(arrays, sample, cores)
        total += sample
        elapsed += dt
    value = total / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "photons/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "CPU oracle port (C float64 restatement of pvtrace's "
                   "algorithm; the reference itself is Python + un-vendored pvtrace and cannot be installed offline)"},
        "cpu_baseline": {"value": value, "unit": "photons/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} photons per step of the same workload, OpenMP over {cores} threads"},
        "e2e": {"value": value, "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_name(args) -> str:
    if WORKLOAD == "config5":
        return (f"BASELINE configs[4]: scaled mini-plant, array of LSC-PM plates with a dense channel grid (46 capillaries at "
                f"10 mm); the plates of the array are optically independent (SURVEY 8d), so one plate = one batch with its own "
                f"RNG stream and tally row; {args.batches} plates x {args.photons} photons per GPU per step")
    return (f"BASELINE configs[1]: standard LSC-PM scene (LR305 plate + 16 MB capillaries), tilt 0 / sun at zenith, "
            f"AM1.5G-like 27-knot SPECTRL2-slice spectrum; {args.batches} batches x {args.photons} photons per GPU per step")


def source_digest() -> str:
    """sha256 over the sources the CUDA library is built from: ties a committed ncu summary to the code that ran."""
    import hashlib

    h = hashlib.sha256()
    for rel in ("lsc-pm_solar_miniplant_b200/csrc/lscpm.cu", "lsc-pm_solar_miniplant_b200/csrc/lscpm_device.cuh",
                "lsc-pm_solar_miniplant_b200/csrc/lscpm_pool.cuh", "include/lscpm.h"):
        with open(os.path.join(ROOT, rel), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def committed_profile():
    """(metrics, reason): the ncu summary committed for the dominant kernel, or None with the reason it is not used.
    A summary taken from other sources than the ones that built the library is stale and must not be replayed."""
    path = os.path.join(ROOT, "profiles", "r2_final_metrics.json")
    if not os.path.exists(path):
        return None, "no committed ncu summary (profiles/r2_final_metrics.json)"
    with open(path) as fh:
        ncu = json.load(fh)
    if ncu.get("source_sha256") != source_digest():
        return None, ("profiles/r2_final_metrics.json was captured from other kernel sources (sha256 differs): "
                      "its counters are not reported for this library")
    return ncu, None


def python_restatement_rate(arrays_scene, seconds: float = 6.0) -> dict:
    """photons/s of the literal Python restatement of pvtrace's algorithm (oracle/pvtrace_oracle.py) on one core: the
    closest thing to the reference's own cost structure that can run offline (a per-photon Python loop over scene-graph
    objects).  Bounded to a few seconds."""
    import pvtrace_oracle as po

    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        po.run(arrays_scene, 200, seed=100 + n)
        n += 200
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "photons/s", "cores": 1, "sample": f"{n} photons in {dt:.1f} s"}


def measure_drivers(device) -> dict:
    """The reference's drivers end to end (BASELINE configs[3]): wall time of ``yearlong_simulation`` for the Eindhoven
    40 deg year — solar stage, batch descriptors, upload, ONE launch over 16 096 batches, download, CSV — at the
    reference's own 120 photons per point and at 1e5; and the latency of the per-point entry point the reference's
    drivers call 16 k times (``run_direct_simulation(num_photons=100)``)."""
    import tempfile

    import numpy as np

    from lscpm_b200.miniplant import full_simulation as fs
    from lscpm_b200.miniplant.locations import EINDHOVEN
    from lscpm_b200.miniplant.simulation_runner import run_direct_simulation
    from lscpm_b200.miniplant.solar_data import solar_data_for_place_and_time

    out = {"workload": "BASELINE configs[3]: full_simulation, Eindhoven tilt 40 deg, 2020 at 30 min resolution, "
                       "direct + sky-diffuse SPECTRL2 spectra per time point, one batch per (time point, light)"}
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)     # the solar stage writes Full_data_<site>.csv into the working directory, as the reference does
        try:
            fs.yearlong_simulation(40, EINDHOVEN, num_photons_per_simulation=120, target_file=os.path.join(tmp, "w.csv"), seed=1, device=device)
            for photons in (120, 100_000):
                t0 = time.perf_counter()
                data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800)
                t1 = time.perf_counter()
                table = fs.year_batches(40, data)
                t2 = time.perf_counter()
                fs.trace_time_points(40, table, photons, True, seed=2, device=device)
                t3 = time.perf_counter()
                res = fs.yearlong_simulation(40, EINDHOVEN, num_photons_per_simulation=photons, target_file=os.path.join(tmp, "r.csv"),
                                             seed=3, device=device)
                t4 = time.perf_counter()
                n_points = len(res)
                out[f"yearlong_{photons}_photons_per_point"] = {
                    "wall_s": t4 - t3, "time_points": n_points, "batches": 2 * n_points, "photons": 2 * n_points * photons,
                    "photons_per_s": 2 * n_points * photons / (t4 - t3),
                    "stages_s": {"solar_stage_and_its_csv": t1 - t0, "batch_descriptors": t2 - t1,
                                 "lower_upload_launch_download": t3 - t2,
                                 "results_csv_and_frame": max((t4 - t3) - (t3 - t0), 0.0)},
                    "yearly_mean_reacted_direct": float(np.mean(res["simulation_direct"])),
                    "yearly_mean_reacted_diffuse": float(np.mean(res["simulation_diffuse"]))}
        finally:
            os.chdir(cwd)
    run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180, num_photons=100, seed=1, device=device)
    laps = []
    for k in range(300):
        t0 = time.perf_counter()
        run_direct_simulation(tilt_angle=40, solar_elevation=20 + (k % 50), solar_azimuth=180, num_photons=100, seed=k, device=device)
        laps.append(time.perf_counter() - t0)
    laps.sort()
    import random

    rng = random.Random(3)

    def two_lines():
        return rng.choice([555, 585])

    run_direct_simulation(tilt_angle=0, solar_elevation=90, solar_spectrum_function=two_lines, num_photons=200, seed=1, device=device)
    t0 = time.perf_counter()
    run_direct_simulation(tilt_angle=0, solar_elevation=90, solar_spectrum_function=two_lines, num_photons=20_000, seed=2, device=device)
    dt = time.perf_counter() - t0
    out["host_emitted_light"] = {"call": "run_direct_simulation with an arbitrary wavelength callable: the rays are emitted by the scene's "
                                         "own Python callables (scene.emit) and traced by follow_kernel, one thread per ray",
                                 "photons": 20_000, "wall_s": dt, "photons_per_s": 20_000 / dt}
    out["latency"] = {"call": "run_direct_simulation(num_photons=100), a new sun position per call", "calls": len(laps),
                      "median_ms": 1e3 * laps[len(laps) // 2], "p90_ms": 1e3 * laps[int(0.9 * len(laps))]}
    return out


def run_config4(args) -> None:
    """Strong scaling of ONE fixed job over N GPUs (BASELINE configs[3]): the Eindhoven 40 deg year, 16 096 batches (8 048
    half-hour points x {direct, sky-diffuse}, each with its own SPECTRL2 spectrum) x `--photons` photons, batches dealt
    round-robin to the ranks, one all-gather of the tally rows per job.  `value`: device-timed, descriptors resident;
    `e2e`: wall time of `yearlong_simulation` under the process group - every rank runs the (replicated) solar stage,
    traces its share, gathers; rank 0 writes the CSV."""
    import tempfile

    import numpy as np
    import torch
    import torch.distributed as dist

    import lscpm_b200

    lscpm_b200.install()
    from lscpm_b200 import _native as nat, backend, distributed
    from lscpm_b200.miniplant import full_simulation as fs
    from lscpm_b200.miniplant.locations import EINDHOVEN
    from lscpm_b200.miniplant.solar_data import solar_data_for_place_and_time

    rank, local_rank, world = distributed.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the traced path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.cuda.current_device()
    dev = f"cuda:{device}"
    lib = nat.load()
    ppb = args.photons
    data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    table = fs.year_batches(40, data)
    arrays = fs.plate_geometry(40, True)
    scene = backend.scene_handle(arrays, device)
    n = len(table)
    rows = distributed.shard_batches(n, rank, world)
    plan = backend.Plan(scene, table.take(rows))
    part = plan.new_tallies()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job(seed):
        part.zero_()
        plan.trace_into(part, ppb, seed)
        return distributed.all_gather_rows(part, n, rank, world)

    for w in range(max(args.warmup, 3)):
        full = job(100 + w)
    barrier()
    launches0 = lib.lscpm_launch_count()
    step_ms = []
    with ClockSampler(local_rank) as sampler:
        barrier()
        t_begin = time.perf_counter()
        for k in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            full = job(500 + k)
            e1.record()
            torch.cuda.synchronize()
            step_ms.append(e0.elapsed_time(e1))
        barrier()
        t_end = time.perf_counter()
    launches = lib.lscpm_launch_count() - launches0
    clocks = sampler.summary(t_begin, t_end)
    tallies = full.cpu().numpy()
    n_bins = scene.n_bins
    assert (tallies[:, :n_bins].sum(axis=1) == ppb).all(), "every batch's bins must sum to its photons"
    events_per_photon = float(tallies[:, n_bins].sum()) / (n * ppb)

    # end to end: the driver a user runs, every step
    e2e_steps = max(1, min(args.steps, 5))
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        fs.yearlong_simulation(40, EINDHOVEN, num_photons_per_simulation=ppb, target_file=os.path.join(tmp, f"w{rank}.csv"), seed=1)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            res = fs.yearlong_simulation(40, EINDHOVEN, num_photons_per_simulation=ppb, target_file=os.path.join(tmp, "year.csv"), seed=10 + k)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    finally:
        os.chdir(cwd)
    t_max = torch.tensor([sum(step_ms) / len(step_ms), e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_per_step, ms_e2e = (float(x) for x in t_max.cpu())
    if rank == 0:
        total = n * ppb
        line = {
            "metric": METRIC, "value": total / (ms_per_step * 1e-3), "unit": "photons/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[3]: full_simulation, Eindhoven tilt 40 deg, 2020 at 30 min resolution: {n} batches "
                                   f"({n // 2} time points x direct + sky-diffuse, own SPECTRL2 spectrum each) x {ppb} photons = "
                                   f"{total:.3g} photons per job, FIXED for every N",
                       "l2": "inputs larger than L2 are not involved: photon state lives on chip; tallies 30 MB per job",
                       "parallelism": f"batch-dp{world} (round-robin)", "collective": "one all-gather of the tally rows per job"},
            "photon_events_per_s": total * events_per_photon / (ms_per_step * 1e-3), "events_per_photon": events_per_photon,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "photons/s", "wall_s_per_job": ms_e2e * 1e-3, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(lib.lscpm_plan_h2d_bytes(plan.handle)), "d2h_bytes_per_step": int(tallies.nbytes),
                    "api": "miniplant.full_simulation.yearlong_simulation (solar stage on every rank, sharded trace, all-gather, CSV on rank 0)",
                    "yearly_mean_reacted_direct": float(np.mean(res["simulation_direct"]))},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_cuda(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from lscpm_b200 import _native as nat, backend, distributed

    rank, local_rank, world = distributed.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the traced path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.cuda.current_device()
    lib = nat.load()

    arrays = workload_scene()
    scene = backend.scene_handle(arrays, device)
    nb, ppb = args.batches, args.photons
    total_batches = nb * world
    my_ids = [rank * nb + i for i in range(nb)]           # weak scaling: every rank owns `nb` distinct batches (plates)
    lights = [arrays.light] * nb
    plan = backend.Plan(scene, lights, my_ids)
    dev = f"cuda:{device}"
    mine = torch.zeros((nb, scene.stride), dtype=torch.int64, device=dev)
    full = torch.zeros((total_batches, scene.stride), dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def combine():
        """The path's one exchange: tallies are additive and every batch has one owner, so the job's steps accumulate
        in the owner's rows and ONE all-gather at the end of the job hands every rank the whole block."""
        if world > 1:
            dist.all_gather_into_tensor(full, mine)
        else:
            full.copy_(mine)

    for w in range(max(args.warmup, 3)):
        plan.trace_into(mine, ppb, 1000 + w)
    combine()
    barrier()

    launches0 = lib.lscpm_launch_count()
    mine.zero_()
    with ClockSampler(local_rank) as sampler:
        barrier()
        t_begin = time.perf_counter()
        events = []
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                            # flush L2 between timed iterations (outside the step's events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.trace_into(mine, ppb, 5000 + k)             # ADDS into this rank's rows
            e1.record()
            events.append((e0, e1))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        combine()
        g1.record()
        barrier()
        t_end = time.perf_counter()
    launches = lib.lscpm_launch_count() - launches0
    clocks = sampler.summary(t_begin, t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in events]
    combine_ms = g0.elapsed_time(g1)
    job_ms = sum(kernel_ms) + combine_ms                     # the K steps and the job's one exchange, device-timed

    tallies = full.cpu().numpy()
    n_bins = scene.n_bins
    assert int(tallies[:, :n_bins].sum()) == total_batches * ppb * args.steps, "tally bins must sum to the photons traced"
    events_per_photon = float(tallies[:, n_bins].sum()) / (total_batches * ppb * args.steps)
    react = sum(int(tallies[:, i].sum()) for i, name in enumerate(scene.bin_names) if name.startswith("react/"))
    reacted_fraction = react / (total_batches * ppb * args.steps)

    # ---- N-rank invariance on the bench workload: one extra step, every rank's share gathered, against ----
    # ---- rank 0 tracing ALL N x nb batches alone with the same seed; and the product's sharded entry point -----------
    verify = None
    if world > 1:
        mine.zero_()
        plan.trace_into(mine, ppb, 424242)
        combine()
        torch.cuda.synchronize()
        ok_bench = ok_batch = ok_photon = True
        if rank == 0:
            alone = backend.Plan(scene, [arrays.light] * total_batches, list(range(total_batches)))
            ref = alone.new_tallies()
            alone.trace_into(ref, ppb, 424242)
            torch.cuda.synchronize()
            ok_bench = bool(torch.equal(ref, full))
        small = 200_000
        lights_all = [arrays.light] * (2 * world + 1)
        by_batch = distributed.trace_lights_distributed(arrays, lights_all, small, 99, by="batch")
        by_photon = distributed.trace_lights_distributed(arrays, lights_all, small + 7, 98, by="photon")
        if rank == 0:
            ok_batch = bool(np.array_equal(by_batch.cpu().numpy(), backend.trace_lights(arrays, lights_all, small, seed=99, device=device)))
            ok_photon = bool(np.array_equal(by_photon.cpu().numpy(), backend.trace_lights(arrays, lights_all, small + 7, seed=98, device=device)))
        verify = {"n_rank_tallies_bit_identical_to_one_gpu": ok_bench, "photons_compared": total_batches * ppb,
                  "trace_lights_distributed_by_batch": ok_batch, "trace_lights_distributed_by_photon": ok_photon,
                  "how": "after the timed region: every rank traces its batches with one more seed and the gathered block is "
                         "compared on rank 0 with rank 0 tracing all N x batches alone; then distributed.trace_lights_distributed "
                         "in both partitions against backend.trace_lights on rank 0"}

    # ---- end to end through the public API: host descriptors in, host tallies out, every step --------
    e2e_steps = max(1, args.steps)
    host_table = backend.batch_table(lights, my_ids)      # LscpmBatchDesc[nb] + spectra in host memory: the API's input format
    backend.trace_lights(arrays, host_table, ppb, seed=1, device=device)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host = backend.trace_lights(arrays, host_table, ppb, seed=7000 + k, device=device)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d_bytes = int(lib.lscpm_plan_h2d_bytes(plan.handle))
    d2h_bytes = int(host.nbytes)

    # max over ranks of the device-timed job, the mean kernel and the end-to-end step; min of the kernel for the skew
    mean_kernel = sum(kernel_ms) / len(kernel_ms)
    t_max = torch.tensor([job_ms / args.steps, mean_kernel, e2e_s * 1e3], dtype=torch.float64, device=dev)
    t_min = torch.tensor([mean_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
    ms_per_step, ms_kernel, ms_e2e = (float(x) for x in t_max.cpu())
    ms_kernel_fastest = float(t_min.cpu()[0])

    if rank == 0:
        photons_per_step = total_batches * ppb
        value = photons_per_step / (ms_per_step * 1e-3)
        peaks = measured_peaks()
        sm_clock = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count
        events_per_s_gpu = nb * ppb * events_per_photon / (ms_kernel * 1e-3)      # one GPU's kernel
        peak_lane = n_sm * 128 * sm_clock * 1e6
        achieved_lane = events_per_s_gpu * I_EVT
        ncu, stale = committed_profile() if WORKLOAD == "config2" else (None, "committed ncu summary is for the config2 workload")
        traffic = None if ncu is None else ncu["dram__bytes_read.sum"] + ncu["dram__bytes_write.sum"]
        line = {
            "metric": METRIC, "value": value, "unit": "photons/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "l2": "flushed between timed steps (256 MiB write, outside the step's events); "
                       "photon state lives on chip, inputs are KB-sized",
                       "seed_policy": "Philox4x32-10 keyed by (seed, batch, photon)", "parallelism": f"batch-dp{world}",
                       "collective": "none per step; one all_gather_into_tensor of the accumulated tally rows at the end of the "
                                     "K-step job, inside the timed region" if world > 1 else "none"},
            "photon_events_per_s": value * events_per_photon, "events_per_photon": events_per_photon,
            "reacted_fraction": reacted_fraction,
            "e2e": {"value": photons_per_step / (ms_e2e * 1e-3), "unit": "photons/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "api": "lscpm_b200.backend.trace_lights(BatchTable) -> lscpm_trace_host: host LscpmBatchDesc[] + spectra lowered and copied H2D, "
                           "launch, D2H of the tallies, every step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernel_ms": {"slowest_rank_mean": ms_kernel, "fastest_rank_mean": ms_kernel_fastest,
                          "exchange_ms_per_job_rank0": combine_ms},
            "roofline": {"bound": "fp32_issue", "achieved": achieved_lane / 1e12, "peak": peak_lane / 1e12,
                         "unit": "Tlane-instr/s", "frac": achieved_lane / peak_lane, "traffic": traffic,
                         "traffic_note": ("DRAM bytes per launch from profiles/r2_final_metrics.json (ncu --set full, same sources): "
                                          "photon state stays on chip") if ncu else stale,
                         "executed": None if ncu is None else {
                             "issue_slot_utilisation": ncu["sm__issue_active.avg.pct_of_peak_sustained_elapsed"] / 100.0,
                             "active_lanes_per_warp_instruction": ncu["smsp__thread_inst_executed_per_inst_executed.ratio"],
                             "warp_instructions_per_launch": ncu["smsp__inst_executed.sum"],
                             "source": "profiles/r2_final_metrics.json", "git_head": ncu.get("git_head")},
                         "kernel": scene.kernel_name,
                         "ms_per_launch": ms_kernel,
                         "how": f"events/s/GPU x {I_EVT} lane-instr/event (SURVEY 8d) / (n_SM {n_sm} x 128 x {sm_clock:.0f} MHz sampled)"},
            "roofline_hbm": {"bound": "hbm", "achieved": events_per_s_gpu * BYTES_PER_EVENT / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": events_per_s_gpu * BYTES_PER_EVENT / 1e9 / peaks["hbm_gbs"],
                             "peak_source": peaks["_source"], "note": "informational: what a state-in-HBM wavefront design "
                             "would have to move; this kernel keeps photon state in shared memory / registers"},
        }
        if verify is not None:
            line["verify"] = verify
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, dt, n_cpu = cpu_oracle_throughput(arrays, args.cpu_photons, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "photons/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_cpu} photons of the same workload in {dt:.1f} s "
                                              f"(C float64 oracle, OpenMP {cores} threads)"}
            rate1, dt1, n1 = cpu_oracle_throughput(arrays, max(int(rate / cores * 4), 100_000), 1)
            line["cpu_baseline"]["one_core"] = {"value": rate1, "unit": "photons/s", "cores": 1,
                                                "sample": f"{n1} photons in {dt1:.1f} s (same C oracle, 1 thread)"}
            if WORKLOAD == "config2":
                from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene

                py = python_restatement_rate(standard_scene(tilt=0, elevation=90, azimuth=180, spectrum=am15_like_photon_spectrum()))
                py["note"] = ("literal Python restatement of pvtrace's per-photon loop over scene-graph objects "
                              "(oracle/pvtrace_oracle.py): the reference's own cost structure; pvtrace itself is not installable offline")
                line["cpu_baseline"]["python_restatement"] = py
        if world == 1 and not args.no_drivers and WORKLOAD == "config2":
            line["drivers"] = measure_drivers(device)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("cuda", "reference"), default="cuda")
    ap.add_argument("--batches", type=int, default=64, help="batches (time points) per GPU per step")
    ap.add_argument("--photons", type=int, default=1_000_000, help="photons per batch")
    ap.add_argument("--cpu-photons", type=int, default=0, help="photons of the bounded CPU sample (0: calibrate to ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-drivers", action="store_true", help="skip the driver-level (config 4) and latency measurements")
    ap.add_argument("--workload", choices=("config2", "config4", "config5"), default="config2",
                    help="config2 (default, the bench contract line), config4 (the Eindhoven year as ONE fixed job: strong "
                         "scaling; use --photons 100000) or config5 (dense-grid plate array)")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif WORKLOAD == "config4":
        run_config4(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
