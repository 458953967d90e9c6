#!/usr/bin/env python
"""Benchmark of the lshrs hot path on B200: vectors hashed/s (dim 768, 256 bits) and rerank queries/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One process per GPU; the work list is partitioned (weak scaling: every rank
hashes its own resident shard of ``--rows`` vectors, 12.5 M = 100 M / 8 by
default -- BASELINE.json config 2 at N = 8), the projection matrix is
replicated, nothing is exchanged on the data path.  A "step" is one pass of the
hot path over one resident shard: the projection kernel over every chunk of
the shard plus the D2H of the signatures into pinned host memory, overlapped on
a copy stream.  Timed with CUDA events on the launching stream, max over ranks.

The JSON line carries, beside the contract keys:
  roofline      dominant (projection) kernel: executed tensor flops / its own
                CUDA-event time inside the timed region vs the measured peak
  e2e           same metric through the C ABI with HOST pinned buffers (H2D of
                the vectors and D2H of the signatures inside the timed region)
  cpu_baseline  the oracle port of the reference's numpy path timed on this
                box's host cores on a bounded sample (rank 0, N = 1 only)
  parity        band-key comparison of the GPU output with the oracle on that sample
  rerank        config 4 (8192 queries x 2000 candidates x 768, k=10 and p=0.2)
``--impl reference`` times the reference's CPU algorithm (oracle port; the
reference is pure Python and cannot travel to the GPU box) on bounded samples.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

DIM, NUM_BANDS, ROWS_PER_BAND, SEED = 768, 16, 16, 42
NUM_PERM = NUM_BANDS * ROWS_PER_BAND
SIG_BYTES = NUM_BANDS * ((ROWS_PER_BAND + 7) // 8)
METRIC = "vectors hashed/sec (dim=768, num_perm=256)"
UNIT = "vectors/s"
# BASELINE.json configs; "hash768" is the metric's configuration, the others are side measurements
WORKLOADS = {
    "hash768": dict(dim=768, bands=16, rows_per_band=16, rows=12_500_000, chunk=781_250, dist="gauss"),
    "hash1536": dict(dim=1536, bands=16, rows_per_band=32, rows=6_250_000, chunk=781_250, dist="gauss"),
    "hash128": dict(dim=128, bands=16, rows_per_band=4, rows=100_000_000, chunk=12_500_000, dist="sift"),
}


def select_workload(args) -> None:
    global DIM, NUM_BANDS, ROWS_PER_BAND, NUM_PERM, SIG_BYTES, METRIC
    w = WORKLOADS[args.workload]
    DIM, NUM_BANDS, ROWS_PER_BAND = w["dim"], w["bands"], w["rows_per_band"]
    NUM_PERM = NUM_BANDS * ROWS_PER_BAND
    SIG_BYTES = NUM_BANDS * ((ROWS_PER_BAND + 7) // 8)
    METRIC = f"vectors hashed/sec (dim={DIM}, num_perm={NUM_PERM})"
    if args.rows is None:
        args.rows = w["rows"]
    if args.chunk is None:
        args.chunk = w["chunk"]
    args.chunk = min(args.chunk, args.rows)
    args.e2e_rows = min(args.e2e_rows, args.rows)
    args.dist = w["dist"]
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel PER ROW, from the ncu --set full
# captures committed under profiles/ (r1_hash_tc_ncu.csv: 2.401 GB + 0.027 GB over 781 250 rows;
# r1_hash_tc_dim128_ncu.csv: 6.400 GB + 0.201 GB over 12 500 000 rows); scaled to the rows of one launch
ROOFLINE_TRAFFIC_PER_ROW = {("hash768", "tcgen05"): 2.428e9 / 781_250, ("hash128", "tcgen05"): 6.601e9 / 12_500_000}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks() -> tuple[dict, str]:
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            if "hbm_gbs" in d and "bf16_tflops" in d:
                d.setdefault("bf16_tflops_sustained", d["bf16_tflops"])
                return d, "measured"
        except Exception:  # noqa: BLE001
            pass
    return dict(FALLBACK_PEAKS), "fallback"


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Pin this rank to the CPUs next to its GPU before any pinned allocation (first-touch NUMA placement).

    With 8 ranks each streaming signatures D2H and vectors H2D, pinned buffers on the wrong socket turn
    the host fabric into the bottleneck.  Best effort: returns a description for the JSON line.
    """
    try:
        import torch

        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        node = int((base / "numa_node").read_text().strip())
        cpulist = (base / "local_cpulist").read_text().strip()
        cpus: set[int] = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            return f"unchanged (gpu {bdf} numa_node={node}, local_cpulist={cpulist})"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} (gpu {bdf}, cpus {cpulist})"
    except Exception as exc:  # noqa: BLE001
        return f"unchanged ({type(exc).__name__}: {exc})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port) on this box's host cores
# --------------------------------------------------------------------------------------

def cpu_hash_sample(rows: int, seed: int = 0) -> np.ndarray:
    x = np.random.default_rng(seed).standard_normal((rows, DIM)).astype(np.float32)
    return x  # the reference arm always runs the metric's configuration (Gaussian, dim 768)


def _mp_hash_worker(job):
    from oracle import lshrs_oracle as oracle

    seed, rows = job
    projs = oracle.make_projections(NUM_BANDS, ROWS_PER_BAND, DIM, SEED)
    X = cpu_hash_sample(rows, seed=seed)
    t0 = time.perf_counter()
    oracle.hash_batch_packed(projs, X)
    return time.perf_counter() - t0


def multiprocess_port_rate(rows_per_worker: int = 4096) -> dict:
    """The same per-vector loop sharded over every host core with one process each -- NOT something the
    reference does (its API is one Python thread), printed so the single-thread figure is not the only one."""
    import multiprocessing as mp

    workers = len(os.sched_getaffinity(0))
    try:
        ctx = mp.get_context("fork")
        t0 = time.perf_counter()
        with ctx.Pool(workers) as pool:
            pool.map(_mp_hash_worker, [(100 + i, rows_per_worker) for i in range(workers)])
        wall = time.perf_counter() - t0
        return {"value": workers * rows_per_worker / wall, "unit": UNIT, "cores": workers,
                "sample": f"{workers} processes x {rows_per_worker} rows, wall clock incl. process start"}
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": workers, "error": f"{type(exc).__name__}: {exc}"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    from oracle import lshrs_oracle as oracle

    projs = oracle.make_projections(NUM_BANDS, ROWS_PER_BAND, DIM, SEED)
    sample = 16384
    X = cpu_hash_sample(sample)
    for _ in range(args.warmup):
        oracle.hash_batch_packed(projs, X[:1024])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.hash_batch_packed(projs, X)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    # not reference code: one sgemm + packbits over every BLAS thread, for honesty (torchrun exports
    # OMP_NUM_THREADS=1, so the BLAS pool is widened explicitly when threadpoolctl is there)
    Xv = cpu_hash_sample(131072, seed=1)
    try:
        from threadpoolctl import threadpool_limits

        blas_threads = threadpool_limits(limits=len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        blas_threads = None
    oracle.hash_batch_vectorized(projs, Xv[:4096])
    t1 = time.perf_counter()
    oracle.hash_batch_vectorized(projs, Xv)
    vec_value = Xv.shape[0] / (time.perf_counter() - t1)
    if blas_threads is not None:
        blas_threads.restore_original_limits()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{sample} Gaussian vectors x {args.steps} steps through the oracle's "
                                   "per-vector, per-band sgemv loop (the reference's LSHHasher.hash_batch path; "
                                   "single Python thread by construction)",
                         "host_cores": os.cpu_count(), "numpy": np.__version__,
                         "vectorized_numpy_not_reference": {"value": vec_value, "unit": UNIT,
                                                            "cores": len(os.sched_getaffinity(0))},
                         "multiprocess_port_not_reference": multiprocess_port_rate()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args) -> dict:
    return {
        "workload": f"LSHHasher dim={DIM} num_perm={NUM_PERM} ({NUM_BANDS}x{ROWS_PER_BAND}); "
                    f"{args.rows} synthetic {'Gaussian' if args.dist == 'gauss' else 'SIFT-like non-negative'} "
                    f"float32 vectors resident per GPU"
                    + (" (BASELINE config 2: 100M vectors / 8 GPUs = 12.5M per GPU, weak scaling)"
                       if args.workload == "hash768" else f" (BASELINE config '{args.workload}', side measurement)"),
        "rows_per_gpu": args.rows, "dim": DIM, "num_perm": NUM_PERM, "signature_bytes": SIG_BYTES,
        "chunk_rows": args.chunk, "e2e_rows_per_gpu": args.e2e_rows,
        "l2": f"inputs larger than L2 ({args.rows * DIM * 4 / 1e9:.0f} GB resident shard, each row read once per step)",
        "parallelism": f"row-sharded x{args.gpus}, projections replicated, no collective",
        "cpu_affinity_rank0": getattr(args, "cpu_affinity", None),
    }


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------

def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    from lshrs_b200 import LSHHasher, _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    args.cpu_affinity = bind_to_gpu_numa_node(local)
    log(f"[rank {rank}] cpu affinity: {args.cpu_affinity}")
    peaks, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hasher = LSHHasher(NUM_BANDS, ROWS_PER_BAND, DIM, seed=SEED, device=local)
    if args.kernel != "auto":
        hasher._ensure_handle()
        hasher.set_kernel(args.kernel)

    # ---- resident shard, generated on device (seeded per rank) --------------------------
    rows, chunk = args.rows, args.chunk
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    X = torch.empty((rows, DIM), dtype=torch.float32, device=dev)
    for r0 in range(0, rows, 1 << 20):
        r1 = min(rows, r0 + (1 << 20))
        X[r0:r1].normal_(generator=gen)
        if args.dist == "sift":  # SIFT-like: non-negative integer-valued floats in [0, 255]
            X[r0:r1].abs_().mul_(40.0).floor_().clamp_(max=255.0)
    out_host = torch.empty((rows, SIG_BYTES), dtype=torch.uint8, pin_memory=True)
    out_dev = [torch.empty((chunk, SIG_BYTES), dtype=torch.uint8, device=dev) for _ in range(2)]
    chunks = [(r0, min(rows, r0 + chunk)) for r0 in range(0, rows, chunk)]
    compute = torch.cuda.Stream(device=dev)
    copy = torch.cuda.Stream(device=dev)

    copied = [None, None]   # per output slot: the event of its last D2H (persists across steps)

    def one_step(kernel_events=None):
        """Hash every chunk on `compute`, D2H each chunk's signatures on `copy` (double-buffered).

        Steps stream into each other: the D2H of a step's last chunks overlaps the next step's first
        kernels; the timed region ends only after every copy has landed (drain())."""
        for ci, (r0, r1) in enumerate(chunks):
            slot = ci & 1
            if copied[slot] is not None:
                compute.wait_event(copied[slot])  # out_dev[slot] is free again
            if kernel_events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(compute)
            hasher.hash_into(X[r0:r1], r1 - r0, out_dev[slot], x_on_device=True, out_on_device=True,
                             stream=compute.cuda_stream)
            done = torch.cuda.Event()
            if kernel_events is not None:
                e1.record(compute)
                kernel_events.append((e0, e1, r1 - r0))
            done.record(compute)
            copy.wait_event(done)
            with torch.cuda.stream(copy):
                out_host[r0:r1].copy_(out_dev[slot][: r1 - r0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
            copied[slot] = ev

    def drain():
        compute.wait_stream(copy)

    for _ in range(args.warmup):
        one_step()
    drain()
    barrier()
    def timed_region():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = _native.launch_count()
        events: list = []
        t_start = torch.cuda.Event(enable_timing=True); t_stop = torch.cuda.Event(enable_timing=True)
        barrier()
        t_start.record(compute)
        for _ in range(args.steps):
            one_step(events)
        drain()                      # every signature of every step is in host memory before the clock stops
        t_stop.record(compute)
        barrier()
        return (t_start, t_stop, events, _native.launch_count() - launches0,
                sampler.stop() if rank == 0 else None)

    start, stop, kernel_events, launches, clocks = timed_region()
    # a run that saw a hardware / thermal slowdown is rejected and taken again, once
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    redo = torch.tensor([1 if (clocks and bad & set(clocks.get("reasons", []))) else 0], device=dev)
    if world > 1:
        dist.broadcast(redo, src=0)
    remeasured = bool(redo.item())
    if remeasured:
        log(f"[rank {rank}] clocks show {clocks and clocks.get('reasons')}: measuring the timed region again")
        start, stop, kernel_events, launches, clocks = timed_region()
    if clocks is not None:
        clocks["remeasured"] = remeasured
    elapsed_ms = max_ranks(start.elapsed_time(stop))
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in kernel_events)
    kern_rows = sum(r for _, _, r in kernel_events)
    value = rows * world * args.steps / (elapsed_ms * 1e-3)
    kernel_name = hasher.last_kernel
    # the same job without the D2H gather: signatures left in HBM (sum of the kernels' own durations)
    kernel_only_value = rows * world * args.steps / (max_ranks(kern_ms) * 1e-3)

    # ---- e2e: host pinned buffers through the C ABI (H2D + kernel + D2H per step) --------
    e2e_rows = args.e2e_rows
    xh = torch.empty((e2e_rows, DIM), dtype=torch.float32, pin_memory=True)
    xh.copy_(X[:e2e_rows])
    oh = torch.empty((e2e_rows, SIG_BYTES), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def e2e_step():
        hasher.hash_into(xh, e2e_rows, oh, x_on_device=False, out_on_device=False)

    e2e_steps = 1 if args.no_e2e else args.steps
    for _ in range(1 if args.no_e2e else args.warmup):
        e2e_step()
    barrier()
    s2 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(e2e_steps):
        e2e_step()
    t2.record()
    barrier()
    e2e_ms = max_ranks(s2.elapsed_time(t2))
    e2e_value = e2e_rows * world * e2e_steps / (e2e_ms * 1e-3)
    # the e2e path must produce the same bytes as the resident path
    same = bool(torch.equal(oh, out_host[:e2e_rows]))
    # the same call on an ordinary (pageable) numpy array: pinned bounce buffers + parallel memcpy in the library
    pageable_value = None
    if not args.no_e2e:
        xp, op = xh.numpy().copy(), np.empty((e2e_rows, SIG_BYTES), dtype=np.uint8)
        hasher.hash_into(xp, e2e_rows, op, x_on_device=False, out_on_device=False)
        t_pg = time.perf_counter()
        for _ in range(3):
            hasher.hash_into(xp, e2e_rows, op, x_on_device=False, out_on_device=False)
        pageable_value = 3 * e2e_rows / (time.perf_counter() - t_pg)
        same = same and bool(np.array_equal(op, oh.numpy()))
        del xp, op
    # a float16 numpy array (what embedding stores often hold): raw rows over PCIe, exact cast on the device
    # (lshx_hash_batch_typed); the signatures must equal those of the same values given as float32
    f16_value = None
    if not args.no_e2e:
        x16 = xh.numpy().astype(np.float16)
        want16 = hasher.hash_batch_packed(x16.astype(np.float32)).reshape(e2e_rows, SIG_BYTES)
        got16 = hasher.hash_batch_packed(x16).reshape(e2e_rows, SIG_BYTES)
        t_16 = time.perf_counter()
        for _ in range(3):
            hasher.hash_batch_packed(x16)
        f16_value = 3 * e2e_rows / (time.perf_counter() - t_16)
        same = same and bool(np.array_equal(got16, want16))
        del x16, want16, got16
    # latency of the per-vector call LSHRS.ingest / query make (reference: one hash_vector per call)
    one = xh[:1].numpy().copy()
    for _ in range(20):
        hasher.hash_vector(one[0])
    t_lat = time.perf_counter()
    for _ in range(200):
        hasher.hash_vector(one[0])
    single_us = (time.perf_counter() - t_lat) / 200 * 1e6

    # ---- roofline of the projection kernel ------------------------------------------------
    ncols = SIG_BYTES * 8
    useful_flop = 2.0 * DIM * NUM_PERM  # what the reference computes; padding columns are not useful work
    # tensor-time units per logical product, in TF32 MMAs: 3xTF32 = 3; TF32 hi.hi + the two BF16 cross
    # terms (one K-doubled BF16 MMA = one TF32 MMA of tensor time) = 2; scaled FP16x3 (three FP16 MMAs,
    # each half a TF32 MMA) = 1.5
    mma_passes = {"tcgen05": 1.5, "tcgen05_tf32bf16": 2, "tcgen05_3xtf32": 3}.get(kernel_name, 1)
    is_tc = kernel_name.startswith("tcgen05")
    per_kernel_ms = kern_ms / max(1, len(kernel_events))
    # ALGORITHMIC flops (2*dim*num_perm per vector) against the ceiling for them: the
    # measured dense TF32 rate (= bf16_tflops_sustained / 2) divided by the tensor-time units the split
    # spends per logical product (SURVEY section 8d "useful ceiling": 3 for 3xTF32, 2 for TF32 + BF16
    # cross terms, 1.5 for the default scaled FP16x3).  The FFMA arm is held to the default arm's
    # ceiling: it is what the hardware can do for this arithmetic.
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    split_units = {"tcgen05_3xtf32": 3.0, "tcgen05_tf32bf16": 2.0}.get(kernel_name, 1.5)
    useful_peak = tf32_peak / split_units
    achieved_tflops = useful_flop * kern_rows / (kern_ms * 1e-3) / 1e12
    ncols_exec = (NUM_PERM + 15) // 16 * 16 if ROWS_PER_BAND % 8 else (ncols + 127) // 128 * 128
    if not is_tc:
        ncols_exec = (ncols + 127) // 128 * 128
    executed_tflops = 2.0 * DIM * ncols_exec * mma_passes * kern_rows / (kern_ms * 1e-3) / 1e12
    bytes_per_vec = 4.0 * DIM + SIG_BYTES
    hbm_gbs = bytes_per_vec * kern_rows / (kern_ms * 1e-3) / 1e9
    tensor_frac, hbm_frac = achieved_tflops / useful_peak, hbm_gbs / peaks["hbm_gbs"]
    common = {
        "kernel": f"hash_{kernel_name}", "traffic": (ROOFLINE_TRAFFIC_PER_ROW[(args.workload, "tcgen05")] * kern_rows / max(1, len(kernel_events))
                    if is_tc and (args.workload, "tcgen05") in ROOFLINE_TRAFFIC_PER_ROW else None),
        "avg_launch_ms": per_kernel_ms, "launches_timed": len(kernel_events),
        "kernel_share_of_step": kern_ms / (start.elapsed_time(stop)),
        "algorithmic_bytes_per_launch": bytes_per_vec * kern_rows / max(1, len(kernel_events)),
        "algorithmic_flops_per_launch": useful_flop * kern_rows / max(1, len(kernel_events)),
        "tensor": {"achieved_tflops": achieved_tflops, "peak_tflops": useful_peak, "frac": tensor_frac,
                   "peak_source": (f"{peak_src}: bf16_tflops_sustained / 2 (dense TF32) / 3 (3xTF32 split: three MMAs "
                                   "per logical product)" if split_units == 3.0 else
                                   f"{peak_src}: bf16_tflops_sustained / 2 (dense TF32) / 2 (split x = hi + lo: one TF32 MMA "
                                   "for hi.hi + one K-doubled BF16 MMA of the same duration for both cross terms)"
                                   if split_units == 2.0 else
                                   f"{peak_src}: bf16_tflops_sustained / 3 (scaled FP16x3 split x = hi + lo: three dense "
                                   "FP16 MMAs -- hi.hi, hi.lo, lo.hi -- per logical product; FP16 runs at the BF16 rate)"),
                   "flops_counted": "algorithmic: 2*dim*num_perm per vector",
                   "executed_tflops": executed_tflops, "executed_vs_tf32_peak": executed_tflops / tf32_peak,
                   "fp32_pipe_nominal_tflops": 148 * 128 * 2 * 1.965e9 / 1e12},
        "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks["hbm_gbs"], "frac": hbm_frac, "peak_source": peak_src,
                "bytes_per_vector": bytes_per_vec},
    }
    if tensor_frac >= hbm_frac:  # the slower of the two ceilings is the bound
        roofline = {"bound": "tensor", "achieved": achieved_tflops, "peak": useful_peak, "unit": "TFLOP/s",
                    "frac": tensor_frac, **common}
    else:
        roofline = {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": hbm_frac, **common}

    # ---- CPU baseline + parity on a bounded sample (rank 0, N = 1 only) ---------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import lshrs_oracle as oracle

        projs = oracle.make_projections(NUM_BANDS, ROWS_PER_BAND, DIM, SEED)
        sample = args.cpu_sample
        idx = torch.linspace(0, rows - 1, sample, device=dev).long()
        Xs = X[idx].cpu().numpy()
        t0 = time.perf_counter()
        ref = oracle.hash_batch_packed(projs, Xs)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        oracle.hash_batch_vectorized(projs, Xs)
        dtv = time.perf_counter() - t1
        got = out_host[idx.cpu()].numpy().reshape(sample, NUM_BANDS, -1)
        parity = oracle.compare_packed(got, ref, oracle.projection_margins(projs, Xs), 1e-5)
        parity["sample_rows"] = sample
        parity["e2e_equals_resident"] = same
        cpu_baseline = {
            "value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample} rows of the same shard through the oracle's per-vector, per-band sgemv loop "
                      f"(reference LSHHasher.hash_batch path, single Python thread by construction), {dt:.1f} s",
            "host_cores": os.cpu_count(), "numpy": np.__version__,
            "vectorized_numpy_not_reference": {"value": sample / dtv, "unit": UNIT,
                                               "cores": len(os.sched_getaffinity(0))},
        }

    # ---- rerank (config 4) ------------------------------------------------------------------
    rerank = None
    api = None
    if rank == 0 and world == 1 and not args.no_e2e and args.workload == "hash768":
        try:
            api = run_api(args, dev)
        except Exception as exc:  # noqa: BLE001 -- a side measurement must not take the headline down
            api = {"error": repr(exc)}
    if not args.no_rerank:
        del X, out_dev
        torch.cuda.empty_cache()
        rerank = run_rerank(args, dev, rank, world, peaks, peak_src, barrier, max_ranks)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"tcgen05": "f16x3", "tcgen05_tf32bf16": "tf32+bf16x2", "tcgen05_3xtf32": "tf32x3"}.get(kernel_name, "f32"),
            "data": "synthetic", "config": workload_config(args), "impl": "b200", "kernel": kernel_name,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_rows * DIM * 4,
                    "d2h_bytes_per_step": e2e_rows * SIG_BYTES, "rows_per_step_per_gpu": e2e_rows,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": "lshx_hash_batch(host pinned X -> host pinned signatures)",
                    "single_vector_call_us": single_us,
                    "pageable_numpy_input_value": pageable_value, "float16_numpy_input_value": f16_value},
            "kernel_only": {"value": kernel_only_value, "unit": UNIT,
                            "note": "signatures left in HBM (no D2H gather); value above includes the overlapped D2H "
                                    "of every signature into pinned host memory"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
            "cpu_baseline": cpu_baseline, "parity": parity, "rerank": rerank, "api": api,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_api(args, dev) -> dict:
    """Throughput of the reference-facing LSHRS calls on one GPU.

    Host data in, Python results out, in-memory bucket storage (the reference's Redis stays a host
    component).  These are dominated by host-side key / bucket handling, not by the kernels; they are here
    so that the drop-in API has numbers next to the kernel figures.  SURVEY section 8a gives the reference's own
    figures with the same kind of storage double, measured in the build container: index 2.2 k vectors/s,
    get_top_k 3.9 k queries/s, get_above_p 2.7 k queries/s.
    """
    import torch

    from lshrs_b200 import LSHRS, InMemoryStorage

    n_index, n_single, n_batch = 50_000, 300, 2_048
    rng = np.random.default_rng(11)
    centers = rng.standard_normal((n_index // 8, DIM)).astype(np.float32)
    Xh = (np.repeat(centers, 8, axis=0) + 0.15 * rng.standard_normal((n_index, DIM))).astype(np.float32)
    extra = (Xh[:n_single] + 0.05 * rng.standard_normal((n_single, DIM))).astype(np.float32)
    allvec = np.concatenate([Xh, extra])          # id -> vector, what vector_fetch_fn serves
    corpus_dev = torch.from_numpy(allvec).to(dev)
    lsh = LSHRS(dim=DIM, num_perm=NUM_PERM, storage=InMemoryStorage(), vector_fetch_fn=lambda ids: allvec[ids])
    ids = list(range(n_index))
    t0 = time.perf_counter()
    lsh.index(ids, Xh)
    index_vps = n_index / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for i in range(n_single):
        lsh.ingest(n_index + i, extra[i])
    lsh.flush()
    ingest_cps = n_single / (time.perf_counter() - t0)
    Q = (Xh[rng.integers(0, n_index, n_batch)] + 0.05 * rng.standard_normal((n_batch, DIM))).astype(np.float32)
    lsh.get_top_k(Q[-1], topk=10)        # first calls create the rerank handle / workspaces: off the clock
    lsh.get_above_p(Q[-1], p=0.2)
    t0 = time.perf_counter()
    got_k = [lsh.get_top_k(Q[i], topk=10) for i in range(n_single)]
    topk_qps = n_single / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    got_p = [lsh.get_above_p(Q[i], p=0.2) for i in range(n_single)]
    above_qps = n_single / (time.perf_counter() - t0)
    lsh.query_batch(Q[:64], top_k=10, top_p=0.2, corpus=corpus_dev)
    t0 = time.perf_counter()
    batch = lsh.query_batch(Q, top_k=10, top_p=0.2, corpus=corpus_dev)
    batch_qps = n_batch / (time.perf_counter() - t0)
    # the batched call must return what the per-query call returns
    same = all([i for i, _ in batch[j]] == [i for i, _ in lsh.query(Q[j], top_k=10, top_p=0.2)]
               for j in range(0, 64))
    return {
        "storage": "InMemoryStorage (bucket sets in a Python dict; no Redis on the box)",
        "index": {"value": index_vps, "unit": "vectors/s", "rows": n_index},
        "ingest": {"value": ingest_cps, "unit": "calls/s", "calls": n_single},
        "get_top_k": {"value": topk_qps, "unit": "queries/s", "calls": n_single,
                      "mean_results": float(np.mean([len(r) for r in got_k]))},
        "get_above_p": {"value": above_qps, "unit": "queries/s", "calls": n_single,
                        "mean_results": float(np.mean([len(r) for r in got_p]))},
        "query_batch": {"value": batch_qps, "unit": "queries/s", "queries": n_batch,
                        "corpus": "device-resident", "equals_per_query_calls": bool(same)},
        "reference_survey_values": {"index": 2.2e3, "get_top_k": 3.9e3, "get_above_p": 2.7e3,
                                    "note": "SURVEY section 8a, reference with its MockStorage in the build container"},
    }


def run_rerank(args, dev, rank, world, peaks, peak_src, barrier, max_ranks) -> dict:
    """BASELINE config 4: 8192 queries x 2000 candidates x 768 from a 1M-row corpus resident in HBM."""
    import torch

    from lshrs_b200 import _native
    from lshrs_b200.utils.similarity import _get_reranker

    N, nq, nc = args.corpus, args.queries, 2000
    gen = torch.Generator(device=dev).manual_seed(1)
    corpus = torch.empty((N, DIM), dtype=torch.float32, device=dev)
    corpus.normal_(generator=gen)
    Q = torch.empty((nq, DIM), dtype=torch.float32, device=dev).normal_(generator=gen)
    # 2000 DISTINCT corpus rows per query: random start, random stride < N / nc
    first = torch.randint(0, N, (nq, 1), generator=gen, device=dev, dtype=torch.int64)
    stride = torch.randint(1, max(2, N // nc), (nq, 1), generator=gen, device=dev, dtype=torch.int64)
    ids = ((first + stride * torch.arange(nc, device=dev, dtype=torch.int64)[None, :]) % N).contiguous()
    offs = torch.arange(nq + 1, device=dev, dtype=torch.int64) * nc
    rer = _get_reranker(DIM, dev.index)
    lib = _native.lib()
    out = {}
    for tag, k, p in (("k10", 10, 0.0), ("p0.2", 0, 0.2)):
        limit = k if k > 0 else int(np.ceil(nc * p))
        pos = torch.empty((nq, limit), dtype=torch.int32, device=dev)
        score = torch.empty((nq, limit), dtype=torch.float32, device=dev)
        count = torch.empty(nq, dtype=torch.int32, device=dev)
        zero = torch.empty(nq, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream

        def step():
            _native.check(lib.lshx_rerank_topk(
                rer._handle, Q.data_ptr(), nq, corpus.data_ptr(), N, offs.data_ptr(), ids.data_ptr(), nc,
                k, p, limit, pos.data_ptr(), score.data_ptr(), count.data_ptr(), zero.data_ptr(), 1, stream))

        for _ in range(args.warmup):
            step()
        barrier()
        s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            step()
        t.record()
        barrier()
        ms = max_ranks(s.elapsed_time(t)) / args.steps
        bytes_per_q = nc * (4 * DIM + 8) + 4 * DIM + 8 * limit
        gbs = bytes_per_q * nq / (ms * 1e-3) / 1e9
        # e2e: PINNED host queries / ids / results through the C ABI, corpus resident (on_device = 2)
        def pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h.numpy()

        Qh, idh, offh = pinned(Q), pinned(ids).reshape(-1), pinned(offs)
        outs = (pinned(pos), pinned(score), pinned(count), pinned(zero))
        for _ in range(2):
            rer.topk(Qh, corpus, offh, idh, k=k, p=p, vectors_on_device=True, out=outs)
        barrier()
        s3 = torch.cuda.Event(enable_timing=True); t3 = torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(args.steps):
            ph, sh, ch, zh = rer.topk(Qh, corpus, offh, idh, k=k, p=p, vectors_on_device=True, out=outs)
        t3.record()
        torch.cuda.synchronize()
        e2e_ms = max_ranks(s3.elapsed_time(t3)) / args.steps
        entry = {
            "value": nq * world / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "results_per_query": limit,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                         "bytes_per_query": bytes_per_q},
            "e2e": {"value": nq * world / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(Qh.nbytes + idh.nbytes + offh.nbytes),
                    "d2h_bytes_per_step": int(ph.nbytes + sh.nbytes + ch.nbytes + zh.nbytes)},
            "device_equals_e2e": bool(np.array_equal(ph, pos.cpu().numpy())),
        }
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle import lshrs_oracle as oracle

            nref = 64
            Ch = corpus.cpu().numpy()
            t1 = time.perf_counter()
            bad = 0
            for i in range(nref):
                ref = oracle.top_k_cosine(Qh[i], Ch[idh[i * nc:(i + 1) * nc]], k=limit)
                got_pos = ph[i, :limit]
                ref_scores = np.array([sc for _, sc in ref])
                if not np.allclose(sh[i, :limit], ref_scores, atol=1e-5, rtol=0):
                    bad += 1
                elif set(got_pos.tolist()) != {pp for pp, _ in ref}:
                    bad += 1
            dt = time.perf_counter() - t1
            entry["cpu_baseline"] = {"value": nref / dt, "unit": "queries/s", "cores": 1, "kind": "port",
                                     "sample": f"{nref} queries through the oracle's top_k_cosine"}
            entry["parity"] = {"queries_checked": nref, "mismatching_queries": bad, "score_tol": 1e-5}
            del Ch
        out[tag] = entry
    out["config"] = {"workload": f"top_k_cosine rerank: {nq} queries x {nc} candidates, dim={DIM}, corpus {N} rows "
                                 "resident in HBM, CSR int64 candidate ids", "l2": "6.1 MB of gathers per query, "
                                 f"{nq * nc * DIM * 4 / 1e9:.1f} GB per step (larger than L2)"}
    return out


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line, on the real stdout (everything else in this process goes to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_OUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_OUT, data)


def _route_stdout_to_stderr() -> None:
    """NCCL / torch print banners on fd 1 (e.g. "NCCL version ..."); keep stdout for the JSON line only."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)


def main() -> None:
    _route_stdout_to_stderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--workload", choices=tuple(WORKLOADS), default="hash768")
    ap.add_argument("--rows", type=int, default=None, help="resident vectors per GPU")
    ap.add_argument("--chunk", type=int, default=None, help="rows per kernel launch / D2H copy")
    ap.add_argument("--e2e-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=65_536)
    ap.add_argument("--kernel", choices=("auto", "ffma", "tcgen05", "tcgen05_3xtf32", "tcgen05_tf32bf16"), default="auto")
    ap.add_argument("--corpus", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=8192)
    ap.add_argument("--no-rerank", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (profiling runs)")
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    if args.impl == "reference":
        args.workload = "hash768"
    select_workload(args)
    if args.workload != "hash768":
        args.no_rerank = True
    args.warmup = max(3, args.warmup) if args.impl == "b200" else max(0, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
