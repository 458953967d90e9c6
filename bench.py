#!/usr/bin/env python
"""Benchmark of the lshrs hot path on B200: vectors hashed/s (dim 768, 256 bits) and rerank queries/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One process per GPU; the work list is partitioned (weak scaling: every rank hashes its own resident shard of
``--rows`` vectors, 12.5 M = 100 M / 8 by default -- BASELINE.json config 2 at N = 8), the projection matrix is
replicated, nothing is exchanged on the data path.  A "step" is one pass of the hot path over one resident shard:
the projection kernel over every chunk of the shard plus the gather of the signatures into pinned host memory,
overlapped on copy streams.  Timed with CUDA events on the launching stream, max over ranks.

Multi-GPU (lshrs_b200/fabric.py): the host links of a box are not equal, so (a) a job on fewer GPUs than the box
has takes the GPUs on the fastest links (rank 0 probes every visible GPU in a child process), and (b) every rank
measures its own link with everybody copying and with only the faster half copying; when that pays, the ranks on
slow links hand their signatures over NVLink (peer-to-peer copy engine, no NCCL) to a partner on a fast link that
writes them into the sender's pinned host buffer.  The ``fabric`` block of the JSON line records the measured
rates, the chosen plan and ``value`` as a fraction of what the measured links allow.

The JSON line carries, beside the contract keys:
  roofline      dominant (projection) kernel: algorithmic flops / its own CUDA-event time inside the timed region
  e2e           same metric through the C ABI with HOST pinned buffers (H2D of the vectors and D2H of the
                signatures inside the timed region); at N > 1 the per-rank row counts follow the measured H2D rates
  cpu_baseline  the reference's own LSHHasher.hash_batch (oracle/_ref/reference, staged by tools/make_ref.py; the
                oracle port when it is not staged) timed on this box's host cores on a bounded sample (N = 1)
  parity        band-key comparison of the GPU output with the CPU baseline's output on that sample
  configs       BASELINE configs 3 and 5 (1536 -> 512 bits, SIFT-like 128 -> 64 bits): value, roofline, parity
  rerank        config 4 (8192 queries x 2000 candidates x 768, k = 10 and p = 0.2)
  fabric        measured host-link rates, device choice, relay plan (N > 1)
``--impl reference`` times the reference's CPU implementation (same source as ``cpu_baseline``) on bounded samples.
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

T_PROCESS_START = time.time() - 2.0
NBUF = 4   # device-side signature buffers per rank: kernels run up to NBUF - 1 chunks ahead of the gather
SEED = 42
UNIT = "vectors/s"
# BASELINE.json configs; "hash768" is the metric's configuration, the others ride along in `configs`
WORKLOADS = {
    "hash768": dict(dim=768, bands=16, rows_per_band=16, rows=12_500_000, chunk=781_250, dist="gauss"),
    "hash1536": dict(dim=1536, bands=16, rows_per_band=32, rows=6_250_000, chunk=781_250, dist="gauss"),
    "hash128": dict(dim=128, bands=16, rows_per_band=4, rows=100_000_000, chunk=12_500_000, dist="sift"),
}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel PER ROW from the committed ncu --set full
# captures (file, bytes per launch, rows per launch); scaled to the rows of one launch of the run
NCU_TRAFFIC = {
    "hash768": ("profiles/r2_hash_tc_ncu.csv", 2.429e9, 781_250),
    "hash128": ("profiles/r2_hash_tc_dim128_ncu.csv", 6.600e9, 12_500_000),
    "rerank": ("profiles/r2_rerank_ncu.csv", 12.345e9, 2048),          # per 2048-query launch of 2000 x 768 candidates
}


class Shape:
    """One hashing configuration (what LSHRS(dim=..., num_perm=...) auto-selects for the BASELINE shapes)."""

    def __init__(self, name: str):
        w = WORKLOADS[name]
        self.name, self.dim, self.bands, self.rows_per_band = name, w["dim"], w["bands"], w["rows_per_band"]
        self.num_perm = self.bands * self.rows_per_band
        self.sig_bytes = self.bands * ((self.rows_per_band + 7) // 8)
        self.rows, self.chunk, self.dist = w["rows"], w["chunk"], w["dist"]
        self.metric = f"vectors hashed/sec (dim={self.dim}, num_perm={self.num_perm})"


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


def workload_config(args, shape: Shape) -> dict:
    """The `config` object: identical for the b200 and the reference arm (the CPU arm's sample size is stated in
    its cpu_baseline.sample, not here)."""
    return {
        "workload": f"LSHHasher dim={shape.dim} num_perm={shape.num_perm} ({shape.bands}x{shape.rows_per_band}); "
                    f"{args.rows} synthetic {'Gaussian' if shape.dist == 'gauss' else 'SIFT-like non-negative'} "
                    f"float32 vectors resident per GPU"
                    + (" (BASELINE config 2: 100M vectors / 8 GPUs = 12.5M per GPU, weak scaling)"
                       if shape.name == "hash768" else f" (BASELINE config '{shape.name}')"),
        "rows_per_gpu": args.rows, "dim": shape.dim, "num_perm": shape.num_perm, "signature_bytes": shape.sig_bytes,
        "chunk_rows": args.chunk, "e2e_rows_per_gpu": args.e2e_rows,
        "l2": f"inputs larger than L2 ({args.rows * shape.dim * 4 / 1e9:.0f} GB resident shard, each row read once per step)",
        "parallelism": f"row-sharded x{args.gpus}, projections replicated, no collective",
    }


# --------------------------------------------------------------------------------------
# the CPU side: the reference itself when it is staged (oracle/_ref/reference), else the oracle port
# --------------------------------------------------------------------------------------

class CpuReference:
    """The reference's CPU implementation of the path, as shipped.

    ``kind == "reference"``: the unmodified mxngjxa/lshrs package staged by tools/make_ref.py under the
    git-ignored oracle/_ref/reference (it travels to the GPU box with the snapshot; /root/reference does not
    exist there), imported with the stub ``redis`` module beside it.  ``kind == "port"``: oracle/lshrs_oracle.py,
    the numpy restatement pinned to the reference's outputs (tests/golden), when nothing is staged.
    """

    def __init__(self):
        ref = REPO / "oracle" / "_ref"
        self.kind = "port"
        if (ref / "reference" / "lshrs" / "hash" / "lsh.py").exists():
            try:
                sys.path[:0] = [str(ref / "reference"), str(ref / "stubs")]
                from lshrs.hash.lsh import LSHHasher as RefHasher
                from lshrs.utils.similarity import top_k_cosine as ref_topk

                self._hasher_cls, self._topk, self.kind = RefHasher, ref_topk, "reference"
            except Exception as exc:  # noqa: BLE001
                log(f"staged reference not importable ({exc!r}); using the oracle port")
                sys.path[:] = [p for p in sys.path if not p.startswith(str(ref))]
        from oracle import lshrs_oracle as oracle

        self.oracle = oracle
        self.source = ("oracle/_ref/reference/lshrs (unmodified reference, tools/make_ref.py)" if self.kind == "reference"
                       else "oracle/lshrs_oracle.py (numpy restatement pinned to tests/golden)")

    def hasher(self, shape: Shape):
        if self.kind == "reference":
            h = self._hasher_cls(num_bands=shape.bands, rows_per_band=shape.rows_per_band, dim=shape.dim, seed=SEED)
            return h, h.projections
        projs = self.oracle.make_projections(shape.bands, shape.rows_per_band, shape.dim, SEED)
        return None, projs

    def hash_batch(self, h, projs, X):
        """What the reference does: LSHHasher.hash_batch (one hash_vector per row, one sgemv per band)."""
        if h is not None:
            return h.hash_batch(X)
        return self.oracle.hash_batch(projs, X)

    @staticmethod
    def packed(sigs, shape: Shape) -> np.ndarray:
        bpb = (shape.rows_per_band + 7) // 8
        flat = b"".join(b for s in sigs for b in s)
        return np.frombuffer(flat, dtype=np.uint8).reshape(len(sigs), shape.bands, bpb)

    def top_k_cosine(self, q, C, k):
        if self.kind == "reference":
            return self._topk(q, C, k=k)
        return self.oracle.top_k_cosine(q, C, k=k)


def cpu_hash_sample(shape: Shape, rows: int, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((rows, shape.dim)).astype(np.float32)


def _mp_hash_worker(job):
    seed, rows = job
    shape = Shape("hash768")
    cpu = CpuReference()
    h, projs = cpu.hasher(shape)
    X = cpu_hash_sample(shape, rows, seed=seed)
    t0 = time.perf_counter()
    cpu.hash_batch(h, projs, X)
    return time.perf_counter() - t0


def multiprocess_rate(rows_per_worker: int = 4096) -> dict:
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


def run_reference(args, shape: Shape) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    cpu = CpuReference()
    h, projs = cpu.hasher(shape)
    sample = 16384
    X = cpu_hash_sample(shape, sample)
    for _ in range(args.warmup):
        cpu.hash_batch(h, projs, X[:1024])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.hash_batch(h, projs, X)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    # not reference code: one sgemm + packbits over every BLAS thread, for honesty (torchrun exports
    # OMP_NUM_THREADS=1, so the BLAS pool is widened explicitly when threadpoolctl is there)
    Xv = cpu_hash_sample(shape, 131072, seed=1)
    try:
        from threadpoolctl import threadpool_limits

        blas_threads = threadpool_limits(limits=len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        blas_threads = None
    cpu.oracle.hash_batch_vectorized(projs, Xv[:4096])
    t1 = time.perf_counter()
    cpu.oracle.hash_batch_vectorized(projs, Xv)
    vec_value = Xv.shape[0] / (time.perf_counter() - t1)
    if blas_threads is not None:
        blas_threads.restore_original_limits()
    line = {
        "impl": "reference", "metric": shape.metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, shape),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": cpu.kind, "source": cpu.source,
                         "sample": f"{sample} Gaussian vectors of the config's shape per step x {args.steps} steps "
                                   "through LSHHasher.hash_batch -- per-vector, per-band sgemv, a single Python "
                                   "thread by construction (reference lshrs/hash/lsh.py:169); a rate, so the "
                                   "bounded sample stands for the config's 12.5 M rows per GPU",
                         "sample_rows_per_step": sample, "host_cores": os.cpu_count(), "numpy": np.__version__,
                         "vectorized_numpy_not_reference": {"value": vec_value, "unit": UNIT,
                                                            "cores": len(os.sched_getaffinity(0))},
                         "multiprocess_not_reference": multiprocess_rate()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
# B200 arm
# --------------------------------------------------------------------------------------

def make_shard(torch, dev, shape: Shape, rows: int, seed: int):
    """The resident shard, generated on the device (seeded per rank): Gaussian, or SIFT-like non-negative
    integer-valued floats in [0, 255] for the 128-dimensional configuration."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    X = torch.empty((rows, shape.dim), dtype=torch.float32, device=dev)
    for r0 in range(0, rows, 1 << 20):
        r1 = min(rows, r0 + (1 << 20))
        X[r0:r1].normal_(generator=gen)
        if shape.dist == "sift":
            X[r0:r1].abs_().mul_(40.0).floor_().clamp_(max=255.0)
    return X


class ResidentRun:
    """One rank's pass over its resident shard: the projection kernel per chunk on `compute`, the gather of each
    chunk's signatures into pinned host memory on a copy stream (a ring of NBUF device buffers) -- or, for a rank on a slow
    host link, into its partner's device slot over NVLink (fabric.RelaySender); a partner rank also drains the
    sender's chunks into the sender's host buffer (fabric.RelayReceiver)."""

    def __init__(self, torch, dev, hasher, X, shape: Shape, chunk: int, out_host):
        self.torch, self.dev, self.hasher, self.X, self.shape = torch, dev, hasher, X, shape
        rows = X.shape[0]
        self.rows = rows
        self.chunks = [(r0, min(rows, r0 + chunk)) for r0 in range(0, rows, chunk)]
        self.out_host = out_host
        self.nbuf = NBUF
        self.out_dev = [torch.empty((chunk, shape.sig_bytes), dtype=torch.uint8, device=dev) for _ in range(self.nbuf)]
        self.compute = torch.cuda.Stream(device=dev)
        self.copy = torch.cuda.Stream(device=dev)
        self.copied = [None] * self.nbuf   # per output slot: the event after which it may be overwritten
        self.sender = None
        self.receiver = None
        self.k = 0                   # relayed chunks issued so far (both partners count alike)
        self.next_slot = 0

    def one_step(self, kernel_events=None):
        """Steps stream into each other: the gather of a step's last chunks overlaps the next step's first
        kernels; the timed region ends only after every signature is in host memory (drain())."""
        torch, compute, sig = self.torch, self.compute, self.shape.sig_bytes
        for r0, r1 in self.chunks:
            slot = self.next_slot
            self.next_slot = (slot + 1) % self.nbuf
            if self.copied[slot] is not None:
                compute.wait_event(self.copied[slot])
            if kernel_events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(compute)
            self.hasher.hash_into(self.X[r0:r1], r1 - r0, self.out_dev[slot], x_on_device=True, out_on_device=True,
                                  stream=compute.cuda_stream)
            if kernel_events is not None:
                e1.record(compute)
                kernel_events.append((e0, e1, r1 - r0))
            done = torch.cuda.Event()
            done.record(compute)
            if self.sender is not None:
                self.copied[slot] = self.sender.send(self.k, self.out_dev[slot][: r1 - r0], done)
            else:
                self.copy.wait_event(done)
                with torch.cuda.stream(self.copy):
                    self.out_host[r0:r1].copy_(self.out_dev[slot][: r1 - r0], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy)
                self.copied[slot] = ev
            if self.receiver is not None:
                self.receiver.drain(self.k, (r1 - r0) * sig, r0 * sig)
            if self.sender is not None or self.receiver is not None:
                self.k += 1

    def drain(self):
        self.compute.wait_stream(self.copy)
        if self.sender is not None:
            self.compute.wait_stream(self.sender.stream)
        if self.receiver is not None:
            self.compute.wait_stream(self.receiver.stream)


def roofline_block(shape: Shape, kernel_name: str, kern_ms: float, kern_rows: int, n_events: int, step_ms: float,
                   peaks: dict, peak_src: str) -> dict:
    """ALGORITHMIC flops (2*dim*num_perm per vector) and bytes (4*dim + signature bytes) over
    the kernel's own CUDA-event time, against the measured ceilings.  Tensor ceiling: the measured dense TF32 rate
    (= bf16_tflops_sustained / 2) divided by the tensor-time units the operand split spends per logical product
    (3 for 3xTF32, 2 for TF32 + BF16 cross terms, 1.5 for the default scaled FP16x3).  The slower ceiling is the
    bound."""
    ncols = shape.sig_bytes * 8
    useful_flop = 2.0 * shape.dim * shape.num_perm
    mma_passes = {"tcgen05": 1.5, "tcgen05_tf32bf16": 2, "tcgen05_3xtf32": 3}.get(kernel_name, 1)
    is_tc = kernel_name.startswith("tcgen05")
    n_events = max(1, n_events)
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    split_units = {"tcgen05_3xtf32": 3.0, "tcgen05_tf32bf16": 2.0}.get(kernel_name, 1.5)
    useful_peak = tf32_peak / split_units
    achieved_tflops = useful_flop * kern_rows / (kern_ms * 1e-3) / 1e12
    ncols_exec = (shape.num_perm + 15) // 16 * 16 if shape.rows_per_band % 8 else (ncols + 127) // 128 * 128
    if not is_tc:
        ncols_exec = (ncols + 127) // 128 * 128
    executed_tflops = 2.0 * shape.dim * ncols_exec * mma_passes * kern_rows / (kern_ms * 1e-3) / 1e12
    bytes_per_vec = 4.0 * shape.dim + shape.sig_bytes
    hbm_gbs = bytes_per_vec * kern_rows / (kern_ms * 1e-3) / 1e9
    tensor_frac, hbm_frac = achieved_tflops / useful_peak, hbm_gbs / peaks["hbm_gbs"]
    traffic, traffic_src = None, None
    if is_tc and kernel_name == "tcgen05" and shape.name in NCU_TRAFFIC:
        f, per_launch, rows_per_launch = NCU_TRAFFIC[shape.name]
        traffic = per_launch / rows_per_launch * kern_rows / n_events
        traffic_src = f"{f}: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled by rows"
    common = {
        "kernel": f"hash_{kernel_name}", "traffic": traffic, "traffic_source": traffic_src,
        "avg_launch_ms": kern_ms / n_events, "launches_timed": n_events,
        "kernel_share_of_step": kern_ms / step_ms if step_ms else None,
        "algorithmic_bytes_per_launch": bytes_per_vec * kern_rows / n_events,
        "algorithmic_flops_per_launch": useful_flop * kern_rows / n_events,
        "tensor": {"achieved_tflops": achieved_tflops, "peak_tflops": useful_peak, "frac": tensor_frac,
                   "peak_source": f"{peak_src}: bf16_tflops_sustained / 2 (dense TF32) / {split_units:g} tensor-time "
                                  "units per logical product of the operand split (3xTF32: 3; TF32 hi.hi + BF16 "
                                  "cross terms: 2; scaled FP16x3: 1.5)",
                   "flops_counted": "algorithmic: 2*dim*num_perm per vector",
                   "executed_tflops": executed_tflops, "executed_vs_tf32_peak": executed_tflops / tf32_peak,
                   "fp32_pipe_nominal_tflops": 148 * 128 * 2 * 1.965e9 / 1e12},
        "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks["hbm_gbs"], "frac": hbm_frac, "peak_source": peak_src,
                "bytes_per_vector": bytes_per_vec},
    }
    if tensor_frac >= hbm_frac:  # the slower of the two ceilings is the bound
        return {"bound": "tensor", "achieved": achieved_tflops, "peak": useful_peak, "unit": "TFLOP/s",
                "frac": tensor_frac, **common}
    return {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_frac, **common}


def link_rate(torch, dev, direction: str, active: bool, barrier, mbytes: int = 256, chunk_mb: int = 25) -> float:
    """This rank's D2H / H2D GB/s while every `active` rank copies at once (0.0 for a rank that sits out)."""
    nbytes, chunk = mbytes << 20, chunk_mb << 20
    if active:
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        s = torch.cuda.Stream(dev)
    best = 0.0
    for rep in range(3):
        barrier()
        if not active:
            continue
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for o in range(0, nbytes, chunk):
                if direction == "d2h":
                    h[o:o + chunk].copy_(d[o:o + chunk], non_blocking=True)
                else:
                    d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        s.synchronize()
        if rep:
            best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    barrier()
    return best


def run_b200(args, shape: Shape) -> None:
    import torch
    import torch.distributed as dist

    from lshrs_b200 import LSHHasher, _native, fabric

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    visible = torch.cuda.device_count()
    fab: dict = {"visible_gpus": visible}

    # ---- which GPU does this rank take?  (fewer ranks than GPUs: the ones on the fastest host links) ----------
    device_index = local
    if not args.no_fabric and world < visible:
        # rank 0 measures every visible GPU's host link in a child process and leaves the result in /dev/shm; the
        # other ranks (which must not touch CUDA before they know their GPU) wait for a file newer than themselves
        probe_path = Path(f"/dev/shm/lshx_probe_{os.environ.get('MASTER_PORT', '0')}.json")
        box = None
        if rank == 0:
            box = fabric.probe_links_subprocess()
            tmp = probe_path.with_suffix(".tmp")
            tmp.write_text(json.dumps({"probe": box, "world": world}))
            os.replace(tmp, probe_path)
        else:
            deadline = time.time() + 240
            while time.time() < deadline:
                try:
                    if probe_path.stat().st_mtime >= T_PROCESS_START:
                        got = json.loads(probe_path.read_text())
                        if got.get("world") == world:
                            box = got["probe"]
                            break
                except (OSError, ValueError):
                    pass
                time.sleep(0.05)
        devices = fabric.choose_devices(box, world, visible)
        device_index = devices[local] if local < len(devices) else local
        fab.update(box_probe=box, devices=devices,
                   device_choice="the GPUs with the fastest D2H links while every visible GPU copies")
    else:
        fab.update(devices=list(range(world)), device_choice="identity (every visible GPU is used)" if world >= visible
                   else "identity (--no-fabric)")
    torch.cuda.set_device(device_index)
    dev = torch.device("cuda", device_index)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL for the barrier and the MAX-reduction of the measured time, gloo for the small Python objects of
        # the set-up (link rates, relay handles); nothing of either is on the data path
        dist.init_process_group("cpu:gloo,cuda:nccl")
    log(f"[rank {rank}] cuda:{device_index} of {visible} visible")
    peaks, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[device_index])
        torch.cuda.synchronize()

    def max_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_obj(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    hasher = LSHHasher(shape.bands, shape.rows_per_band, shape.dim, seed=SEED, device=device_index)
    if args.kernel != "auto":
        hasher._ensure_handle()
        hasher.set_kernel(args.kernel)

    # ---- resident shard + host buffer ---------------------------------------------------------------------
    rows, chunk = args.rows, args.chunk
    X = make_shard(torch, dev, shape, rows, 1000 + rank)
    shm = None
    if world > 1 and not args.no_fabric:
        # POSIX shared memory, pinned: a partner rank may write this rank's signatures over ITS PCIe link
        shm = fabric.SharedHostBuffer(f"lshx_sig_{os.environ.get('MASTER_PORT', '0')}_{rank}", rows * shape.sig_bytes,
                                      create=True)
        out_host = shm.pin().view(rows, shape.sig_bytes)
    else:
        out_host = torch.empty((rows, shape.sig_bytes), dtype=torch.uint8, pin_memory=True)
    run = ResidentRun(torch, dev, hasher, X, shape, chunk, out_host)

    kev: list = []
    for _ in range(args.warmup):
        run.one_step(kev)
    run.drain()
    barrier()
    # signature bytes one GPU produces per second (best launch of the warm-up: the first ones pay module loading)
    kernel_gbs = max(shape.sig_bytes * r / (e0.elapsed_time(e1) * 1e-3) / 1e9 for e0, e1, r in kev)

    # ---- host links: measured, then the gather plan ------------------------------------------------------------
    plan = {"policy": "direct", "pairs": {}}
    if world > 1 and not args.no_fabric:
        d2h_all = gather_obj(link_rate(torch, dev, "d2h", True, barrier))
        h2d_all = gather_obj(link_rate(torch, dev, "h2d", True, barrier))
        order = sorted(range(world), key=lambda r: (-d2h_all[r], r))
        writers = set(order[: world // 2])
        d2h_writers = gather_obj(link_rate(torch, dev, "d2h", rank in writers, barrier))
        plan = fabric.plan_relay(d2h_all, d2h_writers, kernel_gbs)
        if args.relay == "off":
            plan.update(policy="direct", pairs={}, forced="--relay off")
        elif args.relay == "force" and world % 2 == 0 and plan["policy"] != "relay":
            half = world // 2                          # bring-up / tests: relay even where it does not pay
            plan.update(policy="relay", pairs={int(s): int(w) for s, w in zip(order[half:][::-1], order[:half])},
                        writers=sorted(order[:half]), forced="--relay force")
        fab.update(d2h_gbs_all_ranks_copying=[round(x, 2) for x in d2h_all],
                   h2d_gbs_all_ranks_copying=[round(x, 2) for x in h2d_all],
                   d2h_gbs_fast_half_copying=[round(x, 2) for x in d2h_writers],
                   signature_gbs_one_gpu_produces=round(kernel_gbs, 2), plan=plan)
        if plan["policy"] == "relay":
            pairs = {int(s): int(w) for s, w in plan["pairs"].items()}
            tag = f"{os.environ.get('MASTER_PORT', '0')}_{rank}"
            mine, err = None, None
            try:
                if rank in pairs.values():
                    run.receiver = fabric.RelayReceiver(dev, chunk * shape.sig_bytes, tag)
                    mine = ("receiver", run.receiver.export())
                elif rank in pairs:
                    run.sender = fabric.RelaySender(dev, shm, tag)
                    mine = ("sender", run.sender.export())
            except Exception as exc:  # noqa: BLE001
                err = repr(exc)
            everyone = gather_obj((mine, err))
            if not any(e for _, e in everyone):
                try:
                    if run.receiver is not None:
                        partner = [s for s, w in pairs.items() if w == rank][0]
                        run.receiver.attach(everyone[partner][0][1])
                    if run.sender is not None:
                        run.sender.attach(everyone[pairs[rank]][0][1])
                except Exception as exc:  # noqa: BLE001
                    err = repr(exc)
            errors = [e for e in gather_obj(err) if e] + [e for _, e in everyone if e]
            if errors:
                # every rank falls back together: a relay that cannot be set up must not take the headline down
                log(f"[rank {rank}] relay set-up failed somewhere ({errors[0]}); gathering directly")
                run.sender = run.receiver = None
                plan.update(policy="direct", pairs={}, relay_error=errors[0])
            else:
                barrier()
                for _ in range(args.warmup):          # warm the relayed path (peer access, IPC mappings)
                    run.one_step()
                run.drain()
                barrier()

    def timed_region():
        sampler = ClockSampler(device_index)
        if rank == 0:
            sampler.start()
        launches0 = _native.launch_count()
        events: list = []
        t_start = torch.cuda.Event(enable_timing=True); t_stop = torch.cuda.Event(enable_timing=True)
        barrier()
        t_start.record(run.compute)
        for _ in range(args.steps):
            run.one_step(events)
        run.drain()                  # every signature of every step is in host memory before the clock stops
        t_stop.record(run.compute)
        barrier()
        return (t_start, t_stop, events, _native.launch_count() - launches0,
                sampler.stop() if rank == 0 else None)

    start, stop, kernel_events, launches, clocks = timed_region()
    # a run that saw a hardware / thermal slowdown is rejected and taken again, once
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    redo = [bool(clocks and bad & set(clocks.get("reasons", [])))]
    if world > 1:
        dist.broadcast_object_list(redo, src=0)
    remeasured = bool(redo[0])
    if remeasured:
        log(f"[rank {rank}] clocks show {clocks and clocks.get('reasons')}: measuring the timed region again")
        start, stop, kernel_events, launches, clocks = timed_region()
    if clocks is not None:
        clocks["remeasured"] = remeasured
    my_ms = start.elapsed_time(stop)
    elapsed_ms = max_ranks(my_ms)
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in kernel_events)
    kern_rows = sum(r for _, _, r in kernel_events)
    value = rows * world * args.steps / (elapsed_ms * 1e-3)
    kernel_name = hasher.last_kernel
    # the same job without the gather: signatures left in HBM (sum of the kernels' own durations)
    kernel_only_value = rows * world * args.steps / (max_ranks(kern_ms) * 1e-3)

    # every rank: the gathered bytes of its first and last chunk equal a direct recomputation of those chunks
    if world > 1:
        ok = True
        for r0, r1 in (run.chunks[0], run.chunks[-1]):
            fresh = hasher.hash_device(X[r0:r1]).reshape(r1 - r0, shape.sig_bytes).cpu()
            ok = ok and bool(torch.equal(fresh, out_host[r0:r1]))
        checks = gather_obj(ok)
        per_rank_ms = gather_obj(round(my_ms / args.steps, 3))
        total_gbs = shape.sig_bytes * rows * world * args.steps / (elapsed_ms * 1e-3) / 1e9
        per_rank_link = plan.get("per_rank_gbs", {}).get(plan["policy"]) if plan.get("per_rank_gbs") else None
        ceiling = None
        if per_rank_link:
            ceiling = min(per_rank_link, kernel_gbs) * world / shape.sig_bytes * 1e9
        fab.update(gather_check_all_ranks=bool(all(checks)), ms_per_step_per_rank=per_rank_ms,
                   gathered_gbs_all_ranks=round(total_gbs, 1),
                   ceiling_value=ceiling, value_frac_of_ceiling=(value / ceiling) if ceiling else None,
                   ceiling_note="world x min(measured per-rank link rate under the chosen plan, rate at which one "
                                "GPU produces signatures) / signature bytes")

    # ---- e2e: host pinned buffers through the C ABI (H2D + kernel + D2H per step) --------------------------------
    e2e_rows = args.e2e_rows
    e2e_split = None
    if world > 1 and not args.no_fabric:
        # the vectors cross PCIe host -> device here: shards sized to each rank's measured H2D rate
        e2e_split = fabric.weighted_rows(args.e2e_rows * world, fab["h2d_gbs_all_ranks_copying"])
        e2e_rows = e2e_split[rank]
    xh = torch.empty((e2e_rows, shape.dim), dtype=torch.float32, pin_memory=True)
    take = min(e2e_rows, rows)
    xh[:take].copy_(X[:take])
    if take < e2e_rows:
        xh[take:].copy_(X[: e2e_rows - take])
    oh = torch.empty((e2e_rows, shape.sig_bytes), dtype=torch.uint8, pin_memory=True)
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
    e2e_total_rows = sum(e2e_split) if e2e_split else e2e_rows * world
    e2e_value = e2e_total_rows * e2e_steps / (e2e_ms * 1e-3)
    # the e2e path must produce the same bytes as the resident path
    same = bool(torch.equal(oh[:take], out_host[:take]))
    pageable_value = f16_value = None
    if not args.no_e2e:
        # the same call on an ordinary (pageable) numpy array: pinned bounce buffers + parallel memcpy in the library
        pr = min(e2e_rows, 1_000_000)
        xp, op = xh[:pr].numpy().copy(), np.empty((pr, shape.sig_bytes), dtype=np.uint8)
        hasher.hash_into(xp, pr, op, x_on_device=False, out_on_device=False)
        barrier()
        t_pg = time.perf_counter()
        for _ in range(3):
            hasher.hash_into(xp, pr, op, x_on_device=False, out_on_device=False)
        pageable_value = 3 * pr / (time.perf_counter() - t_pg)
        same = same and bool(np.array_equal(op, oh[:pr].numpy()))
        # a float16 numpy array (what embedding stores often hold): raw rows over PCIe, exact cast on the device
        # (lshx_hash_batch_typed); the signatures must equal those of the same values given as float32
        x16 = xp.astype(np.float16)
        want16 = hasher.hash_batch_packed(x16.astype(np.float32)).reshape(pr, shape.sig_bytes)
        got16 = hasher.hash_batch_packed(x16).reshape(pr, shape.sig_bytes)
        t_16 = time.perf_counter()
        for _ in range(3):
            hasher.hash_batch_packed(x16)
        f16_value = 3 * pr / (time.perf_counter() - t_16)
        same = same and bool(np.array_equal(got16, want16))
        del xp, op, x16, want16, got16
    # latency of the per-vector call LSHRS.ingest / query make (reference: one hash_vector per call)
    one = xh[:1].numpy().copy()
    for _ in range(20):
        hasher.hash_vector(one[0])
    t_lat = time.perf_counter()
    for _ in range(200):
        hasher.hash_vector(one[0])
    single_us = (time.perf_counter() - t_lat) / 200 * 1e6

    roofline = roofline_block(shape, kernel_name, kern_ms, kern_rows, len(kernel_events), my_ms, peaks, peak_src)

    # ---- CPU baseline + parity on a bounded sample (rank 0, N = 1 only) ---------------------
    cpu_baseline, parity, cpu = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = CpuReference()
        cpu_baseline, parity = cpu_leg(cpu, shape, X, out_host, rows, args.cpu_sample, dev, torch)
        parity["e2e_equals_resident"] = same

    # ---- BASELINE configs 3 and 5, the API calls, rerank (config 4) -------------------------------------------------
    configs = None
    rerank = None
    api = None
    if rank == 0 and world == 1 and not args.no_e2e and shape.name == "hash768":
        try:
            api = run_api(args, dev, shape)
        except Exception as exc:  # noqa: BLE001 -- a side measurement must not take the headline down
            api = {"error": repr(exc)}
    relay_pair = (run.sender, run.receiver)
    del run, X
    torch.cuda.empty_cache()
    if world == 1 and shape.name == "hash768" and not args.no_configs:
        configs = {}
        for name in ("hash1536", "hash128"):
            try:
                configs[name] = side_config(torch, dev, device_index, Shape(name), args, peaks, peak_src,
                                            cpu if not args.no_cpu else None)
            except Exception as exc:  # noqa: BLE001
                configs[name] = {"error": repr(exc)}
            torch.cuda.empty_cache()
    if not args.no_rerank:
        rerank = run_rerank(args, dev, rank, world, peaks, peak_src, barrier, max_ranks, shape,
                            cpu if (rank == 0 and world == 1 and not args.no_cpu) else None)

    if rank == 0:
        per_gpu_rows = e2e_split if e2e_split else [e2e_rows] * world
        line = {
            "metric": shape.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"tcgen05": "f16x3", "tcgen05_tf32bf16": "tf32+bf16x2", "tcgen05_3xtf32": "tf32x3"}.get(kernel_name, "f32"),
            "data": "synthetic", "config": workload_config(args, shape), "impl": "b200", "kernel": kernel_name,
            "env_overrides": _native.env_overrides(),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(per_gpu_rows) * shape.dim * 4 / world),
                    "d2h_bytes_per_step": int(sum(per_gpu_rows) * shape.sig_bytes / world),
                    "bytes_per_step_are": "per GPU (mean over ranks)", "rows_per_step_per_gpu": per_gpu_rows,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": "lshx_hash_batch(host pinned X -> host pinned signatures)",
                    "single_vector_call_us": single_us,
                    "pageable_numpy_input_value": pageable_value, "float16_numpy_input_value": f16_value},
            "kernel_only": {"value": kernel_only_value, "unit": UNIT,
                            "note": "signatures left in HBM (no gather); value above includes the overlapped gather "
                                    "of every signature into pinned host memory"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "fabric": fab,
            "cpu_baseline": cpu_baseline, "parity": parity, "configs": configs, "rerank": rerank, "api": api,
        }
        emit(line)
    if relay_pair[0] is not None:
        relay_pair[0].close()          # the sender lets go of the partner's device slots first
    if world > 1:
        barrier()
    if relay_pair[1] is not None:
        relay_pair[1].close()
    if shm is not None:
        barrier()
        shm.close()
    if rank == 0 and not args.no_fabric and world < visible:
        try:
            os.unlink(f"/dev/shm/lshx_probe_{os.environ.get('MASTER_PORT', '0')}.json")
        except OSError:
            pass
    if world > 1:
        dist.destroy_process_group()


def cpu_leg(cpu: CpuReference, shape: Shape, X, out_host, rows: int, sample: int, dev, torch):
    """The reference's CPU path on `sample` rows of the shard, and the GPU's bytes for the same rows against it."""
    h, projs = cpu.hasher(shape)
    idx = (torch.arange(sample, device=dev, dtype=torch.int64) * (rows - 1)) // max(1, sample - 1)   # evenly spread, exact
    Xs = X[idx].cpu().numpy()
    t0 = time.perf_counter()
    ref = cpu.packed(cpu.hash_batch(h, projs, Xs), shape)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    cpu.oracle.hash_batch_vectorized(projs, Xs)
    dtv = time.perf_counter() - t1
    got = out_host[idx.cpu()].numpy().reshape(sample, shape.bands, -1)
    parity = cpu.oracle.compare_packed(got, ref, cpu.oracle.projection_margins(projs, Xs), 1e-5)
    parity["sample_rows"] = sample
    parity["against"] = cpu.kind
    base = {
        "value": sample / dt, "unit": UNIT, "cores": 1, "kind": cpu.kind, "source": cpu.source,
        "sample": f"{sample} rows of the same shard through LSHHasher.hash_batch -- per-vector, per-band sgemv, a "
                  f"single Python thread by construction (reference lshrs/hash/lsh.py:169), {dt:.1f} s",
        "host_cores": os.cpu_count(), "numpy": np.__version__,
        "vectorized_numpy_not_reference": {"value": sample / dtv, "unit": UNIT, "cores": len(os.sched_getaffinity(0))},
    }
    return base, parity


def side_config(torch, dev, device_index: int, shape: Shape, args, peaks, peak_src, cpu) -> dict:
    """BASELINE configs 3 (1536 -> 512 bits) and 5 (SIFT-like 128 -> 64 bits) on one GPU: the same resident-shard
    step as the headline (kernel per chunk + gather of the signatures), 5 timed steps, roofline recomputed from
    4*dim + signature bytes / 2*dim*num_perm, and a parity sample against the CPU reference."""
    from lshrs_b200 import LSHHasher

    hasher = LSHHasher(shape.bands, shape.rows_per_band, shape.dim, seed=SEED, device=device_index)
    rows, chunk = shape.rows, shape.chunk
    X = make_shard(torch, dev, shape, rows, 2000)
    out_host = torch.empty((rows, shape.sig_bytes), dtype=torch.uint8, pin_memory=True)
    run = ResidentRun(torch, dev, hasher, X, shape, chunk, out_host)
    for _ in range(3):
        run.one_step()
    run.drain()
    torch.cuda.synchronize()
    steps = 5
    events: list = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(run.compute)
    for _ in range(steps):
        run.one_step(events)
    run.drain()
    t1.record(run.compute)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in events)
    kern_rows = sum(r for _, _, r in events)
    out = {
        "metric": shape.metric, "value": rows * steps / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": 3,
        "ms_per_step": ms / steps, "rows_per_gpu": rows, "chunk_rows": chunk, "kernel": hasher.last_kernel,
        "kernel_only": {"value": kern_rows / (kern_ms * 1e-3), "unit": UNIT},
        "data": "synthetic " + ("Gaussian" if shape.dist == "gauss" else "SIFT-like (non-negative integer-valued floats in [0, 255])"),
        "roofline": roofline_block(shape, hasher.last_kernel, kern_ms, kern_rows, len(events), ms, peaks, peak_src),
        "note": "value includes the gather of every signature into pinned host memory "
                f"({shape.sig_bytes} B per vector over one PCIe link); kernel_only leaves them in HBM",
    }
    if cpu is not None:
        out["cpu_baseline"], out["parity"] = cpu_leg(cpu, shape, X, out_host, rows, 16_384, dev, torch)
    del run, X, out_host
    hasher.close()
    return out


def run_api(args, dev, shape: Shape) -> dict:
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
    centers = rng.standard_normal((n_index // 8, shape.dim)).astype(np.float32)
    Xh = (np.repeat(centers, 8, axis=0) + 0.15 * rng.standard_normal((n_index, shape.dim))).astype(np.float32)
    extra = (Xh[:n_single] + 0.05 * rng.standard_normal((n_single, shape.dim))).astype(np.float32)
    allvec = np.concatenate([Xh, extra])          # id -> vector, what vector_fetch_fn serves
    corpus_dev = torch.from_numpy(allvec).to(dev)
    lsh = LSHRS(dim=shape.dim, num_perm=shape.num_perm, storage=InMemoryStorage(), vector_fetch_fn=lambda ids: allvec[ids],
                device_index=True, device=dev.index)
    ids = list(range(n_index))
    t0 = time.perf_counter()
    lsh.index(ids, Xh)
    index_vps = n_index / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for i in range(n_single):
        lsh.ingest(n_index + i, extra[i])
    lsh.flush()
    ingest_cps = n_single / (time.perf_counter() - t0)
    Q = (Xh[rng.integers(0, n_index, n_batch)] + 0.05 * rng.standard_normal((n_batch, shape.dim))).astype(np.float32)
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
    # the same queries with the candidates generated on the GPU (device mirror of the bucket store):
    # lists must equal the storage path's; 8192 queries per call like BASELINE config 4
    nbig = 8192
    Qbig = (Xh[rng.integers(0, n_index, nbig)] + 0.05 * rng.standard_normal((nbig, shape.dim))).astype(np.float32)
    kw = dict(top_k=10, top_p=0.2, corpus=corpus_dev, device_index=True)
    dev_batch = lsh.query_batch(Q, **kw)
    same_dev = dev_batch == batch
    lsh.query_batch(Qbig, **kw)

    def rate(fn, reps=3):
        best = 0.0
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            best = max(best, nbig / (time.perf_counter() - t))
        return best

    dev_lists_qps = rate(lambda: lsh.query_batch(Qbig, **kw))
    dev_arrays_qps = rate(lambda: lsh.query_batch(Qbig, as_arrays=True, **kw))
    dev_topk_qps = rate(lambda: lsh.query_batch(Qbig, top_k=10, device_index=True, as_arrays=True))
    same_topk = lsh.query_batch(Q, top_k=10, device_index=True) == lsh.query_batch(Q, top_k=10)
    mean_cands = float(np.mean([len(x) for x in lsh.query_batch(Q, top_k=None, device_index=True)]))
    # the same calls with the bucket store itself in HBM (DeviceBucketStorage): index() hands over packed signatures,
    # single queries join on the device; results must equal the dict store's
    from lshrs_b200 import DeviceBucketStorage

    dst = LSHRS(dim=shape.dim, num_perm=shape.num_perm, storage=DeviceBucketStorage(), vector_fetch_fn=lambda ids: allvec[ids],
                device=dev.index)
    dst.index(ids, Xh)                      # first call sizes the segments and staging buffers: off the clock
    dst.clear()
    t0 = time.perf_counter()
    dst.index(ids, Xh)
    d_index_vps = n_index / (time.perf_counter() - t0)
    n_more = 1_000_000                       # a second, larger batch: 1 M x 768 (3 GB of host vectors) in one call
    rep = -(-n_more // n_index)
    Xbig = np.tile(Xh, (rep, 1))[:n_more]
    big = LSHRS(dim=shape.dim, num_perm=shape.num_perm, storage=DeviceBucketStorage(), device=dev.index)
    big.index(np.arange(n_more), Xbig)
    big.clear()
    t0 = time.perf_counter()
    big.index(np.arange(n_more), Xbig)
    big.query_batch(Q[:8], top_k=10)         # the first query sorts the segments: on the clock
    d_index_big_vps = n_more / (time.perf_counter() - t0)
    # the same batch already resident in HBM (a CUDA tensor): hash + append without PCIe
    big.clear()
    xd = torch.from_numpy(Xbig).to(dev)
    idd = torch.arange(n_more, dtype=torch.int64, device=dev)
    big.index(idd[:4096], xd[:4096])
    big.clear()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    big.index(idd, xd)
    big.query_batch(Q[:8], top_k=10)
    d_index_res_vps = n_more / (time.perf_counter() - t0)
    big._storage.index.close()
    big._hasher.close()
    del Xbig, big, xd, idd
    t0 = time.perf_counter()
    for i in range(n_single):
        dst.ingest(n_index + i, extra[i])
    dst.flush()
    d_ingest_cps = n_single / (time.perf_counter() - t0)
    dst.get_top_k(Q[-1], topk=10)
    dst.get_above_p(Q[-1], p=0.2)
    t0 = time.perf_counter()
    d_got_k = [dst.get_top_k(Q[i], topk=10) for i in range(n_single)]
    d_topk_qps = n_single / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    d_got_p = [dst.get_above_p(Q[i], p=0.2) for i in range(n_single)]
    d_above_qps = n_single / (time.perf_counter() - t0)
    dst.set_corpus(corpus_dev)               # candidate vectors gathered from HBM by id instead of vector_fetch_fn
    dst.get_above_p(Q[-1], p=0.2)
    t0 = time.perf_counter()
    d_got_pc = [dst.get_above_p(Q[i], p=0.2) for i in range(n_single)]
    d_above_corpus_qps = n_single / (time.perf_counter() - t0)
    dst.set_corpus(None)
    same_store = (d_got_k == got_k and [[i for i, _ in r] for r in d_got_pc] == [[i for i, _ in r] for r in got_p]
                  and [[i for i, _ in r] for r in d_got_p] == [[i for i, _ in r] for r in got_p]
                  and dst.query_batch(Q, top_k=10, top_p=0.2, corpus=corpus_dev) == dev_batch)
    dst.query_batch(Qbig, as_arrays=True, **kw)
    d_arrays_qps = rate(lambda: dst.query_batch(Qbig, as_arrays=True, **kw))
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
        "query_batch_device_index": {
            "value": dev_arrays_qps, "unit": "queries/s", "queries_per_call": nbig, "mean_candidates": mean_cands,
            "what": "query_batch(top_k=10, top_p=0.2, corpus=<HBM>, device_index=True, as_arrays=True): hash, join on "
                    "the device mirror, rerank and id gather without leaving the GPU; host vectors in, numpy out",
            "python_lists_value": dev_lists_qps, "top_k_only_value": dev_topk_qps,
            "equals_storage_path": bool(same_dev and same_topk)},
        "device_storage": {
            "storage": "DeviceBucketStorage (the bucket store itself in HBM; LSHRS(storage=DeviceBucketStorage()))",
            "index": {"value": d_index_vps, "unit": "vectors/s", "rows": n_index},
            "index_1m_rows": {"value": d_index_big_vps, "unit": "vectors/s", "rows": n_more,
                              "note": "host float32 vectors in: PCIe-bound like e2e; the first query's sort of the "
                                      "segments is inside the timed region"},
            "index_1m_rows_resident": {"value": d_index_res_vps, "unit": "vectors/s", "rows": n_more,
                                       "note": "index(ids, <CUDA tensor>): hash + append + sort in HBM"},
            "ingest": {"value": d_ingest_cps, "unit": "calls/s", "calls": n_single},
            "get_top_k": {"value": d_topk_qps, "unit": "queries/s", "calls": n_single},
            "get_above_p": {"value": d_above_qps, "unit": "queries/s", "calls": n_single},
            "get_above_p_resident_corpus": {"value": d_above_corpus_qps, "unit": "queries/s", "calls": n_single,
                                            "note": "LSHRS.set_corpus(<CUDA tensor>): no vector_fetch_fn round trip"},
            "query_batch": {"value": d_arrays_qps, "unit": "queries/s", "queries_per_call": nbig},
            "equals_dict_store": bool(same_store)},
        "reference_survey_values": {"index": 2.2e3, "get_top_k": 3.9e3, "get_above_p": 2.7e3,
                                    "note": "SURVEY section 8a, reference with its MockStorage in the build container"},
    }


def run_rerank(args, dev, rank, world, peaks, peak_src, barrier, max_ranks, shape: Shape, cpu) -> dict:
    """BASELINE config 4: 8192 queries x 2000 candidates x 768 from a 1M-row corpus resident in HBM."""
    import torch

    from lshrs_b200 import _native
    from lshrs_b200.utils.similarity import _get_reranker

    N, nq, nc = args.corpus, args.queries, 2000
    gen = torch.Generator(device=dev).manual_seed(1)
    corpus = torch.empty((N, shape.dim), dtype=torch.float32, device=dev)
    corpus.normal_(generator=gen)
    Q = torch.empty((nq, shape.dim), dtype=torch.float32, device=dev).normal_(generator=gen)
    # 2000 DISTINCT corpus rows per query: random start, random stride < N / nc
    first = torch.randint(0, N, (nq, 1), generator=gen, device=dev, dtype=torch.int64)
    stride = torch.randint(1, max(2, N // nc), (nq, 1), generator=gen, device=dev, dtype=torch.int64)
    ids = ((first + stride * torch.arange(nc, device=dev, dtype=torch.int64)[None, :]) % N).contiguous()
    offs = torch.arange(nq + 1, device=dev, dtype=torch.int64) * nc
    rer = _get_reranker(shape.dim, dev.index)
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
        bytes_per_q = nc * (4 * shape.dim + 8) + 4 * shape.dim + 8 * limit
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
                         "frac": gbs / peaks["hbm_gbs"],
                         "traffic": (NCU_TRAFFIC["rerank"][1] / NCU_TRAFFIC["rerank"][2] * nq
                                     if (nc, shape.dim) == (2000, 768) else None),
                         "traffic_source": f"{NCU_TRAFFIC['rerank'][0]}: DRAM bytes of one ncu --set full capture, "
                                           "scaled by queries", "peak_source": peak_src,
                         "bytes_per_query": bytes_per_q},
            "e2e": {"value": nq * world / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(Qh.nbytes + idh.nbytes + offh.nbytes),
                    "d2h_bytes_per_step": int(ph.nbytes + sh.nbytes + ch.nbytes + zh.nbytes)},
            "device_equals_e2e": bool(np.array_equal(ph, pos.cpu().numpy())),
        }
        if cpu is not None:
            nref = 256
            Ch = corpus.cpu().numpy()
            t1 = time.perf_counter()
            bad = 0
            for i in range(nref):
                ref = cpu.top_k_cosine(Qh[i], Ch[idh[i * nc:(i + 1) * nc]], limit)
                got_pos = ph[i, :limit]
                ref_scores = np.array([sc for _, sc in ref])
                if not np.allclose(sh[i, :limit], ref_scores, atol=1e-5, rtol=0):
                    bad += 1
                elif set(got_pos.tolist()) != {pp for pp, _ in ref}:
                    bad += 1
            dt = time.perf_counter() - t1
            entry["cpu_baseline"] = {"value": nref / dt, "unit": "queries/s", "cores": 1, "kind": cpu.kind,
                                     "source": cpu.source,
                                     "sample": f"{nref} queries through top_k_cosine (reference "
                                               "lshrs/utils/similarity.py:93-183)"}
            entry["parity"] = {"queries_checked": nref, "mismatching_queries": bad, "score_tol": 1e-5,
                               "against": cpu.kind}
            del Ch
        out[tag] = entry
    out["config"] = {"workload": f"top_k_cosine rerank: {nq} queries x {nc} candidates, dim={shape.dim}, corpus {N} rows "
                                 "resident in HBM, CSR int64 candidate ids", "l2": "6.1 MB of gathers per query, "
                                 f"{nq * nc * shape.dim * 4 / 1e9:.1f} GB per step (larger than L2)"}
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
    import faulthandler

    faulthandler.enable()      # a crash in native code names the Python frame it came from
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
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3 and 5 (the `configs` block)")
    ap.add_argument("--no-fabric", action="store_true",
                    help="multi-GPU: identity rank -> GPU map, equal e2e shards, no link probe, no relay")
    ap.add_argument("--relay", choices=("auto", "off", "force"), default="auto",
                    help="multi-GPU: let ranks on slow host links relay their signatures over NVLink (auto) or not")
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    if args.impl == "reference":
        args.workload = "hash768"
    shape = Shape(args.workload)
    if args.rows is None:
        args.rows = shape.rows
    if args.chunk is None:
        args.chunk = shape.chunk
    args.chunk = min(args.chunk, args.rows)
    args.e2e_rows = min(args.e2e_rows, args.rows)
    if args.workload != "hash768":
        args.no_rerank = True
    args.warmup = max(3, args.warmup) if args.impl == "b200" else max(0, args.warmup)
    if args.impl == "reference":
        run_reference(args, shape)
    else:
        run_b200(args, shape)


if __name__ == "__main__":
    main()
