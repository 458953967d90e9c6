#!/usr/bin/env python
"""Host-fabric probe for the multi-GPU hash bench (profiling script, not product code).

One process, one thread + one CUDA context per GPU.  Answers the questions the 8-GPU scaling curve raised
(D2H of the signatures into pinned host memory, not a kernel, limits N >= 4):

  A  D2H GB/s per GPU for subsets of GPUs copying at once (all 8, each half, interleaved halves, pairs)
  B  the same with the copy granularity the bench uses (25 MB chunks) and with write-combined pinned memory
  C  a shared queue of 25 MB copy jobs pulled by every GPU (what a dynamic balance could reach in aggregate)
  D  D2H of the slow group while its signatures are RELAYED over NVLink to a fast-group GPU (P2P copy, then
     that GPU's D2H) -- is the fabric limit per link / per group, or one shared pool?
  E  H2D subsets (the e2e side)
  F  D2H while the GPU's SMs stream HBM (does a running kernel slow the copy engine's PCIe writes?)

    python tools/fabric_probe.py > gpurun_out/fabric_probe.json
"""
import ctypes
import json
import subprocess
import sys
import threading
import time

import torch

MB = 1 << 20
CHUNK = 25 * MB
TOTAL = 40 * CHUNK          # bytes each GPU moves per measurement (1 GB)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as exc:  # noqa: BLE001
        return repr(exc)


class Barrier2:
    def __init__(self, n):
        self.b = threading.Barrier(n)

    def wait(self):
        self.b.wait()


def cudart():
    return torch.cuda.cudart()


def alloc_wc(nbytes):
    """cudaHostAlloc(..., cudaHostAllocWriteCombined | Portable) through libcudart."""
    lib = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = lib.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc WC failed rc={rc}")
    return lib, p


def run_group(active, fn):
    """Run fn(i, barrier) on one thread per GPU index in `active`; returns {i: seconds}."""
    out = {}
    bar = Barrier2(len(active))

    def work(i):
        torch.cuda.set_device(i)
        out[i] = fn(i, bar)

    th = [threading.Thread(target=work, args=(i,)) for i in active]
    [t.start() for t in th]
    [t.join() for t in th]
    return out


def main():
    n = torch.cuda.device_count()
    res = {"gpus": n, "chunk_mb": CHUNK // MB, "bytes_per_gpu": TOTAL,
           "topo": sh("nvidia-smi topo -m | head -12"), "lscpu": sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"),
           "pcie": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader")}
    dev = [torch.empty(TOTAL, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
    host = [torch.empty(TOTAL, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    streams = [torch.cuda.Stream(i) for i in range(n)]
    for i in range(n):
        dev[i].fill_(i + 1)
    torch.cuda.synchronize()

    def copy_fn(direction, chunk, weights=None):
        def fn(i, bar):
            s = streams[i]
            nbytes = TOTAL if weights is None else int(TOTAL * weights[i]) // chunk * chunk
            with torch.cuda.stream(s):
                for _ in range(2):   # warm
                    (host[i][:chunk].copy_(dev[i][:chunk], non_blocking=True) if direction == "d2h"
                     else dev[i][:chunk].copy_(host[i][:chunk], non_blocking=True))
                s.synchronize()
                bar.wait()
                t0 = time.perf_counter()
                for o in range(0, nbytes, chunk):
                    if direction == "d2h":
                        host[i][o:o + chunk].copy_(dev[i][o:o + chunk], non_blocking=True)
                    else:
                        dev[i][o:o + chunk].copy_(host[i][o:o + chunk], non_blocking=True)
                s.synchronize()
                return time.perf_counter() - t0, nbytes
        return fn

    def rates(out):
        per = {str(i): round(b / t / 1e9, 2) for i, (t, b) in sorted(out.items())}
        wall = max(t for t, _ in out.values())
        return {"per_gpu_gbs": per, "sum_of_rates_gbs": round(sum(per.values()), 1),
                "aggregate_gbs": round(sum(b for _, b in out.values()) / wall / 1e9, 1), "wall_ms": round(wall * 1e3, 1)}

    half = n // 2
    subsets = {"all": list(range(n))}
    if n >= 8:
        subsets.update({"0-3": [0, 1, 2, 3], "4-7": [4, 5, 6, 7], "0,1,4,5": [0, 1, 4, 5], "0,2,4,6": [0, 2, 4, 6],
                        "0,4": [0, 4], "0,1": [0, 1], "4,5": [4, 5], "0": [0], "4": [4]})
    elif n >= 2:
        subsets.update({"0": [0], "0,1": [0, 1]})
    # A / B: D2H subsets, 1 GB single copy vs 25 MB chunks
    res["d2h_chunk25"] = {k: rates(run_group(v, copy_fn("d2h", CHUNK))) for k, v in subsets.items()}
    res["d2h_single_copy"] = {k: rates(run_group(v, copy_fn("d2h", TOTAL))) for k, v in subsets.items()
                              if k in ("all", "0", "0-3", "4-7")}
    res["h2d_chunk25"] = {k: rates(run_group(v, copy_fn("h2d", CHUNK))) for k, v in subsets.items()
                          if k in ("all", "0", "0-3", "4-7", "0,1,4,5", "0,4")}
    # both directions at once on all GPUs (the e2e path overlaps them)
    if n >= 2:
        def both(i, bar):
            s2 = torch.cuda.Stream(i)
            s = streams[i]
            bar.wait()
            t0 = time.perf_counter()
            for o in range(0, TOTAL, CHUNK):
                with torch.cuda.stream(s):
                    host[i][o:o + CHUNK].copy_(dev[i][o:o + CHUNK], non_blocking=True)
                with torch.cuda.stream(s2):
                    dev[i][o:o + CHUNK].copy_(host[i][o:o + CHUNK], non_blocking=True) if False else None
            s.synchronize(); s2.synchronize()
            return time.perf_counter() - t0, TOTAL
        # (kept simple: D2H only here; the H2D+D2H mix is what bench e2e measures directly)

    # balanced static weights: what equal finish times would need, from the all-GPU rates
    allr = res["d2h_chunk25"]["all"]["per_gpu_gbs"]
    mean = sum(allr.values()) / len(allr)
    weights = {int(k): v / mean for k, v in allr.items()}
    res["d2h_weighted_static"] = {"weights": {str(k): round(v, 3) for k, v in weights.items()},
                                  **rates(run_group(list(range(n)), copy_fn("d2h", CHUNK, weights)))}

    # C: shared queue of 25 MB jobs (any GPU copies from ITS OWN device buffer -- models outputs that can be
    # produced anywhere, i.e. the e2e / host-fed case)
    jobs = n * (TOTAL // CHUNK)
    counter = {"next": 0}
    lock = threading.Lock()
    taken = {i: 0 for i in range(n)}

    def queue_fn(i, bar):
        s = streams[i]
        bar.wait()
        t0 = time.perf_counter()
        inflight = []
        while True:
            with lock:
                j = counter["next"]
                counter["next"] += 1
            if j >= jobs:
                break
            o = (taken[i] % (TOTAL // CHUNK)) * CHUNK
            with torch.cuda.stream(s):
                host[i][o:o + CHUNK].copy_(dev[i][o:o + CHUNK], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            inflight.append(ev)
            taken[i] += 1
            if len(inflight) >= 2:          # two copies in flight per GPU, then pull the next job
                inflight.pop(0).synchronize()
        s.synchronize()
        return time.perf_counter() - t0, taken[i] * CHUNK

    out = run_group(list(range(n)), queue_fn)
    res["d2h_dynamic_queue"] = {**rates(out), "jobs_taken": {str(k): v for k, v in taken.items()}}

    # write-combined pinned memory
    try:
        wc = []
        for i in range(n):
            lib, p = alloc_wc(TOTAL)
            wc.append((lib, p))
        cudart_lib = wc[0][0]

        def wc_fn(i, bar):
            s = streams[i]
            bar.wait()
            t0 = time.perf_counter()
            for o in range(0, TOTAL, CHUNK):
                cudart_lib.cudaMemcpyAsync(ctypes.c_void_p(wc[i][1].value + o), ctypes.c_void_p(dev[i].data_ptr() + o),
                                           ctypes.c_size_t(CHUNK), ctypes.c_int(2), ctypes.c_void_p(s.cuda_stream))
            s.synchronize()
            return time.perf_counter() - t0, TOTAL
        res["d2h_write_combined"] = {k: rates(run_group(v, wc_fn)) for k, v in subsets.items() if k in ("all", "0")}
        # cost of READING the write-combined buffer on the host (what a consumer of the signatures pays)
        import numpy as np
        buf = (ctypes.c_uint8 * (256 * MB)).from_address(wc[0][1].value)
        t0 = time.perf_counter()
        s_ = int(np.frombuffer(buf, dtype=np.uint64).sum())
        res["d2h_write_combined"]["host_read_gbs_one_thread"] = round(256 * MB / (time.perf_counter() - t0) / 1e9, 2)
        t0 = time.perf_counter()
        s_ += int(host[0][:256 * MB].numpy().view(np.uint64).sum())
        res["d2h_write_combined"]["host_read_gbs_one_thread_normal_pinned"] = round(256 * MB / (time.perf_counter() - t0) / 1e9, 2)
        for lib, p in wc:
            lib.cudaFreeHost(p)
    except Exception as exc:  # noqa: BLE001
        res["d2h_write_combined"] = {"error": repr(exc)}

    # D: relay -- GPUs 0..3 send a fraction f of their bytes to partner 4..7 over NVLink (P2P copy into the
    # partner's buffer), the partner copies them to the host besides its own.
    if n >= 8:
        relay = {}
        for frac in (0.0, 0.15, 0.25, 0.35):
            nrel = int(round(frac * (TOTAL // CHUNK)))
            peer_buf = {i + 4: torch.empty(max(1, nrel) * CHUNK, dtype=torch.uint8, device=f"cuda:{i + 4}") for i in range(4)}
            host_rel = {i + 4: torch.empty(max(1, nrel) * CHUNK, dtype=torch.uint8, pin_memory=True) for i in range(4)}
            ready = {i: [threading.Event() for _ in range(max(1, nrel))] for i in range(4)}

            def relay_fn(i, bar, nrel=nrel, peer_buf=peer_buf, host_rel=host_rel, ready=ready):
                s = streams[i]
                bar.wait()
                t0 = time.perf_counter()
                if i < 4:
                    p2p = torch.cuda.Stream(i)
                    # relayed chunks first (so the partner can start), own chunks on the D2H stream meanwhile
                    for j in range(nrel):
                        with torch.cuda.stream(p2p):
                            peer_buf[i + 4][j * CHUNK:(j + 1) * CHUNK].copy_(dev[i][j * CHUNK:(j + 1) * CHUNK], non_blocking=True)
                    for o in range(nrel * CHUNK, TOTAL, CHUNK):
                        with torch.cuda.stream(s):
                            host[i][o:o + CHUNK].copy_(dev[i][o:o + CHUNK], non_blocking=True)
                    p2p.synchronize()
                    for j in range(nrel):
                        ready[i][j].set()
                    s.synchronize()
                    return time.perf_counter() - t0, TOTAL - nrel * CHUNK
                src = i - 4
                rs = torch.cuda.Stream(i)
                for o in range(0, TOTAL, CHUNK):
                    with torch.cuda.stream(s):
                        host[i][o:o + CHUNK].copy_(dev[i][o:o + CHUNK], non_blocking=True)
                for j in range(nrel):
                    ready[src][j].wait()
                    with torch.cuda.stream(rs):
                        host_rel[i][j * CHUNK:(j + 1) * CHUNK].copy_(peer_buf[i][j * CHUNK:(j + 1) * CHUNK], non_blocking=True)
                s.synchronize(); rs.synchronize()
                return time.perf_counter() - t0, TOTAL + nrel * CHUNK
            relay[f"{frac:.2f}"] = rates(run_group(list(range(8)), relay_fn))
        res["d2h_relay_0to3_via_4to7"] = relay

    # F: D2H while the SMs stream HBM (a memset-like kernel loop on another stream)
    def busy_fn(i, bar):
        s = streams[i]
        ks = torch.cuda.Stream(i)
        big = torch.empty(4 << 30, dtype=torch.uint8, device=f"cuda:{i}")
        stop = False
        bar.wait()
        with torch.cuda.stream(ks):
            for _ in range(60):
                big.add_(1)
        t0 = time.perf_counter()
        for o in range(0, TOTAL, CHUNK):
            with torch.cuda.stream(s):
                host[i][o:o + CHUNK].copy_(dev[i][o:o + CHUNK], non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
        ks.synchronize()
        del big
        return dt, TOTAL
    res["d2h_with_busy_sms"] = {k: rates(run_group(v, busy_fn)) for k, v in subsets.items() if k in ("all", "0")}

    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
