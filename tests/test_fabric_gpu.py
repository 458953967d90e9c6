"""The relay pair on two real GPUs (two processes, CUDA IPC slots, interprocess events, shared pinned host buffer).

Needs two GPUs: skipped on a one-GPU box (run it with ``gpurun --gpus 2``); the multi-GPU bench exercises the same
classes with ``--relay force``.
"""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]
CHUNKS, CHUNK_BYTES = 64, 1 << 20


def _pattern(torch, k, dev):
    return ((torch.arange(CHUNK_BYTES, device=dev, dtype=torch.int64) * 2654435761 + k * 97) >> 7).to(torch.uint8)


def _proc(rank, tag, q01, q10, result):
    sys.path.insert(0, str(REPO))
    import torch

    from lshrs_b200 import fabric as fb

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    if rank == 0:       # sender (the rank on the slow link)
        host = fb.SharedHostBuffer(f"{tag}_host", CHUNKS * CHUNK_BYTES, create=True)
        host.array[:] = 0
        host.pin()
        snd = fb.RelaySender(dev, host, tag)
        q01.put(snd.export())
        snd.attach(q10.get(timeout=120))
        compute = torch.cuda.Stream(dev)
        bufs = [torch.empty(CHUNK_BYTES, dtype=torch.uint8, device=dev) for _ in range(2)]
        freed = [None, None]
        for k in range(CHUNKS):
            b = k & 1
            if freed[b] is not None:
                compute.wait_event(freed[b])
            with torch.cuda.stream(compute):
                bufs[b].copy_(_pattern(torch, k, dev))
            done = torch.cuda.Event()
            done.record(compute)
            freed[b] = snd.send(k, bufs[b], done)
        snd.stream.synchronize()
        assert q10.get(timeout=120) == "drained"
        want = np.concatenate([_pattern(torch, k, dev).cpu().numpy() for k in range(CHUNKS)])
        result.put(bool(np.array_equal(host.array, want)))
        snd.close()
        q01.put("closed")
        host.close()
    else:               # receiver (the rank on the fast link)
        rcv = fb.RelayReceiver(dev, CHUNK_BYTES, tag)
        q10.put(rcv.export())
        rcv.attach(q01.get(timeout=120))
        for k in range(CHUNKS):
            rcv.drain(k, CHUNK_BYTES, k * CHUNK_BYTES)
        rcv.stream.synchronize()
        q10.put("drained")
        assert q01.get(timeout=120) == "closed"
        rcv.close()


def test_relay_pair_moves_every_chunk_into_the_senders_host_buffer():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    tag = f"lshx_gtest_{os.getpid()}"
    q01, q10, result = ctx.Queue(), ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_proc, args=(r, tag, q01, q10, result)) for r in range(2)]
    [p.start() for p in procs]
    ok = result.get(timeout=300)
    [p.join(120) for p in procs]
    assert ok and all(p.exitcode == 0 for p in procs)
