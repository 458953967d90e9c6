"""Host-fabric planning and the relay protocol's host side (lshrs_b200/fabric.py), without a GPU.

The planning functions are pure; the relay's cross-process ordering (shared-memory handshake counters, shared host
buffer) is exercised by two real processes that move chunks through a two-slot buffer with CPU copies standing in
for the CUDA copies -- what must hold is the protocol: no slot is overwritten before it was drained, nothing is
drained before it was sent, and neither side deadlocks.  The CUDA half (IPC slots, interprocess events) runs in the
multi-GPU bench and in tests/test_fabric_gpu.py.
"""

from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import pytest

from lshrs_b200 import fabric

REPO = Path(__file__).resolve().parents[1]
PROBE = REPO / "profiles" / "r2_fabric_probe_8gpu.json"


def test_choose_devices_takes_the_fast_links():
    probe = {"devices": list(range(8)), "d2h_gbs": [11.9, 11.9, 12.0, 12.0, 18.7, 18.8, 18.8, 18.8],
             "h2d_gbs": [23.6, 23.4, 23.5, 23.4, 36.1, 36.0, 35.9, 36.0]}
    assert fabric.choose_devices(probe, 4, 8) == [4, 5, 6, 7]
    assert fabric.choose_devices(probe, 2, 8) == [5, 6] or set(fabric.choose_devices(probe, 2, 8)) <= {4, 5, 6, 7}
    assert fabric.choose_devices(probe, 1, 8)[0] in (4, 5, 6, 7)
    assert fabric.choose_devices(probe, 8, 8) == list(range(8))       # everything is needed: identity
    assert fabric.choose_devices(None, 4, 8) == [0, 1, 2, 3]          # no probe: identity
    assert fabric.choose_devices({"devices": [0], "d2h_gbs": []}, 2, 8) == [0, 1]


def test_plan_relay_on_the_measured_box():
    """The numbers of profiles/r2_fabric_probe_8gpu.json: all eight copying -> 12 / 18.8 GB/s, GPUs 4-7 alone ->
    44 GB/s; one GPU produces 29 GB/s of signatures.  Direct: min = 12 GB/s per rank; relay: 44 / 2 = 22."""
    d2h_all = [11.95, 11.94, 11.95, 11.97, 18.71, 18.8, 18.84, 18.82]
    writers_only = [0, 0, 0, 0, 44.08, 44.14, 44.08, 44.19]
    plan = fabric.plan_relay(d2h_all, writers_only, kernel_gbs=29.0)
    assert plan["policy"] == "relay" and plan["writers"] == [4, 5, 6, 7]
    assert sorted(plan["pairs"]) == [0, 1, 2, 3] and sorted(plan["pairs"].values()) == [4, 5, 6, 7]
    assert plan["per_rank_gbs"] == {"direct": 11.94, "relay": 22.04}
    # four GPUs on the fast links keep up with the kernel: nothing to gain
    assert fabric.plan_relay([44.1, 44.0, 44.2, 44.1], [56.0, 56.0, 0, 0], 29.0)["policy"] == "direct"
    # a relay that is not faster than writing directly is not taken
    assert fabric.plan_relay([20.0, 20.0, 21.0, 21.0], [0, 0, 30.0, 30.0], 29.0)["policy"] == "direct"
    assert fabric.plan_relay([12.0], None, 29.0)["policy"] == "direct"
    assert fabric.plan_relay([12.0, 12.0, 18.0], [0, 0, 40.0], 29.0)["policy"] == "direct"   # odd world size
    if PROBE.exists():
        p = json.loads(PROBE.read_text())
        allr = [p["d2h_chunk25"]["all"]["per_gpu_gbs"][str(i)] for i in range(8)]
        fast = [0.0] * 4 + [p["d2h_chunk25"]["4-7"]["per_gpu_gbs"][str(i)] for i in range(4, 8)]
        assert fabric.plan_relay(allr, fast, 29.0)["policy"] == "relay"


def test_weighted_rows():
    rows = fabric.weighted_rows(8_000_000, [23.5, 23.5, 23.5, 23.5, 36.0, 36.0, 36.0, 36.0])
    assert sum(rows) == 8_000_000 and all(r % 128 == 0 for r in rows[:-1])
    assert rows[4] > rows[0] and abs(rows[4] / rows[0] - 36.0 / 23.5) < 0.01
    assert fabric.weighted_rows(1000, [1.0]) == [1000]
    with pytest.raises(ValueError):
        fabric.weighted_rows(10, [1.0, 0.0])


# ---------------------------------------------------------------------------------------------------------------
# two processes, the relay protocol with CPU copies
# ---------------------------------------------------------------------------------------------------------------
CHUNKS, CHUNK_BYTES = 200, 4096


def _pattern(k: int) -> np.ndarray:
    return ((np.arange(CHUNK_BYTES, dtype=np.uint32) * 2654435761 + k * 97) >> 7).astype(np.uint8)


def _sender(tag: str, ready) -> None:
    sys.path.insert(0, str(REPO))
    from lshrs_b200 import fabric as fb

    ready.wait()
    slots = fb.SharedHostBuffer(f"{tag}_slots", fb.SLOTS * CHUNK_BYTES, create=False)     # stands in for the IPC slots
    ctr = fb.SharedHostBuffer(f"{tag}_ctr", 8 * 2 * fb.SLOTS, create=False)
    hs = fb.RelayHandshake(ctr.array.view(np.int64), fb.SLOTS, timeout=30)
    uses = [0] * fb.SLOTS
    rng = np.random.default_rng(1)
    for k in range(CHUNKS):
        s = k % fb.SLOTS
        hs.wait_drained(s, uses[s])                  # the slot's previous content has been taken
        if rng.random() < 0.2:
            time.sleep(0.001)
        slots.array[s * CHUNK_BYTES:(s + 1) * CHUNK_BYTES] = _pattern(k)
        hs.mark_sent(s, uses[s])
        uses[s] += 1
    ctr.close()
    slots.close()


def _receiver(tag: str, ready, result) -> None:
    sys.path.insert(0, str(REPO))
    from lshrs_b200 import fabric as fb

    slots = fb.SharedHostBuffer(f"{tag}_slots", fb.SLOTS * CHUNK_BYTES, create=True)
    ctr = fb.SharedHostBuffer(f"{tag}_ctr", 8 * 2 * fb.SLOTS, create=True)
    ctr.array[:] = 0
    host = fb.SharedHostBuffer(f"{tag}_host", CHUNKS * CHUNK_BYTES, create=True)         # the sender's host buffer
    ready.set()
    hs = fb.RelayHandshake(ctr.array.view(np.int64), fb.SLOTS, timeout=30)
    uses = [0] * fb.SLOTS
    rng = np.random.default_rng(2)
    bad = 0
    for k in range(CHUNKS):
        s = k % fb.SLOTS
        hs.wait_sent(s, uses[s])
        if rng.random() < 0.2:
            time.sleep(0.001)
        got = slots.array[s * CHUNK_BYTES:(s + 1) * CHUNK_BYTES].copy()
        host.array[k * CHUNK_BYTES:(k + 1) * CHUNK_BYTES] = got
        bad += int(not np.array_equal(got, _pattern(k)))
        hs.mark_drained(s, uses[s])
        uses[s] += 1
    whole = all(np.array_equal(host.array[k * CHUNK_BYTES:(k + 1) * CHUNK_BYTES], _pattern(k)) for k in range(CHUNKS))
    result.put((bad, whole))
    for b in (slots, ctr, host):
        b.close()


def test_relay_handshake_two_processes():
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    tag = f"lshx_test_{os.getpid()}"
    ready, result = ctx.Event(), ctx.Queue()
    pr = ctx.Process(target=_receiver, args=(tag, ready, result))
    ps = ctx.Process(target=_sender, args=(tag, ready))
    pr.start(); ps.start()
    bad, whole = result.get(timeout=120)
    pr.join(60); ps.join(60)
    assert pr.exitcode == 0 and ps.exitcode == 0
    assert bad == 0 and whole


def test_handshake_times_out_instead_of_hanging():
    c = np.zeros(2 * fabric.SLOTS, dtype=np.int64)
    hs = fabric.RelayHandshake(c, fabric.SLOTS, timeout=0.2)
    hs.wait_drained(0, 0)                    # first use of a slot never waits
    with pytest.raises(TimeoutError, match="sent"):
        hs.wait_sent(1, 0)
    hs.mark_sent(1, 0)
    hs.wait_sent(1, 0)
    with pytest.raises(TimeoutError, match="drained"):
        hs.wait_drained(1, 1)


def test_shared_host_buffer_roundtrip():
    name = f"lshx_test_buf_{os.getpid()}"
    a = fabric.SharedHostBuffer(name, 1 << 16, create=True)
    b = fabric.SharedHostBuffer(name, 1 << 16, create=False)
    a.array[:] = 7
    b.array[100:200] = 9
    assert int(b.array[0]) == 7 and int(a.array[150]) == 9
    b.close()
    a.close()
    assert not os.path.exists(f"/dev/shm/{name}")
