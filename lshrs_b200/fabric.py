"""Host-fabric awareness for the multi-GPU hash path: which GPUs to use, and which PCIe links carry the signatures.

The hash path shards by rows with no exchange step (reference lshrs/hash/lsh.py:200-211: a signature depends on
its own vector only), so at N GPUs the only shared resource is the HOST side: every signature (32 B per vector at
768/256) has to land in pinned host memory.  On the 8-GPU boxes of this pool the PCIe links are not equal
(profiles/r2_fabric_probe_8gpu.json): GPUs 4-7 write 44 GB/s each when they copy alone (176 GB/s together),
GPUs 0-3 18.6 GB/s each (74 GB/s) -- and as soon as GPUs 0-3 copy at all, GPUs 4-7 drop to 18.8 GB/s each
(123 GB/s for all eight).  One B200 produces 29 GB/s of signatures, so with equal shards and every GPU writing
its own, the four slow links set the step time (round 1: 0.40 scaling efficiency at N = 8, 0.61 at N = 4).

Three tools, all driven by a measurement taken at start-up instead of a hard-wired topology:

* :func:`probe_links` / :func:`choose_devices` -- concurrent D2H / H2D rate of every visible GPU; a job that uses
  fewer GPUs than the box has takes the ones on the fastest links (N = 4 on this pool: GPUs 4-7).
* :func:`plan_relay` -- decides whether the GPUs on slow links should hand their signatures over NVLink to a
  partner on a fast link, which writes them to the host ("relay"), or write them themselves ("direct"), from the
  measured rates with everybody copying vs. only the fast half copying.  This is the one place NVLink earns its
  keep on this path; no collective, no NCCL on the data path -- a peer-to-peer copy-engine transfer per chunk.
* :class:`RelaySender` / :class:`RelayReceiver` -- the relay between two PROCESSES (one rank per GPU under
  torchrun): the receiver owns a few device slots (CUDA IPC), the sender owns the pinned host buffer the
  signatures finally land in (POSIX shared memory, registered by both ranks), interprocess CUDA events order the
  copies on the GPUs and :class:`RelayHandshake` orders the two host threads that enqueue them.
* :func:`weighted_rows` -- row counts proportional to measured link rates (the host-fed e2e path, where the
  vectors cross PCIe in the other direction).

Host-side plumbing only: no arithmetic of the hot path lives here.
"""

from __future__ import annotations

import ctypes  # noqa: F401  (annotations)
import json
import mmap
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

__all__ = ["probe_links", "probe_links_subprocess", "choose_devices", "plan_relay", "weighted_rows",
           "SharedHostBuffer", "RelayHandshake", "RelaySender", "RelayReceiver"]


# ------------------------------------------------------------------------------------------------ measurement
def probe_links(devices=None, mbytes: int = 256, chunk_mb: int = 25, reps: int = 2) -> dict:
    """Concurrent D2H and H2D rate (GB/s) of ``devices`` (default: all visible), one thread per GPU.

    Copies only -- no kernel is launched on any GPU, so a GPU that is probed but not chosen stays idle in
    utilisation samplers.  Returns ``{"devices": [...], "d2h_gbs": [...], "h2d_gbs": [...]}``.
    """
    import threading

    import torch

    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    nbytes, chunk = mbytes << 20, chunk_mb << 20
    bufs = {}
    for d in devices:
        bufs[d] = (torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"),
                   torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), torch.cuda.Stream(d))
    out = {"devices": list(devices)}
    for direction in ("d2h", "h2d"):
        rates = {}
        bar = threading.Barrier(len(devices))

        def work(d, direction=direction, rates=rates, bar=bar):
            dev, host, s = bufs[d]
            torch.cuda.set_device(d)
            with torch.cuda.stream(s):
                best = 0.0
                for rep in range(reps + 1):           # first pass is the warm-up
                    bar.wait()
                    t0 = time.perf_counter()
                    for o in range(0, nbytes, chunk):
                        if direction == "d2h":
                            host[o:o + chunk].copy_(dev[o:o + chunk], non_blocking=True)
                        else:
                            dev[o:o + chunk].copy_(host[o:o + chunk], non_blocking=True)
                    s.synchronize()
                    if rep:
                        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
                rates[d] = best

        threads = [threading.Thread(target=work, args=(d,)) for d in devices]
        [t.start() for t in threads]
        [t.join() for t in threads]
        out[f"{direction}_gbs"] = [round(rates[d], 2) for d in devices]
    return out


def probe_links_subprocess(timeout: float = 120.0) -> dict | None:
    """:func:`probe_links` over every visible GPU in a child process (its CUDA contexts die with it)."""
    code = ("import json, sys; sys.path.insert(0, %r); from lshrs_b200.fabric import probe_links; "
            "print('FABRIC ' + json.dumps(probe_links()))" % str(Path(__file__).resolve().parents[1]))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, env=env)
    except (OSError, subprocess.TimeoutExpired):
        return None
    for line in res.stdout.splitlines():
        if line.startswith("FABRIC "):
            return json.loads(line[7:])
    return None


def choose_devices(probe: dict | None, n: int, visible: int) -> list[int]:
    """The ``n`` GPUs whose host links were fastest with every GPU copying (D2H, ties by H2D, then index).

    Without a probe (or when every GPU is needed) the identity map."""
    if n >= visible or not probe or len(probe.get("d2h_gbs", [])) != len(probe.get("devices", [])):
        return list(range(n))
    order = sorted(range(len(probe["devices"])),
                   key=lambda i: (-probe["d2h_gbs"][i], -probe["h2d_gbs"][i], probe["devices"][i]))
    return sorted(probe["devices"][i] for i in order[:n])


def plan_relay(d2h_all: list[float], d2h_writers_only: list[float] | None, kernel_gbs: float) -> dict:
    """Direct or relayed gather of the signatures, from measured link rates (GB/s, indexed by rank).

    ``d2h_all[r]``: rank r's D2H rate with every rank copying.  ``d2h_writers_only[r]``: its rate with only the
    faster half copying (None / 0 for the others).  ``kernel_gbs``: the rate at which one GPU PRODUCES signature
    bytes.  Equal shards, max over ranks: the direct plan runs at ``min(d2h_all)`` per rank; the relay plan pairs
    the slowest rank with the fastest writer (and so on), every writer carries two shards, and runs at
    ``min(writers_only) / 2`` per rank.  The faster plan wins; "direct" on ties or when nothing is gained (e.g.
    links that already keep up with the kernel).
    """
    n = len(d2h_all)
    direct = min(d2h_all)
    plan = {"policy": "direct", "per_rank_gbs": {"direct": round(direct, 2)}, "pairs": {}, "writers": list(range(n))}
    if n < 2 or n % 2 or not d2h_writers_only:
        return plan
    order = sorted(range(n), key=lambda r: (-d2h_all[r], r))
    writers, senders = order[: n // 2], order[n // 2:][::-1]          # slowest sender first, fastest writer first
    if any(not d2h_writers_only[w] or d2h_writers_only[w] <= 0 for w in writers):
        return plan
    relay = min(d2h_writers_only[w] for w in writers) / 2.0
    plan["per_rank_gbs"]["relay"] = round(relay, 2)
    if min(direct, kernel_gbs) >= 0.97 * min(relay, kernel_gbs):      # the links keep up, or nothing to gain
        return plan
    plan.update(policy="relay", pairs={int(s): int(w) for s, w in zip(senders, writers)}, writers=sorted(writers))
    return plan


def weighted_rows(total_rows: int, rates: list[float], align: int = 128) -> list[int]:
    """Split ``total_rows`` proportionally to ``rates`` in multiples of ``align`` (last rank takes the rest)."""
    if total_rows < 0 or not rates or min(rates) <= 0:
        raise ValueError("need positive rates")
    s = float(sum(rates))
    rows = [int(total_rows * r / s) // align * align for r in rates]
    rows[-1] += total_rows - sum(rows)
    return rows


# ------------------------------------------------------------------------------------------- shared host memory
class SharedHostBuffer:
    """A byte buffer in POSIX shared memory (``/dev/shm``) that several ranks map; each may pin it for its GPU.

    The rank that owns the data creates it; a partner rank that writes into it over ITS PCIe link opens the same
    name.  ``pin()`` registers the mapping with CUDA (``cudaHostRegister``) so that copies into it are DMA.
    """

    def __init__(self, name: str, nbytes: int, create: bool) -> None:
        self.name, self.nbytes, self.created = name, int(nbytes), create
        self.path = f"/dev/shm/{name}"
        flags = os.O_RDWR | (os.O_CREAT | os.O_TRUNC if create else 0)
        fd = os.open(self.path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self._map = mmap.mmap(fd, self.nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self._map, dtype=np.uint8)
        self._pinned = False
        self._tensor = None

    def pin(self):
        """Register with CUDA for the current device's context; returns a uint8 torch view of the buffer."""
        import torch

        if self._tensor is None:
            self._tensor = torch.from_numpy(self.array)
        if not self._pinned:
            rc = torch.cuda.cudart().cudaHostRegister(self._tensor.data_ptr(), self.nbytes, 1)   # portable
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister({self.name}, {self.nbytes} bytes) failed: {rc}")
            self._pinned = True
        return self._tensor

    def close(self) -> None:
        if self._pinned:
            import torch

            torch.cuda.cudart().cudaHostUnregister(self._tensor.data_ptr())
            self._pinned = False
        self._tensor = None
        self.array = None
        try:
            self._map.close()
        except BufferError:   # a numpy view is still alive somewhere; the mapping goes with the process
            pass
        if self.created:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass


class RelayHandshake:
    """Orders the two HOST threads of a relay pair; the GPUs are ordered by interprocess CUDA events.

    ``cudaStreamWaitEvent`` on an interprocess event waits for the most recent ``cudaEventRecord`` issued BEFORE
    the wait call, so each side must know that its partner has already *issued* the record it is about to wait
    on.  Two monotonically increasing counters per slot live in shared memory: ``sent[s]`` = transfers into slot
    ``s`` whose full-event the sender has recorded, ``drained[s]`` = transfers out of slot ``s`` whose
    free-event the receiver has recorded.  Use ``u`` (0, 1, 2 ...) of a slot may be *sent* once ``drained >= u``
    and *drained* once ``sent >= u + 1``: the sender runs at most one use per slot ahead, the receiver never
    ahead, and neither waits for anything the other can only do after it -- no cycle.
    """

    def __init__(self, counters: np.ndarray, slots: int, timeout: float = 60.0) -> None:
        assert counters.dtype == np.int64 and counters.size >= 2 * slots
        self.c, self.slots, self.timeout = counters, slots, timeout

    def _spin(self, idx: int, want: int, what: str) -> None:
        t0 = time.perf_counter()
        pause = 0
        while int(self.c[idx]) < want:
            pause += 1
            if pause & 0xFF == 0:
                if time.perf_counter() - t0 > self.timeout:
                    raise TimeoutError(f"relay handshake: waited {self.timeout:.0f} s for {what} >= {want} "
                                       f"(have {int(self.c[idx])})")
                time.sleep(0)

    # sender side
    def wait_drained(self, slot: int, use: int) -> None:
        """Block until the receiver has recorded the free-event of use ``use - 1`` of ``slot``."""
        if use > 0:
            self._spin(self.slots + slot, use, f"drained[{slot}]")

    def mark_sent(self, slot: int, use: int) -> None:
        self.c[slot] = use + 1

    # receiver side
    def wait_sent(self, slot: int, use: int) -> None:
        self._spin(slot, use + 1, f"sent[{slot}]")

    def mark_drained(self, slot: int, use: int) -> None:
        self.c[self.slots + slot] = use + 1


# ----------------------------------------------------------------------------------------------- the relay pair
SLOTS = 4   # device slots per relay pair: the sender may run SLOTS - 1 chunks ahead of the receiver's D2H


def _handle_bytes() -> "ctypes.Array":
    import ctypes

    return (ctypes.c_ubyte * 64)()


class _IpcEvent:
    """An interprocess CUDA event (``lshx_ipc_event_*``): created here, or opened from a partner's 64-byte handle.

    (torch's own ``Event.from_ipc_handle`` objects crash in ``Stream.wait_event`` in the torch of this image, so
    the relay keeps its cross-process ordering inside liblshx.)"""

    def __init__(self, device: int, handle: bytes | None = None) -> None:
        import ctypes

        from lshrs_b200 import _native

        self.device = int(device)
        self._lib = _native.lib()
        ev = ctypes.c_void_p()
        if handle is None:
            buf = _handle_bytes()
            _native.check(self._lib.lshx_ipc_event_create(self.device, ctypes.byref(ev), buf))
            self.handle = bytes(buf)
        else:
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
            _native.check(self._lib.lshx_ipc_event_open(self.device, buf, ctypes.byref(ev)))
            self.handle = handle
        self._ev = ev

    def record(self, stream) -> None:
        from lshrs_b200 import _native

        _native.check(self._lib.lshx_ipc_event_record(self.device, self._ev, stream.cuda_stream))

    def wait(self, stream) -> None:
        from lshrs_b200 import _native

        _native.check(self._lib.lshx_ipc_event_wait(self.device, self._ev, stream.cuda_stream))

    def destroy(self) -> None:
        if self._ev is not None:
            self._lib.lshx_ipc_event_destroy(self._ev)
            self._ev = None


def _copy_async(device: int, dst: int, src: int, nbytes: int, stream) -> None:
    from lshrs_b200 import _native

    _native.check(_native.lib().lshx_memcpy_async(int(device), dst, src, int(nbytes), stream.cuda_stream))


class RelayReceiver:
    """The rank on the fast link: owns the device slots, writes its partner's signatures to the partner's buffer.

    ``export()`` returns what the sender needs (plain bytes / ints, picklable through the process group);
    ``attach()`` takes the sender's half.  ``drain(k, nbytes, host_offset)`` enqueues: wait for the sender's
    full-event of the slot, D2H the slot into the sender's shared host buffer at ``host_offset``, record the
    free-event.
    """

    def __init__(self, device, slot_bytes: int, tag: str) -> None:
        import ctypes

        import torch

        from lshrs_b200 import _native

        self.device = int(torch.device(device).index)
        self.slot_bytes, self.tag = int(slot_bytes), tag
        self._lib = _native.lib()
        self.slots, handles = [], []
        for _ in range(SLOTS):
            p, h = ctypes.c_void_p(), _handle_bytes()
            _native.check(self._lib.lshx_ipc_mem_alloc(self.device, self.slot_bytes, ctypes.byref(p), h))
            self.slots.append(p.value)
            handles.append(bytes(h))
        self.free = [_IpcEvent(self.device) for _ in range(SLOTS)]
        self.stream = torch.cuda.Stream(torch.device("cuda", self.device))
        for ev in self.free:
            ev.record(self.stream)
        self.counters = SharedHostBuffer(f"lshx_relay_{tag}_ctr", 8 * 2 * SLOTS, create=True)
        self.counters.array[:] = 0
        self.hs = RelayHandshake(self.counters.array.view(np.int64), SLOTS)
        self._export = {"slots": handles, "free": [ev.handle for ev in self.free], "device": self.device,
                        "counters": self.counters.name, "slot_bytes": self.slot_bytes}
        self.full = None
        self.host_buf = None
        self.host_ptr = 0
        self.uses = [0] * SLOTS

    def export(self) -> dict:
        return self._export

    def attach(self, sender: dict) -> None:
        self.full = [_IpcEvent(self.device, h) for h in sender["full"]]
        self.host_buf = SharedHostBuffer(sender["host_name"], sender["host_bytes"], create=False)
        self.host_ptr = int(self.host_buf.pin().data_ptr())

    def drain(self, k: int, nbytes: int, host_offset: int) -> None:
        s = k % SLOTS
        use = self.uses[s]
        self.hs.wait_sent(s, use)                 # the sender has ISSUED the full-event of this use
        self.full[s].wait(self.stream)
        _copy_async(self.device, self.host_ptr + host_offset, self.slots[s], nbytes, self.stream)
        self.free[s].record(self.stream)
        self.hs.mark_drained(s, use)
        self.uses[s] = use + 1

    def close(self) -> None:
        self.stream.synchronize()
        for ev in (self.full or []) + self.free:
            ev.destroy()
        self.full, self.free = None, []
        for p in self.slots:
            self._lib.lshx_ipc_mem_free(self.device, p)
        self.slots = []
        if self.host_buf is not None:
            self.host_buf.close()
            self.host_buf = None
        self.counters.close()


class RelaySender:
    """The rank on the slow link: pushes each chunk's signatures into its partner's device slot over NVLink."""

    def __init__(self, device, host_buf: SharedHostBuffer, tag: str) -> None:
        import torch

        from lshrs_b200 import _native

        self.device, self.tag = int(torch.device(device).index), tag
        self._lib = _native.lib()
        self.full = [_IpcEvent(self.device) for _ in range(SLOTS)]
        self.stream = torch.cuda.Stream(torch.device("cuda", self.device))
        for ev in self.full:
            ev.record(self.stream)
        self._export = {"full": [ev.handle for ev in self.full], "host_name": host_buf.name,
                        "host_bytes": host_buf.nbytes, "device": self.device}
        self.slots: list[int] = []
        self.free: list[_IpcEvent] = []
        self.counters = None
        self.uses = [0] * SLOTS

    def export(self) -> dict:
        return self._export

    def attach(self, receiver: dict) -> None:
        import ctypes

        from lshrs_b200 import _native

        for h in receiver["slots"]:
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            _native.check(self._lib.lshx_ipc_mem_open(self.device, buf, ctypes.byref(p)))
            self.slots.append(p.value)
        self.slot_bytes = int(receiver["slot_bytes"])
        self.free = [_IpcEvent(self.device, h) for h in receiver["free"]]
        self.counters = SharedHostBuffer(receiver["counters"], 8 * 2 * SLOTS, create=False)
        self.hs = RelayHandshake(self.counters.array.view(np.int64), SLOTS)

    def send(self, k: int, src, produced_event):
        """Enqueue the transfer of ``src`` (uint8 CUDA tensor, this rank's device) as relayed chunk ``k``.

        ``produced_event``: recorded on the stream that produced ``src``.  Returns an event that fires when
        ``src`` may be overwritten."""
        import torch

        nbytes = int(src.numel() * src.element_size())
        if nbytes > self.slot_bytes:
            raise ValueError(f"chunk of {nbytes} bytes does not fit the partner's {self.slot_bytes}-byte slot")
        s = k % SLOTS
        use = self.uses[s]
        self.stream.wait_event(produced_event)
        self.hs.wait_drained(s, use)               # the receiver has ISSUED the free-event of the previous use
        if use > 0:
            self.free[s].wait(self.stream)
        _copy_async(self.device, self.slots[s], int(src.data_ptr()), nbytes, self.stream)   # P2P over NVLink
        self.full[s].record(self.stream)
        self.hs.mark_sent(s, use)
        self.uses[s] = use + 1
        done = torch.cuda.Event()
        done.record(self.stream)
        return done

    def close(self) -> None:
        self.stream.synchronize()
        for p in self.slots:
            self._lib.lshx_ipc_mem_close(p)
        self.slots = []
        for ev in self.free + self.full:
            ev.destroy()
        self.free, self.full = [], []
        if self.counters is not None:
            self.counters.close()
            self.counters = None
