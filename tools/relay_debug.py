#!/usr/bin/env python
"""Stage-by-stage bring-up of the relay primitives on two GPUs (profiling / debugging script).

    timeout 150 python -X faulthandler tools/relay_debug.py

Two spawned processes (rank 0 = sender on cuda:0, rank 1 = receiver on cuda:1) walk through: shared pinned host
buffer, interprocess events, CUDA-IPC tensor + peer copy, then the RelaySender / RelayReceiver pair.  Every stage
prints before and after, so a hang or a crash names its stage."""
import faulthandler
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))


def say(rank, *a):
    print(f"[{time.strftime('%H:%M:%S')} r{rank}]", *a, file=sys.stderr, flush=True)


def proc(rank, tag, q01, q10):
    faulthandler.enable()
    faulthandler.dump_traceback_later(100, exit=True)
    import numpy as np
    import torch

    from lshrs_b200 import fabric as fb

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    torch.zeros(1, device=dev)
    say(rank, "cuda up", torch.cuda.get_device_name(rank))
    send, recv = (q01, q10) if rank == 0 else (q10, q01)

    # ---- stage 1: shared pinned host buffer
    n = 8 << 20
    if rank == 0:
        host = fb.SharedHostBuffer(f"{tag}_host", n, create=True)
        host.array[:] = 0
        send.put("host ready")
    else:
        assert recv.get(timeout=30) == "host ready"
        host = fb.SharedHostBuffer(f"{tag}_host", n, create=False)
    say(rank, "stage 1: mapped; pinning")
    ht = host.pin()
    say(rank, "stage 1: pinned", ht.is_pinned())
    d = torch.full((n,), rank + 1, dtype=torch.uint8, device=dev)
    if rank == 1:
        ht[: n // 2].copy_(d[: n // 2], non_blocking=True)
        torch.cuda.synchronize()
        send.put("wrote")
    else:
        assert recv.get(timeout=30) == "wrote"
        say(rank, "stage 1: partner's D2H visible here:", int(host.array[0]), int(host.array[n // 2 - 1]), int(host.array[n // 2]))

    # ---- stage 2: interprocess events (liblshx)
    say(rank, "stage 2: creating ipc event")
    s = torch.cuda.Stream(dev)
    ev = fb._IpcEvent(rank)
    ev.record(s)
    send.put((rank, ev.handle))
    prank, ph = recv.get(timeout=30)
    say(rank, "stage 2: opening partner's handle")
    pev = fb._IpcEvent(rank, ph)
    say(rank, "stage 2: opened; waiting on it")
    pev.wait(s)
    s.synchronize()
    say(rank, "stage 2: ok")

    # ---- stage 3: CUDA-IPC memory + peer copy (liblshx)
    import ctypes

    from lshrs_b200 import _native

    lib = _native.lib()
    if rank == 1:
        p, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _native.check(lib.lshx_ipc_mem_alloc(rank, 4 << 20, ctypes.byref(p), h))
        say(rank, "stage 3: allocated, sending handle")
        send.put(bytes(h))
        assert recv.get(timeout=60) == "copied"
        back = torch.empty(4 << 20, dtype=torch.uint8, pin_memory=True)
        fb._copy_async(rank, back.data_ptr(), p.value, 4 << 20, s)
        s.synchronize()
        say(rank, "stage 3: slot now holds", int(back[0]), int(back[-1]))
        send.put("seen")
        lib.lshx_ipc_mem_free(rank, p)
    else:
        h = recv.get(timeout=60)
        say(rank, "stage 3: opening partner's memory")
        p = ctypes.c_void_p()
        _native.check(lib.lshx_ipc_mem_open(rank, (ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p)))
        say(rank, "stage 3: opened; peer copy")
        src = torch.full((4 << 20,), 7, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fb._copy_async(rank, p.value, src.data_ptr(), 4 << 20, s)
        s.synchronize()
        say(rank, "stage 3: copied; peer copy rate %.1f GB/s" % (20 * (4 << 20) / (time.perf_counter() - t0) / 1e9))
        send.put("copied")
        assert recv.get(timeout=60) == "seen"
        lib.lshx_ipc_mem_close(p)

    # ---- stage 4: the pair
    CH, NB = 16, 1 << 20
    if rank == 0:
        out = fb.SharedHostBuffer(f"{tag}_out", CH * NB, create=True)
        out.array[:] = 0
        out.pin()
        say(rank, "stage 4: sender")
        snd = fb.RelaySender(dev, out, tag)
        send.put(snd.export())
        snd.attach(recv.get(timeout=60))
        say(rank, "stage 4: attached")
        comp = torch.cuda.Stream(dev)
        bufs = [torch.empty(NB, dtype=torch.uint8, device=dev) for _ in range(2)]
        freed = [None, None]
        for k in range(CH):
            b = k & 1
            if freed[b] is not None:
                comp.wait_event(freed[b])
            with torch.cuda.stream(comp):
                bufs[b].fill_(k + 1)
            done = torch.cuda.Event()
            done.record(comp)
            freed[b] = snd.send(k, bufs[b], done)
        snd.stream.synchronize()
        say(rank, "stage 4: all sent")
        assert recv.get(timeout=60) == "drained"
        ok = all(int(out.array[k * NB]) == k + 1 and int(out.array[(k + 1) * NB - 1]) == k + 1 for k in range(CH))
        say(rank, "stage 4: host buffer correct:", ok)
        snd.close()
        send.put("closed")
        out.close()
    else:
        say(rank, "stage 4: receiver")
        rcv = fb.RelayReceiver(dev, NB, tag)
        say(rank, "stage 4: exported")
        send.put(rcv.export())
        rcv.attach(recv.get(timeout=60))
        say(rank, "stage 4: attached")
        for k in range(CH):
            rcv.drain(k, NB, k * NB)
        rcv.stream.synchronize()
        say(rank, "stage 4: all drained")
        send.put("drained")
        assert recv.get(timeout=60) == "closed"
        rcv.close()
    host.close()
    say(rank, "done")
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    tag = f"lshx_dbg_{os.getpid()}"
    q01, q10 = ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=proc, args=(r, tag, q01, q10)) for r in range(2)]
    [p.start() for p in ps]
    [p.join(140) for p in ps]
    print("exit codes", [p.exitcode for p in ps], file=sys.stderr)
    for p in ps:
        if p.is_alive():
            p.kill()
