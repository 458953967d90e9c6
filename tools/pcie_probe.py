#!/usr/bin/env python
"""Concurrent H2D / D2H bandwidth of every visible GPU (one thread per GPU), to tell a host-fabric limit
from a kernel limit in the multi-GPU bench.  Prints topology hints too."""
import subprocess, sys, threading, time
import torch

def run(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout
    except Exception as e:
        return str(e)

print(run("nvidia-smi topo -m | head -14"))
print(run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
n = torch.cuda.device_count()
nbytes = 1 << 30
host = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
devs = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
def worker(i, direction, out):
    torch.cuda.set_device(i)
    s = torch.cuda.Stream(i)
    with torch.cuda.stream(s):
        for _ in range(2):
            (devs[i].copy_(host[i], non_blocking=True) if direction == "h2d" else host[i].copy_(devs[i], non_blocking=True))
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            (devs[i].copy_(host[i], non_blocking=True) if direction == "h2d" else host[i].copy_(devs[i], non_blocking=True))
        s.synchronize()
        out[i] = 8 * nbytes / (time.perf_counter() - t0) / 1e9
for direction in ("h2d", "d2h"):
    for active in ([0], list(range(n))):
        out = {}
        th = [threading.Thread(target=worker, args=(i, direction, out)) for i in active]
        [t.start() for t in th]; [t.join() for t in th]
        print(direction, f"{len(active)} gpu(s):", " ".join(f"{out[i]:.1f}" for i in active), "GB/s each; total %.1f" % sum(out.values()))
