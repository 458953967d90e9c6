#!/usr/bin/env python
"""Latency of one hash call vs batch size and kernel (host pinned in/out and device in/out)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from lshrs_b200 import LSHHasher

for kernel in ("auto", "tcgen05", "ffma"):
    h = LSHHasher(16, 16, 768)
    h._ensure_handle(); h.set_kernel(kernel)
    line = [f"{kernel:8s}"]
    for n in (1, 8, 32, 64, 256, 1024, 4096, 16384, 65536):
        x = torch.randn((n, 768), device="cuda")
        o = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
        xh = torch.empty((n, 768), pin_memory=True).normal_(); oh = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
        for _ in range(5):
            h.hash_into(x, n, o, x_on_device=True, out_on_device=True); h.hash_into(xh, n, oh, x_on_device=False, out_on_device=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            h.hash_into(x, n, o, x_on_device=True, out_on_device=True)
        e1.record(); torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) / 20 * 1e3
        t = time.perf_counter()
        for _ in range(20):
            h.hash_into(xh, n, oh, x_on_device=False, out_on_device=False)
        host_us = (time.perf_counter() - t) / 20 * 1e6
        line.append(f"n={n}: dev {dev_us:.0f}us host {host_us:.0f}us [{h.last_kernel}]")
    print(" | ".join(line), flush=True)
