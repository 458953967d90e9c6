"""Does a host->device copy from pinned memory pay for CPU writes that just preceded it?"""
import time
import numpy as np
import torch

dev = torch.device("cuda", 0)
for mb in (1, 12, 64):
    n = mb << 20
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    src = np.random.default_rng(0).integers(0, 255, n, dtype=np.uint8)
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    hv = host.numpy()
    res = {}
    for mode in ("untouched", "cpu-written", "cpu-written+sleep1ms", "untouched"):
        ts = []
        for rep in range(12):
            if mode.startswith("cpu-written"):
                hv[:] = src
            if mode.endswith("sleep1ms"):
                time.sleep(0.001)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[mode] = 1e6 * float(np.median(ts[2:]))
        print(f"{mb:3d} MB {mode:22s}: {res[mode]:8.0f} us  ({n / res[mode] / 1e3:.1f} GB/s)", flush=True)
