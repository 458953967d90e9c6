"""Where does the device band index spend its time?  add (append) / sort / lookup+join of n vectors x 16 bands of
2-byte keys, timed on the host around synchronised calls (run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel split).  Usage: python tools/index_probe.py [n] [nq]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lshrs_b200.storage.device import DeviceIndex  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
nb, bpb = 16, 2
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
sig = torch.randint(0, 256, (n, nb, bpb), dtype=torch.uint8, device=dev, generator=g)
ids = torch.arange(n, dtype=torch.int64, device=dev)
qsig = sig[torch.randint(0, n, (nq,), device=dev, generator=g)].cpu().numpy()
ix = DeviceIndex(nb, bpb, device=0)
for rep in range(3):
    ix.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ix.add_device(sig, ids, torch.cuda.current_stream().cuda_stream)
    t1 = time.perf_counter()
    ix.query(qsig[:8])                      # sorts the segments
    t2 = time.perf_counter()
    total, maxc = ix.query(qsig)
    t3 = time.perf_counter()
    # a small add on top of the sorted index, then a query: what an ingest / query mix pays
    ix.add_device(sig[:1000], ids[:1000] + n, torch.cuda.current_stream().cuda_stream)
    t4 = time.perf_counter()
    ix.query(qsig[:8])
    t5 = time.perf_counter()
    print(f"n={n}: append {1e3 * (t1 - t0):.2f} ms, first query (sort) {1e3 * (t2 - t1):.2f} ms = "
          f"{n * nb / (t2 - t1) / 1e9:.2f} G entries/s, {nq}-query join {1e3 * (t3 - t2):.2f} ms ({total} slots, max {maxc}), "
          f"+1000 vectors: append {1e3 * (t4 - t3):.2f} ms, re-sort {1e3 * (t5 - t4):.2f} ms", flush=True)
