"""Breakdown of LSHRS.index(ids, <CUDA tensor>) on a DeviceBucketStorage, 1 M x 768 resident vectors."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lshrs_b200 import LSHRS, DeviceBucketStorage  # noqa: E402

n, dim = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 768
dev = torch.device("cuda", 0)
xd = torch.randn((n, dim), device=dev)
idd = torch.arange(n, dtype=torch.int64, device=dev)
lsh = LSHRS(dim=dim, num_perm=256, storage=DeviceBucketStorage(), device=0)
q = xd[:8].cpu().numpy()


def tick(label, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    print(f"  {label}: {1e3 * (t - t0):.2f} ms", flush=True)
    return t


for rep in range(3):
    lsh.clear()
    torch.cuda.synchronize()
    print(f"rep {rep}")
    t = time.perf_counter()
    flag = torch.zeros(n, dtype=torch.uint8, device=dev)
    packed = lsh._hasher.hash_device(xd, zero_flag=flag)
    t = tick("hash_device", t)
    bad = bool(((idd < 0) | (flag != 0)).any())
    t = tick("validity check", t)
    lsh._dindex.add_device(packed, idd, torch.cuda.current_stream(dev).cuda_stream)
    t = tick("add_device", t)
    lsh.query_batch(q, top_k=10)
    t = tick("first query_batch (sort)", t)
    lsh.query_batch(q, top_k=10)
    t = tick("second query_batch", t)
    lsh.clear()
    t0 = time.perf_counter()
    lsh.index(idd, xd)
    t = tick("LSHRS.index(resident)", t0)
    lsh.query_batch(q, top_k=10)
    t = tick("query_batch", t)
    print(f"  => {n / (t - t0) / 1e6:.1f} M vectors/s", flush=True)
