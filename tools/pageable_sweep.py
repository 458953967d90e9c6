"""Sweep the pageable-input knobs (LSHX_COPY_THREADS, LSHX_BOUNCE_MB) of lshx_hash_batch; one process per point."""
import os, subprocess, sys
CODE = r'''
import sys, time
sys.path.insert(0, "%s")
import numpy as np
from lshrs_b200 import LSHHasher
h = LSHHasher(16, 16, 768)
n = 1_000_000
X = np.random.default_rng(0).standard_normal((n, 768)).astype(np.float32)
out = np.empty((n, 32), np.uint8)
for _ in range(2): h.hash_into(X, n, out, x_on_device=False, out_on_device=False)
t = time.perf_counter()
for _ in range(4): h.hash_into(X, n, out, x_on_device=False, out_on_device=False)
dt = (time.perf_counter() - t) / 4
print(f"{n/dt/1e6:.1f} M vec/s  {n*3072/dt/1e9:.1f} GB/s")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for threads in (8, 12, 16):
    for mb in (64, 128, 256):
        env = dict(os.environ, LSHX_COPY_THREADS=str(threads), LSHX_BOUNCE_MB=str(mb))
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
        print(f"threads={threads:2d} bounce={mb:3d} MB:", r.stdout.strip() or r.stderr.strip()[-200:], flush=True)
