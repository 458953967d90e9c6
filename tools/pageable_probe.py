import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from lshrs_b200 import LSHHasher
h = LSHHasher(16, 16, 768)
n = 1_000_000
X = np.random.default_rng(0).standard_normal((n, 768)).astype(np.float32)
out = np.empty((n, 32), np.uint8)
for name, (x, o) in {"pageable": (X, out)}.items():
    for _ in range(2): h.hash_into(x, n, o, x_on_device=False, out_on_device=False)
    t=time.perf_counter()
    for _ in range(3): h.hash_into(x, n, o, x_on_device=False, out_on_device=False)
    dt=(time.perf_counter()-t)/3
    print(name, f"{n/dt/1e6:.1f}M vec/s  {n*3072/dt/1e9:.1f} GB/s")
xp = torch.empty((n,768), dtype=torch.float32, pin_memory=True); xp.numpy()[:] = X
op = torch.empty((n,32), dtype=torch.uint8, pin_memory=True)
for _ in range(2): h.hash_into(xp, n, op, x_on_device=False, out_on_device=False)
t=time.perf_counter()
for _ in range(3): h.hash_into(xp, n, op, x_on_device=False, out_on_device=False)
dt=(time.perf_counter()-t)/3
print("pinned", f"{n/dt/1e6:.1f}M vec/s  {n*3072/dt/1e9:.1f} GB/s")
t=time.perf_counter(); sigs = h.hash_batch(X[:200000]); dt=time.perf_counter()-t
print("hash_batch (objects) 200k:", f"{200000/dt/1e3:.0f}k vec/s")
# typed host batches (cast on the device): SIFT-like uint8 at dim 128, fp16 at dim 768
for name, nb, r, dim, dt in (("uint8 dim128", 16, 4, 128, np.uint8), ("float16 dim768", 16, 16, 768, np.float16)):
    hh = LSHHasher(nb, r, dim)
    nn = 4_000_000 if dim == 128 else 1_000_000
    Xt = (np.random.default_rng(1).standard_normal((nn, dim)) * 40).astype(dt)
    Xf = Xt.astype(np.float32)
    for arr, label in ((Xt, name), (Xf, "float32 same data")):
        for _ in range(2): hh.hash_batch_packed(arr)
        t = time.perf_counter()
        for _ in range(3): hh.hash_batch_packed(arr)
        dt_s = (time.perf_counter() - t) / 3
        print(f"{label:18s} dim {dim}: {nn/dt_s/1e6:.1f} M vec/s")
