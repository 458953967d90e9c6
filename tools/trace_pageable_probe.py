import sys, numpy as np
sys.path.insert(0, "/root/repo")
from lshrs_b200 import LSHHasher
h = LSHHasher(16, 16, 768, seed=42, device=0)
X = np.random.default_rng(0).standard_normal((65536, 768)).astype(np.float32)
for n in (65536, 4096, 4096, 4096, 8192, 8192, 16384, 16384, 65536, 65536):
    h.hash_batch_packed(X[:n])
