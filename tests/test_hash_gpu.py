"""GPU parity tests of the hash path (through the C ABI) against the oracle and the golden vectors.

Bar (BASELINE.json north_star): band keys bit-exact for every projection whose
|x.r| exceeds 1e-5 * ||x|| * ||r||; flips inside that margin are counted and
reported, padding bits are always zero.
"""

from __future__ import annotations

import math

import numpy as np
import pytest
from conftest import hash_case_names, load_golden, projections_for

from oracle import lshrs_oracle as oracle

pytestmark = pytest.mark.gpu

KERNELS = ("ffma", "tcgen05", "tcgen05_3xtf32", "tcgen05_tf32bf16")
TC_KERNELS = KERNELS[1:]
REL_MARGIN = 1e-5  # north_star exempt margin
MAX_FLIP_MARGIN = 1e-6  # largest |x.r| / (|x||r|) at which any arm may disagree with the fp32 oracle


def _hasher(nb, r, dim, seed=42, kernel="ffma"):
    from lshrs_b200 import LSHHasher, LshxError

    h = LSHHasher(nb, r, dim, seed=seed)
    try:
        h._ensure_handle()
        h.set_kernel(kernel)
    except LshxError as exc:
        if kernel.startswith("tcgen05") and "does not support" in str(exc):
            pytest.skip(f"tcgen05 kernel does not take this shape: {exc}")
        raise
    return h


def _assert_parity(got, X, projs, what=""):
    want = oracle.hash_batch_vectorized(projs, X)
    rep = oracle.compare_packed(got, want, oracle.projection_margins(projs, X), REL_MARGIN)
    assert rep["flips_outside_margin"] == 0, (what, rep)
    assert rep["nonzero_pad_bits"] == 0, (what, rep)
    # flips inside the margin must be rare -- at most 5 % of the bits that lie inside it (one flip of slack
    # for shapes with a handful of exempt bits) -- and must sit an order of magnitude below the margin itself:
    # a flip at 1e-6 * |x||r| would mean the arithmetic is 10x worse than the fp32 sgemm it stands in for
    assert rep["flips_inside_margin"] <= max(1, math.ceil(0.05 * rep["bits_inside_margin"])), (what, rep)
    assert rep["max_flipped_margin"] <= MAX_FLIP_MARGIN, (what, rep)
    return rep


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", hash_case_names())
def test_golden_signatures(name, kernel):
    case = load_golden(name)
    nb, r, dim, seed = (int(case[k]) for k in ("num_bands", "rows_per_band", "dim", "seed"))
    projs = projections_for(case)
    h = _hasher(nb, r, dim, seed, kernel)
    got = h.hash_batch_packed(case["X"])
    assert h.last_kernel == kernel
    assert got.shape == case["signatures"].shape and got.dtype == np.uint8
    finite = np.isfinite(case["X"]).all(axis=1)
    rep = oracle.compare_packed(got[finite], case["signatures"][finite],
                                oracle.projection_margins(projs, case["X"][finite]), REL_MARGIN)
    assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0, rep
    if name.startswith("hash_kat") or name.startswith("hash_tiny"):
        np.testing.assert_array_equal(got, case["signatures"])  # well-conditioned: byte for byte
    # NaN rows: every projection is NaN, `> 0` is False -> all-zero bytes like the reference
    nan_rows = np.isnan(case["X"]).all(axis=1)
    assert not got[nan_rows].any()
    # object API == packed API
    sigs = h.hash_batch(case["X"])
    assert [s.as_tuple() for s in sigs] == [tuple(bytes(b) for b in row) for row in got]
    first = h.hash_vector(case["X"][0])
    assert first.as_tuple() == sigs[0].as_tuple()
    assert all(isinstance(b, bytes) and len(b) == math.ceil(r / 8) for b in first)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize(
    "nb, r, dim, n, dist",
    [
        (16, 16, 768, 20_000, "gauss"),     # BASELINE config 1 / 2 shape
        (16, 32, 1536, 4_096, "gauss"),     # config 3 shape
        (16, 4, 128, 50_000, "sift"),       # config 5 shape (non-negative SIFT-like)
        (8, 16, 128, 10_000, "gauss"),
        (128, 8, 64, 3_000, "gauss"),       # 1024 bits: many column tiles
        (5, 20, 100, 2_000, "gauss"),       # ragged rows_per_band, dim % 32 != 0
        (3, 5, 7, 1_000, "gauss"),          # unaligned dim (no float4 path)
        (64, 64, 256, 500, "gauss"),        # 4096 bits (reference PRECOMPUTED_CONFIGS shape): 16 passes
        (16, 16, 4096, 300, "gauss"),       # long vectors: 128 K chunks
        (2, 128, 64, 400, "gauss"),         # 128 rows per band: 16-byte band keys
        (1, 1, 4, 100, "gauss"),            # smallest shape the TMA path takes
    ],
)
def test_parity_with_oracle(nb, r, dim, n, dist, kernel):
    rng = np.random.default_rng(1234)
    if dist == "gauss":
        X = rng.standard_normal((n, dim)).astype(np.float32)
    else:
        X = np.minimum(255.0, np.floor(np.abs(rng.standard_normal((n, dim))) * 40.0)).astype(np.float32)
    h = _hasher(nb, r, dim, 42, kernel)
    got = h.hash_batch_packed(X)
    rep = _assert_parity(got, X, h.projections, f"{nb}x{r}x{dim}")
    # the faithful per-vector reference loop on a slice (one sgemv per band, lsh.py:200)
    sl = slice(0, 256)
    rep2 = oracle.compare_packed(got[sl], oracle.hash_batch_packed(h.projections, X[sl]),
                                 oracle.projection_margins(h.projections, X[sl]), REL_MARGIN)
    assert rep2["flips_outside_margin"] == 0, rep2
    print(f"[parity {kernel} {nb}x{r} dim={dim} n={n}] {rep}")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 255, 257, 1000])
def test_ragged_batch_sizes(n, kernel):
    h = _hasher(16, 16, 768, 42, kernel)
    X = np.random.default_rng(n).standard_normal((n, 768)).astype(np.float32)
    _assert_parity(h.hash_batch_packed(X), X, h.projections, f"n={n}")


@pytest.mark.parametrize("kernel", TC_KERNELS)
@pytest.mark.parametrize(
    "nb, r, dim, n",
    [
        (16, 16, 768, 1), (16, 16, 768, 129), (16, 16, 768, 255), (16, 16, 768, 257), (16, 16, 768, 1000),
        (16, 16, 768, 40_000),      # more 256-row tiles than SM pairs: several tiles per pair, ragged tail
        (16, 32, 1536, 3_000),      # two passes per tile
        (16, 8, 32, 700),           # one K chunk, 128 columns (two accumulator stages)
        (48, 8, 96, 5_000),         # 384 columns: three passes of 128
    ],
)
def test_two_cta_kernel(nb, r, dim, n, kernel, monkeypatch):
    """The cta_group::2 kernel on batches smaller than the AUTO threshold (which needs one 256-row tile
    per SM pair): LSHX_TC_FLAGS = B_WARP | CG2 | CG2_ALWAYS, read when the plan is created."""
    monkeypatch.setenv("LSHX_TC_FLAGS", str(1 | 8 | 32))
    h = _hasher(nb, r, dim, 42, kernel)
    X = np.random.default_rng(n).standard_normal((n, dim)).astype(np.float32)
    X[n // 2] = 0.0
    got, flag = h.hash_batch_packed(X, return_zero_flag=True)
    _assert_parity(got, X, h.projections, f"2cta {nb}x{r}x{dim} n={n}")
    want_flag = np.zeros(n, dtype=bool)
    want_flag[n // 2] = True
    np.testing.assert_array_equal(flag.astype(bool), want_flag)


@pytest.mark.parametrize("kernel", TC_KERNELS)
@pytest.mark.parametrize(
    "nb, r, dim, n",
    [
        (16, 16, 768, 3_000),    # 24 K chunks per tile, streamed projections
        (16, 4, 128, 20_000),    # config 5: resident projections, 4 chunks, nibble bands
        (16, 8, 32, 1_500),      # ONE chunk per tile: the groups own alternate tiles
        (5, 20, 96, 2_000),      # three chunks per tile (odd): ownership flips from tile to tile; compact columns
        (16, 32, 1536, 1_000),   # two passes per tile
    ],
)
def test_two_converter_groups_in_the_one_cta_kernel(nb, r, dim, n, kernel, monkeypatch):
    """LSHX_TC_FLAGS = B_WARP | CONV2: the 1-CTA kernel with two converter warpgroups on alternate K
    chunks and the row flags handed to the epilogue through shared memory."""
    monkeypatch.setenv("LSHX_TC_FLAGS", str(1 | 64))
    h = _hasher(nb, r, dim, 42, kernel)
    X = np.random.default_rng(n + dim).standard_normal((n, dim)).astype(np.float32)
    X[::97] = 0.0
    X[5::211, : dim // 2] = 0.0          # leading zeros: the FP16 scale comes from a later chunk
    got, flag = h.hash_batch_packed(X, return_zero_flag=True)
    _assert_parity(got, X, h.projections, f"2conv {nb}x{r}x{dim} n={n}")
    np.testing.assert_array_equal(flag.astype(bool), (np.abs(X) <= 1e-8).all(axis=1))


@pytest.mark.parametrize("flags", [None, 1 | 8 | 32])   # default dispatch; 2-CTA kernel forced
@pytest.mark.parametrize("kernel", TC_KERNELS)
def test_vectors_outside_the_fp16_range_are_recomputed(kernel, flags, monkeypatch):
    """Scale-free parity.  The default tcgen05 arithmetic scales every vector into FP16's range from its
    first non-zero K chunk; vectors whose later elements overflow that range, or whose magnitude has no
    representable scale, must come back bit-exact from the FP32 recomputation (hash_tc.cu redo list)."""
    if flags is not None:
        monkeypatch.setenv("LSHX_TC_FLAGS", str(flags))
    rng = np.random.default_rng(99)
    n, dim = 6_000, 768
    X = rng.standard_normal((n, dim)).astype(np.float32)
    wild = np.arange(n) % 3 == 0          # magnitudes spread over 2^-40 .. 2^40 inside the row
    X[wild] = (X[wild] * np.exp2(rng.integers(-40, 41, size=X[wild].shape))).astype(np.float32)
    late = np.arange(n) % 3 == 1          # tiny first chunk, ordinary rest: the scale comes out too large
    X[late, :32] *= np.float32(1e-9)
    X[5] *= np.float32(1e30)              # uniformly huge / tiny rows are fine for the scaling itself
    X[8] *= np.float32(1e-30)
    X[11] = (X[11] * np.float32(1e-38) * 0.01).astype(np.float32)   # denormals only
    X[14, :700] = 0.0                     # first non-zero element late in the row
    X[17, 3] = np.inf
    X[20, 100] = -np.inf
    h = _hasher(16, 16, dim, 42, kernel)
    got, flag = h.hash_batch_packed(X, return_zero_flag=True)
    finite = np.isfinite(X).all(axis=1)
    dn = np.ones(n, dtype=bool)
    if kernel != "tcgen05":
        dn[11] = False                    # known deviation of the TF32 arms: denormal inputs are flushed
    keep = finite & dn
    _assert_parity(got[keep], X[keep], h.projections, f"{kernel} flags={flags}")
    # +-inf elements: the sign of every projection follows the sign of r at that position (numpy: inf * r)
    if kernel != "tcgen05_3xtf32":        # (3xTF32 multiplies the infinite hi by r_lo too: inf - inf)
        with np.errstate(invalid="ignore"):
            want = oracle.hash_batch_vectorized(h.projections, X[~finite])
        np.testing.assert_array_equal(got[~finite], want)
    np.testing.assert_array_equal(flag.astype(bool), (np.abs(X) <= 1e-8).all(axis=1))


@pytest.mark.parametrize("dtype", [np.float16, np.uint8, np.int8])
@pytest.mark.parametrize("n", [4096, 50_000])
def test_typed_host_batches_are_cast_on_the_device(dtype, n):
    """float16 / uint8 / int8 arrays go through lshx_hash_batch_typed (raw rows over PCIe, exact cast on the
    device): the same bytes as casting on the host first, which is what the reference does (lsh.py:162)."""
    rng = np.random.default_rng(7)
    dim = 128
    if dtype is np.float16:
        X = rng.standard_normal((n, dim)).astype(np.float16)
    elif dtype is np.uint8:
        X = np.minimum(255, np.floor(np.abs(rng.standard_normal((n, dim))) * 40)).astype(np.uint8)  # SIFT-like
    else:
        X = rng.integers(-128, 128, size=(n, dim), dtype=np.int8)
    X[17] = 0
    h = _hasher(16, 4, dim, 42, "auto")
    got, flag = h.hash_batch_packed(X, return_zero_flag=True)
    want, want_flag = h.hash_batch_packed(X.astype(np.float32), return_zero_flag=True)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(flag, want_flag)
    assert flag[17] == 1 and flag.sum() == (X.astype(np.float32) == 0).all(axis=1).sum()
    _assert_parity(got, X.astype(np.float32), h.projections, f"typed {np.dtype(dtype).name}")
    # the object API and a non-contiguous view take the same route
    assert h.hash_batch(X[:5000:1])[3].as_tuple() == tuple(bytes(b) for b in want[3])
    np.testing.assert_array_equal(h.hash_batch_packed(X[::2]), want[::2])


@pytest.mark.parametrize("kernel", KERNELS)
def test_zero_flag_is_prepare_vector_test(kernel):
    h = _hasher(4, 4, 32, 42, kernel)
    X = np.random.default_rng(0).standard_normal((300, 32)).astype(np.float32)
    X[3] = 0.0
    X[17] = 1e-9
    X[40] = 0.0
    X[40, 5] = 2e-8           # above atol: not a zero vector
    X[41] = 0.0
    X[41, 31] = np.nan        # np.allclose is False for NaN
    X[299] = -1e-8            # exactly atol: zero vector
    sig, flag = h.hash_batch_packed(X, return_zero_flag=True)
    want = np.array([oracle.is_zero_vector(x) for x in X], dtype=np.uint8)
    np.testing.assert_array_equal(flag, want)
    assert flag[[3, 17, 299]].all() and not flag[[40, 41]].any()
    assert not sig[3].any()   # zero vector straight into the hasher -> all-zero bytes, no error


@pytest.mark.parametrize("kernel", KERNELS)
def test_device_resident_path_and_properties(kernel):
    """Size-independent properties at a size the oracle would not finish: 1M x 768."""
    import torch

    h = _hasher(16, 16, 768, 42, kernel)
    n = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((n, 768), generator=g, device="cuda", dtype=torch.float32)
    s1 = h.hash_device(x)
    torch.cuda.synchronize()
    assert s1.shape == (n, 16, 2) and s1.dtype == torch.uint8
    # determinism
    assert torch.equal(s1, h.hash_device(x))
    # positive power-of-two scaling is exact in fp32 -> identical keys
    assert torch.equal(s1, h.hash_device(x * 4.0))
    # row permutation commutes with hashing
    perm = torch.randperm(n, device="cuda", generator=g)
    assert torch.equal(h.hash_device(x[perm].contiguous()), s1[perm])
    # negation complements every bit whose projection is non-zero: popcount(s ^ s_neg) ~ all 256 bits
    sneg = h.hash_device(-x)
    # a bit cannot be set for x and for -x (up to near-zero projections inside the exempt margin)
    both = int((s1 & sneg).count_nonzero())
    assert both <= n * 256 * 1e-5, both
    neither = int(torch.bitwise_not(s1 | sneg).count_nonzero())
    assert neither <= n * 256 * 1e-5, neither
    # ~50% of bits set on Gaussian data
    sample = s1[:4096].cpu().numpy()
    frac = np.unpackbits(sample).mean()
    assert 0.49 < frac < 0.51, frac
    # a sampled slice against the oracle
    idx = np.random.default_rng(0).choice(n, 2048, replace=False)
    Xs = x[torch.from_numpy(idx).cuda()].cpu().numpy()
    _assert_parity(s1.cpu().numpy()[idx], Xs, h.projections, "sample of 1M")
    # host-out path from a device input equals the device output
    out = np.empty((n, 32), dtype=np.uint8)
    h.hash_into(x, n, out, x_on_device=True, out_on_device=False)
    np.testing.assert_array_equal(out.reshape(n, 16, 2), s1.cpu().numpy())


def test_kernels_agree_with_each_other():
    X = np.random.default_rng(5).standard_normal((5000, 768)).astype(np.float32)
    a = _hasher(16, 16, 768, 42, "ffma").hash_batch_packed(X)
    b = _hasher(16, 16, 768, 42, "tcgen05").hash_batch_packed(X)
    projs = oracle.make_projections(16, 16, 768, 42)
    rep = oracle.compare_packed(a, b, oracle.projection_margins(projs, X), REL_MARGIN)
    assert rep["flips_outside_margin"] == 0, rep


def test_projections_rebinding_reuploads():
    # LSHRS.load_from_disk / __setstate__ assign hasher.projections (reference main.py:981, 1044)
    h = _hasher(4, 8, 16, seed=1)
    X = np.random.default_rng(0).standard_normal((64, 16)).astype(np.float32)
    before = h.hash_batch_packed(X)
    other = oracle.make_projections(4, 8, 16, seed=2)
    h.projections = other
    after = h.hash_batch_packed(X)
    np.testing.assert_array_equal(after, oracle.hash_batch_vectorized(other, X))
    assert not np.array_equal(before, after)


def test_same_seed_same_signatures_different_seed_differs():
    # reference tests/test_core.py:392-414
    X = np.random.default_rng(0).standard_normal((8, 32)).astype(np.float32)
    a = _hasher(4, 4, 32, seed=42).hash_batch(X)
    b = _hasher(4, 4, 32, seed=42).hash_batch(X)
    c = _hasher(4, 4, 32, seed=43).hash_batch(X)
    assert [s.as_tuple() for s in a] == [s.as_tuple() for s in b]
    assert [s.as_tuple() for s in a] != [s.as_tuple() for s in c]


def test_project_and_pack_single_band():
    h = _hasher(2, 10, 12, seed=3)
    v = np.random.default_rng(1).standard_normal(12).astype(np.float32)
    for proj in h.projections:
        assert h._project_and_pack(proj, v) == oracle.project_and_pack(proj, v)


def test_hashing_is_reentrant_from_threads():
    # reference tests/test_concurrency.py: many threads call ingest() -> hash_vector concurrently
    import threading

    h = _hasher(4, 4, 32)
    X = np.random.default_rng(0).standard_normal((100, 32)).astype(np.float32)
    want = [s.as_tuple() for s in h.hash_batch(X)]
    got: list = [None] * 100

    def work(t):
        for i in range(t * 10, t * 10 + 10):
            got[i] = h.hash_vector(X[i]).as_tuple()

    threads = [threading.Thread(target=work, args=(t,)) for t in range(10)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert got == want


def test_sharded_hasher_uses_every_visible_gpu():
    from lshrs_b200.sharding import ShardedHasher

    sh = ShardedHasher(16, 16, 768, seed=42)
    X = np.random.default_rng(2).standard_normal((3000, 768)).astype(np.float32)
    got = sh.hash_batch_packed(X)
    _assert_parity(got, X, sh.projections, "sharded")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("nb, r, dim, n", [(16, 16, 768, 129), (16, 4, 128, 130), (5, 20, 100, 257),
                                           (16, 32, 1536, 200), (40, 7, 32, 300), (3, 100, 256, 77)])
def test_no_writes_outside_the_output(nb, r, dim, n, kernel):
    """Canary rows around the signature / zero-flag buffers stay untouched (compute-sanitizer is
    closed on this pool, so out-of-bounds stores are caught this way)."""
    import torch

    h = _hasher(nb, r, dim, 42, kernel)
    sig = h.signature_bytes
    x = torch.randn((n, dim), device="cuda", dtype=torch.float32)
    out = torch.full((n + 4, sig), 0xAB, dtype=torch.uint8, device="cuda")
    flag = torch.full((n + 64,), 0xCD, dtype=torch.uint8, device="cuda")
    h.hash_into(x, n, out[2:], x_on_device=True, out_on_device=True, zero_flag=flag[32:],
                stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool((out[:2] == 0xAB).all()) and bool((out[n + 2:] == 0xAB).all())
    assert bool((flag[:32] == 0xCD).all()) and bool((flag[32 + n:] == 0xCD).all())
    assert not bool(flag[32:32 + n].any())
    got = out[2:n + 2].cpu().numpy().reshape(n, nb, -1)
    _assert_parity(got, x.cpu().numpy(), h.projections, "canary")


@pytest.mark.parametrize("name", hash_case_names())
def test_auto_kernel_and_per_vector_latency_path(name):
    """AUTO: a handful of host rows (what LSHRS.ingest / query send) take the small-batch kernel,
    larger batches the tiled kernels; both must reproduce the reference's bytes."""
    case = load_golden(name)
    nb, r, dim, seed = (int(case[k]) for k in ("num_bands", "rows_per_band", "dim", "seed"))
    projs = projections_for(case)
    h = _hasher(nb, r, dim, seed, "auto")
    X = case["X"]
    finite = np.isfinite(X).all(axis=1)
    margins = oracle.projection_margins(projs, X[finite])
    whole = h.hash_batch_packed(X)
    rep = oracle.compare_packed(whole[finite], case["signatures"][finite], margins, REL_MARGIN)
    assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0, rep
    # row by row (n = 1) and in threes: the latency path
    rows = np.stack([np.frombuffer(b"".join(h.hash_vector(x).as_tuple()), dtype=np.uint8) for x in X])
    assert h.last_kernel == "small"
    rep1 = oracle.compare_packed(rows.reshape(whole.shape)[finite], case["signatures"][finite], margins, REL_MARGIN)
    assert rep1["flips_outside_margin"] == 0 and rep1["nonzero_pad_bits"] == 0, rep1
    sig3, flag3 = h.hash_batch_packed(X[:3], return_zero_flag=True)
    rep3 = oracle.compare_packed(sig3[finite[:3]], case["signatures"][:3][finite[:3]],
                                 oracle.projection_margins(projs, X[:3][finite[:3]]), REL_MARGIN)
    assert rep3["flips_outside_margin"] == 0, rep3
    np.testing.assert_array_equal(flag3, [oracle.is_zero_vector(x) for x in X[:3]])


def _random_shapes(count, seed):
    rng = np.random.default_rng(seed)
    shapes = []
    for _ in range(count):
        nb = int(rng.integers(1, 41))
        r = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 24, 31, 32, 33, 64, 70]))
        dim = int(rng.choice([4, 8, 12, 20, 32, 36, 64, 100, 128, 132, 200, 256, 300]))
        n = int(rng.integers(1, 700))
        shapes.append((nb, r, dim, n))
    return shapes


@pytest.mark.parametrize("kernel", ("tcgen05", "ffma", "auto"))
def test_random_shapes_against_oracle(kernel):
    """Shape fuzz: band counts, ragged rows_per_band (compact / padded column layouts, one or more
    passes), K padding, ragged batch sizes -- every combination must reproduce the oracle's bytes."""
    for nb, r, dim, n in _random_shapes(40, seed=2024):
        X = np.random.default_rng(nb * 1000 + r * 10 + dim).standard_normal((n, dim)).astype(np.float32)
        h = _hasher(nb, r, dim, 7, kernel)
        got, flag = h.hash_batch_packed(X, return_zero_flag=True)
        assert got.shape == (n, nb, (r + 7) // 8)
        _assert_parity(got, X, h.projections, f"{kernel} {nb}x{r} dim={dim} n={n}")
        assert not flag.any()
        h.close()
