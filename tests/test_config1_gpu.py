"""BASELINE config 1 in full: 100 000 x 768 -> 16 x 16 bits against the faithful reference loop."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import lshrs_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["tcgen05", "tcgen05_3xtf32", "tcgen05_tf32bf16", "ffma"])
def test_config1_all_band_keys(kernel):
    from lshrs_b200 import LSHHasher

    # R = LSHHasher(16, 16, 768, seed=42); X = default_rng(0) standard normal
    X = np.random.default_rng(0).standard_normal((100_000, 768)).astype(np.float32)
    h = LSHHasher(16, 16, 768, seed=42)
    h._ensure_handle()
    h.set_kernel(kernel)
    got = h.hash_batch_packed(X)
    assert h.last_kernel == kernel
    # the golden fixture holds the reference's own bytes for the first 64 of these rows
    from conftest import load_golden

    case = load_golden("hash_cfg_768_16x16")
    np.testing.assert_array_equal(case["X"], X[:64])
    margins64 = oracle.projection_margins(h.projections, X[:64])
    rep64 = oracle.compare_packed(got[:64], case["signatures"], margins64, 1e-5)
    assert rep64["flips_outside_margin"] == 0, rep64
    # all 1.6 M band keys: the one-sgemm oracle everywhere, the per-vector reference loop on 8192 rows
    margins = oracle.projection_margins(h.projections, X)
    rep = oracle.compare_packed(got, oracle.hash_batch_vectorized(h.projections, X), margins, 1e-5)
    assert rep["band_keys"] == 1_600_000
    assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0, rep
    sl = slice(40_000, 48_192)
    rep_ref = oracle.compare_packed(got[sl], oracle.hash_batch_packed(h.projections, X[sl]), margins[sl], 1e-5)
    assert rep_ref["flips_outside_margin"] == 0, rep_ref
    # near-zero flips are counted and must stay a small fraction of the exempt bits
    assert rep["flips_inside_margin"] <= 0.05 * rep["bits_inside_margin"], rep
    # empirical headroom: the largest relative margin at which this arm's bit differs from the fp32 oracle's
    # (and from the reference loop's) must be far inside the 1e-5 band -- measured on the hardware, not modelled
    assert rep["max_flipped_margin"] <= 1e-6 and rep_ref["max_flipped_margin"] <= 1e-6, (rep, rep_ref)
    print(f"[config 1 {kernel}] {rep}  reference-loop slice: {rep_ref}")


@pytest.mark.parametrize("kernel, n", [("tcgen05", 128), ("tcgen05", 24_000), ("tcgen05_tf32bf16", 24_000),
                                       ("tcgen05_3xtf32", 128), ("tcgen05_3xtf32", 24_000)])
def test_tensor_core_accumulators_against_fp64(kernel, n):
    """The split arithmetic measured ON THE HARDWARE: the raw fp32 TMEM accumulators of one tile (1-CTA kernel
    for 128 rows, 2-CTA kernel for a batch of 24 000) against float64 dot products.

    For the scaled FP16x3 arm an entry is s_x[i] * s_r[c] * (x_i . r_c) with exact power-of-two scales; the
    scales are recovered from the dump itself (entries with a comfortable margin), checked to be powers of two
    that factor into row x column, and divided out.  Bar: |acc / (s_x s_r) - x.r| <= 1e-6 * |x||r| everywhere --
    ten times inside the parity margin; tools/sim_splits.py predicts ~2e-8 for the split alone, this adds the
    tensor core's fp32 accumulation in TMEM.
    """
    from lshrs_b200 import LSHHasher

    X = np.random.default_rng(7).standard_normal((n, 768)).astype(np.float32)
    X[3] *= 1e-6          # rows of very different magnitude: the per-vector scale must absorb it
    X[5] *= 3e4
    h = LSHHasher(16, 16, 768, seed=42)
    h._ensure_handle()
    h.set_kernel(kernel)
    acc = h.debug_accumulators(X)
    rows, cols = acc.shape
    assert cols == 256 and rows == (128 if n == 128 else 256), acc.shape
    R = np.concatenate(h.projections, axis=0).astype(np.float64)           # column c = projection row c
    Xd = X[:rows].astype(np.float64)
    dots = Xd @ R.T
    norm = np.linalg.norm(Xd, axis=1)[:, None] * np.linalg.norm(R, axis=1)[None, :]
    margin = np.abs(dots) / norm
    with np.errstate(divide="ignore", invalid="ignore"):
        lg = np.log2(np.abs(acc.astype(np.float64) / dots))
    good = margin > 0.01
    # per-row and per-column log2 scale from well-conditioned entries; they must be integers (exact powers of 2)
    lg_good = np.where(good, lg, np.nan)
    assert good.sum(axis=1).min() >= 8 and good.sum(axis=0).min() >= 8      # every row / column has anchors
    sx = np.rint(np.nanmedian(lg_good, axis=1))                             # alternate: rows given columns, ...
    for _ in range(3):
        base = np.rint(np.nanmedian(lg_good - sx[:, None], axis=0))         # ... columns given rows
        sx = np.rint(np.nanmedian(lg_good - base[None, :], axis=1))
    scale = np.exp2(sx[:, None] + base[None, :])
    assert np.nanmax(np.abs(lg_good - np.log2(scale))) < 1e-3, "scales are not exact powers of two per row x column"
    if kernel != "tcgen05":
        assert np.all(scale == 1.0)                                         # the TF32 arms are scale-free
    err = np.abs(acc.astype(np.float64) / scale - dots) / norm
    print(f"[accumulators {kernel} n={n}] max |acc/(s_x s_r) - x.r| / (|x||r|) = {err.max():.3e}, "
          f"mean {err.mean():.3e}, log2 s_x range {sx.min():.0f}..{sx.max():.0f}")
    assert err.max() <= 1e-6, err.max()
    # and the signs the kernel emitted are the signs of these accumulators
    got = h.hash_batch_packed(X)[:rows]
    bits = np.unpackbits(got.reshape(rows, -1), axis=1, bitorder="little")[:, :cols]
    np.testing.assert_array_equal(bits.astype(bool), acc > 0)


def test_projection_rows_without_an_fp16_scale_take_the_fp32_kernel():
    """a projection row whose largest |r| is below 2^-114 (or infinite / NaN) cannot be scaled
    into FP16.  AUTO must hash such a hasher with the FP32 kernel (same bits as numpy), and an explicit request
    for the FP16x3 arm must fail loudly -- never flush the row to zero bits."""
    from lshrs_b200 import LSHHasher, LshxError

    rng = np.random.default_rng(3)
    X = rng.standard_normal((5_000, 64)).astype(np.float32)
    for poison in ("tiny", "inf"):
        h = LSHHasher(4, 8, 64, seed=1)
        projs = [p.copy() for p in h.projections]
        if poison == "tiny":
            projs[2][3] *= np.float32(2.0 ** -126)       # largest |r| ~ 2^-125: no FP16 scale exists
        else:
            projs[1][0, 5] = np.inf
        h.projections = projs
        got = h.hash_batch_packed(X)
        assert h.last_kernel == "ffma"
        with np.errstate(invalid="ignore", over="ignore"):
            want = oracle.hash_batch_vectorized(projs, X)
        rep = oracle.compare_packed(got, want, oracle.projection_margins(
            [np.nan_to_num(p, posinf=3e38) for p in projs], X), 1e-5)
        assert rep["flips_outside_margin"] == 0, (poison, rep)
        assert got[:, 2 if poison == "tiny" else 1].any()    # the poisoned band still carries information
        with pytest.raises(LshxError, match="FP16"):
            h.set_kernel("tcgen05")
        h.set_kernel("tcgen05_tf32bf16")                      # the scale-free arms remain selectable
        h.close()
    # back to ordinary planes: the tensor-core arm is the default again
    h = LSHHasher(4, 8, 64, seed=1)
    h.hash_batch_packed(X)
    assert h.last_kernel == "tcgen05"


def test_recompute_lists_of_concurrent_launches_do_not_mix():
    """The FP16x3 arm's FP32-recompute tile lists live in a handle-owned ring of scratch slots (no allocation per
    launch): many back-to-back launches on two streams, each with out-of-range rows in different tiles."""
    import torch

    from lshrs_b200 import LSHHasher

    dev = torch.device("cuda", 0)
    h = LSHHasher(16, 16, 768, seed=42, device=0)
    rng = np.random.default_rng(9)
    batches, wants = [], []
    for i in range(12):
        X = rng.standard_normal((3_000 + 128 * i, 768)).astype(np.float32)
        X[17 + 130 * i, 0] = np.float32(1e-40)             # denormal first chunk -> FP32 recompute of that tile
        X[17 + 130 * i, 1:32] = 0
        X[900 + i, 40] = np.inf                            # infinite element -> recompute as well
        batches.append(torch.from_numpy(X).to(dev))
        with np.errstate(invalid="ignore"):
            wants.append(oracle.hash_batch_vectorized(h.projections, X))
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        outs = []
        for i, xb in enumerate(batches):
            with torch.cuda.stream(streams[i & 1]):
                outs.append(h.hash_device(xb))
    torch.cuda.synchronize()
    for i, (o, want) in enumerate(zip(outs, wants)):
        X = batches[i].cpu().numpy()
        finite = np.isfinite(X).all(axis=1)
        rep = oracle.compare_packed(o.cpu().numpy()[finite], want[finite],
                                    oracle.projection_margins(h.projections, X[finite]), 1e-5)
        assert rep["flips_outside_margin"] == 0 and rep["nonzero_pad_bits"] == 0, (i, rep)
        inf_row = 900 + i                                   # inf * r: +inf where r > 0 -> bit 1, like numpy
        np.testing.assert_array_equal(o.cpu().numpy()[inf_row], want[inf_row])
