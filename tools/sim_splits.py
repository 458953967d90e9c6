#!/usr/bin/env python
"""numpy model of the three operand splits of the tcgen05 hash kernel.

    python tools/sim_splits.py            # the table of section 3.1 on BASELINE-shaped data

Products of the split parts are exact in the tensor cores' fp32 accumulators up to the accumulation
order, so the model evaluates them in float64: what is left is exactly the error of the SPLIT (the
rounding of hi / lo parts and the dropped lo.lo term), which is the quantity that decides whether a
sign can flip outside the 1e-5 * |x||r| parity margin.  Used by tests/test_split_numerics.py.
"""
from __future__ import annotations

import numpy as np


def tf32(v: np.ndarray) -> np.ndarray:
    """Round to nearest (ties away) to TF32, like hash_tc.cu tf32_rna."""
    u = np.ascontiguousarray(v, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16(v: np.ndarray) -> np.ndarray:
    """Round to nearest even to BF16 (cvt.rn.bf16x2.f32), returned as float32."""
    u = np.ascontiguousarray(v, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def pow2_scale(m: np.ndarray, lo_exp: int) -> np.ndarray:
    """Power of two s with m * s in [2^lo_exp, 2^(lo_exp+1)) (m > 0), like row_scale_for / split_f16_kernel."""
    _, e = np.frexp(m.astype(np.float64))          # m = f * 2^e, f in [0.5, 1)
    return np.exp2(lo_exp - (e - 1)).astype(np.float32)


def dot_3xtf32(X, R):
    xh, rh = tf32(X), tf32(R)
    xl, rl = tf32(X - xh), tf32(R - rh)
    f = lambda a: a.astype(np.float64)  # noqa: E731
    return f(xl) @ f(rh).T + f(xh) @ f(rl).T + f(xh) @ f(rh).T


def dot_tf32_bf16(X, R):
    xh, rh = tf32(X), tf32(R)
    xl, rl = X - xh, R - rh
    f = lambda a: a.astype(np.float64)  # noqa: E731
    return f(bf16(xl)) @ f(bf16(rh)).T + f(bf16(xh)) @ f(bf16(rl)).T + f(xh) @ f(rh).T


def dot_fp16x3(X, R, first_chunk: int = 32):
    """Scaled FP16x3.  Returns (dots in the ORIGINAL scale, rows that overflow FP16 and are recomputed)."""
    n, dim = X.shape
    m0 = np.zeros(n, dtype=np.float32)
    for c in range(0, dim, first_chunk):           # first K chunk holding a non-zero element
        cm = np.abs(X[:, c:c + first_chunk]).max(axis=1)
        m0 = np.where(m0 == 0, cm, m0)
    sx = np.where(m0 > 0, pow2_scale(np.where(m0 > 0, m0, 1), 1), 1).astype(np.float32)[:, None]
    rm = np.abs(R).max(axis=1)
    sr = np.where(rm > 0, pow2_scale(np.where(rm > 0, rm, 1), 13), 1).astype(np.float32)[:, None]
    Y, Q = X * sx, R * sr
    redo = (np.abs(Y) > 65504).any(axis=1)
    with np.errstate(over="ignore", invalid="ignore"):
        yh = Y.astype(np.float16)
        yl = (Y - yh.astype(np.float32)).astype(np.float16)
        qh = Q.astype(np.float16)
        ql = (Q - qh.astype(np.float32)).astype(np.float16)
    f = lambda a: a.astype(np.float64)  # noqa: E731
    with np.errstate(over="ignore", invalid="ignore"):
        d = f(yl) @ f(qh).T + f(yh) @ f(ql).T + f(yh) @ f(qh).T
    return d / (f(sx) * f(sr).T), redo


def report(X, R) -> dict:
    X64, R64 = X.astype(np.float64), R.astype(np.float64)
    truth = X64 @ R64.T
    scale = np.linalg.norm(X64, axis=1)[:, None] * np.linalg.norm(R64, axis=1)[None, :]
    out = {}
    f16, redo = dot_fp16x3(X, R)
    arms = {"3xtf32": (dot_3xtf32(X, R), None), "tf32+bf16": (dot_tf32_bf16(X, R), None),
            "fp16x3": (f16, redo), "fp32 sgemm": ((X @ R.T).astype(np.float64), None)}
    for name, (d, skip) in arms.items():
        keep = np.ones(len(X), dtype=bool) if skip is None else ~skip
        err = np.abs(d[keep] - truth[keep]) / scale[keep]
        flips = (d[keep] > 0) != (truth[keep] > 0)
        out[name] = {"max_rel_err": float(err.max()), "flips": int(flips.sum()),
                     "flips_outside_margin": int((flips & (np.abs(truth[keep]) > 1e-5 * scale[keep])).sum()),
                     "recomputed_rows": 0 if skip is None else int(skip.sum())}
    return out


def main() -> None:
    for dim, nperm, n, kind in [(768, 256, 20000, "gauss"), (1536, 512, 5000, "gauss"), (128, 64, 100000, "gauss"),
                                (128, 64, 100000, "sift"), (768, 256, 5000, "wide")]:
        R = np.random.default_rng(42).standard_normal((nperm, dim)).astype(np.float32)
        rng = np.random.default_rng(0)
        X = rng.standard_normal((n, dim)).astype(np.float32)
        if kind == "sift":
            X = np.minimum(255, np.floor(np.abs(X) * 40)).astype(np.float32)
        if kind == "wide":   # magnitudes spread over 2^-12 .. 2^5 inside every row
            X = (X * np.exp2(rng.integers(-12, 6, size=X.shape))).astype(np.float32)
        for name, rep in report(X, R).items():
            print(f"dim {dim:5d} bits {nperm:4d} {kind:6s} {name:11s} max err {rep['max_rel_err']:.3g} |x||r|  "
                  f"flips {rep['flips']} (outside margin {rep['flips_outside_margin']})  recomputed rows {rep['recomputed_rows']}")


if __name__ == "__main__":
    main()
