"""Band / row selection for ``LSHRS(dim=..., num_perm=...)`` when ``num_bands`` / ``rows_per_band`` are not given.

Host-only arithmetic (no kernel): it decides the SHAPE the projection kernel runs, so it has to pick what
the reference picks for every ``(num_perm, similarity_threshold)`` -- ``get_optimal_config`` of reference
lshrs/utils/br.py:325-395: (1) a table of pre-searched shapes for 2^12 .. 2^16 bits (br.py:38-79), used when
a tabulated threshold lies within 0.05 of the target; (2) otherwise every factorisation ``b * r == num_perm``
whose S-curve midpoint ``(1/b)^(1/r)`` is within 0.05 of the target, scored by the area of the detection
curve ``P(s) = 1 - (1 - s^r)^b`` on the wrong side of the target (false positives below it + false negatives
above it, br.py:205-220); (3) otherwise the factor pair closest to the square root (br.py:386-395).
``tests/golden/manifest.json`` holds the reference's answers for a grid of inputs (tools/make_golden.py) and
``tests/test_host_logic.py`` compares against all of them.
"""

from __future__ import annotations

import math
from typing import Optional

import numpy as np

__all__ = ["compute_lsh_threshold", "compute_collision_probability", "compute_false_rates",
           "find_optimal_br", "get_optimal_config", "print_config_analysis", "PRECOMPUTED_CONFIGS"]

# rows_per_band the reference tabulates per (signature bits, threshold); bands = bits // rows (br.py:38-79)
_TABLE_ROWS = {
    4096: {0.5: 8, 0.7: 16, 0.85: 32, 0.9: 64, 0.95: 128},
    8192: {0.4: 8, 0.7: 16, 0.8: 32, 0.85: 32, 0.9: 64, 0.95: 128},
    16384: {0.4: 8, 0.6: 16, 0.8: 32, 0.85: 32, 0.9: 64, 0.95: 128},
    32768: {0.4: 8, 0.6: 16, 0.8: 32, 0.85: 32, 0.9: 64, 0.95: 128},
    65536: {0.3: 8, 0.6: 16, 0.8: 32, 0.85: 64, 0.9: 64, 0.95: 128},
}
PRECOMPUTED_CONFIGS = {bits: {t: (bits // r, r) for t, r in rows.items()} for bits, rows in _TABLE_ROWS.items()}


def compute_lsh_threshold(b: int, r: int) -> float:
    """Similarity at which a ``b x r`` banding detects a pair about half of the time (br.py:119)."""
    return (1 / b) ** (1 / r)


def compute_collision_probability(similarity: float, b: int, r: int) -> float:
    """``1 - (1 - s^r)^b``: probability that two items of similarity ``s`` share at least one band."""
    return 1 - (1 - similarity ** r) ** b


def _integrate(fn, lo: float, hi: float) -> float:
    try:
        from scipy.integrate import quad  # the reference's integrator (br.py:217-218): identical scores
    except ImportError:  # pragma: no cover - scipy is a hard dependency of the reference
        # composite Gauss-Legendre, 64 panels x 16 nodes: ~1e-12 on these smooth integrands
        x, w = np.polynomial.legendre.leggauss(16)
        edges = np.linspace(lo, hi, 65)
        total = 0.0
        for a, b in zip(edges[:-1], edges[1:]):
            half = 0.5 * (b - a)
            total += half * float(np.dot(w, [fn(half * xi + 0.5 * (a + b)) for xi in x]))
        return total
    return quad(fn, lo, hi, limit=100)[0]


def compute_false_rates(b: int, r: int, threshold: float) -> tuple[float, float]:
    """(area of the detection curve below ``threshold``, area of its complement above it) -- br.py:205-220."""
    fp = _integrate(lambda s: 1 - (1 - s ** r) ** b, 0, threshold)
    fn = _integrate(lambda s: (1 - s ** r) ** b, threshold, 1)
    return fp, fn


def find_optimal_br(num_perm: int, target_threshold: float, tolerance: float = 0.05) -> Optional[tuple[int, int]]:
    """Best-scoring factorisation whose midpoint is within ``tolerance`` of the target, else ``None``.

    Candidate order matters for ties (strict ``<``): small ``r`` first (r = 1 .. floor(sqrt)), then small ``b``
    (br.py:270-320).
    """
    root = int(np.sqrt(num_perm))
    candidates = [(num_perm // r, r) for r in range(1, root + 1) if num_perm % r == 0]
    candidates += [(b, num_perm // b) for b in range(1, root + 1) if num_perm % b == 0]
    best, best_score = None, math.inf
    for b, r in candidates:
        if abs(compute_lsh_threshold(b, r) - target_threshold) > tolerance:
            continue
        score = sum(compute_false_rates(b, r, target_threshold))
        if score < best_score:
            best, best_score = (b, r), score
    return best


def get_optimal_config(num_perm: int, target_threshold: float = 0.5) -> tuple[int, int]:
    """``(num_bands, rows_per_band)`` the reference selects for ``num_perm`` signature bits (br.py:325-395)."""
    table = PRECOMPUTED_CONFIGS.get(num_perm)
    if table:
        nearest = min(table, key=lambda t: abs(t - target_threshold))
        if abs(nearest - target_threshold) <= 0.05:
            return table[nearest]
    found = find_optimal_br(num_perm, target_threshold)
    if found:
        return found
    b = int(np.sqrt(num_perm))          # largest divisor not above the square root
    while num_perm % b:
        b -= 1
    return b, num_perm // b


def print_config_analysis(num_perm: int, threshold: float = 0.5) -> None:
    """Print the selected shape, its midpoint, both error areas and four points of the detection curve
    (diagnostic helper the reference exports, br.py:398-460)."""
    b, r = get_optimal_config(num_perm, threshold)
    fp, fn = compute_false_rates(b, r, threshold)
    print("LSH Configuration Analysis")
    print("=" * 50)
    print(f"Number of permutations: {num_perm}")
    print(f"Target threshold: {threshold:.2f}")
    print(f"\nOptimal configuration:\n  Bands (b): {b}\n  Rows per band (r): {r}")
    print(f"\nPerformance metrics:\n  Actual threshold: {compute_lsh_threshold(b, r):.4f}")
    print(f"  False positive rate: {fp:.2%}\n  False negative rate: {fn:.2%}")
    print("\nDetection probabilities:")
    for sim in (0.3, 0.5, 0.7, 0.9):
        print(f"  Similarity {sim:.1f}: {compute_collision_probability(sim, b, r):.2%} chance of detection")
