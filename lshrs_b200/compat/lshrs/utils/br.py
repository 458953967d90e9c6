"""``lshrs.utils.br`` -> ``lshrs_b200.utils.br`` (replaces reference lshrs/utils/br.py)."""

from lshrs_b200.utils.br import (  # noqa: F401
    PRECOMPUTED_CONFIGS,
    compute_collision_probability,
    compute_false_rates,
    compute_lsh_threshold,
    find_optimal_br,
    get_optimal_config,
    print_config_analysis,
)
