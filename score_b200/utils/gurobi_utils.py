"""Names the reference exposes from ``score.utils.gurobi_utils`` that callers import
(/root/reference/score/utils/gurobi_utils.py:26-34; examples/solve_goats_example_score.py:22)."""
from ..lowering import (  # noqa: F401
    ACCEPTABLE_RELAXATIONS,
    QCQP_RELAXATION,
    SOCP_RELAXATION,
    check_valid_relaxation as _check_valid_relaxation,
)

RANDOM_INIT = "random"
ZERO_INIT = "zero"
ODOM_INIT = "odom"
GT_INIT = "gt"
ACCEPTABLE_INIT = [RANDOM_INIT, ZERO_INIT, ODOM_INIT, GT_INIT]
