"""Alias of score_b200.solve_score under the reference's module path (score/solve_score.py)."""
from score_b200.solve_score import (  # noqa: F401
    _check_factor_graph,
    solve_problem_with_intermediate_iterates,
    solve_score,
    solve_and_refine,
    solve_score_batch,
)
