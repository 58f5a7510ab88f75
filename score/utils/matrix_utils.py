"""Alias of score_b200.utils.matrix_utils under the reference's module path."""
from score_b200.utils.matrix_utils import get_matrix_determinant, round_to_special_orthogonal  # noqa: F401
