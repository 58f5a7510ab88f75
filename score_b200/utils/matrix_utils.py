"""``round_to_special_orthogonal`` on the GPU (score/utils/matrix_utils.py:59-79 in the reference)."""
import numpy as np

from ..solver import round_to_special_orthogonal_batch


def _check_square(mat: np.ndarray) -> None:
    assert mat.shape[0] == mat.shape[1], "matrix must be square"


def _check_rotation_matrix(R: np.ndarray, assert_test: bool = False) -> None:
    """matrix_utils.py:293-318 — orthogonality and determinant within 1e-3."""
    d = R.shape[0]
    if not np.allclose(R @ R.T, np.eye(d), rtol=1e-3, atol=1e-3):
        if assert_test:
            raise ValueError(f"R is not orthogonal {R @ R.T}")
    if not abs(np.linalg.det(R) - 1) < 1e-3:
        if assert_test:
            raise ValueError(f"R det incorrect {np.linalg.det(R)}")


def get_matrix_determinant(mat: np.ndarray) -> float:
    _check_square(mat)
    return float(np.linalg.det(mat))


def round_to_special_orthogonal(mat: np.ndarray, device: int = 0) -> np.ndarray:
    mat = np.asarray(mat, dtype=np.float64)
    _check_square(mat)
    try:
        R = round_to_special_orthogonal_batch(mat[None], device=device)[0]
        _check_rotation_matrix(R, assert_test=True)
    except ValueError:
        raise ValueError(f"Could not round matrix to special orthogonal form: {mat}")
    return R
