"""Small matrix helpers of the PyFactorGraph surface (host side, numpy).

``round_to_special_orthogonal`` is the function the reference extraction step
imports (/root/reference/score/utils/gurobi_utils.py:20-23; local twin at
/root/reference/score/utils/matrix_utils.py:59-79).  The GPU solver does its own
rounding in a CUDA kernel; this host version exists for users of the data model
and for tests.
"""
import numpy as np


def get_matrix_determinant(mat: np.ndarray) -> float:
    mat = np.asarray(mat)
    assert mat.shape[0] == mat.shape[1], "matrix must be square"
    return float(np.linalg.det(mat))


def get_rotation_matrix_from_theta(theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s], [s, c]], dtype=np.float64)


def get_theta_from_rotation_matrix(mat: np.ndarray) -> float:
    return float(np.arctan2(mat[1, 0], mat[0, 0]))


def check_rotation_matrix(R: np.ndarray, tol: float = 1e-3) -> None:
    """Same acceptance test as the reference (matrix_utils.py:293-318)."""
    d = R.shape[0]
    if not np.allclose(R @ R.T, np.eye(d), rtol=tol, atol=tol):
        raise ValueError(f"R is not orthogonal {R @ R.T}")
    if not abs(np.linalg.det(R) - 1.0) < tol:
        raise ValueError(f"R det incorrect {np.linalg.det(R)}")


def round_to_special_orthogonal(mat: np.ndarray) -> np.ndarray:
    mat = np.asarray(mat, dtype=np.float64)
    assert mat.shape[0] == mat.shape[1], "matrix must be square"
    dim = mat.shape[0]
    try:
        U, _, Vh = np.linalg.svd(mat)
        R = U @ Vh
        if np.linalg.det(R) < 0:
            R = U @ np.diag([1.0] * (dim - 1) + [-1.0]) @ Vh
        check_rotation_matrix(R)
    except ValueError:
        raise ValueError(f"Could not round matrix to special orthogonal form: {mat}")
    return R
