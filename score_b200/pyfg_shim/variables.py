"""Pose / landmark variable records (PyFactorGraph ``variables`` surface).

Field names follow the pickles shipped with the reference:
2D poses carry ``name, true_position, true_theta, timestamp``.
"""
from typing import Optional, Sequence, Tuple, Union

import numpy as np


def _rot2(theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s], [s, c]], dtype=np.float64)


class PoseVariable2D:
    def __init__(
        self,
        name: str,
        true_position: Tuple[float, float],
        true_theta: float,
        timestamp: Optional[float] = None,
    ):
        self.name = name
        self.true_position = tuple(float(v) for v in true_position)
        self.true_theta = float(true_theta)
        self.timestamp = timestamp

    @property
    def rotation_matrix(self) -> np.ndarray:
        return _rot2(self.true_theta)

    @property
    def position_vector(self) -> np.ndarray:
        return np.asarray(self.true_position, dtype=np.float64)

    @property
    def transformation_matrix(self) -> np.ndarray:
        T = np.eye(3)
        T[:2, :2] = self.rotation_matrix
        T[:2, 2] = self.position_vector
        return T

    def __repr__(self):
        return f"PoseVariable2D({self.name!r}, {self.true_position}, {self.true_theta})"


class PoseVariable3D:
    def __init__(
        self,
        name: str,
        true_position: Sequence[float],
        true_rotation: np.ndarray,
        timestamp: Optional[float] = None,
    ):
        self.name = name
        self.true_position = tuple(float(v) for v in true_position)
        self.true_rotation = np.asarray(true_rotation, dtype=np.float64).reshape(3, 3)
        self.timestamp = timestamp

    @property
    def rotation_matrix(self) -> np.ndarray:
        return self.true_rotation

    @property
    def position_vector(self) -> np.ndarray:
        return np.asarray(self.true_position, dtype=np.float64)

    @property
    def transformation_matrix(self) -> np.ndarray:
        T = np.eye(4)
        T[:3, :3] = self.true_rotation
        T[:3, 3] = self.position_vector
        return T


class LandmarkVariable2D:
    def __init__(self, name: str, true_position: Union[Sequence[float], np.ndarray]):
        self.name = name
        self.true_position = tuple(float(v) for v in true_position)

    @property
    def position_vector(self) -> np.ndarray:
        return np.asarray(self.true_position, dtype=np.float64)


class LandmarkVariable3D(LandmarkVariable2D):
    pass


POSE_VARIABLE_TYPES = Union[PoseVariable2D, PoseVariable3D]
LANDMARK_VARIABLE_TYPES = Union[LandmarkVariable2D, LandmarkVariable3D]
