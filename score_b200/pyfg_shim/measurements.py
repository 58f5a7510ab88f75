"""Relative-pose and range measurement records (PyFactorGraph ``measurements``).

Only the attributes the reference hot path reads are provided
(/root/reference/score/utils/gurobi_utils.py:288,398-401,463-469,487-500,514-523):
``base_pose, to_pose, translation_vector, rotation_matrix,
translation_precision, rotation_precision`` for pose measurements and
``first_key, second_key, dist, precision`` for ranges.
"""
from typing import Optional, Tuple, Union

import numpy as np


class PoseMeasurement2D:
    def __init__(
        self,
        base_pose: str,
        to_pose: str,
        x: float,
        y: float,
        theta: float,
        translation_precision: float,
        rotation_precision: float,
        timestamp: Optional[float] = None,
    ):
        self.base_pose = base_pose
        self.to_pose = to_pose
        self.x = float(x)
        self.y = float(y)
        self.theta = float(theta)
        self.translation_precision = float(translation_precision)
        self.rotation_precision = float(rotation_precision)
        self.timestamp = timestamp

    @property
    def rotation_matrix(self) -> np.ndarray:
        c, s = np.cos(self.theta), np.sin(self.theta)
        return np.array([[c, -s], [s, c]], dtype=np.float64)

    @property
    def translation_vector(self) -> np.ndarray:
        return np.array([self.x, self.y], dtype=np.float64)

    @property
    def transformation_matrix(self) -> np.ndarray:
        T = np.eye(3)
        T[:2, :2] = self.rotation_matrix
        T[:2, 2] = self.translation_vector
        return T


class PoseMeasurement3D:
    def __init__(
        self,
        base_pose: str,
        to_pose: str,
        translation: np.ndarray,
        rotation: np.ndarray,
        translation_precision: float,
        rotation_precision: float,
        timestamp: Optional[float] = None,
    ):
        self.base_pose = base_pose
        self.to_pose = to_pose
        self.translation = np.asarray(translation, dtype=np.float64).reshape(3)
        self.rotation = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
        self.translation_precision = float(translation_precision)
        self.rotation_precision = float(rotation_precision)
        self.timestamp = timestamp

    @property
    def rotation_matrix(self) -> np.ndarray:
        return self.rotation

    @property
    def translation_vector(self) -> np.ndarray:
        return self.translation

    @property
    def transformation_matrix(self) -> np.ndarray:
        T = np.eye(4)
        T[:3, :3] = self.rotation
        T[:3, 3] = self.translation
        return T


class FGRangeMeasurement:
    def __init__(
        self,
        association: Tuple[str, str],
        dist: float,
        stddev: float,
        timestamp: Optional[float] = None,
    ):
        self.association = (association[0], association[1])
        self.dist = float(dist)
        self.stddev = float(stddev)
        self.timestamp = timestamp

    @property
    def first_key(self) -> str:
        return self.association[0]

    @property
    def second_key(self) -> str:
        return self.association[1]

    @property
    def pose_key(self) -> str:
        return self.association[0]

    @property
    def landmark_key(self) -> str:
        return self.association[1]

    @property
    def variance(self) -> float:
        return self.stddev**2

    @property
    def weight(self) -> float:
        return 1.0 / (self.stddev**2)

    @property
    def precision(self) -> float:
        return 1.0 / (self.stddev**2)


POSE_MEASUREMENT_TYPES = Union[PoseMeasurement2D, PoseMeasurement3D]
