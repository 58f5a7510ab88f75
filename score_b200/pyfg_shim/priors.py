"""Pose / landmark prior records (PyFactorGraph ``priors`` surface).

The reference cost reads only landmark priors
(/root/reference/score/utils/gurobi_utils.py:433-446); pose priors are carried
but ignored.  ``PosePrior2D`` is pickled by the upstream
slotted class as a plain tuple, hence the ``__setstate__`` below.
"""
from typing import Optional, Sequence

import numpy as np


class _TupleStateMixin:
    _FIELDS: tuple = ()

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            # (None, slots-dict) form
            if state[0]:
                self.__dict__.update(state[0])
            self.__dict__.update(state[1])
        else:
            for key, val in zip(self._FIELDS, state):
                self.__dict__[key] = val


class PosePrior2D(_TupleStateMixin):
    _FIELDS = (
        "name",
        "position",
        "theta",
        "translation_precision",
        "rotation_precision",
        "timestamp",
    )

    def __init__(
        self,
        name: str,
        position: Sequence[float],
        theta: float,
        translation_precision: float,
        rotation_precision: float,
        timestamp: Optional[float] = None,
    ):
        self.name = name
        self.position = tuple(position)
        self.theta = float(theta)
        self.translation_precision = float(translation_precision)
        self.rotation_precision = float(rotation_precision)
        self.timestamp = timestamp

    @property
    def translation_vector(self) -> np.ndarray:
        return np.asarray(self.position, dtype=np.float64)

    @property
    def rotation_matrix(self) -> np.ndarray:
        c, s = np.cos(self.theta), np.sin(self.theta)
        return np.array([[c, -s], [s, c]], dtype=np.float64)


class PosePrior3D(_TupleStateMixin):
    _FIELDS = (
        "name",
        "position",
        "rotation",
        "translation_precision",
        "rotation_precision",
        "timestamp",
    )

    def __init__(
        self,
        name: str,
        position: Sequence[float],
        rotation: np.ndarray,
        translation_precision: float,
        rotation_precision: float,
        timestamp: Optional[float] = None,
    ):
        self.name = name
        self.position = tuple(position)
        self.rotation = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
        self.translation_precision = float(translation_precision)
        self.rotation_precision = float(rotation_precision)
        self.timestamp = timestamp

    @property
    def translation_vector(self) -> np.ndarray:
        return np.asarray(self.position, dtype=np.float64)

    @property
    def rotation_matrix(self) -> np.ndarray:
        return self.rotation


class LandmarkPrior2D(_TupleStateMixin):
    _FIELDS = ("name", "position", "translation_precision", "timestamp")

    def __init__(
        self,
        name: str,
        position: Sequence[float],
        translation_precision: float,
        timestamp: Optional[float] = None,
    ):
        self.name = name
        self.position = tuple(float(v) for v in position)
        self.translation_precision = float(translation_precision)
        self.timestamp = timestamp

    @property
    def translation_vector(self) -> np.ndarray:
        return np.asarray(self.position, dtype=np.float64)


class LandmarkPrior3D(LandmarkPrior2D):
    pass
