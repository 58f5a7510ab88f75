"""``FactorGraphData`` — the input container of ``solve_score``.

Attribute names match the pickled ``__dict__`` of the upstream class
 so the shipped graphs load by plain ``BUILD``.
The properties are the ones the reference path reads
(/root/reference/score/solve_score.py:28-32,
/root/reference/score/utils/gurobi_utils.py:181,199,236-258,288,398-401,425-427,438-444).
"""
from typing import Dict, List, Optional, Set

from .measurements import FGRangeMeasurement
from .variables import (
    LandmarkVariable2D,
    LandmarkVariable3D,
    PoseVariable2D,
    PoseVariable3D,
)

# PyFactorGraph reserves "L" for landmarks; robot chains use the other letters.
_ROBOT_CHARS = "ABCDEFGHIJKMNOPQRSTUVWXYZ"


def get_robot_char_from_number(idx: int) -> str:
    if idx < len(_ROBOT_CHARS):
        return _ROBOT_CHARS[idx]
    # beyond 25 chains: two-letter prefix (still never starts with "L")
    return _ROBOT_CHARS[idx // len(_ROBOT_CHARS) - 1] + _ROBOT_CHARS[idx % len(_ROBOT_CHARS)].lower()


class FactorGraphData:
    def __init__(self, dimension: int):
        if dimension not in (2, 3):
            raise ValueError(f"Value {dimension} is not 2 or 3")
        self.dimension = dimension
        self.pose_variables: List[List] = []
        self.landmark_variables: List = []
        self.existing_pose_variables: Set[str] = set()
        self.existing_landmark_variables: Set[str] = set()
        self.odom_measurements: List[List] = []
        self.loop_closure_measurements: List = []
        self.ambiguous_loop_closure_measurements: List = []
        self.range_measurements: List[FGRangeMeasurement] = []
        self.ambiguous_range_measurements: List = []
        self.pose_priors: List = []
        self.landmark_priors: List = []
        self.x_min: Optional[float] = None
        self.x_max: Optional[float] = None
        self.y_min: Optional[float] = None
        self.y_max: Optional[float] = None
        self.z_min: Optional[float] = None
        self.z_max: Optional[float] = None
        self.max_measure_weight: Optional[float] = None
        self.min_measure_weight: Optional[float] = None

    # ---- construction helpers (used by the synthetic generators) -------------
    def add_pose_variable(self, pose_var, chain: Optional[int] = None) -> None:
        if chain is None:
            chain = _ROBOT_CHARS.index(pose_var.name[0])
        while len(self.pose_variables) <= chain:
            self.pose_variables.append([])
        self.pose_variables[chain].append(pose_var)
        self.existing_pose_variables.add(pose_var.name)

    def add_landmark_variable(self, landmark_var) -> None:
        self.landmark_variables.append(landmark_var)
        self.existing_landmark_variables.add(landmark_var.name)

    def add_odom_measurement(self, robot_idx: int, odom_meas) -> None:
        while len(self.odom_measurements) <= robot_idx:
            self.odom_measurements.append([])
        self.odom_measurements[robot_idx].append(odom_meas)

    def add_loop_closure(self, loop_closure) -> None:
        self.loop_closure_measurements.append(loop_closure)

    def add_range_measurement(self, range_meas: FGRangeMeasurement) -> None:
        self.range_measurements.append(range_meas)

    def add_pose_prior(self, pose_prior) -> None:
        self.pose_priors.append(pose_prior)

    def add_landmark_prior(self, landmark_prior) -> None:
        self.landmark_priors.append(landmark_prior)

    # ---- read-side API -----------------------------------------------------
    @property
    def num_robots(self) -> int:
        return len(self.pose_variables)

    @property
    def num_poses(self) -> int:
        return sum(len(c) for c in self.pose_variables)

    @property
    def num_landmarks(self) -> int:
        return len(self.landmark_variables)

    @property
    def num_odom_measurements(self) -> int:
        return sum(len(c) for c in self.odom_measurements)

    @property
    def num_loop_closures(self) -> int:
        return len(self.loop_closure_measurements)

    @property
    def num_range_measurements(self) -> int:
        return len(self.range_measurements)

    @property
    def pose_variables_dict(self) -> Dict[str, object]:
        return {p.name: p for chain in self.pose_variables for p in chain}

    @property
    def landmark_variables_dict(self) -> Dict[str, object]:
        return {l.name: l for l in self.landmark_variables}

    @property
    def all_variable_names(self) -> List[str]:
        names = [p.name for chain in self.pose_variables for p in chain]
        names += [l.name for l in self.landmark_variables]
        return names

    @property
    def unconnected_variable_names(self) -> Set[str]:
        """Variables appearing in no odometry, loop-closure or range factor."""
        touched: Set[str] = set()
        for chain in self.odom_measurements:
            for m in chain:
                touched.add(m.base_pose)
                touched.add(m.to_pose)
        for m in self.loop_closure_measurements:
            touched.add(m.base_pose)
            touched.add(m.to_pose)
        for m in self.range_measurements:
            touched.add(m.association[0])
            touched.add(m.association[1])
        return set(self.all_variable_names) - touched

    def get_pose_chain_names(self) -> List[List[str]]:
        return [[p.name for p in chain] for chain in self.pose_variables]

    def pose_exists(self, name: str) -> bool:
        return name in self.existing_pose_variables

    def landmark_exists(self, name: str) -> bool:
        return name in self.existing_landmark_variables

    @property
    def pose_to_range_measures_dict(self) -> Dict[str, List[FGRangeMeasurement]]:
        out: Dict[str, List[FGRangeMeasurement]] = {}
        for m in self.range_measurements:
            out.setdefault(m.association[0], []).append(m)
        return out

    def __repr__(self):
        return (
            f"FactorGraphData(dim={self.dimension}, poses={self.num_poses}, "
            f"landmarks={self.num_landmarks}, odom={self.num_odom_measurements}, "
            f"loops={self.num_loop_closures}, ranges={self.num_range_measurements})"
        )


__all__ = [
    "FactorGraphData",
    "PoseVariable2D",
    "PoseVariable3D",
    "LandmarkVariable2D",
    "LandmarkVariable3D",
    "get_robot_char_from_number",
]
