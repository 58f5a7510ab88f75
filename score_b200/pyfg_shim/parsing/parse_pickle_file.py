"""``parse_pickle_file`` — load a pickled ``FactorGraphData``.

Mirrors the call the reference example makes
(/root/reference/examples/solve_goats_example_score.py:18,40).
"""
import pickle

from ..factor_graph import FactorGraphData


def parse_pickle_file(filepath: str) -> FactorGraphData:
    with open(filepath, "rb") as f:
        data = pickle.load(f)
    if not isinstance(data, FactorGraphData):
        raise ValueError(f"{filepath} does not hold a FactorGraphData (got {type(data)})")
    return data
