"""Minimal stand-in for the un-vendored ``py_factor_graph`` dependency.

The reference imports its input/output data model from PyFactorGraph
(/root/reference/score/solve_score.py:17-21,
/root/reference/score/utils/gurobi_utils.py:6-23), which is neither vendored nor
pinned.  This shim provides exactly the surface the hot path touches (SURVEY.md
App. D) under the same module paths so the shipped pickles unpickle.  If a real
``py_factor_graph`` is importable it always wins; ``install()`` is then a no-op.
"""
import importlib
import importlib.util
import sys

_SUBMODULES = [
    "factor_graph",
    "variables",
    "measurements",
    "priors",
    "utils",
    "utils.solver_utils",
    "utils.matrix_utils",
    "parsing",
    "parsing.parse_pickle_file",
]


def install() -> bool:
    """Alias this package as ``py_factor_graph`` when the real one is absent.

    Returns True when the shim is (now) the provider.
    """
    if "py_factor_graph" in sys.modules:
        return getattr(sys.modules["py_factor_graph"], "__score_b200_shim__", False)
    try:
        if importlib.util.find_spec("py_factor_graph") is not None:
            return False
    except (ImportError, ValueError):
        pass
    me = sys.modules[__name__]
    me.__score_b200_shim__ = True
    sys.modules["py_factor_graph"] = me
    for sub in _SUBMODULES:
        mod = importlib.import_module(f"{__name__}.{sub}")
        sys.modules[f"py_factor_graph.{sub}"] = mod
        # classes pickle under the real package's paths, so files written here load with PyFactorGraph
        for obj in vars(mod).values():
            if isinstance(obj, type) and obj.__module__ == mod.__name__:
                obj.__module__ = f"py_factor_graph.{sub}"
    return True
