"""Alias of score_b200.utils.gurobi_utils under the reference's module path."""
from score_b200.utils.gurobi_utils import *  # noqa: F401,F403
from score_b200.utils.gurobi_utils import QCQP_RELAXATION, SOCP_RELAXATION, ACCEPTABLE_RELAXATIONS  # noqa: F401
