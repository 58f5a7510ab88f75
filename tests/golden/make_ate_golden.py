"""Regenerate tests/golden/ate.json: the SE(d)-aligned absolute trajectory error of the oracle's tight optimum x*
(stored in the graph fixtures) against the ground truth the reference ships with its graphs, computed with the
oracle's Kabsch alignment (oracle/score_oracle.py::align_trajectory) — per graph and per robot chain.

    python tests/golden/make_ate_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import score_oracle as so  # noqa: E402
from score_b200.evaluate import chain_offsets, ground_truth_positions  # noqa: E402
from score_b200.graph_io import load_graph_npz  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for name in ["goats", "man4", "man1", "mc0_small"]:
    fg, extra = load_graph_npz(os.path.join(HERE, name + ".npz"))
    d = int(fg.dimension)
    blk = d * (d + 1)
    gt = ground_truth_positions(fg)
    est = np.asarray(extra["x_star"])[: len(gt) * blk].reshape(len(gt), d, d + 1)[:, :, d]
    off = chain_offsets(fg)
    rmse, R, t = so.align_trajectory(est, gt)
    out[name] = {
        "rmse": rmse,
        "R": R.tolist(),
        "t": t.tolist(),
        "rmse_unaligned": so.align_trajectory(est, gt, align=False)[0],
        "rmse_per_chain": [so.align_trajectory(est[a:b], gt[a:b])[0] for a, b in zip(off[:-1], off[1:])],
    }
    print(name, out[name]["rmse"], out[name]["rmse_per_chain"])
with open(os.path.join(HERE, "ate.json"), "w") as f:
    json.dump(out, f, indent=1)
