"""Time-to-solve of the single-graph configs (bench.py's per_config table) on their own: python scripts/per_config.py [--config5]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import types
import bench
from score_b200 import build
build.build()
args = types.SimpleNamespace(no_config5="--config5" not in sys.argv)
t = bench.run_per_config(bench.load_per_config_inputs(args), 0)
for k, v in t.items():
    print(f"{k:40s} {v['time_to_solve_ms']:8.2f} ms  solve {v['solve_ms']:8.2f}  newton {v['newton']:3d} cg {v['cg']:5d} ticks {v['ticks']:5d} us/tick {v['us_per_tick']:.1f} kkt {v['rel_kkt']:.1e}")
