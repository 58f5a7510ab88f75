"""Single large 3D graph (or a scaled-down version) on one GPU.

    python scripts/large3d.py robots steps landmarks ranges [grid] [key=value solver params ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from score_b200 import build, generators
from score_b200.lowering import lower_grid3d_arrays

R, S, L, K = (int(a) for a in sys.argv[1:5])
rest = sys.argv[5:]
grid = int(rest.pop(0)) if rest and "=" not in rest[0] else 100
kw = {}
for a in rest:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
t0 = time.time()
arr = generators.grid_3d_arrays(generators.MC_BASE_SEED, n_robots=R, n_steps=S, grid=grid, n_landmarks=L, n_ranges=K)
prob = lower_grid3d_arrays(arr)
print(f"generated in {time.time() - t0:.1f}s: P={prob.P} L={prob.L} E={prob.E} K={prob.K}", flush=True)
build.build()
from score_b200.solver import ScoreSolver

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
repeat = kw.pop("repeat", 1)
with ScoreSolver(prob, device=local_rank) as s:
    if world > 1:
        box = [ScoreSolver.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        s.comm_init(world, rank, box[0])
    for _ in range(repeat):
        t0 = time.time()
        st = s.solve(**kw)
    rec = st.instances[0]
    if world > 1:
        poses = s.solution()[0]
        chk = torch.tensor([float(poses.sum()), float(rec["objective"])], device="cuda", dtype=torch.float64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"ranks agree bitwise: {bool((lo == hi).all().item())}  world={world}", flush=True)
    if rank == 0 and kw.get("profile_cycles"):
        from score_b200.solver import KERNEL_NAMES
        tot = st.kernel_ms.sum()
        for k, ms, c in zip(KERNEL_NAMES, st.kernel_ms, st.kernel_count):
            print(f"  {k:16s} {ms:9.1f} ms {100 * ms / max(tot, 1e-9):5.1f}%  launches {int(c):6d}  avg {1e3 * ms / max(1, c):8.1f} us")
    if rank == 0:
      print(f"solved={rec['solved']} kkt={rec['rel_kkt']:.3e} f={rec['objective']:.6f} newton={rec['newton_iters']} "
          f"cg={rec['cg_iters']} lsfail={rec['ls_failures']} ticks={st.ticks} cycles={st.cycles} "
          f"asm={st.assemble_ms:.1f}ms setup={st.setup_ms:.1f}ms solve={st.solve_ms:.1f}ms wall={time.time() - t0:.1f}s "
          f"nnz={st.nnz_reduced} rows={st.rows} cols={st.cols} GB/s={st.algorithmic_bytes / st.solve_ms / 1e6:.0f}", flush=True)
if world > 1:
    dist.destroy_process_group()
