"""In-tree build of libscore_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libscore_b200.so")
SOURCES = ["api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + ["../../include/score_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libscore_b200.so if it is missing or older than its sources.  Safe to call from several processes
    at once (one rank per GPU): an exclusive file lock serialises them and the output is renamed into place."""
    import fcntl

    if not force and not _stale():
        return OUT
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():  # another process built it while we waited
                return OUT
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = OUT + f".tmp{os.getpid()}"
            cmd = ([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
                   + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp, "-ldl"])
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libscore_b200.so")
            os.replace(tmp, OUT)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
