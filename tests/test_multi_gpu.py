"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the x-slab partition with the in-kernel
peer-memory halo exchange and all-reduce gives the single-GPU result (owned dofs, ghost dofs, cell-model states,
iteration counts) for every kernel variant (matrix-in-smem / resident / streaming; cg / pipecg; x0 zero / v_;
Godunov / Strang).  Run with:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    dict(dx=0.5, dt=0.05, nsteps=12, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8),
    dict(dx=0.5, dt=0.05, nsteps=12, theta=0.5, ksp="pipecg", x0_prev=True, rtol=1e-11, tol=1e-8),
    dict(dx=0.25, dt=0.05, nsteps=6, theta=1.0, ksp="pipecg", x0_prev=False, rtol=1e-11, tol=1e-8),
    dict(dx=0.125, dt=0.05, nsteps=4, theta=1.0, ksp="pipecg", x0_prev=False, rtol=1e-11, tol=1e-8),  # several rows per thread
    dict(dx=0.5, dt=0.05, nsteps=6, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8, env={"MONO_PDE_STREAM": "1"}),
    dict(dx=0.25, dt=0.05, nsteps=4, theta=1.0, ksp="pipecg", x0_prev=False, rtol=1e-11, tol=1e-8, env={"MONO_PDE_STREAM": "1"}),
    # streaming KSPCG in its other modes: dictionary through the L1 (no rings), SELL stream (no dictionary), x0 = v_ with the
    # rings, the tagged-exchange kernel; and a finer mesh with several ring tiles per CTA
    dict(dx=0.5, dt=0.05, nsteps=6, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8, env={"MONO_PDE_STREAM": "1", "MONO_PDE_NO_RING": "1"}),
    dict(dx=0.5, dt=0.05, nsteps=6, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8, env={"MONO_PDE_STREAM": "1", "MONO_PDE_DICT": "0"}),
    dict(dx=0.5, dt=0.05, nsteps=6, theta=0.5, ksp="cg", x0_prev=True, rtol=1e-12, tol=1e-8, env={"MONO_PDE_STREAM": "1"}),
    dict(dx=0.5, dt=0.05, nsteps=6, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8, env={"MONO_PDE_STREAM": "1", "MONO_PDE_TAGGED_STREAM": "1"}),
    dict(dx=0.125, dt=0.05, nsteps=4, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8, env={"MONO_PDE_STREAM": "1"}),
    dict(dx=0.5, dt=0.05, nsteps=12, theta=1.0, ksp="pipecg", pc="chebyshev", x0_prev=False, rtol=1e-11, tol=1e-8),
    dict(dx=0.25, dt=0.05, nsteps=6, theta=0.5, ksp="pipecg", pc="chebyshev", x0_prev=True, rtol=1e-11, tol=1e-8),
    dict(dx=0.0, lv=[3, 12, 32], dt=0.05, nsteps=12, theta=1.0, ksp="pipecg", x0_prev=False, rtol=1e-11, tol=1e-8),
    dict(dx=0.0, lv=[3, 10, 24], dt=0.05, nsteps=8, theta=1.0, ksp="cg", x0_prev=False, rtol=1e-12, tol=1e-8),
]


def _ngpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_run_matches_single_gpu(world):
    n = _ngpus()
    if n < world:
        pytest.skip(f"needs {world} GPUs, box has {n}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29511 + world), os.path.join(ROOT, "tests", "_mgpu_worker.py"), json.dumps(CASES)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [json.loads(l[5:]) for l in r.stdout.splitlines() if l.startswith("MGPU ")]
    bad = [l for l in lines if not l["ok"]]
    assert r.returncode == 0 and not bad and len(lines) == world * len(CASES), (bad, r.stdout[-3000:], r.stderr[-3000:])
