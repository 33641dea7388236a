"""N > 1 path.  CPU: world_size 2 and 4 over gloo (host logic: replicated deterministic setup, halo
plans used across real process boundaries).  GPU (needs >= 2 devices): one rank per GPU under
torchrun with CUDA-IPC peer memory — parity against the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, port, timeout, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    env.setdefault("OMP_NUM_THREADS", "2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_distributed_halo_and_setup(world):
    r = _torchrun(world, os.path.join(ROOT, "tests", "mp_worker_cpu.py"), 29511 + world, 300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MP_CPU_OK {world}" in r.stdout


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.timeout(700, method="thread")
@pytest.mark.skipif(_gpu_count() < 2, reason="needs >= 2 GPUs (one rank per GPU; ranks must not share a device)")
def test_one_rank_per_gpu_parity():
    world = 2 if _gpu_count() < 4 else 4
    r = _torchrun(world, os.path.join(ROOT, "scripts", "mp_gpu_check.py"), 29533, 680)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MP_GPU_OK" in r.stdout
