"""Launches the multi-GPU strip-partition check when the box has >= 2 GPUs (logs of the last runs at 2 / 4 / 8 GPUs
for every CG driver are kept under profiles/r2_dist_*.log)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("cg", ["graph", "fused", "persistent_fused"])
def test_strip_partition_matches_single_gpu(cg, world):
    import torch
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs >= {world} GPUs (run with gpurun --gpus {world})")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_strip_check.py"),
                          "--soak", "20"],
                         capture_output=True, text=True, timeout=900, env=dict(os.environ, SRPS_CG=cg))
    sys.stdout.write(res.stdout[-6000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0 and "DIST_OK" in res.stdout
