"""Node-partitioned step on real GPUs (BASELINE config 4): spawns one process per GPU with
torchrun and checks, on the small case, that the gathered partitioned output equals the
single-GPU step bit for bit.  Needs >= 2 GPUs (the single-GPU boxes skip it; host logic is covered
on the CPU by tests/test_partitioned_host.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("masters", [False, True], ids=["image_only_latents", "fp32_masters"])
@pytest.mark.parametrize("workload", ["sample_2deg_13lvl"])
def test_partitioned_step_is_bit_identical_on_two_gpus(workload, masters):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs at least 2 GPUs")
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(REPO, "bench.py"),
         "--gpus", "2", "--mode", "partitioned", "--check", "--steps", "2", "--warmup", "3",
         "--workload", workload] + (["--masters"] if masters else [])
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
  assert out.returncode == 0, out.stderr[-2000:]
  line = json.loads(out.stdout.strip().splitlines()[-1])
  assert line["config"]["mode"] == "partitioned" and line["n_gpus"] == 2
  assert line["config"]["image_residual"] == (not masters)
  assert line["check"]["bitwise_equal"], line["check"]
  for key in ("e2e", "gpu_launches", "roofline"):
    assert key in line
