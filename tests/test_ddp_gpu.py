"""Multi-GPU path on real hardware (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`):
two NCCL ranks, the CUDA-graph-replayed step on two shards == the eager single-process step on the concatenated batch
(gradients in the flat buckets, post-step weights of G and D, mean losses), identical weights on both ranks, and a
clean process-group teardown with the captured graph released first."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_graph_step_equals_full_batch_step(tmp_path):
    env = dict(os.environ, DDP_TEST_OUT=str(tmp_path))
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "ddp_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    reps = [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in range(2)]
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "ddp_gpu_2rank.json").write_text(json.dumps(reps, indent=1))
    for rep in reps:
        # same kernels on different batch sizes: fp32 summation order (+ TF32 tile order) only
        assert rep["gG"] < 2e-3 and rep["gD"] < 2e-3, rep
        assert rep["wG"] < 1e-4 and rep["wD"] < 1e-4, rep
        assert rep["rank_weight_divergence"] == 0.0, rep
        for k, v in rep.items():
            if k.startswith("loss."):
                assert v < 2e-3, (k, v)
