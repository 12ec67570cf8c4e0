"""Two ranks on NCCL (skipped on a box with one GPU): TrainerPredictor at world size 2 -- queries sharded, KG replicated,
one flat gradient all-reduce per step with DDP-mean semantics (src/trainer.py:52-60) -- ends with the same parameters
and the same valid MRR as ONE process that takes the same two batches per optimizer step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_training_equals_single_process(tmp_path):
    script = os.path.join(ROOT, "scripts", "check_ddp_equivalence.py")
    one, two = str(tmp_path / "single.pt"), str(tmp_path / "ddp.pt")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    subprocess.run([sys.executable, script, "single", one], check=True, timeout=600, env=env, cwd=ROOT)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29613", script, "ddp", two], check=True, timeout=600, env=env, cwd=ROOT)
    out = subprocess.run([sys.executable, script, "compare", one, two], check=True, timeout=120, env=env, cwd=ROOT,
                         capture_output=True, text=True).stdout
    assert "DDP equivalence OK" in out, out
