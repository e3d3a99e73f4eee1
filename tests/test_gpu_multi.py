"""Multi-GPU reassembly on real hardware: spawns 2 ranks (one per GPU) running tools/check_reassembly.py, which compares, on
every rank, every slot of the gathered buffer of every route (pipelined push narrow / wide, stores fused into the kernels over
P2P / packed / NVSwitch multicast) bit for bit against the NCCL all-gather and the C oracle.  Skipped below 2 devices."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spawn(world, extra):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "check_reassembly.py"), "--time", "0"] + extra
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ALL ROUTES OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MISMATCH" not in out.stdout


def test_two_ranks_every_route_n16():
    _spawn(2, ["--records", "100000", "--cands", "16"])


def test_two_ranks_every_route_n40_wide_words():
    _spawn(2, ["--records", "40000", "--cands", "40"])
