"""Multi-GPU (NCCL, one process per GPU) parity; needs >= 2 GPUs on the box, skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env", [{}, {"DOTSOCP_TCHUNKS": "1"}, {"DOTSOCP_TPUSH": "0"}, {"DOTSOCP_TSOLVE": "transpose"},
                                 {"DOTSOCP_TSOLVE": "transpose", "DOTSOCP_NO_IPC": "1"},
                                 {"DOTSOCP_TSOLVE": "transpose", "DOTSOCP_XCHG": "direct"}],
                         ids=["pipelined-thomas", "pipelined-1-chunk", "pipelined-nccl-handoff", "transpose-ipc-push", "transpose-nccl-sendrecv", "transpose-direct-stores"])
def test_nccl_time_slab_parity(gpu, env):
    if gpu < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    nproc = 4 if gpu >= 4 else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_parity.py")],
                         capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert out.stdout.count("dist parity ok") == 4


def test_nccl_parity_at_baseline_grids(gpu):
    """configs[3] (512x512x256, 12 checked iterations against the CPU oracle's golden) and the 3-level 256x256x128 solve through
    the multilevel driver on time slabs over NCCL, on all GPUs of the box (2, 4 or 8)."""
    if gpu < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    nproc = 8 if gpu >= 8 else 4 if gpu >= 4 else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_parity_big.py")],
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert out.stdout.count("dist parity ok") == 2
