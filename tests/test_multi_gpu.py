"""Multi-GPU parity (one process per GPU over torch.distributed): the NVLink peer-memory training exchange, the NCCL
path and entity-sharded evaluation against the oracle.  Needs >= 2 B200s (gpurun --gpus 2); skipped on one GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run_worker(n, tag, **env):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):                  # keep the workers' full output next to the other logs of the run
        with open(os.path.join(out_dir, f"multi_gpu_worker_{tag}.log"), "w") as f:
            f.write(res.stdout + "\n==== stderr ====\n" + res.stderr)
    if not (res.returncode == 0 and res.stdout.strip().endswith("ok")):
        sys.stderr.write(res.stdout[-3000:] + "\n" + res.stderr[-8000:])
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), "multi-GPU worker failed (output on stderr)"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("backend", ["auto", "ipc"])     # auto = symmetric memory (+ NVSwitch multicast) when available
def test_peer_exchange_and_sharded_eval_against_oracle(backend):
    run_worker(min(torch.cuda.device_count(), 8), backend, KGE_PEER_BACKEND=backend)


def _gpu0_shareable():
    """Two processes can hold contexts on GPU 0 only in the default compute mode."""
    try:
        import pynvml
        pynvml.nvmlInit()
        mode = pynvml.nvmlDeviceGetComputeMode(pynvml.nvmlDeviceGetHandleByIndex(0))
        return mode == pynvml.NVML_COMPUTEMODE_DEFAULT
    except Exception:
        return True


@pytest.mark.skipif(torch.cuda.is_available() and not _gpu0_shareable(), reason="GPU 0 is in an exclusive compute mode")
def test_two_ranks_sharing_one_gpu_against_oracle():
    """The multi-GPU train paths (entity-sharded optimizer, dense peer exchange, all-reduce) and entity-sharded evaluation
    with TWO RANKS ON ONE GPU: runs on the 1-GPU test box too.  Peer memory is cudaIpc between two processes on the same
    device, the cross-GPU barriers spin while the contexts time-slice (bounded wait), torch.distributed is gloo."""
    run_worker(2, "same_gpu", KGE_TEST_SAME_GPU="1", KGE_PEER_BACKEND="ipc", KGE_PEER_TIMEOUT_S="30")
