"""The real multi-process path: torchrun, one rank per GPU, the NCCL communicator of libibk.so (csrc/ibk_comm.cu).
Needs two GPUs: marked `mgpu`, NOT `gpu` -- a one-GPU lease runs the same plan / pack / message / unpack / migration code
through the loopback communicator instead (tests/test_gpu_parity.py::test_ranks_as_contexts_of_one_process), so `-m gpu`
has nothing to skip.  Run with:  python -m pytest tests -m mgpu  on a 2-GPU lease."""
import os

import pytest

pytestmark = pytest.mark.mgpu


def test_multi_gpu_parity_two_ranks():
    """Patch-partitioned level on 2 GPUs with the NCCL halo exchange (tests/mgpu_worker.py) against the
    oracle; needs >= 2 visible GPUs (the driver's single-GPU tier skips it)."""
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (-m mgpu on a 2-GPU lease); one GPU: test_ranks_as_contexts_of_one_process")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for kernel in ("IB_4", "IB_6"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "mgpu_worker.py"), kernel],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "MGPU_PARITY" in r.stdout
    # the same with the markers starting on arbitrary ranks: marker migration first (halo.MarkerMigration)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29518", os.path.join(root, "tests", "mgpu_worker.py"), "IB_4", "migrate"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MGPU_PARITY" in r.stdout and "migrate=1" in r.stdout
    # the exchange overlapped with the interior tiles (ibk_*_part)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29519", os.path.join(root, "tests", "mgpu_worker.py"), "IB_4", "overlap"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MGPU_PARITY" in r.stdout and "overlap=1" in r.stdout
    # the pipelined sequence of bench.py (each exchange behind the other operation's kernel)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29520", os.path.join(root, "tests", "mgpu_worker.py"), "IB_4", "pipelined"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MGPU_PARITY" in r.stdout and "pipelined=1" in r.stdout
