"""Multi-GPU parity: 2 (and 4, if present) ranks, one per GPU, NCCL halo + all-reduce, against the CPU
oracle.  Skipped on boxes with a single GPU (the host-side plans are covered on CPU by
tests/test_partition.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    from oasisx_b200 import _lib

    return _lib.load_library().b2_device_count()


@pytest.mark.parametrize("nranks,mode", [(2, "lu"), (2, "krylov"), (2, "mg"), (2, "bench"), (2, "pbc"), (4, "krylov")])
def test_multirank_matches_oracle(nranks, mode):
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    port = 29700 + nranks * 10 + {"lu": 0, "krylov": 1, "mg": 2, "bench": 3, "pbc": 4}[mode]
    steps = "7" if mode == "bench" else "3"  # long enough for the three-deep solution histories to be in use
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "mr_worker.py"), "8", steps, mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("MR_OK") == nranks
