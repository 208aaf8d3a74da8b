"""Multi-GPU parity: 2, 4 and 8 ranks, one per GPU, against the CPU oracle / CPU port; peer-memory collectives and the
NCCL fallback.  Every rank builds only its slab of the mesh (oasisx_b200/slab.py); one case keeps the replicated mesh
+ partition route covered.  Skipped on boxes with fewer GPUs (the host-side plans are covered on CPU by
tests/test_partition.py and tests/test_slab.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    from oasisx_b200 import _lib

    return _lib.load_library().b2_device_count()


CASES = [
    # ranks, mode, mesh, steps, B200_PEER
    (2, "lu", 8, 3, "1"), (2, "krylov", 8, 3, "1"), (2, "mg", 8, 3, "1"), (2, "bench", 8, 7, "1"), (2, "pbc", 8, 3, "1"),
    (2, "bench", 8, 7, "0"),   # the NCCL path (peer memory switched off) must stay correct: it is the fallback
    (2, "bench", 16, 5, "1"),  # >= 16^3 against the CPU port, rotated field (three live components)
    (4, "krylov", 8, 3, "1"), (4, "bench", 16, 5, "1"), (4, "bench", 16, 5, "0"),
    (8, "bench", 16, 5, "1"),
]


@pytest.mark.parametrize("nranks,mode,mesh,steps,peer", CASES)
def test_multirank_matches_oracle(nranks, mode, mesh, steps, peer):
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    port = 29700 + nranks * 20 + CASES.index((nranks, mode, mesh, steps, peer))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "mr_worker.py"), str(mesh), str(steps), mode]
    env = dict(os.environ, B200_PEER=peer)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("MR_OK") == nranks
    print(res.stdout[-1500:])


def test_multirank_replicated_mesh_route():
    """B200_GLOBAL_MESH=1: every rank builds the whole mesh and oasisx_b200.partition cuts it (the route the slab-local
    provider replaced as the default; still what a user-constructed Mesh object takes)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29799", os.path.join(HERE, "mr_worker.py"), "8", "3", "krylov"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, B200_PEER="1", B200_GLOBAL_MESH="1"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("MR_OK") == 2
