"""N > 1 correctness on hardware: bricks of ONE global seeded problem, ghost exchange through
the product's C ABI (mx_halo_exchange, NCCL and NVLink-P2P transports), fused density pass per
brick, compared with the CPU oracle on the undivided problem across the brick interfaces.
Also the C++ (non-Python) host driving two GPUs through the same ABI."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", [0, 1], ids=["nccl", "p2p"])
@pytest.mark.parametrize("world,ppm", [(2, 1), (2, 2), (4, 1), (8, 0)])
def test_bricks_match_undivided_problem(world, ppm, transport):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + 10 * world + transport
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_halo_worker.py"), str(transport), str(ppm)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("transport", [0, 1], ids=["nccl", "p2p"])
def test_cpp_host_two_gpus(transport):
    """tests/cpp/halo2.cpp: a C++ program (no Python, no torch) with one host thread per GPU."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "tests", "cpp", "halo2")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/halo2 not built (build() compiles it)")
    r = subprocess.run([exe, str(transport)], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "HALO2 OK" in r.stdout
