"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): element partitions + NCCL halo
exchange + distributed PCG against the serial oracle.  See tests/dist_worker.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2])
def test_distributed_assembly_halo_and_pcg_vs_oracle(world, transport):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "dist_worker.py")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OB200_P2P="1" if transport == "p2p" else "0"))
    import re
    lines = [json.loads(m) for m in re.findall(r"\{[^{}]*\}", r.stdout)]      # ranks may share a line
    assert r.returncode == 0 and len(lines) == world and all(l["ok"] for l in lines), (r.stdout[-3000:], r.stderr[-3000:])
    assert all(l["p2p"] == (transport == "p2p") for l in lines), lines
