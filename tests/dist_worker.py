"""Worker of tests/test_gpu_dist.py: run under torch.distributed.run, one rank per GPU.

Runs the parity cases of tests/dist_cases.py (LSpace / LTRSpace, x-slabs / box partition with dofs shared by
4 or 8 ranks) for the transport OB200_P2P selects and prints one JSON line per case on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oofem_b200 import capi  # noqa: E402
import dist_cases  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    transport = "p2p" if os.environ.get("OB200_P2P", "1") != "0" else "nccl"
    ctx = capi.Context(local)
    res = dist_cases.run_all(ctx, dev, [(e, k, transport) for e in ("lspace", "ltrspace") for k in ("slab", "box")])
    if rank == 0:
        for r in res:
            print(json.dumps(r), flush=True)
    ok = all(r["ok"] for r in res)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
