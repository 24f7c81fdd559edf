"""Host-side handle of the multi-GPU communicator (ob200_comm, include/oofem_b200.h): one process
per GPU, NCCL for the shared-dof halo exchange and the dot-product all-reduce.  The 128-byte NCCL
unique id is created by rank 0 and handed round by the caller's own process group
(torch.distributed here, MPI_Bcast in OOFEM's parallel mode)."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import capi
from .capi import check, lib, ptr


class Comm:
    def __init__(self, ctx: capi.Context, rank: int, nranks: int, unique_id: bytes | None):
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        self.h = C.c_void_p()
        raw = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        check(lib().ob200_comm_create(ctx.h, nranks, rank, raw, C.byref(self.h)))

    @staticmethod
    def unique_id() -> bytes:
        raw = (C.c_char * 128)()
        check(lib().ob200_comm_unique_id(raw))
        return bytes(raw.raw)

    @classmethod
    def from_torch_distributed(cls, ctx: capi.Context, device=None):
        """Bootstrap over an initialised torch.distributed process group."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8).clone()
        if device is not None:
            buf = buf.to(device)
        dist.broadcast(buf, 0)
        return cls(ctx, rank, world, bytes(buf.cpu().numpy().tobytes()))

    def set_halo(self, neq: int, neigh_rank, neigh_offset, shared_eq, owned):
        """Arrays as produced by oofem_b200.partition.halo_arrays."""
        neigh_rank = np.ascontiguousarray(neigh_rank, dtype=np.int32)
        neigh_offset = np.ascontiguousarray(neigh_offset, dtype=np.int64)
        shared_eq = np.ascontiguousarray(shared_eq, dtype=np.int32)
        owned = np.ascontiguousarray(owned, dtype=np.uint8)
        check(lib().ob200_comm_set_halo(self.h, int(neq), int(neigh_rank.size), ptr(neigh_rank), ptr(neigh_offset),
                                        ptr(shared_eq), ptr(owned)))
        if self.nranks > 1 and os.environ.get("OB200_P2P", "1") != "0":
            self.enable_p2p(int(np.diff(neigh_offset).max()) if neigh_rank.size else 0)
        return self

    def enable_p2p(self, pair_count: int) -> bool:
        """Map the mailboxes of all ranks of the node (CUDA IPC over the initialised torch.distributed
        group) so that the halo sum and the CG reductions go through peer memory instead of NCCL.
        Collective: every rank must call it.  Returns False (NCCL stays in use) if any rank cannot map."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() != self.nranks:
            return False
        dev = torch.device("cuda", torch.cuda.current_device())
        cap = torch.tensor([pair_count], dtype=torch.int64, device=dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        handle = (C.c_char * 64)()
        ok = lib().ob200_comm_p2p_export(self.h, int(cap.item()), handle) == 0
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).clone().to(dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(self.nranks)]
        dist.all_gather(allh, mine)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return False
        raw = b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh)
        buf = (C.c_char * len(raw)).from_buffer_copy(raw)
        rc = lib().ob200_comm_p2p_open(self.h, buf)
        flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            # some rank could not map: nobody may use the mailboxes, the NCCL transport stays in use
            lib().ob200_comm_p2p_disable(self.h)
            if self.rank == 0:
                print("oofem_b200: peer-memory transport not available on every rank (%s); using NCCL"
                      % lib().ob200_last_error().decode(), file=sys.stderr)
            return False
        return True

    @property
    def p2p(self) -> bool:
        return bool(lib().ob200_comm_p2p_enabled(self.h))

    def exchange_add(self, y_dev):
        """y <- y + the neighbours' contributions on shared dofs (device vector, in place)."""
        check(lib().ob200_comm_exchange_add(self.h, ptr(y_dev)))
        return y_dev

    def close(self):
        if self.h:
            lib().ob200_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
