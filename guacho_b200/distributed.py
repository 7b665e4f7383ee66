"""One process per GPU: wires a :class:`Block` into a NCCL communicator using
torch.distributed only as the bootstrap plumbing (it plays the role of ``mpi_init`` +
``MPI_Bcast`` in the reference's Fortran host; the halo exchange and the CFL all-reduce
run inside libguacho_gx.so on the solver's own stream).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .config import Params
from .decomp import choose_decomposition, coords_of


def env_rank() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: Optional[str] = None):
    """torch.distributed bootstrap (reads RANK/WORLD_SIZE/MASTER_* from the env)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def broadcast_bytes(payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    """Broadcast a small byte string from `src` (MPI_Bcast of the NCCL unique id)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload[:nbytes].ljust(nbytes, b"\0")), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def make_rank_block(p: Params, rank: int, world: int, local_rank: int, nb=None):
    """Create this rank's Block (device = local_rank) and attach the communicator."""
    from .solver import Block
    if nb is None:
        nb = choose_decomposition(p, world)
    pr = p.replace(MPI_NBX=nb[0], MPI_NBY=nb[1], MPI_NBZ=nb[2], device=local_rank)
    coords = coords_of(rank, nb)
    blk = Block(pr, coords)
    if world > 1:
        uid = Block.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 128, 0)
        blk.comm_attach(uid, rank, world)
    return blk
