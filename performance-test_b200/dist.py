"""Host-side plumbing for the multi-GPU bootstrap (one process per GPU).

The reference does this with MPI inside DOLFINx (IndexMap / Scatterer construction); here
torch.distributed carries the few bytes that have to be exchanged once, before the timed regions:
the NCCL unique id, or -- for the NCCL-free peer-memory path -- the CUDA IPC handles and the
owners' send lists. Nothing here is on the hot path.
"""
from __future__ import annotations

import numpy as np


def init_nccl(ctx, abi, dist, rank, world):
    uid = [abi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])


def connect_peers(ctx, P, dist, rank, world):
    """Exchange IPC handles and halo send lists, then ptb_peer_connect."""
    mine = dict(handles=ctx.peer_export(), nbr=np.array(P["nbr_ranks"]).tolist(),
                send_displ=np.array(P["send_displ"]).tolist(),
                local_indices=np.array(P["local_indices"], dtype=np.int32))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    src = []
    for r in np.array(P["nbr_ranks"]).tolist():
        other = everyone[r]
        j = other["nbr"].index(rank)  # my slot in the owner's neighbour list
        src.append(other["local_indices"][other["send_displ"][j]:other["send_displ"][j + 1]])
    src_index = np.concatenate(src).astype(np.int32) if src else np.zeros(0, np.int32)
    rd = np.array(P["recv_displ"])
    assert len(src_index) == (rd[-1] if len(rd) else 0), "halo lists of neighbours do not match"
    ctx.peer_connect(rank, world, b"".join(e["handles"] for e in everyone), src_index)
