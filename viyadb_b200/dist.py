"""One process per GPU: plumbing between torch.distributed and the C ABI's communicator.

The data path has exactly one exchange step — the merge of the per-GPU partial group tables inside
vgpu_query_agg (NCCL, see csrc/vgpu.cu: nccl_merge_*). Everything here is host plumbing: which
segments a rank owns (segment i -> GPU i mod G, whole segments only) and how the
ncclUniqueId reaches every rank.
"""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def segments_for_rank(nsegments, rank, world_size):
    """Global segment indices owned by `rank`: round-robin, whole segments."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, nsegments, world_size))


def local_to_global(local_idx, rank, world_size):
    return local_idx * world_size + rank


def share_unique_id(make_id, rank, dist):
    """Rank 0 creates the 128-byte ncclUniqueId (vgpu_comm_unique_id); torch.distributed (any backend)
    hands it to everybody."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad ncclUniqueId")
    return bytes(uid)


def init_database_comm(db, dist):
    """Attach `db` (already created on this rank's GPU) to the job-wide NCCL communicator."""
    from .db import Database
    rank, world_size, _ = world()
    if world_size == 1:
        return
    uid = share_unique_id(Database.comm_unique_id, rank, dist)
    db.init_comm(rank, world_size, uid)
