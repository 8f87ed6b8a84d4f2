"""Host-side wiring of a column-strip pipeline with one process per GPU (the reference's decomposition:
src/mpi/mpi-vert-driver.cpp:35-36, src/mpi/mpi-vert.cpp:17-105, with MPI ranks replaced by torch.distributed ranks).

torch.distributed is only plumbing here: it carries the 64-byte CUDA IPC handles of the halo mailboxes and the strip
height every rank must share.  The boundary column itself never goes through it: the producing kernel stores it
straight into the consumer GPU's mailbox over NVLink (nw_plan_import_mailbox).
"""


def agree_rows_per_lane(dist, rank, choose):
    """Rank 0 decides the strip height (rows per lane), every rank uses it."""
    box = [choose() if rank == 0 else 0]
    dist.broadcast_object_list(box, src=0)
    return int(box[0])


def exchange_mailboxes(dist, plan, rank, world):
    """Every rank > 0 exports its halo mailbox; every rank < world-1 imports its right neighbour's.
    Returns the list of handles (index = owning rank; rank 0 has none)."""
    handles = [None] * world
    dist.all_gather_object(handles, plan.export_mailbox() if rank > 0 else b"")
    if rank + 1 < world:
        if len(handles[rank + 1]) != 64:
            raise RuntimeError(f"rank {rank + 1} published a {len(handles[rank + 1])}-byte mailbox handle")
        plan.import_mailbox(handles[rank + 1], rank + 1)
    return handles
