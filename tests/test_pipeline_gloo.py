"""CPU, world_size 2 (gloo): the host-side logic of the one-process-per-GPU column-strip pipeline -- partition,
strip-height agreement, mailbox handle exchange -- with the strips themselves computed by the oracle and the boundary
column carried by gloo send/recv (on the GPU box it travels through peer memory instead)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakePlan:
    """Stands in for nw.Plan where there is no GPU: same export/import calls, records what it was given."""

    def __init__(self, rank):
        self.rank, self.imported = rank, None

    def export_mailbox(self):
        return bytes([self.rank]) * 64

    def import_mailbox(self, handle, consumer):
        self.imported = (bytes(handle), consumer)


def _worker(rank, world, port, n1, n2, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle, synth_pair
    nw = importlib.import_module("fast-needleman-wunsch_b200")
    pipeline = importlib.import_module("fast-needleman-wunsch_b200.pipeline")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        orc = Oracle()
        s1, s2 = synth_pair(77, n1, n2, 5)
        R = pipeline.agree_rows_per_lane(dist, rank, lambda: 8)
        assert R == 8
        plan = FakePlan(rank)
        handles = pipeline.exchange_mailboxes(dist, plan, rank, world)
        if rank + 1 < world:
            assert plan.imported == (bytes([rank + 1]) * 64, rank + 1)
        else:
            assert plan.imported is None
        assert handles[0] == b""
        # the data path, on the CPU: halo in from the left, right column out to the right
        start, owned = nw.strip_partition(n1, world, rank)
        assert (start, owned) == orc.strip_partition(n1, world, rank)
        halo = None
        if rank > 0:
            buf = torch.empty(n2 + 1, dtype=torch.int32)
            dist.recv(buf, src=rank - 1)
            halo = buf.numpy()
        right, last = orc.strip(s1, s2, world, rank, halo)
        if rank + 1 < world:
            dist.send(torch.from_numpy(right.copy()), dst=rank + 1)
        np.save(os.path.join(out_dir, f"right{rank}.npy"), right)
        if rank == world - 1:
            np.save(os.path.join(out_dir, "score.npy"), np.array([last]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_pipeline_host_logic(tmp_path, world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle, synth_pair
    n1, n2 = 1201, 640
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, n1, n2, str(tmp_path)), nprocs=world, join=True)
    orc = Oracle()
    s1, s2 = synth_pair(77, n1, n2, 5)
    t = orc.fill(s1, s2)
    assert int(np.load(tmp_path / "score.npy")[0]) == t[-1, -1]
    for r in range(world):
        start, owned = orc.strip_partition(n1, world, r)
        assert np.array_equal(np.load(tmp_path / f"right{r}.npy"), t[:, start + owned - 1])
