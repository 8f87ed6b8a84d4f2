"""Worker of tests/test_gpu_parity.py::test_multi_process_ipc_pipeline: one process per GPU (torchrun), column strips with
CUDA-IPC halo mailboxes -- the path bench.py --gpus N uses -- checked against the oracle on every rank.
    python -m torch.distributed.run --nproc-per-node N tests/mgpu_worker.py {synthetic|64gb}"""
import importlib
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from conftest import GOLDEN, Oracle, pair_paths, synth_pair      # noqa: E402

nw = importlib.import_module("fast-needleman-wunsch_b200")
pipeline = importlib.import_module("fast-needleman-wunsch_b200.pipeline")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    nw.init(local)
    if what == "64gb":
        a, b = pair_paths("64gb")
        s1, s2 = np.fromfile(a, dtype=np.int8), np.fromfile(b, dtype=np.int8)
        modes = [nw.NW_MODE_BOUNDARY]
    else:
        s1, s2 = synth_pair(61, 20011, 9000, 5)
        modes = [nw.NW_MODE_BOUNDARY, nw.NW_MODE_FULL]
    orc = Oracle()
    for mode in modes:
        plan = nw.Plan(s1.size, s2.size, mode=mode, device=local, part=rank, nparts=world, rows_per_lane=8)
        pipeline.exchange_mailboxes(dist, plan, rank, world)
        plan.upload(s1, s2)
        for rep in range(3):                     # three epochs: double-buffered mailboxes, ack words
            plan.run()
        plan.sync()
        dist.barrier()
        if what == "64gb":
            if rank == world - 1:
                assert plan.score() == GOLDEN["fixtures"]["64gb"]["score"], plan.score()
        else:
            t = orc.fill(s1, s2)
            assert np.array_equal(plan.last_col(), t[:, plan.jstart + plan.ncols]), f"rank {rank}: last column differs"
            assert np.array_equal(plan.last_row(), t[-1, plan.jstart:plan.jstart + plan.ncols + 1]), f"rank {rank}: last row"
            if rank == world - 1:
                assert plan.score() == t[-1, -1]
            if mode == nw.NW_MODE_FULL:
                out = np.zeros_like(t)
                plan.table_to_host(out)
                lo = plan.jstart + (1 if rank > 0 else 0)
                assert np.array_equal(out[:, lo:plan.jstart + plan.ncols + 1], t[:, lo:plan.jstart + plan.ncols + 1]), f"rank {rank}: table"
        dist.barrier()
        plan.close()
    if rank == 0:
        print(f"mgpu_worker ok: {what}, {world} ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
