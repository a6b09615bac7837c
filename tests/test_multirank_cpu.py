"""World-size-2 gloo test of the offline multi-GPU plumbing (hrbffusion3d_b200/multigpu.py): sequences are
scattered from rank 0, each rank 'tracks' its own, trajectories are gathered back ragged.  CPU only."""
import json
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hrbffusion3d_b200 import multigpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 owns the "logs": ragged blobs, one per rank (the second one longer, one empty case is covered below)
        blobs = None
        if rank == 0:
            blobs = [json.dumps({"seq": r, "frames": 5 + 3 * r, "payload": "x" * (17 * r)}).encode() for r in range(world)]
        mine = json.loads(multigpu.scatter_blobs(blobs, device="cpu"))
        assert mine["seq"] == rank and len(mine["payload"]) == 17 * rank
        # an empty blob must survive the padded scatter
        empties = [b"" if r == 1 else b"abc" for r in range(world)] if rank == 0 else None
        got = multigpu.scatter_blobs(empties, device="cpu")
        assert got == (b"" if rank == 1 else b"abc")
        # each rank produces a trajectory whose length and content depend on its sequence
        n = mine["frames"]
        traj = torch.arange(n * 12, dtype=torch.float32).reshape(n, 12) + 1000.0 * rank
        gathered = multigpu.gather_trajectories(traj)
        ms = multigpu.max_over_ranks(10.0 + rank)
        assert ms == 10.0 + (world - 1)
        if rank == 0:
            assert len(gathered) == world
            for r, g in enumerate(gathered):
                assert g.shape == (5 + 3 * r, 12)
                assert torch.equal(g, torch.arange((5 + 3 * r) * 12, dtype=torch.float32).reshape(-1, 12) + 1000.0 * r)
            np.save(os.path.join(outdir, "ok.npy"), np.array([len(gathered)]))
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


def test_scatter_gather_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0] == world


def test_single_process_paths():
    assert multigpu.assign_sequences(5, 2) == [[0, 2, 4], [1, 3]]
    assert multigpu.scatter_blobs([b"hello"]) == b"hello"
    t = torch.zeros(3, 12)
    assert multigpu.gather_trajectories(t)[0].shape == (3, 12)
    assert multigpu.max_over_ranks(1.5) == 1.5
