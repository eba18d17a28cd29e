"""world_size-2 run of the sharding + gather logic over gloo on the CPU.  The compute inside each
rank is the ORACLE here (tests may use it as a stand-in; the product path has no CPU compute) --
what is under test is the partitioning, the rebased offsets and the gather."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import vcb200
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gm = vcb200.synth.random_joint_gmm(21, 4, 16)
        fm, off = vcb200.synth.trajectory_utterances(gm, 7, (8, 20), 22)
        (b, e), rel = vcb200.shard.shard_ragged(off, rank, world)
        g = O.GMMMap(*gm)
        local_fm = np.asfortranarray(fm[:, off[b]:off[e]])
        local = O.vc_traj_batch(g, 6, local_fm, rel) if e > b else np.zeros((5, 0), order="F")
        gathered = vcb200.shard.gather_frames(torch.from_numpy(np.ascontiguousarray(local.T)), dist)
        # the same gather with the shard sizes known up front (what bench.py --path traj does: no size exchange)
        sizes = [int(off[vcb200.shard.shard_ragged(off, r, world)[0][1]] - off[vcb200.shard.shard_ragged(off, r, world)[0][0]])
                 for r in range(world)]
        again = vcb200.shard.gather_frames(torch.from_numpy(np.ascontiguousarray(local.T)), dist, sizes=sizes)
        assert (again is None) == (rank != 0) and (rank != 0 or torch.equal(again, gathered))
        # frame-by-frame: contiguous frame ranges
        gm2 = vcb200.synth.random_joint_gmm(23, 3, 8)
        fm2 = vcb200.synth.fbf_feature_matrix(gm2, 101, 24)
        fb, fe = vcb200.shard.frame_range(101, rank, world)
        g2 = O.GMMMap(*gm2)
        loc2 = g2.vc(np.asfortranarray(fm2[:, fb:fe]))
        gathered2 = vcb200.shard.gather_frames(torch.from_numpy(np.ascontiguousarray(loc2.T)), dist)
        if rank == 0:
            full = O.vc_traj_batch(g, 6, fm, off)
            full2 = g2.vc(fm2)
            q.put((bool(np.array_equal(gathered.numpy().T, full)), bool(np.array_equal(gathered2.numpy().T, full2))))
        else:
            assert gathered is None and gathered2 is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shard_and_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == (True, True)
