"""CPU tier of context parallelism: the sharding arithmetic and the gather order, with a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_frames_and_head_owner():
    from drb200.context_parallel import head_owner, shard_frames
    assert [shard_frames(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert shard_frames(8, 0, 1) == (0, 8)
    with pytest.raises(ValueError):
        shard_frames(8, 0, 3)                   # uneven split: every rank must run the same kernels
    with pytest.raises(ValueError):
        shard_frames(8, 2, 2)
    owners = [head_owner(h, 32, 8) for h in range(32)]
    assert owners[0] == (0, 0) and owners[3] == (0, 3) and owners[4] == (1, 0) and owners[31] == (7, 3)
    assert sorted(set(o for o, _ in owners)) == list(range(8))
    with pytest.raises(ValueError):
        head_owner(0, 32, 5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drb200.context_parallel import gather_frames, shard_frames
        full = torch.arange(16 * 4 * 3 * 5, dtype=torch.float32).reshape(16, 4, 3, 5)
        t0, t1 = shard_frames(4, rank, world)
        got = gather_frames(full[:, t0:t1].contiguous())
        # N batched passes (N, C, T/P, H, W): the frame axis is third from the right whatever leads it
        full5 = torch.arange(3 * 16 * 4 * 3 * 5, dtype=torch.float32).reshape(3, 16, 4, 3, 5)
        got5 = gather_frames(full5[:, :, t0:t1].contiguous())
        # the work split of the batched pipeline: pass p is decoded by rank p mod P, every pass exactly once
        owners = [p % world for p in range(5)]
        ret[rank] = bool(torch.equal(got, full)) and bool(torch.equal(got5, full5)) and sorted(set(owners)) == list(range(world))
    finally:
        dist.destroy_process_group()


def test_gather_frames_restores_frame_order_gloo_world2():
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_context_parallel_needs_a_process_group():
    from drb200.context_parallel import ContextParallel
    with pytest.raises(RuntimeError):
        ContextParallel()
