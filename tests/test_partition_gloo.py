"""N > 1 host-side logic on CPU: partition arithmetic, and a world_size-2 gloo run in which each rank renders only its
interleaved row tiles (device trace code compiled for the host, tests/hostemu) and rank 0 gathers + assembles the frame
with the same pack/assemble code bench.py's NCCL-gather path uses. The result must equal the oracle's full frame."""
import os
import socket
import sys

import numpy as np
import pytest

import partition
import scenes


@pytest.mark.parametrize("h,tile_rows,world", [(2160, 8, 8), (131, 3, 4), (720, 16, 2), (5, 8, 4), (17, 1, 8)])
def test_partition_covers_every_row_once(h, tile_rows, world):
    seen = np.concatenate([partition.rows_of_rank(h, tile_rows, r, world) for r in range(world)])
    assert sorted(seen.tolist()) == list(range(h))
    for r in range(world):
        rows = partition.rows_of_rank(h, tile_rows, r, world)
        assert np.all((rows // tile_rows) % world == r)
    assert partition.max_rows_per_rank(h, tile_rows, world) >= h // world


def _worker(rank, world, port, w, h, tile_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.join(os.path.dirname(here), "uu-infogr-raytracer_b200"))
    import torch
    import torch.distributed as dist
    import hostemu_lib as E
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)
    full = E.render(sc, cam, w, h, 8, tiny=2)["pixels"]          # stands in for this rank's GPU; only own rows are used
    rows = partition.rows_of_rank(h, tile_rows, rank, world)
    local = np.zeros_like(full); local[rows] = full[rows]
    pad = partition.max_rows_per_rank(h, tile_rows, world)
    payload = partition.pack_rows(torch.from_numpy(local), rows, pad)
    gathered = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, gathered, dst=0)
    if rank == 0:
        frame = partition.assemble(gathered, h, w, tile_rows).numpy()
        q.put(frame)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tile_rows", [8, 3])
def test_two_rank_gloo_gather_equals_oracle(built, tile_rows):
    import torch.multiprocessing as mp
    import oracle_lib as O
    w, h, world = 96, 61, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, tile_rows, q)) for r in range(world)]
    for p in procs: p.start()
    frame = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    sc = scenes.default_scene()
    cam = scenes.make_camera(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15, width=w, height=h)
    assert np.array_equal(frame, O.render(sc, cam, w, h, 8)["pixels"])
