"""CPU, world_size 2 over gloo: the multi-GPU plumbing (sample split, one film reduce, clamp on the root).
The per-rank partial films come from the C oracle here (no GPU in this container); on the GPU box the same
functions run over NCCL (tests/test_gpu_distributed.py, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_ranges_tile_the_job():
    from ky_b200.distributed import sample_range
    for spp in (0, 1, 2, 7, 64, 16384):
        for world in (1, 2, 3, 4, 8):
            r = [sample_range(spp, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == max(1, spp)
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ky_b200 as ky
    import kyo
    from ky_b200.distributed import reduce_film, sample_range
    w, h, spp = 40, 30, 5
    scene = ky.Scene(ky.SCENE_CORNELL, w, h)
    b, e = sample_range(spp, world, rank)
    part, _ = kyo.render(scene, ky.render_desc(w, h, spp, sample_begin=b, sample_end=e, flags=0))
    film = reduce_film(torch.from_numpy(part.copy()))
    ordered = reduce_film(torch.from_numpy(part.copy()), ordered=True)
    if rank == 0:
        np.save(out, film.numpy())
        np.save(out + ".ordered.npy", ordered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_split_reduce_clamp_equals_the_single_process_film(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    import ky_b200 as ky
    import kyo
    scene = ky.Scene(ky.SCENE_CORNELL, 40, 30)
    want, _ = kyo.render(scene, ky.render_desc(40, 30, 5))
    # the per-pixel FP32 sum is re-associated across ranks: SURVEY.md 8(e) bounds it by ~1e-6 relative
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= 2e-6 * np.maximum(1.0, np.abs(want)))
    assert (got >= 0).all() and (got <= 1).all()
    # the rank-ordered sum: exactly clamp(partial_0 + partial_1), whatever the backend's reduce algorithm
    from ky_b200.distributed import sample_range
    parts = [kyo.render(scene, ky.render_desc(40, 30, 5, sample_begin=b, sample_end=e, flags=0))[0] for b, e in (sample_range(5, 2, 0), sample_range(5, 2, 1))]
    expect = np.clip((parts[0] + parts[1]).astype(np.float32), 0.0, 1.0)
    assert np.array_equal(np.load(out + ".ordered.npy").view(np.uint32), expect.view(np.uint32))
