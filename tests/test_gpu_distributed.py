"""GPU: the multi-GPU plumbing over NCCL (needs >= 2 GPUs; skipped otherwise) and its single-GPU degenerate case."""
import os
import socket
import sys
import time

import numpy as np
import pytest

import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP = 64, 36, 6


def test_render_job_single_process(device):
    import torch
    from ky_b200.distributed import render_job
    scene = ky.Scene(ky.SCENE_VEACH, W, H)
    device.upload(scene)
    film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    render_job(device, scene, ky.render_desc(W, H, SPP), film)
    torch.cuda.synchronize()
    want, _ = kyo.render(scene, ky.render_desc(W, H, SPP))
    assert np.array_equal(film.cpu().numpy().view(np.uint32), want.view(np.uint32))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    for k in [k for k in os.environ if k.startswith("TORCHELASTIC") or k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "GROUP_RANK")]:
        del os.environ[k]  # a launcher's environment must not leak into this private two-rank group
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ky_b200 as ky
    from ky_b200.distributed import render_job
    dev = ky.Device(rank)
    scene = ky.Scene(ky.SCENE_CORNELL, W, H)
    dev.upload(scene)
    film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    render_job(dev, scene, ky.render_desc(W, H, SPP), film)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out, film.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_split_reduce_clamp(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "film.npy")
    ctx = mp.spawn(_worker, args=(2, port, out), nprocs=2, join=False)
    deadline = time.time() + 300
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for proc in ctx.processes:
                proc.kill()
            pytest.fail("two-rank NCCL job did not finish within 300 s")
    got = np.load(out)
    scene = ky.Scene(ky.SCENE_CORNELL, W, H)
    want, _ = kyo.render(scene, ky.render_desc(W, H, SPP))
    assert np.all(np.abs(got - want) <= 2e-6 * np.maximum(1.0, np.abs(want)))
