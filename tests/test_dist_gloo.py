"""CPU: the N > 1 path (image strips + one framebuffer gather) over torch.distributed with the gloo backend, world 2."""
import os
import socket

import numpy as np
import pytest

from ray_b200 import dist as rdist


def test_strip_rects_partition_the_frame():
    for world in (1, 2, 3, 4, 8):
        for h in (1080, 1081, 7, 8 * 1080):
            if h < world:
                continue
            rects = rdist.all_rects(world, 1920, h)
            assert rects[0][1] == 0 and sum(r[3] for r in rects) == h
            for a, b in zip(rects, rects[1:]):
                assert a[1] + a[3] == b[1]
            assert max(r[3] for r in rects) - min(r[3] for r in rects) <= 1


def _worker(rank, world, port, w, h, out_dir):
    import torch
    import torch.distributed as tdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, ww, hh = rdist.strip_rect(rank, world, w, h)
    rows = torch.arange(y, y + hh, dtype=torch.float32).view(hh, 1, 1)
    cols = torch.arange(w, dtype=torch.float32).view(1, w, 1)
    strip = (rows * 1000 + cols).expand(hh, w, 4).contiguous() + torch.tensor([0.0, 0.25, 0.5, 0.75])
    frame = rdist.gather_strips(strip, w, h, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "frame.npy"), frame.numpy())
    else:
        assert frame is None
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.parametrize("h", [12, 13])
def test_gather_strips_world2_gloo(tmp_path, h):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w, world = 9, 2
    mp.spawn(_worker, args=(world, port, w, h, str(tmp_path)), nprocs=world, join=True)
    frame = np.load(tmp_path / "frame.npy")
    rows = np.arange(h, dtype=np.float32).reshape(h, 1, 1)
    cols = np.arange(w, dtype=np.float32).reshape(1, w, 1)
    want = np.broadcast_to(rows * 1000 + cols, (h, w, 4)) + np.array([0.0, 0.25, 0.5, 0.75], np.float32)
    assert frame.shape == (h, w, 4)
    assert np.array_equal(frame, want.astype(np.float32))


def _shared_worker(rank, world, port, w, h, out_dir):
    import torch.distributed as tdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    frame = rdist.SharedHostFrame(w, h, pin=False)
    x, y, ww, hh = rdist.strip_rect(rank, world, w, h)
    rows = np.arange(y, y + hh, dtype=np.float32).reshape(hh, 1, 1)
    cols = np.arange(w, dtype=np.float32).reshape(1, w, 1)
    frame.rows(y, hh)[...] = rows * 1000 + cols + np.array([0.0, 0.25, 0.5, 0.75], np.float32)  # the rank's "D2H copy"
    tdist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "shared.npy"), np.array(frame.array))
    frame.close()
    tdist.destroy_process_group()


@pytest.mark.parametrize("h", [12, 13])
def test_shared_host_frame_world2_gloo(tmp_path, h):
    """End-to-end delivery path of N > 1: each rank writes its own strip into one shared host frame."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w, world = 9, 2
    mp.spawn(_shared_worker, args=(world, port, w, h, str(tmp_path)), nprocs=world, join=True)
    frame = np.load(tmp_path / "shared.npy")
    rows = np.arange(h, dtype=np.float32).reshape(h, 1, 1)
    cols = np.arange(w, dtype=np.float32).reshape(1, w, 1)
    want = np.broadcast_to(rows * 1000 + cols, (h, w, 4)) + np.array([0.0, 0.25, 0.5, 0.75], np.float32)
    assert np.array_equal(frame, want.astype(np.float32))


def _private_worker(rank, world, port, w, h, out_dir):
    import torch.distributed as tdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RAY_B200_NO_SHARED_FRAME"] = "1"
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    frame = rdist.SharedHostFrame(w, h, pin=False)
    assert frame.shared is False and frame.array.shape == (h, w, 4)
    x, y, ww, hh = rdist.strip_rect(rank, world, w, h)
    frame.rows(y, hh)[...] = float(rank + 1)
    tdist.barrier()
    # private frames: the other rank's strip was NOT written here
    other = rdist.strip_rect(1 - rank, world, w, h)
    assert float(np.abs(frame.rows(other[1], other[3])).max()) == 0.0
    np.save(os.path.join(out_dir, f"private{rank}.npy"), np.array(frame.array))
    frame.close()
    tdist.destroy_process_group()


def test_shared_host_frame_falls_back_to_private_strips(tmp_path):
    """No room in /dev/shm (forced here): every rank keeps a private frame and delivers its own strip; nobody hangs."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w, h, world = 7, 10, 2
    mp.spawn(_private_worker, args=(world, port, w, h, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        a = np.load(tmp_path / f"private{rank}.npy")
        x, y, ww, hh = rdist.strip_rect(rank, world, w, h)
        assert (a[y:y + hh] == float(rank + 1)).all()


def test_comm_bands_partition_any_rect():
    """rc_comm_strip (pure arithmetic of the C-ABI): bands of the frame tile it exactly; heights differ by <= 1 row."""
    import ctypes as C
    from ray_b200 import capi, cuda
    lib = cuda.load_library()
    for n in (1, 2, 3, 8):
        for h in (1080, 1081, 9):
            full = capi.rc_rect(0, 0, 1920, h)
            y = 0
            hs = []
            for r in range(n):
                out = capi.rc_rect()
                assert lib.rc_comm_strip(C.byref(full), n, r, C.byref(out)) == 0
                assert (out.x, out.w, out.y) == (0, 1920, y)
                y += out.h
                hs.append(out.h)
            assert y == h and max(hs) - min(hs) <= 1
