"""Multi-GPU sharding of the render: image strips per rank + ONE framebuffer gather per sample batch.

The reference has no multi-device story; the path shards trivially because a pixel's estimate
depends only on (x, y, iteration, scene) (section 8(e)): the RNG is keyed by the absolute pixel coordinate
(reference internal/CoreRef.cpp:1477-1480), jitter is sampled inside the pixel, and every ray carries its `xy`.  So each
rank renders a horizontal strip of the frame through the ordinary RegionContext{rect} API into its own replica of the
scene, with no communication between bounces, and the only exchange is a gather of the accumulated strips to rank 0.
This module is backend-agnostic torch.distributed code (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import List, Tuple


def strip_rect(rank: int, world: int, w: int, h: int) -> Tuple[int, int, int, int]:
    """Rect (x, y, w, h) of the rows rank `rank` owns: contiguous strips, heights differ by at most one row."""
    base, rem = divmod(h, world)
    y0 = rank * base + min(rank, rem)
    rows = base + (1 if rank < rem else 0)
    return (0, y0, w, rows)


def all_rects(world: int, w: int, h: int) -> List[Tuple[int, int, int, int]]:
    return [strip_rect(r, world, w, h) for r in range(world)]


def gather_strips(local_strip, w: int, h: int, dst: int = 0):
    """Gather every rank's (rows_r, w, 4) float32 strip to `dst` and assemble the (h, w, 4) frame there.

    Strip heights may differ by one row, so strips are padded to the maximum height for the collective (NCCL gather
    needs equal shapes) and cropped on arrival.  Returns the assembled frame on `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    rects = all_rects(world, w, h)
    max_rows = max(r[3] for r in rects)
    rows = rects[rank][3]
    assert tuple(local_strip.shape) == (rows, w, 4), (tuple(local_strip.shape), (rows, w, 4))
    if rows < max_rows:
        pad = torch.zeros((max_rows - rows, w, 4), dtype=local_strip.dtype, device=local_strip.device)
        send = torch.cat([local_strip, pad], dim=0).contiguous()
    else:
        send = local_strip.contiguous()
    if rank == dst:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, gather_list=parts, dst=dst)
        frame = torch.empty((h, w, 4), dtype=send.dtype, device=send.device)
        for r, (x, y, ww, hh) in enumerate(rects):
            frame[y:y + hh] = parts[r][:hh]
        return frame
    dist.gather(send, gather_list=None, dst=dst)
    return None


class DeviceImage:
    """Zero-copy torch view of one of the renderer's device frame buffers (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, h: int, w: int):
        self.__cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def device_frame_tensor(renderer, which, device_index: int):
    """torch.Tensor aliasing frame buffer `which` (ray_b200.capi.RC_BUF_*) of a ray_b200.host.Renderer."""
    import torch

    from . import cuda

    lib = cuda.load_library()
    ptr = lib.rc_device_ptr(renderer.native_context(), which)
    if not ptr:
        raise RuntimeError("rc_device_ptr returned NULL")
    return torch.as_tensor(DeviceImage(ptr, renderer.hh, renderer.w), device=torch.device("cuda", device_index))


class SharedHostFrame:
    """ONE host frame shared by all ranks of a node: every rank copies its own strip device->host straight into it, so N
    PCIe links work in parallel and rank 0 never has to pull N strips through its own link (the end-to-end delivery path;
    the NCCL gather above stays for consumers that want the frame on a device).

    The block is POSIX shared memory created by rank 0 and attached by the others (name passed through the process
    group); on a CUDA machine each rank page-locks its mapping (cudaHostRegister) so the strip copy is a full-speed DMA.
    `rows(y, h)` is the (h, w, 4) float32 numpy view of rows [y, y + h); `ptr(y)` the address of row y."""

    def __init__(self, w: int, h: int, pin: bool = True):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.w, self.h = w, h
        self.nbytes = w * h * 16
        rank = dist.get_rank() if dist.is_initialized() else 0
        self.rank = rank
        self.shm = None
        if rank == 0:
            # tmpfs allocates on first touch: a block larger than what /dev/shm has left would be created fine and
            # kill the writers with SIGBUS later, so check the space first and tell every rank when sharing is off
            try:
                import os
                if os.environ.get("RAY_B200_NO_SHARED_FRAME"):
                    raise OSError("disabled by RAY_B200_NO_SHARED_FRAME")
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < self.nbytes + (64 << 20):
                    raise OSError("/dev/shm has %d MB free, the frame needs %d MB" % (st.f_bavail * st.f_frsize >> 20, self.nbytes >> 20))
                self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
                name = [self.shm.name]
            except Exception as e:  # noqa: BLE001
                self.why_not_shared = str(e)
                name = [""]
        else:
            name = [None]
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast_object_list(name, src=0)
        self.shared = bool(name[0])
        if rank != 0 and self.shared:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # rank 0 owns the block: keep this process's resource tracker from unlinking / warning about it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        if not self.shared:
            # no shared block: every rank keeps a private page-locked frame and fills only its own strip -- the same N
            # parallel device->host copies, but the strips do not land in one address space
            self.array = np.zeros((h, w, 4), dtype=np.float32)
        else:
            self.array = np.ndarray((h, w, 4), dtype=np.float32, buffer=self.shm.buf)
        self.pinned = False
        if pin:
            try:
                import torch
                if torch.cuda.is_available():
                    rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 0)
                    self.pinned = (int(rc) == 0)
            except Exception:
                self.pinned = False

    def ptr(self, y: int = 0) -> int:
        return self.array.ctypes.data + y * self.w * 16

    def rows(self, y: int, h: int):
        return self.array[y:y + h]

    def close(self):
        import torch.distributed as dist
        if self.pinned:
            try:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            except Exception:
                pass
        self.array = None
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
        try:
            if self.shm is not None:
                self.shm.close()
                if self.rank == 0:
                    self.shm.unlink()
        except Exception:
            pass


def deliver_strip(renderer, which, rect, frame: SharedHostFrame):
    """Copy this rank's strip `rect` of plane `which` device->host into the shared frame (blocking)."""
    import ctypes as C

    from . import capi, cuda

    lib = cuda.load_library()
    ctx = renderer.native_context()
    x, y, w, h = rect
    r = capi.rc_rect(x, y, w, h)
    if lib.rc_readback_async(ctx, which, C.byref(r), C.c_void_p(frame.ptr(y) + x * 16), frame.w) != 0:
        raise RuntimeError(lib.rc_last_error(ctx).decode())
    if lib.rc_sync(ctx) != 0:
        raise RuntimeError(lib.rc_last_error(ctx).decode())
