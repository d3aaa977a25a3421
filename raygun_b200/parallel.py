"""Screen-space partitioning of one frame across GPUs (one process per GPU) and the gather hand-shake.

The reference is single-device (raygun/vulkan_context.cpp:165); this is the one place the B200 path adds parallelism:
pixels are independent through ray generation and shading, and the post chain has a bounded dependency radius
(1 px rough_prepare + 10 px blur + 28 px FXAA search = 39 px), so every rank renders its band plus a 40 px halo against a
replicated scene and the final kernel stores the band straight into rank 0's frame buffer over NVLink.
"""
from __future__ import annotations

HALO = 40


def band_region(width: int, height: int, rank: int, world: int, split: str = "columns"):
    """(x0, y0, x1, y1) owned by `rank`: contiguous, disjoint, covering the frame; sizes differ by at most one pixel."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    if split == "columns":
        if world > width:
            raise ValueError("more ranks than columns")
        return (width * rank) // world, 0, (width * (rank + 1)) // world, height
    if split == "rows":
        if world > height:
            raise ValueError("more ranks than rows")
        return 0, (height * rank) // world, width, (height * (rank + 1)) // world
    raise ValueError(f"unknown split {split!r}")


def rendered_rect(region, width, height, halo: int = HALO):
    """Rectangle a rank actually traces: region grown by the halo, clipped to the frame."""
    x0, y0, x1, y1 = region
    return max(0, x0 - halo), max(0, y0 - halo), min(width, x1 + halo), min(height, y1 + halo)


def overdraw(width, height, world, split="columns", halo: int = HALO) -> float:
    """Traced pixels over all ranks / frame pixels."""
    total = 0
    for r in range(world):
        x0, y0, x1, y1 = rendered_rect(band_region(width, height, r, world, split), width, height, halo)
        total += (x1 - x0) * (y1 - y0)
    return total / float(width * height)


def share_gather_handle(dist, rank: int, handle: bytes | None) -> bytes:
    """Rank 0 publishes the 64-byte CUDA IPC handle of its full-frame buffer; every rank returns it."""
    obj = [handle if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    if not isinstance(obj[0], (bytes, bytearray)) or len(obj[0]) != 64:
        raise RuntimeError("gather handle exchange failed")
    return bytes(obj[0])


def attach_partition(dist, rt, rank: int, world: int):
    """Partitioned mode over processes: every rank exports its G-buffer descriptor (CUDA IPC handles), all ranks gather
    the descriptors and attach each other.  `rt` must already have its region (rt.set_region)."""
    rt.set_partition(rank, world)
    descs = [None] * world
    dist.all_gather_object(descs, rt.peer_export())
    for q in range(world):
        if q != rank:
            rt.peer_attach(q, descs[q], open_ipc=True)
    dist.barrier()


def attach_partition_in_process(rts):
    """Same for several contexts of ONE process (tests, single-process multi-GPU hosts): raw device pointers, no IPC."""
    world = len(rts)
    for r, rt in enumerate(rts):
        rt.set_partition(r, world)
    descs = [rt.peer_export() for rt in rts]
    for r, rt in enumerate(rts):
        for q in range(world):
            if q != r:
                rt.peer_attach(q, descs[q], open_ipc=False)


class SharedHostFrame:
    """A width x height RGBA8 frame in POSIX shared memory, page-locked and mapped into the CUDA address space of the calling process
    (rg_host_frame_register).  Every rank of a node opens the SAME segment and makes it its gather target: each GPU's final kernel then
    stores its band of the frame into host memory over its own PCIe link, and the consumer (rank 0) reads the assembled frame from
    `array` without any device -> host copy.  No reference counterpart (the reference presents through its swapchain)."""

    def __init__(self, rt, name: str, width: int, height: int, create: bool):
        import mmap
        import os
        import numpy as np
        self.rt, self.path, self.owner = rt, os.path.join("/dev/shm", name), create
        nbytes = width * height * 4
        fd = os.open(self.path, (os.O_CREAT | os.O_RDWR) if create else os.O_RDWR, 0o600)
        try:
            if create:
                os.ftruncate(fd, nbytes)
            self.mm = mmap.mmap(fd, nbytes)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self.mm, dtype=np.uint8).reshape(height, width, 4)
        self.host_ptr = self.array.ctypes.data
        self.device_ptr = 0
        try:
            self.device_ptr = rt.host_frame_register(self.host_ptr, nbytes)
        except Exception:
            self.close()    # nothing is left behind in /dev/shm when the memory cannot be page-locked
            raise

    def close(self):
        import os
        if self.device_ptr:
            self.rt.host_frame_unregister(self.host_ptr)
            self.device_ptr = 0
        self.array = None
        try:
            self.mm.close()
        except BufferError:
            pass
        if self.owner and os.path.exists(self.path):
            os.unlink(self.path)


def open_shared_host_frame(dist, rt, rank: int, width: int, height: int, agree=None):
    """Rank 0 creates the segment and publishes its name; every rank maps and registers it (collective).  Returns None on EVERY rank
    when any rank could not (no /dev/shm, locked-memory limit ...): `agree(ok) -> bool` is the all-ranks AND (bench.py passes an
    all_reduce; without it a failure raises)."""
    import os
    frame, err = None, None
    name = [None]
    if rank == 0:
        try:
            frame = SharedHostFrame(rt, f"rgb200_frame_{os.getpid()}", width, height, create=True)
            name[0] = os.path.basename(frame.path)
        except Exception as e:   # noqa: BLE001 -- any failure means "fall back", decided together below
            err = e
    dist.broadcast_object_list(name, src=0)
    if rank != 0 and name[0] is not None:
        try:
            frame = SharedHostFrame(rt, name[0], width, height, create=False)
        except Exception as e:   # noqa: BLE001
            err = e
    ok = frame is not None
    all_ok = agree(ok) if agree is not None else ok
    if agree is None and err is not None:
        raise err
    if not all_ok:
        if frame is not None:
            frame.close()
        return None
    dist.barrier()
    return frame
