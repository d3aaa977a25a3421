"""CPU tests (-m "not gpu"): the N > 1 host logic -- band partition, halo rectangles, and the gather hand-shake over a
world_size-2 gloo group (the data path itself has no collective: bands are stored by the final kernel into rank 0's
frame buffer)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from raygun_b200.parallel import HALO, band_region, overdraw, rendered_rect, share_gather_handle


@pytest.mark.parametrize("W,H,world,split", [(1920, 1080, 1, "columns"), (1920, 1080, 8, "columns"), (3840, 2160, 8, "rows"), (101, 57, 4, "columns"), (7, 5, 5, "rows")])
def test_bands_partition_the_frame(W, H, world, split):
    cover = np.zeros((H, W), np.int32)
    for r in range(world):
        x0, y0, x1, y1 = band_region(W, H, r, world, split)
        assert x0 < x1 and y0 < y1
        cover[y0:y1, x0:x1] += 1
        rx0, ry0, rx1, ry1 = rendered_rect((x0, y0, x1, y1), W, H)
        assert rx0 == max(0, x0 - HALO) and rx1 == min(W, x1 + HALO) and ry0 == max(0, y0 - HALO) and ry1 == min(H, y1 + HALO)
    assert np.all(cover == 1)


def test_overdraw_numbers():
    assert overdraw(1920, 1080, 1) == 1.0
    assert abs(overdraw(3840, 2160, 8) - (2 * 520 + 6 * 560) / 3840) < 1e-9
    with pytest.raises(ValueError):
        band_region(4, 4, 0, 8)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    handle = bytes(range(64)) if rank == 0 else None
    got = share_gather_handle(dist, rank, handle)
    W, H = 640, 360
    region = band_region(W, H, rank, world)
    regions = [None] * world
    dist.all_gather_object(regions, region)
    # every rank "stores" its band into rank 0's frame: emulate with a gather of band ids
    band = np.full((region[3] - region[1], region[2] - region[0]), rank, np.int32)
    bands = [None] * world
    dist.all_gather_object(bands, band)
    frame = np.full((H, W), -1, np.int32)
    for r, (reg, b) in enumerate(zip(regions, bands)):
        frame[reg[1]:reg[3], reg[0]:reg[2]] = b
    q.put((rank, got == bytes(range(64)), regions, bool(np.all(frame >= 0)), int(frame[0, 0]), int(frame[-1, -1])))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_handshake_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, regions, full, first, last in res:
        assert ok and full and first == 0 and last == 1
        assert regions == [(0, 0, 320, 360), (320, 0, 640, 360)]


class _FakeRt:
    """Stands in for raygun_b200.Raytracer on a machine without a GPU: "registration" returns the host address itself (what
    unified addressing gives), or fails on request."""
    def __init__(self, fail=False):
        self.fail, self.registered = fail, set()

    def host_frame_register(self, host_ptr, nbytes):
        if self.fail:
            raise RuntimeError("cudaHostRegister failed (test)")
        self.registered.add(host_ptr)
        return host_ptr

    def host_frame_unregister(self, host_ptr):
        self.registered.discard(host_ptr)


def _shared_frame_worker(rank, world, port, q, fail_rank):
    import torch
    from raygun_b200.parallel import open_shared_host_frame
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def agree(ok):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    W, H = 64, 16
    rt = _FakeRt(fail=(rank == fail_rank))
    frame = open_shared_host_frame(dist, rt, rank, W, H, agree)
    if frame is None:
        q.put((rank, None, None, len(rt.registered)))
    else:
        x0, y0, x1, y1 = band_region(W, H, rank, world)
        frame.array[y0:y1, x0:x1] = rank + 1      # what this rank's final kernel stores: its band of the ONE frame
        dist.barrier()
        seen = frame.array[:, :, 0].copy()         # every rank (rank 0 is the consumer) sees the assembled frame, no copy
        path = frame.path
        dist.barrier()
        frame.close()
        dist.barrier()
        q.put((rank, seen, os.path.exists(path), len(rt.registered)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1, 0])
def test_shared_host_frame_world2_gloo(fail_rank):
    """The e2e frame of N > 1 (bench.py): one POSIX shared-memory frame mapped by every rank; all ranks fall back together when one
    of them cannot register it."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shared_frame_worker, args=(r, 2, port, q, fail_rank)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, seen, still_there, registered in res:
        assert registered == 0                      # nothing stays page-locked
        if fail_rank >= 0:
            assert seen is None
        else:
            assert not still_there                  # the owner unlinked the segment
            assert (seen[:, :32] == 1).all() and (seen[:, 32:] == 2).all()
    for p in procs:   # nothing of THIS test's ranks is left in /dev/shm (also when registration failed after the segment was created)
        assert not os.path.exists(f"/dev/shm/rgb200_frame_{p.pid}")
