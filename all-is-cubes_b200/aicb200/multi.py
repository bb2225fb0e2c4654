"""Row-strip sharding of one frame across ranks — host-side plumbing over
torch.distributed.  One process per GPU; the scene is replicated; strip s (16 rows) belongs to rank
s % world.  Two ways to deliver the strips to rank 0:

  * gather_frame(): NCCL (or gloo on CPU) gather of the packed per-rank strips + reassembly;
  * PeerFrame: rank 0 exports its frame buffer through CUDA IPC, every rank maps it, and the trace
    kernel stores its pixels straight into rank 0's memory over NVLink (no collective in the data
    path, only a barrier) — see aicb_render_srgb8_device_frame().

The reference has no multi-process path at all ("Collectives: none"); the contract here is
"N-rank frame == 1-rank frame byte for byte"."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

STRIP_ROWS = 16


def shard_rows(height: int, index: int, count: int, strip_rows: int = STRIP_ROWS) -> np.ndarray:
    """Framebuffer rows owned by shard `index` of `count`, in increasing order (== the packed order the
    renderer writes them in)."""
    y = np.arange(height)
    if count <= 1:
        return y
    return y[(y // strip_rows) % count == index]


def max_shard_rows(height: int, count: int, strip_rows: int = STRIP_ROWS) -> int:
    return max(len(shard_rows(height, r, count, strip_rows)) for r in range(count))


def gather_frame(local: torch.Tensor, height: int, width: int, rank: int, world: int, frame: torch.Tensor = None,
                 strip_rows: int = STRIP_ROWS, scratch=None):
    """`local`: [max_shard_rows*width, 4] uint8 (packed rows of this rank, padded to the common size).
    Returns the [height, width, 4] frame on rank 0 (None elsewhere)."""
    if world == 1:
        rows = height
        return local[: rows * width].view(height, width, 4)
    gather_list = None
    if rank == 0:
        gather_list = scratch if scratch is not None else [torch.empty_like(local) for _ in range(world)]
    dist.gather(local, gather_list, dst=0)
    if rank != 0:
        return None
    if frame is None:
        frame = torch.empty((height, width, 4), dtype=torch.uint8, device=local.device)
    fr = frame.view(height, width * 4)
    for r in range(world):
        rows = torch.as_tensor(shard_rows(height, r, world, strip_rows), device=local.device)
        fr[rows] = gather_list[r][: rows.numel() * width].view(rows.numel(), width * 4)
    return frame


class PeerFrame:
    """A full-frame sRGB8 buffer that lives on rank 0's GPU and is mapped (CUDA IPC, peer access over
    NVLink) into every rank of the node.  The 64-byte IPC handle travels over torch.distributed."""

    def __init__(self, ctx, height: int, width: int, rank: int, world: int):
        import ctypes as C
        from . import _check, load_library
        self.ctx, self.rank, self.world = ctx, rank, world
        self.height, self.width = height, width
        self.lib = load_library()
        self.ptr = C.c_void_p()
        handle = torch.zeros(64, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 64)()
            _check(self.lib.aicb_frame_create(ctx.handle, height * width, C.byref(self.ptr), buf))
            handle = torch.tensor(list(buf), dtype=torch.uint8)
        if world > 1:
            h = handle.cuda() if dist.get_backend() == "nccl" else handle
            dist.broadcast(h, src=0)
            handle = h.cpu()
        if rank != 0:
            buf = (C.c_uint8 * 64)(*handle.tolist())
            _check(self.lib.aicb_frame_open(ctx.handle, buf, C.byref(self.ptr)))
        self.opened = rank != 0
        self.frames = 0   # frames delivered through the arrival counter so far

    # ---- delivery without a collective: the counters behind the frame's pixels (aicb_frame_signal & co.) ----------
    def begin_frame(self, stream_ptr: int = 0):
        """Before this rank stores into the frame again: the owner must be through with the previous frame."""
        from . import _check
        if self.frames:
            _check(self.lib.aicb_frame_wait_consumed(self.ctx.handle, self.ptr, self.height * self.width, self.frames, stream_ptr))

    def end_frame(self, stream_ptr: int = 0, release: bool = True):
        """After aicb_render_srgb8_device_frame on the same stream: signal; the owner waits for every rank's signal and
        (release=True: the frame stays on the device) marks it consumed."""
        from . import _check
        n = self.height * self.width
        self.frames += 1
        _check(self.lib.aicb_frame_signal(self.ctx.handle, self.ptr, n, stream_ptr))
        if self.rank == 0:
            _check(self.lib.aicb_frame_wait_arrived(self.ctx.handle, self.ptr, n, self.frames * self.world, stream_ptr))
            if release:
                _check(self.lib.aicb_frame_release(self.ctx.handle, self.ptr, n, self.frames, stream_ptr))

    def release(self, stream_ptr: int = 0):
        from . import _check
        if self.rank == 0:
            _check(self.lib.aicb_frame_release(self.ctx.handle, self.ptr, self.height * self.width, self.frames, stream_ptr))

    def timed_out(self) -> bool:
        import ctypes as C
        from . import _check
        v = C.c_uint32(0)
        _check(self.lib.aicb_frame_timed_out(self.ctx.handle, self.ptr, self.height * self.width, C.byref(v)))
        return bool(v.value)

    def read(self, out: torch.Tensor, stream_ptr: int = 0):
        """Rank 0: device -> host copy of the frame into a (pinned) uint8 tensor of height*width*4 bytes."""
        from . import _check
        _check(self.lib.aicb_frame_read(self.ctx.handle, self.ptr, out.data_ptr(), self.height * self.width, stream_ptr))

    def close(self):
        if self.ptr:
            self.lib.aicb_frame_close(self.ctx.handle, self.ptr, 1 if self.opened else 0)
            self.ptr = None
