"""Work partitioning for one-process-per-GPU rendering.  The scene is replicated on
every GPU; frames are split by camera view or by image tile; there is no collective on the data path
(only the finished framebuffer pieces are gathered)."""
from __future__ import annotations

import numpy as np

TILE_COLUMNS = 32   # tile = 32 pixel columns x full height: contiguous in the (W,H,3) i-major image
STRIPE_COLUMNS = TILE_COLUMNS   # = the kernels' macro-tile width (RTGS_OPT_STRIPE)


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPU cores next to GPU ``device_index`` (NVML's ideal-CPU mask for the
    device, i.e. its NUMA node), so that the pinned host image buffers it allocates afterwards are local to the
    GPU's PCIe root and eight ranks do not pile their framebuffer copies onto one socket.  Call before the first
    pinned allocation.  Never raises: returns {"bound": False, "why": ...} when NVML or the mask is
    unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"bound": False, "why": "empty NVML affinity mask within the allowed CPUs"}
        if cpus == allowed:
            return {"bound": False, "why": "single NUMA domain (mask = all allowed CPUs)", "cpus": len(cpus)}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "first_cpu": min(cpus)}
    except Exception as e:   # no NVML, no permission, ...: rendering does not depend on it
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


def views_for_rank(n_views: int, rank: int, world: int) -> list[int]:
    """View k -> rank k mod world."""
    return list(range(rank, n_views, world))


def tiles_for_rank(W: int, H: int, rank: int, world: int) -> list[tuple[int, int, int, int]]:
    """32-column strips dealt round-robin to the ranks (interleaving balances dense and empty image
    regions).  Returns (x0, y0, w, h) regions."""
    strips = [(x0, 0, min(TILE_COLUMNS, W - x0), H) for x0 in range(0, W, TILE_COLUMNS)]
    return strips[rank::world]


def assemble_tiles(W: int, H: int, world: int, payloads) -> np.ndarray:
    """Rebuild the (W,H,3) frame from the per-rank concatenated strip buffers (rank order)."""
    frame = np.empty((W, H, 3), dtype=np.float32)
    for rank, buf in enumerate(payloads):
        off = 0
        for (x0, y0, w, h) in tiles_for_rank(W, H, rank, world):
            n = w * h * 3
            frame[x0:x0 + w, y0:y0 + h] = np.asarray(buf[off:off + n]).reshape(w, h, 3)
            off += n
    return frame


def stripe_columns(W: int, rank: int, world: int):
    """Pixel columns of the stripes owned by `rank` (what ``Scene.set_stripe(world, rank)`` renders)."""
    return np.concatenate([np.arange(x0, x0 + w) for (x0, _, w, _) in tiles_for_rank(W, 1, rank, world)] or
                          [np.zeros(0, np.int64)]).astype(np.int64)


class StripeGather:
    """Gathers a tile-sharded frame on rank 0 with one collective (torch.distributed gather of the packed
    stripes; NCCL on GPUs, gloo in the CPU tests).  Only finished pixels move - nothing on the render path."""

    def __init__(self, W: int, H: int, rank: int, world: int, device):
        import torch
        self.W, self.H, self.rank, self.world = W, H, rank, world
        self.cols = [torch.from_numpy(stripe_columns(W, r, world)).to(device) for r in range(world)]
        self.max_cols = max(int(c.numel()) for c in self.cols)
        self.send = torch.zeros((self.max_cols, H, 3), dtype=torch.float32, device=device)
        self.recv = [torch.zeros_like(self.send) for _ in range(world)] if rank == 0 else None

    def __call__(self, frame, dist):
        """frame: (W,H,3) tensor holding this rank's stripes; on rank 0 the other ranks' stripes are filled in."""
        import torch
        mine = self.cols[self.rank]
        if self.world == 1:
            return frame
        torch.index_select(frame, 0, mine, out=self.send[:mine.numel()])
        dist.gather(self.send, self.recv, dst=0)
        if self.rank == 0:
            for r in range(1, self.world):
                frame.index_copy_(0, self.cols[r], self.recv[r][:self.cols[r].numel()])
        return frame


class PeerFrame:
    """Framebuffers in rank 0's device memory that every rank's render kernels store into directly over NVLink
    (CUDA IPC + peer access), handed over with two device-side counters - no collective, no host
    synchronisation, no extra kernel on the producing ranks:

    * ``slots`` images of (W,H,3) float32 per buffer: 1 for tile sharding (every rank stores its stripes of the one
      frame, ``Scene.set_stripe(world, rank)``), ``world`` for view sharding (rank r stores its whole frame into
      slot r), ``buffers`` deep so that the producers may run ahead of the consumer;
    * ``arrive[b]``: the last CTA of a rank's frame increments it once all the frame's stores are performed
      (release, system scope; csrc/render_common.cuh: frame_complete).  Rank 0 waits on its stream until the buffer
      has ``world`` arrivals per use (``wait()``: a one-thread kernel, rtgs_stream_wait_counter);
    * ``consumed``: rank 0 stores the number of frames it has finished with (``release()``); a producer's warps
      check it before their first store into a buffer that is being reused (RenderParams::grant).

    Per frame, every rank calls ``begin()`` (returns the tensor to pass as ``out`` of a full-frame
    ``RayTracer.render_device``) and renders; rank 0 then calls ``wait()``, consumes ``frame()`` on the same
    stream and calls ``release()``."""

    class _Raw:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 3}

    CTRL_BYTES = 1024   # arrive[b] at 128 b, consumed at 128 * buffers

    def __init__(self, W: int, H: int, rank: int, world: int, device: int, dist=None, slots: int = 1,
                 buffers: int = 2):
        import ctypes as C

        import torch

        from . import _native
        lib = _native.load()
        assert 1 <= buffers <= 4 and slots >= 1
        self.rank, self.world, self.device = rank, world, device
        self.W, self.H, self.slots, self.buffers = W, H, slots, buffers
        self._lib, self._owner_ptr, self._peer_ptr = lib, None, None
        self.image_bytes = W * H * 3 * 4
        nbytes = self.CTRL_BYTES + buffers * slots * self.image_bytes
        handle = torch.zeros(64, dtype=torch.uint8)
        p = C.c_void_p()
        if rank == 0:
            _native.check(lib.rtgs_device_alloc(device, nbytes, C.byref(p)))
            self._owner_ptr = p.value
            ctrl = torch.as_tensor(PeerFrame._Raw(p.value, (self.CTRL_BYTES // 4,)), device=torch.device("cuda", device))
            ctrl.zero_()
            torch.cuda.synchronize(device)
            buf = (C.c_ubyte * 64)()
            _native.check(lib.rtgs_ipc_export(device, p, buf))
            handle = torch.tensor(list(buf), dtype=torch.uint8)
        if world > 1:
            h = handle.to(torch.device("cuda", device))
            dist.broadcast(h, src=0)
            handle = h.cpu()
        if rank != 0:
            buf = (C.c_ubyte * 64)(*handle.tolist())
            _native.check(lib.rtgs_ipc_open(device, buf, C.byref(p)))
            self._peer_ptr = p.value
        self.base = p.value
        dev = torch.device("cuda", device)
        self._images = [[torch.as_tensor(PeerFrame._Raw(self.base + self.CTRL_BYTES + (b * slots + k) * self.image_bytes,
                                                        (W, H, 3)), device=dev) for k in range(slots)]
                        for b in range(buffers)]
        self.frames = 0          # frames begun on this rank
        self.released = 0        # rank 0: frames released

    # -- addresses of the counters (device pointers valid on this rank)
    def _arrive(self, b):
        return self.base + 128 * b

    @property
    def _consumed(self):
        return self.base + 128 * self.buffers

    def begin(self, scene, slot: int = 0):
        """Start this rank's next frame: arm the scene's hand-over counters and return the (W,H,3) tensor to render
        into (rank 0's memory, peer-mapped here)."""
        self.frames += 1
        f = self.frames
        b = f % self.buffers
        scene.set_frame_sync(arrive=self._arrive(b), grant=self._consumed if f > self.buffers else 0,
                             grant_value=f - self.buffers)
        return self._images[b][slot]

    def begin_copy(self):
        """Bulk-copy form of the tile gather: start this rank's next frame WITHOUT arming the render kernels' hand-over
        (the frame is rendered into local memory); returns nothing.  Follow the render with ``deliver_stripes``."""
        self.frames += 1

    def deliver_stripes(self, local_image, stream=None):
        """Copy this rank's stripes of ``local_image`` (a (W,H,3) CUDA tensor) into the shared frame with one strided
        copy over NVLink, then signal the arrival - both stream-ordered."""
        import torch

        from . import _native
        f = self.frames
        b = f % self.buffers
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if f > self.buffers:      # the buffer is being reused: wait (on the stream) until its previous frame is consumed
            _native.check(self._lib.rtgs_stream_wait_counter(self.device, self._consumed, f - self.buffers, st))
        _native.check(self._lib.rtgs_copy_stripes_d2d(self.device, self._images[b][0].data_ptr(), local_image.data_ptr(),
                                                      self.W, self.H, self.world, self.rank, st))
        _native.check(self._lib.rtgs_stream_add_counter(self.device, self._arrive(b), st))

    def frame(self, slot: int = 0):
        """The tensor of the frame begun last (on rank 0: complete after ``wait()``)."""
        return self._images[self.frames % self.buffers][slot]

    @staticmethod
    def uses_of_buffer(frame: int, buffers: int) -> int:
        """How many of the frames 1 .. `frame` (this one included) went into the buffer frame `frame` uses
        (frame f uses buffer f % buffers): what every rank's arrive counter of that buffer must have reached."""
        b = frame % buffers
        return len(range(b if b else buffers, frame + 1, buffers))

    def wait(self, stream=None):
        """Rank 0: make `stream` (default: torch's current stream) wait until every rank's part of the frame begun
        last has landed."""
        import torch
        assert self.rank == 0
        f = self.frames
        b = f % self.buffers
        uses = self.uses_of_buffer(f, self.buffers)
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        from . import _native
        _native.check(self._lib.rtgs_stream_wait_counter(self.device, self._arrive(b), self.world * uses, st))

    def release(self, stream=None):
        """Rank 0: the consumer is done with the frame begun last (stream-ordered)."""
        import torch
        assert self.rank == 0
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        from . import _native
        self.released = self.frames
        _native.check(self._lib.rtgs_stream_set_counter(self.device, self._consumed, self.released, st))

    def close(self):
        self._images = None
        if self._peer_ptr:
            self._lib.rtgs_ipc_close(self.device, self._peer_ptr)
            self._peer_ptr = None
        if self._owner_ptr:
            self._lib.rtgs_device_free(self.device, self._owner_ptr)
            self._owner_ptr = None


class HostFrame:
    """Tile sharding, delivery to the HOST: one (W,H,3) float32 image per buffer in POSIX shared memory, page-locked by
    every rank (rtgs_host_register), into which each rank's GPU copies its own stripes over its OWN PCIe link
    (rtgs_copy_stripes_d2h: one strided DMA per rank and frame).  A 1080p frame that takes 0.44 ms over one link
    takes 1/world of that; nothing is gathered on a device first and there is no collective.

    Hand-over is by flags in the same shared memory: behind its copy every rank's stream stores the frame number to
    ``done[buffer][rank]`` (rtgs_stream_store_u32, one writer per flag); rank 0's host thread polls them in
    ``wait()`` and marks the buffer free again with ``release()``; a producer that would overwrite a buffer still in
    use spins in ``deliver()`` (with ``buffers`` >= 3 it never does in a steady sweep).

    ``dist`` is only used to hand the segment's name to the other ranks at construction."""

    FLAG_BYTES = 4096

    def __init__(self, W: int, H: int, rank: int, world: int, device: int, dist=None, buffers: int = 3):
        import ctypes as C
        import mmap
        import os

        from . import _native
        self._lib = _native.load()
        self.W, self.H, self.rank, self.world, self.device, self.buffers = W, H, rank, world, device, buffers
        self.image_bytes = W * H * 12
        nbytes = self.FLAG_BYTES + buffers * self.image_bytes
        names = [f"/dev/shm/rtgs_hostframe_{os.getpid()}_{id(self) & 0xffffff:x}"] if rank == 0 else [None]
        if rank == 0:
            fd = os.open(names[0], os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, nbytes)
        if world > 1:
            dist.broadcast_object_list(names, src=0)
        self._path = names[0]
        if rank != 0:
            fd = os.open(self._path, os.O_RDWR)
        self._mm = mmap.mmap(fd, nbytes)
        os.close(fd)
        self._addr = C.addressof(C.c_char.from_buffer(self._mm))
        _native.check(self._lib.rtgs_host_register(self._addr, nbytes))
        dp = C.c_void_p()
        _native.check(self._lib.rtgs_host_device_pointer(self._addr, C.byref(dp)))
        self._flags_dev = dp.value
        self.flags = np.frombuffer(self._mm, dtype=np.uint32, count=self.FLAG_BYTES // 4)   # [b*64 + r] done, [1000] consumed
        self.images = [np.frombuffer(self._mm, dtype=np.float32, count=W * H * 3,
                                     offset=self.FLAG_BYTES + b * self.image_bytes).reshape(W, H, 3) for b in range(buffers)]
        assert world <= 64 and buffers * 64 <= 1000
        if rank == 0:
            self.flags[:] = 0
        if world > 1:
            dist.barrier()
        self.frames = 0        # frames delivered by this rank
        self.collected = 0     # rank 0: frames waited for

    def deliver(self, dev_image, stream=None):
        """Queue the copy of this rank's stripes of ``dev_image`` (a (W,H,3) CUDA tensor holding them at their
        full-frame positions) into the next host buffer, followed by the done flag."""
        import torch

        from . import _native
        self.frames += 1
        f = self.frames
        b = f % self.buffers
        while f > self.buffers and int(self.flags[1000]) < f - self.buffers:     # the buffer's previous frame is still in use
            pass
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        host = self._addr + self.FLAG_BYTES + b * self.image_bytes
        _native.check(self._lib.rtgs_copy_stripes_d2h(self.device, host, dev_image.data_ptr(), self.W, self.H,
                                                      self.world, self.rank, st))
        _native.check(self._lib.rtgs_stream_store_u32(self.device, self._flags_dev + 4 * (b * 64 + self.rank), f, st))

    def wait(self):
        """Rank 0: block until every rank's stripes of the oldest frame not yet collected are in host memory; returns
        that frame as a (W,H,3) float32 array over the shared buffer (valid until ``release()``)."""
        assert self.rank == 0
        self.collected += 1
        f = self.collected
        b = f % self.buffers
        done = self.flags[b * 64: b * 64 + self.world]
        while int(done.min()) < f:
            pass
        return self.images[b]

    def release(self):
        """Rank 0: the frame returned by the last ``wait()`` may be overwritten."""
        self.flags[1000] = self.collected

    def close(self):
        import os
        if self._mm is not None:
            self._lib.rtgs_host_unregister(self._addr)
            self.flags = self.images = None
            try:
                self._mm.close()
            except BufferError:
                pass
            self._mm = None
            if self.rank == 0:
                try:
                    os.unlink(self._path)
                except OSError:
                    pass
