#!/usr/bin/env python
"""Device-to-host ceiling of the box, nothing else: N processes (torchrun, one per GPU), each with one device
buffer and one pinned host buffer of a 1080p float32 RGB frame (24.9 MB), plain cudaMemcpyAsync loops, no kernels.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        tools/d2h_probe.py [--mb 24.8832] [--seconds 1.5]

Prints one JSON line on rank 0: per-rank and aggregate GB/s for three kinds of host memory - cudaHostAlloc default,
cudaHostAllocWriteCombined, and a cudaHostRegister'ed POSIX shared-memory mapping (what a multi-process gather into
one host image uses).  VERDICT r1 item 2: is ~100 GB/s in total really the platform's limit?"""
import argparse
import ctypes as C
import json
import mmap
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=1920 * 1080 * 12 / 1e6)
    ap.add_argument("--seconds", type=float, default=1.5)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    rt = C.CDLL("libcudart.so.12")
    nbytes = int(args.mb * 1e6) // 4096 * 4096
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.fill_(1)
    stream = C.c_void_p()
    assert rt.cudaStreamCreateWithFlags(C.byref(stream), 1) == 0
    D2H = 2

    def bench(host_ptr):
        def burst(n):
            for _ in range(n):
                assert rt.cudaMemcpyAsync(C.c_void_p(host_ptr), C.c_void_p(dev.data_ptr()), C.c_size_t(nbytes), D2H, stream) == 0
            assert rt.cudaStreamSynchronize(stream) == 0
        burst(4)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < args.seconds:      # every rank copies for the same wall-clock window
            burst(8)
            n += 8
        dt = time.perf_counter() - t0
        return n * nbytes / dt / 1e9

    res = {}
    for kind, flags in (("pinned", 0), ("write_combined", 4)):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), flags) == 0
        res[kind] = bench(p.value)
        rt.cudaFreeHost(p)
    fd = os.open(f"/dev/shm/rtgs_d2h_probe_{rank}", os.O_CREAT | os.O_RDWR, 0o600)
    os.ftruncate(fd, nbytes)
    mm = mmap.mmap(fd, nbytes)
    addr = C.addressof(C.c_char.from_buffer(mm))
    C.memset(addr, 0, nbytes)
    assert rt.cudaHostRegister(C.c_void_p(addr), C.c_size_t(nbytes), 1 | 2) == 0   # portable | mapped
    res["registered_shm"] = bench(addr)
    rt.cudaHostUnregister(C.c_void_p(addr))
    os.unlink(f"/dev/shm/rtgs_d2h_probe_{rank}")

    vals = torch.tensor([res[k] for k in sorted(res)], dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        out = {"gpus": world, "mb_per_copy": nbytes / 1e6, "host_cores": os.cpu_count()}
        for i, k in enumerate(sorted(res)):
            per = [float(v[i]) for v in allv]
            out[k] = {"total_gbs": round(sum(per), 1), "per_gpu_gbs": [round(x, 1) for x in per]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
