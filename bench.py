#!/usr/bin/env python
"""bench.py — Mrays/s of the fused per-ray render path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

Workload (config 3 of BASELINE.json, the one the metric is quoted on): synthetic 1 M random
Gaussians, SH degree 3, seed 1002, 1920x1080, vertical fov 60 deg, orbit radius 2.2, depth 16.
A "step" = ONE full 1080p frame (ray generation + LBVH traversal + intersection + 16-nearest
k-buffer + SH compositing + framebuffer write) of one view of the 64-view orbit (config 5); the
step s on rank r renders view (s*N + r) mod 64, view 0 being config 3's camera.  Multi-GPU is
view-sharded with the scene replicated on every GPU and no collective on the data path, so
per-GPU work is fixed as N grows ("scaling": "weak").  One ray = one finished pixel.

value  = whole-job Mrays/s with everything resident in HBM (device-timed, CUDA events, max over ranks)
e2e    = the same through the public API with HOST buffers: the orbit sweep as a user writes it with
         RayTracer.render_async()/result() (camera struct in, image delivered to pinned host memory, every step,
         inside the timed region; two frames in flight so that step s+1 renders while the tail of step s crosses
         PCIe); e2e.sync_value = the same loop with the blocking RayTracer.render()
roofline = the dominant kernel, k_shade_tiles (intersection + k-buffer + SH compositing): algorithmic bytes per
         ray (16 + kbar*(64 + 192[sh]), all of them consumed by this kernel) x rays per launch /
         its mean duration, measured with CUDA events the library records around each kernel on the render
         stream during the timed region, against the measured HBM copy bandwidth in MEASURED_PEAKS.json;
         "kernels" lists every kernel of the step with its mean time and share
cpu_baseline = oracle/ref_cpu.cpp (reference-shaped C++/OpenMP port, float32) on a pixel subsample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "rt-gaussian-splat-renderer_b200"))

N_VIEWS = 64
DEPTH = 16
T_CUT = 1e-4
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="1m_deg3_1080p")
    ap.add_argument("--sharding", default="views", choices=["views", "tiles"],
                    help="multi-GPU partition: camera views (weak scaling, default) or 32-column stripes of every "
                         "frame gathered on rank 0 (strong scaling)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="--sharding tiles: every rank stores its stripes straight into rank 0's framebuffer over "
                         "NVLink (CUDA IPC peer mapping, default) or packs them for one NCCL gather")
    ap.add_argument("--h-target", type=float, default=None,
                    help="diagnostic: expected ellipsoid crossings per cube-spanning ray of the synthetic scene "
                         "(default 16, SURVEY.md 8d); larger = bigger Gaussians, denser tiles")
    ap.add_argument("--morton-bits", type=int, default=0, choices=[0, 30, 63],
                    help="LBVH code width: 0 = automatic (30, the north-star spec, unless the codes are degenerate), "
                         "30, 63 (the wide variant for scenes with outliers)")
    ap.add_argument("--outliers", type=int, default=0,
                    help="diagnostic: move this many Gaussians ~4000 scene radii away (what stray points of a "
                         "trained scene do to 10-bit-per-axis Morton codes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-GPU: do not pin each rank to the CPU cores next to its GPU")
    ap.add_argument("--cpu-stride", type=int, default=0, help="pixel subsample stride of the CPU legs (0 = auto)")
    return ap.parse_args()


def make_views(W, H):
    from rtgs.orbit import focal_from_fov, orbit_pose
    from rtgs.synthetic import FOV_DEG, ORBIT_R
    f = focal_from_fov(H, FOV_DEG)
    return f, [orbit_pose(2 * np.pi * k / N_VIEWS, np.pi / 2, ORBIT_R) for k in range(N_VIEWS)]


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few ms DURING the timed region
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints)."""

    def __init__(self, index, period_s=0.002):
        self.index, self.period, self.samples, self.reasons = index, period_s, [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.power_w = []

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                if len(self.samples) % 8 == 1:   # the power query can take tens of milliseconds
                    self.power_w.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power_w) if self.power_w else None}


def load_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu summary (profiles/), if any."""
    p = ROOT / "profiles" / "render_kernel_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())[kernel]["dram_bytes_per_launch"]
        except Exception:
            return None
    return None


def cpu_leg(cfg_name, scene_arrays, W, H, focal, views, steps, warmup, stride):
    """The reference-shaped CPU port (float32, all host threads) on a pixel subsample."""
    from oracle import ref_cpu
    from oracle import ref_numpy as O
    ref_cpu.build()
    t0 = time.perf_counter()
    cs = ref_cpu.CpuScene(scene_arrays["pos"], scene_arrays["rot"], scene_arrays["scale"], scene_arrays["color"],
                          scene_arrays["opacity"], scene_arrays["sh"])
    build_s = time.perf_counter() - t0
    pix = ref_cpu.all_pixels(W, H, stride)
    times = []
    for s in range(warmup + steps):
        pos, rot = views[s % N_VIEWS]
        cam = O.CameraParams(np.asarray(pos), np.asarray(rot), W, H, (focal, focal))
        t0 = time.perf_counter()
        cs.render(cam, DEPTH, pixels=pix, precision="float")
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    return {"mrays": pix.shape[0] * len(times) / total / 1e6, "ms_per_step": 1e3 * total / len(times),
            "cores": ref_cpu.max_threads(), "rays_per_step": int(pix.shape[0]), "build_s": build_s,
            "sample": f"every {stride}th column and row of each {W}x{H} view ({pix.shape[0]} rays/step), "
                      f"{len(times)} steps, float32, K={DEPTH} closest-hit restarts over the LBVH"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from rtgs.synthetic import CONFIGS, make_scene
    n_g, seed, sh_deg, (W, H) = CONFIGS[args.config]
    config = {"workload": f"synthetic {n_g} random Gaussians, SH degree {sh_deg}, seed {seed}, {W}x{H}, fov 60, "
                          f"orbit r=2.2, depth {DEPTH}, t_cut {T_CUT}; 64-view orbit, view (step*N+rank)%64",
              "gaussians": n_g, "sh_degree": sh_deg, "resolution": [W, H], "depth": DEPTH,
              "sharding": "camera views (scene replicated, no collective)" if args.sharding == "views" else
                          "32-column stripes of every frame, dealt round-robin (scene replicated; the finished "
                          "stripes are gathered on rank 0 after the render)",
              "l2": "inputs larger than L2 (packed scene 528 MB > 126 MB) and a new view every step"}

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return 0
        arrays = make_scene(n_g, seed, sh_deg)
        focal, views = make_views(W, H)
        # bounded sample: every 4th column and row (1/16 of the rays) keeps 100 steps within ~1 minute of CPU time
        stride = args.cpu_stride or (4 if W * H <= 1920 * 1080 else 8)
        r = cpu_leg(args.config, arrays, W, H, focal, views, args.steps, args.warmup, stride)
        line = {"impl": "reference", "metric": "Mrays/s", "value": r["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["mrays"], "unit": "Mrays/s", "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"]},
                "e2e": {"value": r["mrays"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "the reference is Python+Taichi and cannot run here; this is oracle/ref_cpu.cpp, a C++/OpenMP "
                        "port with the reference's algorithm shape, on the host cores"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    from rtgs.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else {"bound": False,
                                                                                         "why": "not requested"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from rtgs.camera import Camera
    from rtgs.ray_tracer import RayTracer
    from rtgs.scene import Scene

    arrays = make_scene(n_g, seed, sh_deg) if args.h_target is None else make_scene(n_g, seed, sh_deg, args.h_target)
    if args.h_target is not None:
        config["workload"] += f" [diagnostic: h_target {args.h_target}]"
    if args.outliers > 0:
        rng = np.random.default_rng(99)
        arrays["pos"][:args.outliers] = (rng.uniform(-1, 1, (args.outliers, 3)) * 4000.0).astype(np.float32)
        config["workload"] += f" [diagnostic: {args.outliers} outliers at ~4000 scene radii]"
    if args.morton_bits == 63:
        config["workload"] += " [LBVH with 63-bit Morton codes]"
    focal, views = make_views(W, H)
    t0 = time.perf_counter()
    scene = Scene(device=local_rank, morton_bits=args.morton_bits or "auto").from_arrays(arrays["pos"], arrays["rot"], arrays["scale"], arrays["color"],
                                                 arrays["opacity"], arrays["sh"])
    torch.cuda.synchronize()
    build_ms = 1e3 * (time.perf_counter() - t0)
    cam = Camera(views[0][0], views[0][1], (W, H), (focal, focal), device=local_rank)
    rt = RayTracer((W, H), scene, cam, t_cut=T_CUT)
    out = torch.empty((W, H, 3), dtype=torch.float32, device="cuda")

    tiles = args.sharding == "tiles"
    gather = None
    peer = None
    if tiles:
        from rtgs.sharding import PeerFrame, StripeGather
        scene.set_stripe(world, rank)
        if args.gather == "peer":
            peer = PeerFrame(W, H, rank, world, local_rank, dist)
            out = peer.tensor            # rank 0's image, peer-mapped on the other ranks
        else:
            gather = StripeGather(W, H, rank, world, torch.device("cuda", local_rank))
        config["sharding"] += f" [{args.gather}]"

    def set_view(step):
        v = step % N_VIEWS if tiles else (step * world + rank) % N_VIEWS
        cam.position, cam.rotation = views[v]
        return v

    def render_step():
        rt.render_device(DEPTH, out=out)
        if tiles and dist is not None:
            if peer is not None:
                peer.finish(dist)
            else:
                gather(out, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # stats pass (untimed): kbar, hit fraction, traversal counters for the timed views (whole frames)
    scene.set_stripe()
    agg = {}
    for s in range(min(args.steps, N_VIEWS)):
        set_view(s)
        rt.render_device(DEPTH, out=out, collect_stats=True)
        for k, v in rt.last_stats.items():
            agg[k] = agg.get(k, 0) + v
    if tiles:
        scene.set_stripe(world, rank)
    kbar = agg["layers"] / agg["rays"]
    bytes_ray = 16 + kbar * (64 + (192 if sh_deg > 0 else 0))

    for s in range(args.warmup):
        set_view(s)
        render_step()
    barrier()
    tiles_verified = None
    if tiles:
        # untimed check: the frame assembled from all ranks' stripes == rank 0's own full-frame render, bit for bit
        set_view(0)
        render_step()
        barrier()
        if rank == 0:
            scene.set_stripe()
            full = rt.render_device(DEPTH)
            scene.set_stripe(world, rank)
            torch.cuda.synchronize()
            tiles_verified = bool(torch.equal(full, out))
        barrier()
    timed_frames = min(args.steps, 4096)
    scene.set_option("kernel_timing", timed_frames)   # events around every kernel of the timed steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    e_beg.record()
    for s in range(args.steps):
        set_view(s)
        ev[s][0].record()
        render_step()
        ev[s][1].record()
        launches += 3     # k_tile_lists, k_shade_tiles, k_render (fallback list; returns at once when empty)
    e_end.record()
    barrier()
    total_ms = e_beg.elapsed_time(e_end)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    per_kernel = scene.read_kernel_times(timed_frames).astype(np.float64).mean(axis=0)   # ms: lists, shade, fused
    scene.set_option("kernel_timing", 0)

    # end-to-end: public API call with host buffers (camera in, image out on the host)
    if tiles:
        host = torch.empty((W, H, 3), dtype=torch.float32, pin_memory=True) if rank == 0 else None

        def e2e_step():
            render_step()
            if rank == 0:
                host.copy_(out, non_blocking=True)
                torch.cuda.synchronize()
    else:
        def e2e_step():
            return rt.render(DEPTH)
    for s in range(min(args.warmup, 2)):
        set_view(s)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        set_view(s)
        e2e_step()
    barrier()
    e2e_s = e2e_sync_s = time.perf_counter() - t0
    e2e_kernels = None
    if not tiles:
        # the sweep API: every step still uploads its camera and delivers its image to pinned host memory, but
        # two frames are in flight (RayTracer.render_async), so step s+1 renders while the tail of step s is
        # copied out.  Every image is collected inside the timed region.
        def pipelined(n):
            prev = None
            for s in range(n):
                set_view(s)
                cur = rt.render_async(DEPTH)
                if prev is not None:
                    prev.result()
                prev = cur
            prev.result()
        pipelined(min(args.warmup, 3))
        barrier()
        t0 = time.perf_counter()
        pipelined(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
    # the sampler has been running through all timed loops (device-timed steps and both end-to-end loops)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "device-timed steps + end-to-end loops"
    if not tiles:
        # untimed repeat with the library's per-kernel events switched on: what the kernels cost in this mode
        scene.set_option("kernel_timing", timed_frames)
        pipelined(timed_frames)
        e2e_kernels = scene.read_kernel_times(timed_frames).astype(np.float64).mean(axis=0)
        scene.set_option("kernel_timing", 0)

    if dist is not None:
        t = torch.tensor([total_ms, e2e_s, e2e_sync_s, float(np.mean(kern_ms)), *per_kernel.tolist()],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_sync_s, kern_mean, *pk = t.tolist()
        per_kernel = np.asarray(pk)
    else:
        kern_mean = float(np.mean(kern_ms))

    if rank == 0:
        rays_step = W * H * (1 if tiles else world)
        value = rays_step * args.steps / (total_ms * 1e-3) / 1e6
        e2e_val = rays_step * args.steps / e2e_s / 1e6
        peak, peak_src = load_peak()
        from rtgs._native import KERNEL_NAMES
        dom = int(np.argmax(per_kernel))
        dom_ms = float(per_kernel[dom])
        rays_launch = W * H / (world if tiles else 1)      # rays one launch of the kernel processes
        achieved = rays_launch * bytes_ray / (dom_ms * 1e-3) / 1e9
        step_ms = float(per_kernel.sum())
        kernels = [{"kernel": KERNEL_NAMES[k], "ms": float(per_kernel[k]), "share": float(per_kernel[k] / step_ms)}
                   for k in range(len(KERNEL_NAMES))]
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if tiles else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": 44 * world,
                    "d2h_bytes_per_step": W * H * 3 * 4 * (1 if tiles else world),
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "stripes gathered on rank 0, then one copy to pinned host memory" if tiles else
                           "RayTracer.render_async()/result(): two frames in flight, every image collected",
                    "sync_value": rays_step * args.steps / e2e_sync_s / 1e6,
                    "sync_api": "RayTracer.render() per step (blocking)",
                    "kernels_ms": None if e2e_kernels is None else [round(float(v), 5) for v in e2e_kernels]},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(KERNEL_NAMES[dom]), "peak_source": peak_src,
                         "kernel": KERNEL_NAMES[dom], "kernel_ms": dom_ms, "step_ms": kern_mean,
                         "bytes_per_ray": bytes_ray, "kbar": kbar,
                         "note": "algorithmic bytes assume zero reuse between rays; neighbouring rays share "
                                 "Gaussians through L1/L2, so achieved may exceed what DRAM moved (see traffic)"},
            "kernels": kernels,
            "clocks": clocks,
            "scene_stats": {"hit_fraction": agg["rays_hit"] / agg["rays"], "kbar": kbar,
                            "child_boxes_tested_per_ray": agg["nodes_tested"] / agg["rays"],
                            "candidates_per_tile": agg["candidates"] / max(agg["tiles"], 1),
                            "pair_tests_per_ray": agg["pair_tests"] / agg["rays"],
                            "f64_refinements_per_Mray": 1e6 * agg["f64_refinements"] / agg["rays"],
                            "traversal_steps_per_tile": agg["traversal_steps"] / max(agg["tiles"], 1),
                            "insert_rounds_per_tile": agg["insert_rounds"] / max(agg["tiles"], 1),
                            "useful_candidates_per_tile": agg["useful_candidates"] / max(agg["tiles"], 1),
                            "fallback_tiles": agg["fallback_tiles"]},
            "bvh_build_ms": scene.build_ms,          # device time of the LBVH build kernels
            "morton_bits": scene.morton_bits,
            "scene_load_ms": build_ms,               # wall: upload of the arrays + build (+ CUDA start-up on first use)
            "numa_bind_rank0": numa,
        }
        if tiles:
            line["tiles_verified_bit_identical"] = tiles_verified
        if world == 1 and not args.no_cpu_baseline:
            stride = args.cpu_stride or (4 if W * H <= 1920 * 1080 else 8)
            r = cpu_leg(args.config, arrays, W, H, focal, views, min(args.steps, 32), 1, stride)   # ~10-20 s of CPU work
            line["cpu_baseline"] = {"value": r["mrays"], "unit": "Mrays/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
