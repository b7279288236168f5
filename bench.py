#!/usr/bin/env python
"""bench.py -- Mrays/s of the wavefront path tracer on the Sponza-sized synthetic atrium, 1920x1080, 8 bounces.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one PathTracer.Compute() (1 sample per pixel, RayDepth 9) over the whole frame. With N GPUs the frame is
cut into 8-row stripes dealt round-robin to the ranks (scene replicated), followed by ONE all-gather of tile radiance.
Rank 0 prints ONE JSON line (see DESIGN.md "Measurement" for every field).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOAD_TRIS = 262144
WIDTH, HEIGHT = 1920, 1080
RAY_DEPTH = 9            # RayDepth = 1 + bounces (PathTracer.cs:228)
STRIPE = 8
SKY = (0.6, 0.7, 0.9)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--tris", type=int, default=WORKLOAD_TRIS)
    ap.add_argument("--ray-depth", type=int, default=RAY_DEPTH)
    ap.add_argument("--sort", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_traffic():
    """DRAM bytes per launch of the roofline kernel from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(args):
    from idkengine_b200 import scenes
    scene, cam = scenes.atrium(args.tris)
    frame = scenes.camera_frame(cam, args.width, args.height)
    return scene, cam, frame


def workload_config(args, world, scene):
    info = scene.build_info[0]
    return {
        "workload": f"atrium-{args.tris // 1000}k (Sponza-sized synthetic, seed 0x1D4E) {args.width}x{args.height} RayDepth {args.ray_depth} "
                    f"({args.ray_depth - 1} bounces) 1spp RR on, sort {'on' if args.sort else 'off'}, constant sky, constant textures",
        "triangles": int(info["source_triangles"]), "blas_triangles": int(info["triangles"]), "blas_nodes": int(info["nodes"]),
        "blas_stack_size": int(scene.blas_stack_size),
        "parallelism": f"screen tiles: {STRIPE}-row stripes round-robin over {world} GPU(s), scene replicated, tile gather fused into "
                       f"FinalDraw over NVLink peer memory (IDKPT_GATHER=nccl: one NCCL all-gather per frame)",
        "l2": "per-step working set (2x64B path state + hits, ~330 MB at 1080p) exceeds the 126 MB L2; the 22 MB BVH is "
              "re-read every bounce and is L2-resident by design (SURVEY 8d)",
    }


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args, rank, world):
    """The CPU restatement of the path (oracle port: FirstHit/NHit/FinalDraw + BVHIntersect on host cores, all threads)
    on the same workload; each step = a bounded sample (1/8 of the frame as an interleaved stripe tile)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    from idkengine_b200 import capi
    scene, cam, frame = build_scene(args)
    s = capi.default_settings()
    s.RayDepth = args.ray_depth
    s.DoRaySorting = args.sort
    threads = os.cpu_count() or 1
    tile = (STRIPE, 0, 8)
    img = np.zeros((args.height, args.width, 4), np.float32)
    acc, rays, secs = 0, 0, 0.0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o = ol.path_trace(scene, frame, s, args.width, args.height, sky=SKY, tile=tile, accumulated=acc, result=img,
                          want_rays=False, threads=threads)
        dt = time.perf_counter() - t0
        acc = o.accumulated
        if i >= args.warmup:
            rays += o.stats.Rays
            secs += dt
    value = rays / secs / 1e6
    sample = f"1/8 of the frame per step (stripe tile {tile}), {args.steps} steps, {rays} rays"
    line = {
        "impl": "reference", "metric": "Mrays/s Sponza-sized synthetic 1920x1080 8-bounce", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1, scene),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(scene, frame, args):
    """The reference's own CPU traversal shape, Gui.Test (Gui.cs:1484-1503): one primary ray per pixel through
    BVH.Intersect -> BLAS.Intersect (C# semantics), rows in parallel on every host core."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    threads = os.cpu_count() or 1
    rays = ol.gui_test_rays(frame, args.width, args.height)
    best = None
    for _ in range(3):
        _, secs = ol.cpu_intersect(scene, rays, threads=threads, want_hits=False)
        best = secs if best is None else min(best, secs)
    return {"value": len(rays) / best / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"{len(rays)} primary rays (one full {args.width}x{args.height} frame, Gui.Test shape), best of 3, "
                      f"restated BVH.Intersect/BLAS.Intersect (C# semantics)"}


# ----------------------------------------------------------------------------------------------- our arm (GPU)
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from idkengine_b200 import capi, multigpu
    from idkengine_b200.pathtracer import PathTracer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libidkpt has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    scene, cam, frame = build_scene(args)
    s = capi.default_settings()
    s.RayDepth = args.ray_depth
    s.DoRaySorting = args.sort
    pt = PathTracer(args.width, args.height, s, device=local_rank, tile=(STRIPE, rank, world))
    pt.SetScene(scene)
    pt.SetSky(SKY)
    pt.SetFrame(frame)
    rows = pt.TileRows()
    ptr, nbytes = pt.ResultDevicePtr()
    local = torch.as_tensor(multigpu.DeviceArray(ptr, (len(rows), args.width, 4)), device=dev)
    pinned = [torch.empty((args.height, args.width, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    # N > 1: the tile all-gather is fused into Compute() over NVLink peer memory (CUDA IPC); IDKPT_GATHER=nccl selects the
    # torch.distributed all_gather + de-interleave fallback instead.
    peer_gather = world > 1 and os.environ.get("IDKPT_GATHER", "peer") != "nccl"
    gatherer = multigpu.TileGatherer(args.height, args.width, 4, STRIPE, world, dev) if (world > 1 and not peer_gather) else None
    if peer_gather:
        def exchange(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        pt.EnablePeerGather(rank, world, exchange)

    lanes = int(os.environ.get("IDKPT_LANES", "8"))      # the library default
    pipelined = lanes > 1 and (world == 1 or peer_gather)      # the NCCL fallback gathers between steps: one sample at a time

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e, k=0):
        """e2e: the frame's GpuPerFrameData goes host->device inside Compute(); the finished image of EVERY step is read
        back into pinned host memory (double-buffered, asynchronously, so the transfer overlaps the next step)."""
        if pipelined:
            pt.ComputeAsync()         # stats == NULL: queued, up to `lanes` samples in flight
            st = None
        else:
            st = pt.Compute()
        if world == 1:
            if e2e:
                pt.PresentAsync(pinned[k & 1].data_ptr(), pinned[k & 1].numel() * 4)
        elif peer_gather:
            if e2e and rank == 0:
                pt.PresentAsync(pinned[k & 1].data_ptr(), pinned[k & 1].numel() * 4, which=capi.IDKPT_IMAGE_GATHERED)
        else:
            torch.cuda.current_stream().wait_stream(copy_stream)     # the previous read-back still reads gatherer.full
            full = gatherer.gather(local)
            if e2e and rank == 0:
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    pinned[k & 1].copy_(full, non_blocking=True)
        return st

    def e2e_drain():
        if world == 1 or peer_gather:
            pt.PresentWait()
            pt.Sync()
        else:
            copy_stream.synchronize()

    # ---- stats replay (untimed): exact S/T/I for the sample sequence the timed region will run
    pt.CollectStats = 1
    pt.ResetAccumulation()
    S = T = I = R = 0
    for _ in range(args.steps):
        rs = pt.Compute()
        S += rs.NodePairFetches; T += rs.TriangleTests; I += rs.InstanceVisits; R += rs.Rays
    pt.CollectStats = 0

    for k in range(max(args.warmup, 3)):     # warm-up exercises the end-to-end path too (first present allocates the snapshot
        step(True, k)                        # buffer, the copy stream and touches the pinned pages)
    e2e_drain()

    # ---- timed region 1a: inputs resident in HBM, no read-back, ONE sample at a time with per-kernel CUDA events
    # (the roofline's kernel durations come from here; with pipelining off this is also `value`)
    pt.ResetAccumulation()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms = trav_ms = shade_ms = 0.0
    rays = launches = trav_launches = 0
    gather_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = pt.Compute()
        dev_ms += st.TotalMs; trav_ms += st.TraverseMs; shade_ms += st.ShadeMs
        rays += st.Rays; launches += st.KernelLaunches; trav_launches += st.TraverseLaunches
        if world > 1 and not peer_gather:
            ev0.record()
            gatherer.gather(local)
            ev1.record()
            ev1.synchronize()
            gather_ms += ev0.elapsed_time(ev1)
            launches += 1
        elif world > 1:
            gather_ms += st.OtherMs        # ray-gen + fused accumulate/scatter + arrival wait (inside Compute)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    assert rays == R, "timed region traced a different ray set than the stats replay"

    # ---- timed region 1b: the same K steps queued asynchronously (several samples in flight). Timed on the device with
    # events on the context's main stream: every sample's FinalDraw runs there in submission order.
    pipe_ms = None
    if pipelined:
        ext = torch.cuda.ExternalStream(pt.StreamHandle(), device=dev)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pt.ResetAccumulation()
        barrier()
        pe0.record(ext)
        for _ in range(args.steps):
            pt.ComputeAsync()
        pe1.record(ext)
        pt.Sync()
        barrier()
        pipe_ms = pe0.elapsed_time(pe1)
        assert pt.AccumulatedSamples == args.steps

    # ---- timed region 2: end to end through the public API with host buffers (frame H2D, result D2H every step)
    pt.ResetAccumulation()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(True, k)
    e2e_drain()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)
    serial_ms = reduce(dev_ms + (0.0 if peer_gather else gather_ms), MAX)     # device time of region 1a, max over ranks
    job_ms = reduce(pipe_ms, MAX) if pipelined else serial_ms                 # device time of the K timed steps, max over ranks
    # nvidia-smi samples every 100 ms but K steps may last only tens of ms: keep the identical load running (untimed) until
    # the sampler has seen ~0.6 s of it, then stop it. The step count is derived from the reduced time, i.e. equal on all ranks.
    extra_steps = int(min(600, max(0, 600.0 / max(job_ms / args.steps, 1e-3) - args.steps)))
    for k in range(extra_steps):
        step(False)
    if pipelined:
        pt.Sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = f"the {args.steps} timed steps + {extra_steps} identical untimed steps (nvidia-smi -lms 100)"
    per_rank = None
    if world > 1:      # per-rank device times: shows tile imbalance / a slow GPU behind the max-over-ranks figure
        mine = {"rank": rank, "pipelined_total": (pipe_ms / args.steps) if pipelined else None, "total": dev_ms / args.steps, "traverse": trav_ms / args.steps, "shade": shade_ms / args.steps,
                "gather": gather_ms / args.steps, "rays": rays // args.steps}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    total_rays = reduce(rays, SUM)
    total_launches = int(reduce(launches, SUM))
    e2e_ms = reduce(e2e_ms, MAX)
    wall_ms = reduce(wall_ms, MAX)

    if rank == 0:
        peak, peak_src = measured_peaks()
        trav_bytes = 64 * S + 52 * T + 48 * I + 52 * R      # algorithmic bytes of k_traverse (rank 0's tile)
        achieved = trav_bytes / (trav_ms * 1e-3) / 1e9
        line = {
            "metric": "Mrays/s Sponza-sized synthetic 1920x1080 8-bounce", "value": total_rays / (job_ms * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": job_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, scene),
            "rays_per_step": total_rays / args.steps,
            "serial_ms_per_step": serial_ms / args.steps,
            "pipeline": (f"{lanes} samples in flight: asynchronous Compute (stats == NULL), per-sample results and accumulation order identical to the "
                         f"serial path" if pipelined else "off: one sample at a time"),
            "wall_ms_per_step": wall_ms / args.steps,
            "kernel_ms_per_step": {"measured_in": "serial pass (region 1a: one sample in flight, per-kernel CUDA events)", "traverse": trav_ms / args.steps, "shade": shade_ms / args.steps,
                                   "all_gather": gather_ms / args.steps, "total_device": dev_ms / args.steps},
            "roofline": {"kernel": "k_traverse2 (bounces) + k_traverse (primary rays)", "measured_in": "serial pass of the same K steps (kernels of different samples overlap in the pipelined pass)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(), "traffic_source": "profiles/traffic.json (ncu --set full, heaviest launches)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": trav_bytes / max(trav_launches, 1),
                         "launch_ms": trav_ms / max(trav_launches, 1),
                         "per_ray": {"node_pair_fetches": S / R, "triangle_tests": T / R, "instances": I / R}},
            "e2e": {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 544 + 44, "d2h_bytes_per_step": args.width * args.height * 16},
            "gpu_launches": total_launches,
            "clocks": clocks,
        }
        if per_rank is not None:
            line["per_rank_ms"] = per_rank
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(scene, frame, args)
        print(json.dumps(line))
    pt.Dispose()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
