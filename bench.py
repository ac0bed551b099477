#!/usr/bin/env python
"""Benchmark of the stereo hot path (BASELINE.json metric: 1080p stereo frames/s, Mpix/s, % of HBM roofline,
host-CPU reference beside it).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU cores

Workload (N = 1 and per GPU for N > 1, weak scaling): BASELINE.json configs[1] -- 1920x1080, 'Fill - Polylines
Sharp', depth_map_blur on (20/20/2.0/6), convergence_point 0.5, divergence 3.5, left-right -- as a batch of
--frames frames per GPU per step (the roofline target is stated on a batch; a single 1080p
frame is 25 us at roofline and only measures launch latency).  A "step" is one pass of the hot path over the
batch.  Synthetic seeded frames, random image + ramp/discs/noise depth.

Printed JSON line (rank 0):
  value        frames/s with inputs and outputs resident in HBM (CUDA events, max over ranks, summed over GPUs)
  e2e          the same metric through StereoImageNode.generate with pinned HOST tensors in and host tensors
               out (H2D + kernels + D2H inside the timed region)
  roofline     dominant kernel: its algorithmic bytes per launch / its mean CUDA-event duration, against the
               measured HBM peak; `path` = the whole step's 80 B/px against the same peak
  cpu_baseline the oracle port of the reference algorithm on this box's cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

NODE_PARAMS = dict(divergence=3.5, separation=0.0, modes="left-right", stereo_balance=0.0, convergence_point=0.5,
                   stereo_offset_exponent=2.0, fill_technique="Fill - Polylines Sharp",
                   depth_blur_edge_threshold=20.0, depth_blur_strength=20.0, depth_map_blur=True,
                   depth_blur_falloff=2.0, depth_blur_vert_smooth=6, batch_size=12)
BYTES_PER_PX = {"cpu_sbs": 80, "gw_sbs": 76, "anaglyph": 64}  # inputs read once, outputs written once


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--fill", default=NODE_PARAMS["fill_technique"])
    ap.add_argument("--mode", default="left-right")
    ap.add_argument("--divergence", type=float, default=3.5)
    ap.add_argument("--balance", type=float, default=0.0, help="stereo_balance (BASELINE config 5: 0.5)")
    ap.add_argument("--separation", type=float, default=0.0)
    ap.add_argument("--convergence", type=float, default=0.5)
    ap.add_argument("--strong-frames", type=int, default=96, help="frames of the fixed batch of the strong-scaling leg (0 = skip)")
    ap.add_argument("--no-extra-e2e", action="store_true", help="skip the pageable-input / large-output e2e legs")
    ap.add_argument("--chunk", type=int, default=0, help="frames per kernel sequence (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=0,
                    help="frames of the CPU baseline sample (0 = 32 for the cpu_baseline leg, about 7 s of host work on 16 cores; "
                         "4 per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_frames(n, h, w, seed, first=0):
    """Frames [first, first + n) of the synthetic video: four distinct seeded frames, repeated with a horizontal roll that
    grows with the global frame index -- cheap to generate, different content in every frame, and every rank can produce
    any frame of the global batch on its own."""
    from comfystereo_b200 import synthetic as syn
    base = 4
    img = syn.make_image(base, h, w, seed=seed)
    dep = syn.make_depth(base, h, w, "scene", seed=seed)
    oi = np.empty((n, h, w, 3), np.float32)
    od = np.empty((n,) + dep.shape[1:], np.float32)
    for i in range(n):
        g = first + i
        roll = (g // base) * 8 % w
        oi[i] = np.roll(img[g % base], roll, axis=1) if roll else img[g % base]
        od[i] = np.roll(dep[g % base], roll, axis=1) if roll else dep[g % base]
    return oi, od


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(args, frames):
    """The oracle port of the reference algorithm (same node-level work: gray, blur, 2 x fill, compose, mask) on
    the host cores, `frames` frames of the benchmark workload.  One untimed call first (page-in, thread pool)."""
    import oracle as orc
    orc.lib()
    orc.set_threads(len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    img, dep = make_frames(frames, args.height, args.width, seed=100)
    params = node_params(args)
    orc.node_generate(img[:1], dep[:1], **params)
    t0 = time.perf_counter()
    orc.node_generate(img, dep, **params)
    dt = time.perf_counter() - t0
    return frames / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    cores = orc.set_threads(len(os.sched_getaffinity(0)))   # the threads the OpenMP loops really get
    sample = args.cpu_frames or 4
    times = []
    for i in range(args.warmup + args.steps):
        fps, dt = cpu_baseline(args, sample)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = sample / t
    line = {
        "impl": "reference", "metric": "1080p stereo frames/s", "value": val, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample),
        "mpix_per_s": val * args.height * args.width / 1e6,
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} frame(s) of the workload per step, oracle/stereo_oracle.c (OpenMP over rows), "
                                   "the reference itself is Python and does not exist on this box"},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def node_params(args):
    return dict(NODE_PARAMS, fill_technique=args.fill, modes=args.mode, divergence=args.divergence,
                stereo_balance=args.balance, separation=args.separation, convergence_point=args.convergence)


def workload_config(args, frames):
    fill = args.fill + (" [mesh warp]" if os.environ.get("COMFYSTEREO_GPU_WARP", "").lower() == "mesh" and
                        args.fill == "GPU Warp (Fast)" else "")
    return {"workload": f"{args.width}x{args.height} {fill}, depth_map_blur on (strength 20, threshold 20, "
                        f"falloff 2.0, vert 6), convergence {args.convergence}, divergence {args.divergence}, "
                        f"stereo_balance {args.balance}, separation {args.separation}, exponent 2, {args.mode} "
                        f"(BASELINE.json configs[1] as a batch)",
            "frames_per_gpu_per_step": frames, "l2": "inputs+outputs per step exceed the 126 MB L2 many times over "
                                                     "(no flush needed)"}


def host_copy_bandwidth(lib):
    """GB/s (read + write bytes) the host's memory system gives a team of threads streaming a large buffer into
    another with non-temporal stores -- the same loop the host path's result expansion runs."""
    return float(lib.cs_host_stream_bandwidth(1 << 30, 0))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import ctypes
    import torch
    import torch.distributed as dist
    from comfystereo_b200 import StereoImageNode, _lib, engine, stereoimage_generation

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    _lib.check(lib.cs_device_check())

    n, h, w = args.frames, args.height, args.width
    key = engine.FILL_NAME_TO_KEY.get(args.fill, "gpu_warp")
    if key == "gpu_warp" and stereoimage_generation.MODERNGL_AVAILABLE:     # COMFYSTEREO_GPU_WARP=mesh, as the node decides
        key = "gpu_warp_mesh"
    gpu_warp = key in engine.GPU_WARP_KEYS
    group = min(NODE_PARAMS["batch_size"], n) if gpu_warp else 0
    p = engine.make_params(key, args.mode, args.divergence, args.separation, args.balance, args.convergence, 2.0, True,
                           20.0, 20.0, 2.0, 6, group_size=group)
    # Frame-wise sharding: the global batch of n * world frames is cut into contiguous ranges with
    # engine.shard_range; this rank generates and processes only its own slice, no collective on the data path.
    glo, ghi = engine.shard_range(n * world, rank, world)
    img_np, dep_np = make_frames(ghi - glo, h, w, seed=0, first=glo)
    img_h = torch.from_numpy(img_np).pin_memory()
    dep_h = torch.from_numpy(dep_np).pin_memory()
    img_d, dep_d = img_h.to(dev), dep_h.to(dev)
    s_shape, d_shape, m_shape = engine.output_shapes(p, n, h, w)
    outs = tuple(torch.empty(s, dtype=torch.float32, device=dev) for s in (s_shape, d_shape, d_shape, m_shape))
    chunk = args.chunk if args.chunk > 0 else None

    def step():
        engine.stereo_batch_device(img_d, dep_d, p, out=outs, chunk=chunk)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.15)   # every rank: let rank 0's nvidia-smi start, then one more untimed step
    step()
    barrier()
    lib.cs_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = lib.cs_launch_count(0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / args.steps
    fps = world * n / (ms_per_step * 1e-3)

    # ---- in-order assembly check (N > 1): two frames owned by other ranks travel to rank 0 (the only transfers of the
    # whole run, outside the timed region) and must equal rank 0's own single-GPU result for the same global frame
    sharding = {"global_frames": n * world, "partition": "engine.shard_range: contiguous frame ranges, no collective"}
    if world > 1:
        samples = sorted({(n * world) // 2, n * world - 1})
        ok = True
        for g in samples:
            owner = g // n
            if rank == owner and owner != 0:
                dist.send(outs[0][g - glo].contiguous(), dst=0)
            if rank == 0:
                if owner == 0:
                    got = outs[0][g - glo]
                else:
                    got = torch.empty(s_shape[1:], dtype=torch.float32, device=dev)
                    dist.recv(got, src=owner)
                fi, fd = make_frames(1, h, w, seed=0, first=g)
                want = engine.stereo_batch_device(torch.from_numpy(fi).to(dev), torch.from_numpy(fd).to(dev), p)[0][0]
                ok = ok and bool(torch.equal(got, want))
        if rank == 0:
            assert ok, "a frame computed by another rank differs from the single-GPU result"
            sharding["verified_frames"] = samples
            sharding["verified"] = "frames owned by other ranks == rank 0's single-GPU result, bit for bit"
        barrier()

    # ---- strong scaling: one fixed batch (BASELINE config 3's shape: a video batch sharded frame-wise over the GPUs)
    strong = None
    if args.strong_frames > 0:
        gs = args.strong_frames
        slo, shi = engine.shard_range(gs, rank, world)
        if shi > slo:
            si, sd = make_frames(shi - slo, h, w, seed=0, first=slo)
            si_d, sd_d = torch.from_numpy(si).to(dev), torch.from_numpy(sd).to(dev)
            so = engine.stereo_batch_device(si_d, sd_d, p)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        reps = 3
        for _ in range(reps):
            if shi > slo:
                engine.stereo_batch_device(si_d, sd_d, p, out=so)
        a1.record()
        barrier()
        t = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"global_frames": gs, "value": gs * reps / (float(t.item()) * 1e-3), "unit": "frames/s",
                  "frames_per_gpu": -(-gs // world), "note": "fixed batch, frame-sharded; divide by the N=1 line's value "
                  "of this key for strong-scaling efficiency"}
        if shi > slo:
            del si_d, sd_d, so
        torch.cuda.empty_cache()

    # ---- per-kernel CUDA-event durations over the same steps (separate pass so `value` carries no event overhead)
    lib.cs_profile_enable(1)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    lib.cs_profile_enable(0)
    nk = lib.cs_profile_kernel_count()
    k_ms = (ctypes.c_double * nk)()
    k_n = (ctypes.c_longlong * nk)()
    lib.cs_profile_collect(k_ms, k_n)
    kernels = {lib.cs_profile_kernel_name(i).decode(): (k_ms[i], k_n[i]) for i in range(nk) if k_n[i] > 0}
    peak, peak_src = peaks()
    px_step = n * h * w
    cls = "gw_sbs" if gpu_warp else "cpu_sbs"
    if args.mode.endswith("anaglyph") or args.mode in ("left-only", "only-right"):
        cls = "anaglyph"
    path_bytes = BYTES_PER_PX[cls] * px_step
    # algorithmic bytes per pixel of each kernel (both eyes), DESIGN.md section 4
    k_bytes_px = {"k_prepare": 24 + 8, "k_edge_dist": 4 + 2, "k_blur_blend": 4 + 2 + 8 + 24, "k_depth_out": 4 + 24,
                  "k_warp_rows": 24, "k_polylines": 24, "k_polylines_exact": 24, "k_hybrid_splat": 24,
                  "k_hybrid_gapfill": 16, "k_gpuwarp": 12 + 8 + 24 + 4, "k_compose": 8 + 24 + 8}
    if args.mode in ("left-right", "right-left", "top-bottom", "bottom-top") and abs(args.balance) < 0.999:
        # these kernels write the composed float32 tensor + mask themselves (24 + 8 B/px) instead of 8 B/px of RGBX8 eyes
        for k in ("k_polylines", "k_warp_rows"):
            k_bytes_px[k] = 8 + 8 + 32
    timed = {k: v for k, v in kernels.items() if k != "misc"}
    total_k_ms = sum(v[0] for v in timed.values()) or 1.0
    dom = max(timed, key=lambda k: timed[k][0]) if timed else None
    traffic_px = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic_px = json.load(f)["dram_bytes_per_input_pixel"]   # dram read+write per input pixel, from an ncu capture
        if key == "gpu_warp_mesh" and "k_gpuwarp_mesh" in traffic_px:     # the mesh kernels are timed under k_gpuwarp's id
            traffic_px["k_gpuwarp"] = traffic_px["k_gpuwarp_mesh"]
    except Exception:  # noqa: BLE001
        pass
    roofline = None
    if dom:
        d_ms, d_n = timed[dom]
        units_px = px_step * args.steps / d_n          # pixels one launch processes
        per_launch_bytes = k_bytes_px.get(dom, 24) * units_px
        achieved = per_launch_bytes / (d_ms / d_n * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    "traffic": (traffic_px[dom] * units_px) if dom in traffic_px else None,
                    "traffic_source": "profiles/ncu_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel, "
                                      "scaled to the pixels of one launch)" if dom in traffic_px else None,
                    "peak_source": peak_src,
                    "bytes_per_launch": per_launch_bytes, "ms_per_launch": d_ms / d_n,
                    "share_of_step": d_ms / total_k_ms,
                    "path": {"bytes_per_step": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9,
                             "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                             "note": f"whole step: {BYTES_PER_PX[cls]} B/px algorithmic I/O (SURVEY.md 8d) / step time"},
                    "kernels": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                                    "share": v[0] / total_k_ms,
                                    "gbs": k_bytes_px.get(k, 0) * px_step * args.steps / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None}
                                for k, v in timed.items()}}

    # ---- single-frame latency of BASELINE.json configs[0] and configs[1] (report them as latency)
    latency = None
    if rank == 0:
        latency = {}
        for name, (lh, lw, lfill, lblur) in {"config0_512x512_naive": (512, 512, "naive", False),
                                             "config1_1080p_polylines_sharp_blur": (1080, 1920, "polylines_sharp", True)}.items():
            li, ld = make_frames(1, lh, lw, seed=7)
            li, ld = torch.from_numpy(li).to(dev), torch.from_numpy(ld).to(dev)
            lp = engine.make_params(lfill, "left-right", 3.5, 0.0, 0.0, 0.5, 2.0, lblur, 20.0, 20.0, 2.0, 6)
            lo = engine.stereo_batch_device(li, ld, lp)
            ts = []
            for _ in range(30):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                engine.stereo_batch_device(li, ld, lp, out=lo)
                a1.record()
                a1.synchronize()
                ts.append(a0.elapsed_time(a1) * 1e3)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in range(30):
                engine.stereo_batch_device(li, ld, lp, out=lo)
            b1.record()
            b1.synchronize()
            latency[name] = {"median_us": float(np.median(ts)), "min_us": float(np.min(ts)),
                             "back_to_back_us": float(b0.elapsed_time(b1) * 1e3 / 30),
                             "note": "one call, synchronised each time (median/min) and 30 calls queued back to back; repeated "
                                     "identical calls replay a CUDA graph (COMFYSTEREO_GRAPHS=0: direct launches)"}

    # ---- end to end through the node with host tensors
    e2e = None
    if not args.no_e2e:
        node = StereoImageNode()
        params = node_params(args)
        # the node shards over every visible device from one process; under torchrun each rank owns one GPU
        os.environ.setdefault("COMFYSTEREO_SINGLE_DEVICE", "1")

        def e2e_rate(ih, dh, steps):
            def one():
                o = node.generate(ih, dh, **params)
                return float(o[3][0, 0, 0])   # touch the result on the host
            for _ in range(2):
                one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                one()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())

        dt = e2e_rate(img_h, dep_h, args.e2e_steps)
        h2d = img_h.numel() * 4 + dep_h.numel() * 4
        d2h = sum(int(np.prod(s)) for s in (s_shape, d_shape, d_shape, m_shape)) * 4
        e2e = {"value": world * n * args.e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "api": "StereoImageNode.generate (CPU tensors in/out) -> cs_stereo_batch_host", "inputs": "page-locked"}
        bus = d2h
        if _lib.lib().cs_host_compact_enabled():
            # the result tensors hold d2h_bytes_per_step; of those, the depth outputs (3 identical channels) and the
            # mask (0/1) crossed PCIe as one channel / one byte per pixel and were expanded by host threads
            # (depth: the byte k of k/255 for the CPU techniques, one float for GPU Warp; mask: one byte)
            dbytes = 4 if gpu_warp else 1
            bus = int(np.prod(s_shape)) * 4 + 2 * int(np.prod(d_shape)) // 3 * dbytes + int(np.prod(m_shape))
            e2e["d2h_bus_bytes_per_step"] = bus
            e2e["transport"] = "compact depth/mask"
        # how close the call is to what the host's memory system can move: DMA reads of the inputs, DMA writes of what
        # crossed the bus, and the expansion's reads and writes of the compacted outputs
        host_bw = host_copy_bandwidth(lib) if rank == 0 else None
        if host_bw:
            host_bytes = h2d + bus + ((bus - int(np.prod(s_shape)) * 4) + (d2h - int(np.prod(s_shape)) * 4) if bus != d2h else 0)
            e2e["host_bytes_per_step"] = host_bytes
            e2e["host_copy_gbs"] = host_bw
            e2e["host_frac"] = (host_bytes * world / (dt / args.e2e_steps)) / 1e9 / host_bw
            e2e["host_note"] = "host DRAM bytes the call moves per second (all ranks) / the host's streaming-copy bandwidth " \
                               "measured in this run (one process, all threads, non-temporal stores)"
        if world == 1 and not args.no_extra_e2e:
            # what ComfyUI really passes: pageable inputs; and a result too large to be page-locked (> 4 GiB)
            ip, dp = torch.from_numpy(img_np.copy()), torch.from_numpy(dep_np.copy())
            dtp = e2e_rate(ip, dp, args.e2e_steps)
            e2e["pageable_inputs"] = {"value": n * args.e2e_steps / dtp, "unit": "frames/s"}
            big = 96
            bi, bd = make_frames(big, h, w, seed=0)
            bi, bd = torch.from_numpy(bi), torch.from_numpy(bd)
            dtl = e2e_rate(bi, bd, 1)
            e2e["large_batch"] = {"value": big / dtl, "unit": "frames/s", "frames": big,
                                  "note": "pageable inputs, 11 GB of results (beyond the 4 GiB page-lock limit of the python host)"}
            del bi, bd
            # one process, every visible device: the node's own frame sharding (COMFYSTEREO_MULTI_GPU), which a
            # one-rank-per-GPU launch never exercises
            ndev = torch.cuda.device_count()
            if ndev > 1:
                frames = 16 * ndev
                mi, md = make_frames(frames, h, w, seed=0)
                mi, md = torch.from_numpy(mi).pin_memory(), torch.from_numpy(md).pin_memory()
                mp = p
                one = engine.stereo_batch_host(mi[:16], md[:16], mp, device=0)
                for _ in range(2):
                    many = engine.stereo_batch_multi_gpu(mi, md, mp, list(range(ndev)))
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    many = engine.stereo_batch_multi_gpu(mi, md, mp, list(range(ndev)))
                dtm = time.perf_counter() - t0
                same = all(torch.equal(a[:16], b) for a, b in zip(many, one)) if group == 0 else None
                e2e["one_process_multi_gpu"] = {"value": frames * args.e2e_steps / dtm, "unit": "frames/s", "devices": ndev,
                                                "frames": frames, "first_shard_equals_single_gpu": same}
                del mi, md, many, one

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_frames = args.cpu_frames or 32
        val, dt = cpu_baseline(args, cpu_frames)
        import oracle as orc
        cpu = {"value": val, "unit": "frames/s", "cores": orc.set_threads(0), "kind": "port",
               "sample": f"{cpu_frames} frame(s) of the same workload, oracle/stereo_oracle.c with OpenMP over rows, "
                         f"{dt:.2f} s"}

    if rank == 0:
        cfg = workload_config(args, n)
        cfg["sharding"] = sharding
        line = {
            "metric": "1080p stereo frames/s", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "mpix_per_s": fps * h * w / 1e6,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "single_frame_latency": latency, "strong_scaling": strong,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
