#!/usr/bin/env python
"""Headline benchmark: path-traced Msamples/s (and Mrays/s) on bunny.json.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

One "step" is one complete render of the workload scene (S1 of SURVEY 8(d):
examples/bunny.json with the path_tracing integrator, 512 x 384, 100 spp, depth
8; the missing bunny.obj is replaced by the deterministic stand-in mesh of
goblin_b200/bin/scene_gen, 81,920 triangles) through the C ABI, scene resident
in HBM.  With N > 1 every rank renders the same scene with its own sample set
(weak scaling) and the film buffers are summed with one NCCL all-reduce per
step, inside the timed region.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENES = {
    # name: (scene_gen kind, args, json file)
    "bunny": ("bunny", [], "bunny_pt.json"),
    "bunny_ao": ("bunny", [], "bunny_ao.json"),
    "spheres": ("spheres", [], "spheres_pt.json"),
    "grid": ("grid", [], "grid_pt.json"),
    "grid_small": ("grid", [512], "grid_pt.json"),
    "field": ("field", [], "field_pt.json"),
}
NOTES = {
    "bunny": ("S1: examples/bunny.json layout, render_method path_tracing, 512x384, 100 spp, max_ray_depth 8, "
              "gaussian filter r=2; stand-in bunny mesh (icosphere 6x + hashed noise, 81,920 tris): "
              "examples/models/bunny.obj is absent from the reference mount"),
    "bunny_ao": "S2: S1 geometry, render_method ao, ao_sample_num 25, 1920x1080, 16 spp",
    "spheres": "S3: 768 spheres + 256 disks (Lambert / mirror / glass), 4 disk area lights, 2048x2048, 64 spp, depth 8",
    "grid": "S4: 9,999,392-triangle displaced grid mesh with vn, sphere area light + point light, 1920x1080, 64 spp, depth 8",
    "grid_small": "S4 at 524,288 triangles",
    "field": "S5: 27x27 instances of the 81,920-triangle stand-in bunny, 3840x2160, 64 spp, depth 8",
}
METRIC = {"bunny": "path-traced Msamples/s (bunny.json)", "bunny_ao": "AO Msamples/s (bunny.json, ao integrator)"}
SCENE_GEN = os.path.join(ROOT, "goblin_b200", "bin", "scene_gen")
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")


def scene_path(name):
    kind, args, fname = SCENES[name]
    d = os.path.join(ROOT, "scenes", "_gen", kind + ("_" + "_".join(map(str, args)) if args else ""))
    marker = os.path.join(d, ".done")
    if not os.path.exists(marker):
        os.makedirs(d, exist_ok=True)
        subprocess.run([SCENE_GEN, kind, d, *map(str, args)], check=True)
        open(marker, "w").close()
    return os.path.join(d, fname)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, through NVML in-process (a poll every
    10 ms; spawning nvidia-smi takes longer than a short timed region), nvidia-smi as fallback."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.max_mhz, self.reasons = [], None, set()
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def poll(self):
        if self.nvml is not None:
            self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            mask = int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
        if out:
            r = [c.strip() for c in out[0].split(",")]
            self.sm.append(float(r[0]))
            self.max_mhz = float(r[1])
            for k, name in enumerate(["sw_power_cap", "hw_slowdown", "sw_thermal_slowdown", "hw_thermal_slowdown"]):
                if r[2 + k].lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.poll()
            except Exception:
                pass
            self.stop_flag.wait(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def run_ref_tool(scene, spp, threads=0):
    """Timed RenderContext::render() of the unmodified reference (oracle/_ref/ref_tool render)."""
    cmd = [REF_TOOL, "render", scene, "-", "--spp", str(spp)]
    if threads:
        cmd += ["--threads", str(threads)]
    with tempfile.TemporaryDirectory() as td:  # the reference writes its image next to the CWD
        out = subprocess.run(cmd, check=True, capture_output=True, text=True, cwd=td).stdout
    return json.loads(out.split("REF_RESULT", 1)[1])


_REF_SPP = {}


def cpu_reference_sample(scene_json, budget_s=12.0):
    """A bounded sample of the workload on the host cores: the reference itself when its binary
    travelled here (kind "reference"), else the oracle port (kind "port").  The spp of the sample is
    chosen once per scene from a 1-spp probe so that one sample costs about budget_s seconds."""
    if os.path.exists(REF_TOOL):
        if scene_json not in _REF_SPP:
            probe = run_ref_tool(scene_json, 1)
            rate = probe["camera_samples"] / max(probe["seconds"], 1e-3)
            _REF_SPP[scene_json] = max(1, min(10, int((budget_s * rate / probe["camera_samples"]) ** 0.5)))
        root = _REF_SPP[scene_json]
        res = run_ref_tool(scene_json, root * root)
        return {"value": res["msamples_per_s"], "unit": "Msamples/s", "cores": res["cores"], "kind": "reference",
                "sample": f"same scene, {res['spp']} spp ({res['camera_samples']} camera samples) in "
                          f"{res['seconds']:.2f} s, g_ray thread pool on all host cores",
                "mrays_per_s": res["mrays_per_s"], "seconds": res["seconds"]}
    from goblin_b200 import api
    from tests import oracle_port as op
    scene = api.Scene(scene_json)
    t0 = time.perf_counter()
    _, c, calls = op.render(scene, seed=1, spp_total=1)
    dt = time.perf_counter() - t0
    return {"value": c["camera_samples"] / dt * 1e-6, "unit": "Msamples/s", "cores": op.hardware_threads(),
            "kind": "port", "sample": f"same scene, 1 spp ({c['camera_samples']} camera samples) in {dt:.2f} s",
            "mrays_per_s": (calls[0] + calls[1]) / dt * 1e-6, "seconds": dt}


def workload_config(scene_name, scene_json, world, spp_override=0):
    """The `config` of the JSON line: what is rendered, read from the scene file alone, so that the CUDA arm and
    the reference arm (--impl reference) describe the same workload with the same keys and values."""
    import math
    sc = json.load(open(scene_json))
    rs, cam = sc["render_setting"], sc["camera"]
    xres, yres = cam["film"]["resolution"]
    fw = cam.get("filter", {}).get("width", [2.0, 2.0])
    spp = spp_override or int(rs["sample_per_pixel"])
    root = int(math.ceil(math.sqrt(spp)))  # sample_per_pixel rounded up to a square (src/GoblinSampler.cpp:72)
    samples = (xres + 2 * math.ceil(fw[0])) * (yres + 2 * math.ceil(fw[1])) * root * root  # Film::getSampleRange
    return {"workload": NOTES[scene_name], "scene": scene_name, "resolution": [xres, yres], "spp": root * root,
            "max_ray_depth": int(rs.get("max_ray_depth", 5)), "render_method": rs.get("render_method", "path_tracing"),
            "camera_samples_per_gpu_step": samples, "camera_samples_per_step": samples * world,
            "parallelism": (f"weak scaling: {world} GPUs x one full frame each (scene replica per GPU, own sample set, NCCL film "
                            f"all-reduce per step)") if world > 1 else "1 GPU",
            "l2": "CUDA arm: L2 flushed between steps (256 MB fill, its time subtracted), per-step path state exceeds L2; "
                  "reference arm: host CPU, not applicable"}


def bench_reference(args, rank):
    """The reference's own CPU path (oracle/_ref: the unmodified sources) on this box's host cores: every step is one
    render of the same frame at the same spp as the CUDA arm's step when that takes <= ~8 s (the headline scene does),
    else a bounded spp sample of it."""
    if rank != 0:
        return
    scene_json = scene_path(args.scene)
    config = workload_config(args.scene, scene_json, args.gpus, args.spp)
    vals = []
    last = None
    full_spp = None
    if os.path.exists(REF_TOOL) and not args.ref_budget:
        probe = run_ref_tool(scene_json, 4)  # 4 spp: enough work that the thread pool's start-up does not dominate the rate
        if config["camera_samples_per_gpu_step"] / max(probe["camera_samples"] / max(probe["seconds"], 1e-3), 1.0) <= 10.0:
            full_spp = config["spp"]
    for step in range(args.warmup + args.steps):
        if full_spp:
            res = run_ref_tool(scene_json, full_spp)
            last = {"value": res["msamples_per_s"], "cores": res["cores"], "kind": "reference", "seconds": res["seconds"],
                    "mrays_per_s": res["mrays_per_s"],
                    "sample": f"the whole frame: {res['spp']} spp ({res['camera_samples']} camera samples) in {res['seconds']:.2f} s, "
                              f"g_ray thread pool on all host cores"}
        else:
            last = cpu_reference_sample(scene_json, budget_s=args.ref_budget or 4.0)
        if step >= args.warmup:
            vals.append(last)
    value = sum(v["value"] for v in vals) / len(vals)
    secs = sum(v["seconds"] for v in vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC.get(args.scene, "path-traced Msamples/s (%s)" % args.scene), "value": value,
            "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config,
            "reference_step": last["sample"],
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "mrays_per_s": sum(v["mrays_per_s"] for v in vals) / len(vals),
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def strong_record(ctx, scene, spp, rank, world, stream, flush, barrier, steps):
    """One frame of `scene` at `spp` samples per pixel, the sample indices split over the ranks: per
    frame film clear + render of this rank's slice + NCCL all-reduce, timed with CUDA events on the context's
    stream, max over ranks.  Rank 0 then renders the same frames alone (same seeds): the 1-GPU frame time on this
    box for the efficiency, and the film the all-reduced one must equal."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from goblin_b200 import distributed
    b, e = distributed.spp_shard(spp, rank, world)
    sr = scene.sample_range()
    frame_samples = (sr[1] - sr[0]) * (sr[3] - sr[2]) * spp

    def frames(n, begin, end, merge, seed0):
        evs = []
        for i in range(n):
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xFF)
                a, z, m = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record()
                ctx.film_clear()
                ctx.render(seed=seed0 + i, spp_total=spp, spp_begin=begin, spp_end=end)
                m.record()
                if merge:
                    ctx.film_allreduce()
                z.record()
                evs.append((a, m, z))
        ctx.synchronize()
        torch.cuda.synchronize()
        return evs

    frames(2, b, e, True, 7000)  # warm-up (allocates the wave buffers of this slice size)
    barrier()
    ctx.reset_kernel_times()
    ctx.enable_kernel_timing(True)
    evs = frames(steps, b, e, True, 7100)
    ctx.enable_kernel_timing(False)
    kt = ctx.kernel_times()
    merged = ctx.film_download().copy() if rank == 0 else None  # the last frame, all-reduced
    ms = sum(a.elapsed_time(z) for a, _, z in evs) / steps
    ar_ms = sum(m.elapsed_time(z) for _, m, z in evs) / steps
    t = torch.tensor([ms, ar_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0].item()), float(t[1].item())
    barrier()
    rec = None
    if rank == 0:
        frames(1, 0, spp, False, 7000)
        one = frames(steps, 0, spp, False, 7100)
        one_ms = sum(a.elapsed_time(z) for a, _, z in one) / steps
        whole = ctx.film_download()
        denom = np.maximum(np.abs(whole), 1e-3)
        rel = float((np.abs(merged - whole) / denom).max())
        ok = bool(np.allclose(merged, whole, rtol=5e-4, atol=1e-4))
        overhead = ms - one_ms / world
        rec = {"spp_total": spp, "spp_per_gpu": [e - b if r == rank else distributed.spp_shard(spp, r, world)[1] - distributed.spp_shard(spp, r, world)[0] for r in range(world)],
               "camera_samples_per_frame": frame_samples, "ms_per_frame": ms, "value": frame_samples / (ms * 1e-3) * 1e-6,
               "unit": "Msamples/s", "steps": steps, "one_gpu_ms_per_frame": one_ms, "speedup": one_ms / ms,
               "efficiency": one_ms / (world * ms), "allreduce_ms": ar_ms, "film_bytes": int(whole.nbytes),
               "allreduce_check": ok, "allreduce_max_rel_diff": rel,
               "kernel_ms_per_frame_rank0": {k: v[0] / steps for k, v in kt.items() if v[1]},
               "limiter": (f"{overhead:.2f} ms per frame do not shrink with N: {ar_ms:.2f} ms film all-reduce, the rest are the late-bounce "
                           f"launches (a few thousand paths on 148 SMs) and per-frame fixed costs (film clear, launch latency)")}
    barrier()
    return rec


class DevBuf:
    """A device allocation exposed through __cuda_array_interface__ (for torch.as_tensor)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def scene_bytes(desc):
    d = desc
    return (32 * (d.n_top_nodes + d.n_model_nodes) + 4 * d.n_instances + 128 * d.n_instances + 12 * d.n_tris * 2 +
            4 * d.n_tris + 32 * d.n_verts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="bunny", choices=sorted(SCENES))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--accel", default="equal_count", choices=["equal_count", "middle", "sah"],
                    help="BVH split method: equal_count = the reference's tree (the headline, parity mode); "
                         "sah = the non-parity fast tree (SURVEY 8(f) rank 1)")
    ap.add_argument("--trace-mode", default="pair", choices=["pair", "wide", "exact"],
                    help="pair = the pair-node walk (default; 'exact' is its older name); wide = 4-wide nodes (measured slower)")
    ap.add_argument("--spp", type=int, default=0, help="override the scene's sample_per_pixel")
    ap.add_argument("--wave-paths", type=int, default=0)
    ap.add_argument("--tune", default="", help="comma list for gb_set_tuning (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-stats", action="store_true", help="skip the counters child process (ncu runs: no roofline)")
    ap.add_argument("--no-fast-tree", action="store_true", help="skip the extra --accel sah leg of the default run")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling records")
    ap.add_argument("--no-strong-grid", action="store_true", help="N > 1: skip the 10 M-triangle strong-scaling record")
    ap.add_argument("--stats-only", default="", help=argparse.SUPPRESS)  # internal: "seed,spp_total,begin,end"
    ap.add_argument("--ref-budget", type=float, default=0.0,
                    help="seconds of CPU work per reference sample (default: 12 for the cpu_baseline of the CUDA arm, 4 per step of --impl reference)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        bench_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from goblin_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    # scene (rank 0 generates, everybody loads: full replica per GPU)
    if rank == 0:
        scene_json = scene_path(args.scene)
    barrier()
    scene_json = scene_path(args.scene)
    t0 = time.perf_counter()
    scene = api.Scene(scene_json, accel=args.accel)
    load_s = time.perf_counter() - t0
    if args.stats_only:
        seed, spp_total, b, e = (int(v) for v in args.stats_only.split(","))
        ctx = api.Context(local_rank)
        ctx.upload_scene(scene)
        if args.wave_paths:
            ctx.set_wave_paths(args.wave_paths)
        ctx.enable_counters(True)
        ctx.reset_counters()
        ctx.film_clear()
        ctx.render(seed=seed, spp_total=spp_total, spp_begin=b, spp_end=e)
        ctx.synchronize()
        print("STATS " + json.dumps(ctx.counters()), flush=True)
        return
    stats_pass = None
    if rank == 0 and not args.no_stats:
        spp0 = scene.spp_squared(args.spp or None)
        if args.scaling == "strong":
            per0 = (spp0 + world - 1) // world
            rng0 = (0, min(spp0, per0))
        else:
            rng0 = (0, spp0)
        cmd = [sys.executable, os.path.abspath(__file__), "--scene", args.scene, "--accel", args.accel, "--stats-only",
               f"1000,{spp0},{rng0[0]},{rng0[1]}"]
        if args.wave_paths:
            cmd += ["--wave-paths", str(args.wave_paths)]
        env = dict(os.environ, RANK="0", LOCAL_RANK=str(local_rank), WORLD_SIZE="1")
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, env=env)
            line = [l for l in out.stdout.splitlines() if l.startswith("STATS ")]
            stats_pass = json.loads(line[-1][6:]) if line else None
        except Exception:
            stats_pass = None
    ctx = api.Context(local_rank)
    ctx.upload_scene(scene)
    ctx.set_trace_mode(args.trace_mode)
    if args.wave_paths:
        ctx.set_wave_paths(args.wave_paths)
    if args.tune:
        ctx.set_tuning([int(v) for v in args.tune.split(",")])
    spp = scene.spp_squared(args.spp or None)
    if args.scaling == "strong":
        per = (spp + world - 1) // world
        spp_begin, spp_end = min(spp, rank * per), min(spp, (rank + 1) * per)
        seed_of = lambda step: 1000 + step  # noqa: E731  one sample set, split by index range
    else:
        spp_begin, spp_end = 0, spp
        seed_of = lambda step: 1000 + step + 7919 * rank  # noqa: E731  independent sample sets
    my_samples = (scene.sample_range()[1] - scene.sample_range()[0]) * \
        (scene.sample_range()[3] - scene.sample_range()[2]) * (spp_end - spp_begin)

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    film_ptr, film_floats = ctx.film_device_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    if world > 1:
        # Film::mergeTile across GPUs is the library's own NCCL all-reduce (gb_film_allreduce); torch.distributed only
        # carries the communicator id to the other ranks
        from goblin_b200 import distributed
        distributed.init_film_comm(ctx, rank, world)

    def step(i, flush_l2=True):
        with torch.cuda.stream(stream):
            if flush_l2:
                flush.fill_(i & 0xFF)
            ctx.film_clear()
            ctx.render(seed=seed_of(i), spp_total=spp, spp_begin=spp_begin, spp_end=spp_end)
            if world > 1:
                ctx.film_allreduce()

    # traversal statistics of one step (counters on: a different, slower instantiation of the
    # traversal kernels) give the algorithmic bytes.  They are gathered by a short-lived child
    # process before the timed run, so that the measurement process only ever runs the kernels it
    # times.
    st = stats_pass if rank == 0 else None

    for i in range(args.warmup):
        step(i)
    ctx.synchronize()
    torch.cuda.synchronize()
    ctx.reset_counters()
    ctx.reset_kernel_times()
    ctx.enable_kernel_timing(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the L2 flush is part of the loop; its fill (~40 us) is timed separately and subtracted
    fl0, fl1 = [], []
    with torch.cuda.stream(stream):
        ev0.record()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            flush.fill_(i & 0xFF)
            b.record()
            fl0.append(a)
            fl1.append(b)
        step(args.warmup + i, flush_l2=False)
    with torch.cuda.stream(stream):
        ev1.record()
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    clocks.stop_flag.set()
    total_ms = ev0.elapsed_time(ev1) - sum(a.elapsed_time(b) for a, b in zip(fl0, fl1))
    ctx.enable_kernel_timing(False)
    ktimes_overlapped = ctx.kernel_times()
    timed = ctx.counters()
    timed_launches = int(timed["kernel_launches"])
    # The timed region runs two wave lanes and the shadow kernels side by side on their own streams (tail overlap,
    # DESIGN.md 4): an event pair around one launch there spans the kernels it shares the machine with.  The dominant
    # kernel's own launch duration, which the roofline needs, comes from a short extra pass of the very same step
    # with both features off (one lane, shadow kernel in line), L2 flushed before every step.
    base_tune = [int(v) for v in args.tune.split(",")] if args.tune else []
    base_tune = (base_tune + [20, 6, 4, 10, 0][len(base_tune):])[:5]
    ctx.set_tuning(base_tune + [1, 0])
    for i in range(2):
        step(30000 + i)
    ctx.synchronize()
    ctx.reset_kernel_times()
    ctx.enable_kernel_timing(True)
    excl_steps = max(2, min(args.steps, 5))
    for i in range(excl_steps):
        step(30100 + i)
    ctx.synchronize()
    torch.cuda.synchronize()
    ctx.enable_kernel_timing(False)
    kt_excl = ctx.kernel_times()
    ktimes = {k: (v[0] * args.steps / excl_steps, v[1] * args.steps // excl_steps) for k, v in kt_excl.items()}
    user_tune = [int(v) for v in args.tune.split(",")] if args.tune else []
    ctx.set_tuning(base_tune + (user_tune[5:7] if len(user_tune) >= 7 else [2, 1]))
    ctx.reset_counters()
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(my_samples), float(timed["rays_closest"] + timed["rays_any"])], dtype=torch.float64,
                       device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    total_ms = float(t.item())
    samples_per_step_all = float(tot[0].item())
    rays_all = float(tot[1].item())
    ms_per_step = total_ms / args.steps
    value = samples_per_step_all / (ms_per_step * 1e-3) * 1e-6
    mrays = rays_all / (total_ms * 1e-3) * 1e-6

    # ---- end to end through the C ABI with host buffers: scene upload (H2D) + render + film download (D2H)
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 20))
    host_film = np.zeros(1, np.float32)
    for i in range(2 if e2e_steps else 0):  # warm both scene slots of the asynchronous path (their arenas are allocated on first use)
        ctx.upload_scene_async(scene)
        ctx.synchronize()
    host_buf = np.zeros((scene.desc.film.yres, scene.desc.film.xres, 4), np.float32)
    barrier()
    torch.cuda.synchronize()
    # Every step: the scene goes host -> device from the caller's arrays (gb_upload_scene_async: staged and copied on the
    # library's copy stream), the frame is rendered (+ all-reduced), the film comes back device -> host into the caller's
    # buffer.  Software-pipelined like a frame loop: the upload for step i + 1 is issued while step i renders, so the
    # host staging and the PCIe copy overlap device work; every step still pays its own H2D and D2H inside the timed region.
    w0 = time.perf_counter()
    if e2e_steps:
        ctx.upload_scene_async(scene)
    for i in range(e2e_steps):
        step(10000 + i, flush_l2=False)
        if i + 1 < e2e_steps:
            ctx.upload_scene_async(scene)
        host_film = ctx.film_download(out=host_buf)  # the caller's film buffer, reused like Film::mPixels
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    # the unpipelined figure beside it: upload, render, download strictly one after the other
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        ctx.upload_scene(scene)
        step(11000 + i, flush_l2=False)
        host_film = ctx.film_download(out=host_buf)
    torch.cuda.synchronize()
    e2e_serial_ms = (time.perf_counter() - w0) * 1e3
    up_ms = None
    if e2e_steps:  # one synchronous upload by itself: host validation + staging + PCIe copy + device-side derivation
        ctx.synchronize()
        u0 = time.perf_counter()
        for i in range(3):
            ctx.upload_scene(scene)
        up_ms = (time.perf_counter() - u0) * 1e3 / 3
    te = torch.tensor([e2e_ms, e2e_serial_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_serial_ms = float(te[0].item()), float(te[1].item())
    e2e_value = samples_per_step_all / (e2e_ms / e2e_steps * 1e-3) * 1e-6 if e2e_steps else None
    assert np.isfinite(host_film).all()

    # ---- the same workload on the non-parity fast tree, reported beside the headline
    fast_tree = None
    if world == 1 and args.accel == "equal_count" and not args.no_fast_tree and not args.no_e2e:
        try:
            t0 = time.perf_counter()
            sah_scene = api.Scene(scene_json, accel="sah")
            sah_load_s = time.perf_counter() - t0
            ctx.upload_scene(sah_scene)
            for i in range(3):
                step(20000 + i)
            ctx.synchronize()
            fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fsteps = max(1, min(args.steps, 5))
            with torch.cuda.stream(stream):
                flush.fill_(1)
                fa.record()
            for i in range(fsteps):
                step(20100 + i, flush_l2=False)
            with torch.cuda.stream(stream):
                fb.record()
            ctx.synchronize()
            torch.cuda.synchronize()
            fms = fa.elapsed_time(fb) / fsteps
            fast_tree = {"accel": "sah", "value": samples_per_step_all / (fms * 1e-3) * 1e-6, "unit": "Msamples/s",
                         "ms_per_step": fms, "steps": fsteps, "scene_load_s": sah_load_s,
                         "note": "binned-SAH tree behind --accel sah: same kernels, same closest-hit distances, "
                                 "not the reference's node order; the headline value is the parity tree"}
            ctx.upload_scene(scene)
        except Exception as e:  # an extra, never fatal
            fast_tree = {"accel": "sah", "value": None, "note": f"failed: {e}"}

    # ---- strong scaling, driver-visible (N > 1): ONE sample set split N ways by sample index, film all-reduce inside the
    # timed frame, the all-reduced frame checked on rank 0 against the same frame rendered by rank 0 alone
    strong = None
    if world > 1 and not args.no_strong:
        strong = {}
        strong[args.scene] = strong_record(ctx, scene, spp, rank, world, stream, flush, barrier, steps=max(3, min(args.steps, 10)))
        if args.scene == "bunny" and not args.no_strong_grid:  # BASELINE.json config 4: the 10 M-triangle grid, 64 spp split N ways
            try:
                if rank == 0:
                    scene_path("grid")
                barrier()
                grid = api.Scene(scene_path("grid"), accel=args.accel)
                ctx.upload_scene(grid)
                ctx.set_trace_mode(args.trace_mode)
                strong["grid"] = strong_record(ctx, grid, grid.spp_squared(), rank, world, stream, flush, barrier, steps=3)
                if strong["grid"] is not None:
                    strong["grid"]["workload"] = NOTES["grid"]
                ctx.upload_scene(scene)
                ctx.set_trace_mode(args.trace_mode)
                del grid
            except Exception as e:  # an extra, never fatal -- but every rank must fail alike or the barriers hang
                strong["grid"] = {"failed": str(e)}

    if rank == 0 and st is None:
        st = {k: 0 for k in ("nodes_visited", "nodes_visited_any", "prims_tested", "prims_tested_any", "instances_entered",
                             "instances_entered_any", "rays_closest", "rays_any")}
        stats_failed = True
    else:
        stats_failed = False
    if rank == 0:
        # ---- roofline of the dominant traversal kernel (closest-hit extend vs any-hit shadow / ao)
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak, peak_src = (peaks["hbm_gbs"], "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
            else (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")
        # With the tail overlap the shadow kernel of a bounce shares the machine with the next extend on a
        # second stream: its event pair then spans both, so the shadow class is a span, not exclusive time, and only
        # the ambient-occlusion kernel can outweigh the extend kernel as the dominant one.
        any_cls = "ao" if ktimes["ao"][1] else "shadow"
        bytes_closest = (32 * (st["nodes_visited"] - st["nodes_visited_any"]) + 48 * (st["prims_tested"] - st["prims_tested_any"]) +
                         64 * (st["instances_entered"] - st["instances_entered_any"]) + 48 * st["rays_closest"])
        bytes_any = (32 * st["nodes_visited_any"] + 48 * st["prims_tested_any"] + 64 * st["instances_entered_any"] +
                     48 * st["rays_any"])
        if any_cls == "shadow" or ktimes["extend"][0] >= ktimes[any_cls][0]:
            kname, kms, kbytes, krays = "k_extend (closest-hit two-level BVH traversal)", ktimes["extend"], bytes_closest, st["rays_closest"]
        else:
            kname, kms, kbytes, krays = f"k_{any_cls} (any-hit traversal)", ktimes[any_cls], bytes_any, st["rays_any"]
        launches_per_step = kms[1] / args.steps
        avg_launch_ms = kms[0] / max(kms[1], 1)
        achieved = (kbytes / max(launches_per_step, 1)) / (avg_launch_ms * 1e-3) * 1e-9 if avg_launch_ms > 0 else 0.0
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tr_path):
            traffic = json.load(open(tr_path)).get(args.scene)
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src,
                    "frac_of_nominal_8000": achieved / 8000.0,  # north_star quotes the 8 TB/s nominal figure as well
                    "traffic": traffic,
                    # DRAM bytes actually moved per launch / launch time / peak: what fraction of HBM bandwidth the kernel uses.
                    # `frac` above counts ALGORITHMIC bytes, most of which L1 / L2 serve on scenes that fit there.
                    "dram_frac": (traffic / (avg_launch_ms * 1e-3) * 1e-9 / peak) if (traffic and avg_launch_ms > 0) else None,
                    "algorithmic_bytes_per_ray": kbytes / max(krays, 1),
                    "algorithmic_bytes_per_launch": kbytes / max(launches_per_step, 1),
                    "avg_launch_ms": avg_launch_ms, "launches_per_step": launches_per_step,
                    "nodes_per_ray": (st["nodes_visited"]) / max(st["rays_closest"] + st["rays_any"], 1),
                    "kernel_share_of_step": kms[0] / total_ms,
                    "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items() if v[1]},
                    "kernel_ms_per_step_overlapped": {k: v[0] / args.steps for k, v in ktimes_overlapped.items() if v[1]},
                    "kernel_ms_note": "kernel_ms_per_step / avg_launch_ms: each class by itself, from an extra pass of the same step "
                                      "with one wave lane and the shadow kernel in line (event pairs per launch); "
                                      "kernel_ms_per_step_overlapped: the same event pairs inside the timed region, where two wave "
                                      "lanes and the shadow kernels share the machine, so the classes add up to more than the step"}
        if roofline.get("dram_frac") is not None:
            roofline["reading"] = (f"the kernel's algorithmic bytes are served by L1 / L2 on this scene: DRAM moves {roofline['dram_frac']:.1%} of the peak, "
                                   f"so `frac` is a throughput in HBM units (it can exceed 1), not distance to the HBM wall; the kernel is bound by "
                                   f"SM issue slots at ~18 of 32 active lanes and by the L1 gather pipe (profiles/README.md, DESIGN.md 4)")
        if stats_failed:
            roofline.update({"achieved": None, "frac": None, "frac_of_nominal_8000": None,
                             "note": "no algorithmic bytes: the counters pass was skipped (--no-stats)" if args.no_stats
                             else "the counters pass failed; no algorithmic bytes"})
        line = {"metric": METRIC.get(args.scene, "path-traced Msamples/s (%s)" % args.scene), "value": value, "unit": "Msamples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": workload_config(args.scene, scene_json, world, args.spp),
                "run": {"accel": args.accel, "trace_mode": ctx.trace_mode(), "spp_range_rank0": [spp_begin, spp_end],
                        "camera_samples_per_step_measured": samples_per_step_all, "triangles": int(scene.desc.n_tris),
                        "scene_load_s": load_s,
                        "film_merge": "gb_film_allreduce (NCCL, library-owned communicator)" if world > 1 else "none (1 GPU)"},
                "mrays_per_s": mrays, "rays_per_sample": rays_all / (samples_per_step_all * args.steps),
                "gpu_launches": timed_launches,
                "clocks": clocks.summary(),
                "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(ctx.upload_bytes()),
                        "d2h_bytes_per_step": int(film_floats * 4), "steps": e2e_steps,
                        "serial_value": samples_per_step_all / (e2e_serial_ms / e2e_steps * 1e-3) * 1e-6 if e2e_steps else None,
                        "upload_ms": up_ms, "upload_gb_per_s": (ctx.upload_bytes() / (up_ms * 1e-3) * 1e-9) if up_ms else None,
                        "what": "per step: gb_upload_scene_async (host arrays -> pinned staging -> H2D on the copy stream, issued one step "
                                "ahead so it overlaps the previous step's kernels) + gb_film_clear + gb_render (+ gb_film_allreduce) + "
                                "gb_film_download into the caller's buffer; serial_value = the same with gb_upload_scene, nothing overlapped"},
                "roofline": roofline}
        if strong is not None:
            line["strong"] = strong
            line["allreduce_check"] = all(v.get("allreduce_check") is True for v in strong.values() if v and "failed" not in v)
        if fast_tree is not None:
            line["fast_tree"] = fast_tree
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_sample(scene_json, budget_s=args.ref_budget or 12.0)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["mrays_per_s_reference_equivalent"] = cb["mrays_per_s"]
            except Exception as e:  # the baseline is reported, never fatal
                line["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
