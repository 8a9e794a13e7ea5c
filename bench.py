#!/usr/bin/env python
"""bench.py — headline benchmark of the render hot path (contract: see the task statement / DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU renderer (oracle/_ref), rank 0 only

Workload (BASELINE.json configs[4]): the synthetic 10 M-triangle terrain inside a
five-wall Lambert box under one 4x4 RectLight, path traced (gi on, depth 8) at 1920x1080 x 1024 samples
per pixel: ONE FIXED FRAME whatever the GPU count (strong scaling). Its sample passes are dealt over the
GPUs (GPU g renders the samples s % N == g of every pixel) and the per-GPU sums are reduced once per frame:
  * under torchrun (one process per GPU, the driver's launch): torch.distributed NCCL reduce
    (hexray_b200.distributed.render_frame, the function the tests exercise);
  * in one process (`python bench.py --gpus N` without torchrun): the library's own multi-GPU context
    (hxr_config.devices: one kernel on GPU 0 sums the peers' frames over NVLink).
`--spp S` instead renders S samples per pixel PER GPU per step (weak scaling; quick A/B runs).

A step = one frame: ray generation -> closest hit -> shading -> shadow rays -> ... -> accumulate
(-> reduce -> resolve). metric = Mrays/s, rays = closest-hit queries past the depth guard +
visible() queries, the same definition the reference counters use.
Prints ONE JSON line on rank 0.
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

TERRAIN_SCENE = """GlobalSettings {{
	frameWidth {W}
	frameHeight {H}
	ambientLight (0, 0, 0)
	maxTraceDepth 8
	gi 1
	numPaths {spp}
}}
Camera camera {{
	pos (0, 150, -600)
	aspectRatio 1.77778
	pitch -15
	fov 70
}}
RectLight {{
	scale (300, 1, 300)
	translate (0, 400, 0)
	xSubd 4
	ySubd 4
	power 1100000
}}
Mesh terrain {{
	file "{mesh}"
}}
Plane wall {{
	limit 520
}}
Lambert grey {{
	color (0.75, 0.75, 0.75)
}}
Lambert ground {{
	color (0.55, 0.6, 0.4)
}}
Node terrain {{
	geometry terrain
	shader ground
}}
Node floor {{
	geometry wall
	shader grey
	translate (0, -40, 0)
}}
Node left {{
	geometry wall
	shader grey
	rotate (0, 0, 90)
	translate (-520, 0, 0)
}}
Node right {{
	geometry wall
	shader grey
	rotate (0, 0, -90)
	translate (520, 0, 0)
}}
Node back {{
	geometry wall
	shader grey
	rotate (0, 90, 0)
	translate (0, 0, 520)
}}
Node front {{
	geometry wall
	shader grey
	rotate (0, -90, 0)
	translate (0, 0, -700)
}}
"""

# algorithmic bytes per ray of the traversal roofline: 64 B ray+hit record,
# 32 B per inner node visited, 48 B per triangle tested, 16 B per leaf entered
B_RECORD, B_INNER, B_TRI, B_LEAF = 64, 32, 48, 16
# what the walk kernel actually reads and writes per ray: 48 B entry record + 16 B candidate record,
# 32 B per block step (two tree levels), 48 B triangle + 4 B index per test, 4 B leaf count per leaf
F_RECORD, F_INNER, F_TRI, F_LEAF = 64, 32, 52, 4

DATA = os.path.join(ROOT, "assets", "data")  # the reference's scene assets, staged by __graft_entry__.build()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="terrain", choices=["terrain", "soup", "cornell_box", "kdtree_test", "smallpt"])
    ap.add_argument("--soup-triangles", type=int, default=10000000, help="soup workload: random triangles in [-500,500]^3 (SURVEY.md 8d, worst-case incoherence)")
    ap.add_argument("--grid-side", type=int, default=2237, help="terrain vertices per side (2237 -> 9 999 392 triangles)")
    ap.add_argument("--spp-total", type=int, default=1024, help="samples per pixel of the WHOLE frame, dealt over the GPUs (strong scaling; BASELINE config 5)")
    ap.add_argument("--spp", type=int, default=0, help="if > 0: samples per pixel PER GPU per step instead (weak scaling, quick runs)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--ref-width", type=int, default=384, help="reference arm: bounded sample resolution")
    ap.add_argument("--ref-height", type=int, default=216)
    ap.add_argument("--ref-spp", type=int, default=8, help="reference arm: samples per pixel of the bounded sample (the cpu_baseline leg of our arm uses 4x)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the bundled-scene block (configs C1-C4) of the N=1 line")
    ap.add_argument("--queue-capacity", type=int, default=64 << 20,
                    help="rays per wave of the wavefront renderer: 64 Mi holds one whole 1080p x 32 spp pass (~19 GB of queues out of 180 GB HBM)")
    return ap.parse_args()


def workload_name(a):
    if a.workload == "terrain":
        ntri = 2 * (a.grid_side - 1) ** 2
        return "synthetic terrain %d triangles (grid %d^2, seed 0x5EED) in a 5-wall Lambert box, 1 RectLight 4x4, GI depth 8" % (ntri, a.grid_side)
    if a.workload == "soup":
        return "synthetic soup %d random triangles (seed 0x5EEE) in the same 5-wall Lambert box, 1 RectLight 4x4, GI depth 8" % a.soup_triangles
    return "data/%s.hexray" % a.workload


def scene_text(a, mesh_file, W, H, spp):
    return TERRAIN_SCENE.format(W=W, H=H, spp=spp, mesh=mesh_file)


def synthetic_mesh_name(a):
    if getattr(a, "workload", "terrain") == "soup":
        return "synthetic:soup:%d:0x5EEE" % a.soup_triangles
    return "synthetic:terrain:%d:0x5EED" % a.grid_side


def frame_spp(a, world):
    """(samples per pixel of the whole frame, scaling label)"""
    if a.spp > 0:
        return a.spp * world, "weak"
    return a.spp_total, "strong"


def base_config(a, world):
    spp_total, scaling = frame_spp(a, world)
    return {"workload": workload_name(a), "width": a.width, "height": a.height, "spp_total": spp_total,
            "spp_per_gpu": spp_total / world, "integrator": "path tracing (gi), maxTraceDepth 8"}, scaling


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------ reference arm
# (this arm never loads libhexray_b200.so: the procedural mesh is written by tools/bin/hxr_objgen, a host-only program)
def reference_binary():
    ref = os.path.join(ROOT, "oracle", "_ref", "hexray_ref_count")
    return ref if os.path.exists(ref) else None


def run_reference(scene, cwd, width, height, spp, threads=0, repeat=1):
    ref = reference_binary()
    if ref is None:
        return None, {"unavailable": "oracle/_ref/hexray_ref_count not built (run `make -C oracle` where /root/reference exists)"}
    cmd = [ref, "render", scene, "--width", str(width), "--height", str(height), "--repeat", str(repeat)]
    if spp:
        cmd += ["--spp", str(spp)]
    if threads:
        cmd += ["--threads", str(threads)]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if p.returncode != 0:
        return None, {"unavailable": "reference run failed: " + p.stderr[-300:]}
    info = json.loads(p.stdout.strip().splitlines()[-1])
    info["wall_s"] = time.time() - t0
    info["rays"] = info["rays_closest"] + info["rays_shadow"]
    return info["rays"] / info["best_ms"] / 1e3, info


def run_reference_sample(a, workdir, threads=0, repeat=1):
    """Time the reference's own CPU renderer (oracle/_ref/hexray_ref_count: the unmodified reference sources behind a
    headless shim) on a bounded sample of the workload. Returns (Mrays/s best, info dict)."""
    if a.workload in ("terrain", "soup"):
        size = a.grid_side if a.workload == "terrain" else a.soup_triangles
        obj = os.path.join(workdir, "%s_%d.obj" % (a.workload, size))
        scene = os.path.join(workdir, "%s_ref_%d.hexray" % (a.workload, size))
        if not os.path.exists(obj):
            gen = os.path.join(ROOT, "tools", "bin", "hxr_objgen")
            if not os.path.exists(gen):
                return None, {"unavailable": "tools/bin/hxr_objgen not built (python -c 'import __graft_entry__ as g; g.build()')"}
            seed = "0x5EED" if a.workload == "terrain" else "0x5EEE"
            p = subprocess.run([gen, a.workload, str(size), seed, obj + ".tmp"], capture_output=True, text=True)
            if p.returncode != 0:
                return None, {"unavailable": "hxr_objgen failed: " + p.stderr[-200:]}
            os.replace(obj + ".tmp", obj)
        with open(scene, "w") as f:
            f.write(scene_text(a, os.path.basename(obj), a.ref_width, a.ref_height, a.ref_spp))
        cwd = workdir
    else:
        scene = os.path.join(DATA, a.workload + ".hexray")
        cwd = os.path.dirname(DATA)
    return run_reference(scene, cwd, a.ref_width, a.ref_height, a.ref_spp, threads, repeat)


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    workdir = os.path.join(tempfile.gettempdir(), "hexray_b200_bench")
    os.makedirs(workdir, exist_ok=True)
    total = a.steps + a.warmup
    mr, info = run_reference_sample(a, workdir, repeat=total)
    cfg, scaling = base_config(a, max(world, a.gpus))
    line = {"impl": "reference", "metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg}
    if mr is None:
        line["unavailable"] = info["unavailable"]
        print(json.dumps(line))
        return
    times = info["render_ms"][a.warmup:]
    ms = sum(times) / len(times)
    value = info["rays"] / ms / 1e3
    sample = "%dx%d x %d spp of the same scene (%d rays per step), %d host threads; KD build %.1f s and OBJ parse %.1f s outside the timed region" % (
        a.ref_width, a.ref_height, a.ref_spp, info["rays"], info["threads"], info["begin_render_ms"] / 1e3, info["parse_ms"] / 1e3)
    line.update({"value": value, "ms_per_step": ms,
                 "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": info["threads"], "kind": "reference", "sample": sample},
                 "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ bundled scenes (configs C1-C4)
EXTRA_SCENES = [
    # (scene, config tag, GPU spp override, CPU spp for the bounded sample (0 = the scene's own), also at 1080p on the GPU)
    ("simple", "C1", 0, 0, True),
    ("meshes", "C2", 0, 0, True),
    ("kdtree_test", "C2/C4", 0, 0, True),
    ("heightfield", "C4", 0, 0, True),
    ("beer", "C4", 0, 0, True),
    ("cornell_box", "C3", 256, 16, False),
]


def measured_l2_gbs():
    """L2 read bandwidth of this GPU model as measured by tools/microbench.cu (profiles/microbench_r1.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "microbench_r1.json")) as f:
            return float(json.load(f)["l2_read_gbs"]), "profiles/microbench_r1.json (tools/microbench.cu)"
    except Exception:
        return 23000.0, "fallback"


def scene_traversal_roofline(hx, r, sf, W, H, spp, mrays, sm_ghz, sms=148):
    """SURVEY 8d for a scene whose geometry sits in L2: T_ray = max(T_issue, T_mem) with
    B_ray = 64 + 32 N_inner + 48 N_tri + 16 N_leaf bytes at the measured L2 bandwidth and
    I_ray = 30 per node tested + analytic primitives (plane 10, sphere 25, cube 60, rect light 35) + 10 N_inner + 30 N_tri
    FP32-pipe instructions at SMs x 128 lanes x clock. N_* come from a counting render of the same frame (per ray, rays =
    closest-hit + visible() queries)."""
    _, st = r.render(width=W, height=H, spp=min(spp, 16) if spp else 0, seed=0, flags=hx.RENDER_COUNT_TRAVERSAL)
    rays = max(1, st["rays_closest"] + st["rays_shadow"])
    n_inner, n_tri, n_leaf = st["kd_inner"] / rays, st["tri_tests"] / rays, st["kd_leaves"] / rays
    pod = sf.pod.contents
    prim = {0: 10.0, 1: 25.0, 2: 60.0}  # hxr_geometry_type: plane, sphere, cube

    def geom_cost(gi, depth=0):
        g = pod.geometries[gi]
        if g.type == 3 and depth < 8:  # CSG: both children
            return geom_cost(g.b, depth + 1) + geom_cost(g.c, depth + 1)
        return prim.get(g.type, 0.0)  # meshes are counted through N_inner / N_tri; the heightfield DDA has no constant in 8d
    inst = sum(30.0 + geom_cost(pod.nodes[i].geom) for i in range(pod.n_nodes))
    inst += 35.0 * sum(1 for i in range(pod.n_lights) if pod.lights[i].type == 1)
    inst += 10.0 * n_inner + 30.0 * n_tri
    b_ray = 64.0 + 32.0 * n_inner + 48.0 * n_tri + 16.0 * n_leaf
    l2, l2_src = measured_l2_gbs()
    t_issue = inst / (sms * 128 * sm_ghz * 1e9)
    t_mem = b_ray / (l2 * 1e9)
    peak = 1e-6 / max(t_issue, t_mem)
    note = "texture / shading bytes and instructions are excluded (traversal roofline)"
    if pod.n_heightfields:
        note += "; the heightfield DDA has no constant in SURVEY 8d: only its node transform is counted"
    return {"note": note, "bound": "l2" if t_mem >= t_issue else "fp32 issue", "bytes_per_ray": b_ray, "inst_per_ray": inst, "N_inner": n_inner, "N_tri": n_tri,
            "N_leaf": n_leaf, "l2_gbs": l2, "l2_source": l2_src, "sm_ghz": sm_ghz, "peak_mrays_per_s": peak, "frac": mrays / peak}


def extra_scenes(hx, device, want_cpu, sm_ghz=1.965):
    """Mrays/s and ms/frame of the bundled scenes on one GPU (median of 3 frames after 2 warm-up frames, CUDA-event time of
    hxr_render), at their own resolution and - Whitted scenes - at 1920x1080, next to the reference on this box's host cores."""
    import numpy as np
    out = []
    for scene, tag, spp, cpu_spp, hd in EXTRA_SCENES:
        path = os.path.join(DATA, scene + ".hexray")
        if not os.path.exists(path):
            continue
        sf = hx.SceneFile(path)
        r = hx.Renderer(device=device)
        r.load(sf)
        W0, H0 = r.frame_size()
        rec = {"scene": "data/%s.hexray" % scene, "config": tag}
        sizes = [(W0, H0, "own")] + ([(1920, int(round(1920 * H0 / W0)), "1080p-class")] if hd else [])
        for W, H, label in sizes:
            ms, rays = [], 0
            for i in range(5):
                _, st = r.render(width=W, height=H, spp=spp, seed=i)
                if i >= 2:
                    ms.append(st["render_ms"])
                    rays = st["rays_closest"] + st["rays_shadow"]
            m = float(np.median(ms))
            rec[label] = {"width": W, "height": H, "spp": spp or None, "rays": rays, "ms_per_frame": m, "mrays_per_s": rays / m / 1e3,
                          "kernel_launches": st["kernel_launches"]}
            try:  # the scene as a fraction of its traversal roofline (reporting only: never fails the bench)
                rec[label]["roofline"] = scene_traversal_roofline(hx, r, sf, W, H, spp, rays / m / 1e3, sm_ghz)
            except Exception as e:
                rec[label]["roofline"] = {"error": str(e)[:200]}
        r.close()
        sf.close()
        if want_cpu and reference_binary():
            mr, info = run_reference(path, os.path.dirname(DATA), W0, H0, cpu_spp, repeat=2)
            if mr is not None:
                rec["cpu_reference"] = {"mrays_per_s": mr, "ms_per_frame": info["best_ms"], "cores": info["threads"], "rays": info["rays"],
                                        "sample": "%dx%d%s" % (W0, H0, " x %d spp (bounded sample of the 256-spp frame)" % cpu_spp if cpu_spp else "")}
        out.append(rec)
    return out


# ------------------------------------------------------------------------------------------ our arm
def ours(a):
    import numpy as np
    import torch
    import hexray_b200 as hx

    os.environ.setdefault("HEXRAY_DATA", DATA)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # one process that owns several GPUs (no torchrun): the library's multi-GPU context
    in_library = world == 1 and a.gpus > 1
    n_gpus = a.gpus if in_library else world
    dev = torch.device("cuda", local)
    W, H = a.width, a.height
    spp_total, scaling = frame_spp(a, n_gpus)

    workdir = os.path.join(tempfile.gettempdir(), "hexray_b200_bench")
    os.makedirs(workdir, exist_ok=True)
    if a.workload in ("terrain", "soup"):
        path = os.path.join(workdir, "%s_%d_rank%d.hexray" % (a.workload, a.grid_side if a.workload == "terrain" else a.soup_triangles, rank))
        with open(path, "w") as f:
            f.write(scene_text(a, synthetic_mesh_name(a), W, H, spp_total))
    else:
        path = os.path.join(DATA, a.workload + ".hexray")
    t0 = time.time()
    sf = hx.SceneFile(path)
    r = hx.Renderer(device=local, queue_capacity=a.queue_capacity, devices=list(range(a.gpus)) if in_library else None)
    r.load(sf)
    setup_s = time.time() - t0
    accel = r.accel_info(0) if sf.pod.contents.n_meshes > 0 else {}
    cam = sf.camera()
    mode = hx.MODE_MONTECARLO
    geometry_bytes = sum(r.accel_info(i)["bytes_nodes"] + r.accel_info(i)["bytes_tris"] for i in range(sf.pod.contents.n_meshes))
    flush = None
    if geometry_bytes < (256 << 20):
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > L2: written between timed iterations

    acc = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    host = torch.empty(H * W * 3, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from hexray_b200 import distributed

    def step(i, to_host, flags=0):
        """one frame: this rank's sample passes -> reduce -> resolve [-> host]. Returns (stats, device ms of the part that
        runs on torch's stream: reduce + resolve + copy)."""
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize(dev)
        if to_host:
            r.set_camera(cam)  # the per-frame input of the C ABI (hxr_set_camera), host -> device
        if world == 1:
            # one GPU, or the library's multi-GPU context: ONE call; the reduce + resolve are inside (stats: reduce_ms)
            if to_host:
                _, st = r.render(width=W, height=H, mode=mode, spp=spp_total, seed=i, out=host.numpy().reshape(H, W, 3))
            else:
                st = r.render_device(acc.data_ptr(), width=W, height=H, mode=mode, spp=spp_total, seed=i, flags=flags)
            return st, 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = distributed.render_frame(r, acc, W, H, spp_total=spp_total, seed=i, rank=rank, world=world, mode=mode, dist=dist,
                                      sync=lambda: torch.cuda.synchronize(dev), on_rendered=e0.record, flags=flags)
        if to_host and rank == 0:
            host.copy_(acc, non_blocking=False)
        e1.record()
        torch.cuda.synchronize(dev)
        return st, e0.elapsed_time(e1)

    KEYS = ("trace_closest_ms", "trace_shadow_ms", "walk_ms", "setup_ms", "finish_ms", "shade_ms", "shadow_resolve_ms", "gen_ms", "other_ms",
            "render_ms", "reduce_ms")

    def run(n, to_host, first_seed, flags=0):
        rays = np.zeros(2, dtype=np.float64)
        prof = {k: 0.0 for k in KEYS}
        prof.update({"walk_launches": 0, "kernel_launches": 0, "post_ms": 0.0, "cand_overflow": 0})
        for i in range(n):
            st, post_ms = step(first_seed + i, to_host, flags)
            rays += (st["rays_closest"], st["rays_shadow"])
            for k in prof:
                prof[k] += post_ms if k == "post_ms" else st[k]
        return rays, prof

    # ---- device-resident throughput ("value"): CUDA-event time of the frames (render on the library's stream(s) +
    # reduce/resolve), max over ranks; the wall clock around the same region is reported beside it
    run(a.warmup, False, 1000)
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        rays, timed = run(a.steps, False, 0)
        barrier()
        dt_wall = time.perf_counter() - t0
    dt = (timed["render_ms"] + timed["post_ms"]) * 1e-3
    # ---- per-kernel device times (the roofline's denominator): the same K steps once more with every kernel on ONE stream
    # (HXR_RENDER_ONE_LANE) -- in the timed region above the shadow chain of a bounce runs beside the next bounce's closest-hit
    # chain on a second stream, so event pairs around a launch there would also count its neighbour
    r.set_profiling(True)
    _, prof = run(a.steps, False, 0, hx.RENDER_ONE_LANE)
    r.set_profiling(False)
    barrier()
    # ---- end to end through the C ABI with host buffers ("e2e"): wall clock around blocking calls that end with the
    # frame in host memory
    run(min(a.warmup, 1), True, 2000)
    barrier()
    t0 = time.perf_counter()
    rays_e, _ = run(a.steps, True, 0)
    barrier()
    dt_e = time.perf_counter() - t0

    def allsum(x):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return t.cpu().numpy()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rays = allsum(rays)
    rays_e = allsum(rays_e)
    dt = allmax(dt)
    dt_wall = allmax(dt_wall)
    dt_e = allmax(dt_e)

    # ---- traversal counters (a separate, un-timed counting pass of the same workload at 1 spp per GPU)
    if world == 1:
        cst = r.render_device(acc.data_ptr(), width=W, height=H, mode=mode, spp=n_gpus, seed=0, flags=hx.RENDER_COUNT_TRAVERSAL)
    else:
        cst = r.render_device(acc.data_ptr(), width=W, height=H, mode=mode, spp=world, seed=0, shard=(rank, world), flags=hx.RENDER_COUNT_TRAVERSAL)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        total_rays = float(rays.sum())
        value = total_rays / dt / 1e6
        n_rays_cnt = cst["rays_closest"] + cst["rays_shadow"]
        per_ray = {"inner": cst["kd_inner"] / max(1, n_rays_cnt), "tri": cst["tri_tests"] / max(1, n_rays_cnt),
                   "leaf": cst["kd_leaves"] / max(1, n_rays_cnt), "mesh_queries": cst["mesh_queries"] / max(1, n_rays_cnt)}
        b_ray = B_RECORD + B_INNER * per_ray["inner"] + B_TRI * per_ray["tri"] + B_LEAF * per_ray["leaf"]
        f_ray = F_RECORD + F_INNER * per_ray["inner"] + F_TRI * per_ray["tri"] + F_LEAF * per_ray["leaf"]
        walk_ms = prof["walk_ms"]  # this GPU's k_walk launches (in-library multi-GPU: the slowest GPU's)
        walk_launches = prof["walk_launches"] / (n_gpus if in_library else 1)
        my_rays = total_rays / n_gpus  # one GPU's share (shards are equal)
        achieved = my_rays * b_ray / (walk_ms * 1e-3) / 1e9 if walk_ms > 0 else None
        hbm_bound = geometry_bytes > (126 << 20)
        peak = peaks["hbm_gbs"] if hbm_bound else 23149.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "walk_traffic_r2.json")  # written by tools/ncu_traffic.py from an ncu pass over the timed-size k_walk launches
        if os.path.exists(tp) and a.workload == "terrain" and a.grid_side == 2237:
            with open(tp) as f:
                traffic = json.load(f)
        # context for the roofline: what this GPU delivers to RANDOM gathers (tools/microbench.cu, profiles/microbench_r1_v12.json) --
        # the copy bandwidth in `peak` is not reachable by a tree walk, whatever the record size
        gather = None
        mp = os.path.join(ROOT, "profiles", "microbench_r1_v12.json")
        if os.path.exists(mp):
            try:
                with open(mp) as f:
                    g = json.loads(f.readline())["random_gather_G_per_s"]["l2gran64(got 64)"]
                gather = {"sectors32B_G_per_s": g["32B"], "gbs_64B_records": g["64B"] * 64.0, "source": "profiles/microbench_r1_v12.json"}
                if traffic and walk_ms > 0:
                    gather["walk_dram_gbs"] = traffic["dram_bytes_per_launch"] / (walk_ms / max(1, walk_launches) * 1e-3) / 1e9
            except Exception:
                gather = None
        cfg, _ = base_config(a, n_gpus)
        cfg.update({"sharding": "sample passes s % N == gpu; " + ("torch.distributed NCCL reduce of the sum buffers (one process per GPU)" if world > 1 else
                                                                  ("library-owned multi-GPU context, reduce: " + r.reduce_backend() if in_library else "one GPU")),
                    "l2": "geometry %.2f GB >> 126 MB L2" % (geometry_bytes / 1e9) if flush is None else "512 MB buffer written between timed iterations",
                    "rays_per_step": total_rays / a.steps, "setup_s": setup_s, "kd": accel, "queue_capacity_rays": a.queue_capacity})
        path_peak = peak * 1e9 / b_ray / 1e6  # Mrays/s one GPU could do if the whole path ran at the traversal roofline
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32 walk + f64 exact tests", "data": "synthetic", "config": cfg,
            "ms_per_frame": dt / a.steps * 1e3, "wall_ms_per_step": dt_wall / a.steps * 1e3,
            "timing": "CUDA events (render on the library's stream + reduce/resolve), max over ranks; wall clock beside it",
            "e2e": {"value": float(rays_e.sum()) / dt_e / 1e6, "unit": "Mrays/s", "ms_per_step": dt_e / a.steps * 1e3,
                    "h2d_bytes_per_step": 256, "d2h_bytes_per_step": W * H * 12},
            "gpu_launches": int(timed["kernel_launches"]),
            "clocks": clk.summary(),
            "N_inner": per_ray["inner"], "N_tri": per_ray["tri"], "N_leaf": per_ray["leaf"],
            "roofline": {"bound": "hbm" if hbm_bound else "l2", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_source": traffic,
                         "kernel": "k_walk (KD-tree walk: closest-hit + shadow launches)", "peak_source": peak_src if hbm_bound else "tools/microbench L2 read (profiles/microbench_r1.json)",
                         "bytes_per_ray": b_ray, "bytes_per_ray_fetched": f_ray,
                         "frac_fetched": (my_rays * f_ray / (walk_ms * 1e-3) / 1e9 / peak) if walk_ms > 0 else None,
                         "path_frac": (value / n_gpus) / path_peak, "path_peak_mrays_per_gpu": path_peak,
                         "bytes_per_launch": my_rays * b_ray / max(1, walk_launches), "per_ray": per_ray,
                         "launches": int(walk_launches), "avg_launch_ms": walk_ms / max(1, walk_launches),
                         "random_gather": gather,
                         "walk_share_of_step": walk_ms / max(1e-9, prof["render_ms"])},
            "kernel_ms_per_step": {k: prof[k] / a.steps for k in KEYS},
            "kernel_ms_source": "the same K steps repeated after the timed region with every kernel on ONE stream (HXR_RENDER_ONE_LANE, CUDA events "
                                "around each launch); render_ms there is the one-lane step, ms_per_step above the two-lane one",
            "cand_overflow_per_step": prof["cand_overflow"] / a.steps,
        }
        if n_gpus == 1 and not a.no_cpu_baseline:
            b = argparse.Namespace(**vars(a))
            b.ref_spp = a.ref_spp * 4  # ~10 s of CPU rendering on the 1 M-triangle variant
            note = ""
            if a.workload == "terrain" and a.grid_side > 709:
                b.grid_side = 709  # the reference needs ~1 min to parse + build the 10 M mesh: that is what --impl reference times
                note = " (1 002 528-triangle variant of the terrain; the full mesh is timed by --impl reference)"
            if a.workload == "soup" and a.soup_triangles > 1000000:
                b.soup_triangles = 1000000
                note = " (1 M-triangle variant of the soup; the full mesh is timed by --impl reference)"
            mr, info = run_reference_sample(b, workdir)
            if mr is None:
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": info["unavailable"]}
            else:
                line["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": info["threads"], "kind": "reference",
                                        "sample": "%dx%d x %d spp, %d rays, %.1f s wall incl. load%s" % (b.ref_width, b.ref_height, b.ref_spp, info["rays"], info["wall_s"], note)}
        if n_gpus == 1 and not a.no_extra:
            r.close()
            line["extra"] = {"scenes": extra_scenes(hx, local, not a.no_cpu_baseline, sm_ghz=(line["clocks"].get("sm_mhz") or 1965.0) / 1e3)}
            if sf.pod.contents.n_meshes > 0:
                # the same mesh's KD-tree built on the GPU (HXR_CFG_DEVICE_KD_BUILD) next to the host build above
                t0 = time.time()
                rd = hx.Renderer(device=local, queue_capacity=1 << 20, flags=hx.CFG_DEVICE_KD_BUILD)
                rd.load(sf)
                ai = rd.accel_info(0)
                line["extra"]["device_kd_build"] = {"build_ms": ai["build_ms"], "device_passes_ms": ai["device_ms"], "upload_scene_s": time.time() - t0,
                                                    "nodes": ai["nodes"], "leaves": ai["leaves"], "tri_refs": ai["tri_refs"], "max_depth": ai["max_depth"],
                                                    "host_build_ms": accel.get("build_ms"), "host_tree_from_cache": accel.get("from_cache")}
                rd.close()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()
    sf.close()


class JsonOnlyStdout:
    """The contract is ONE JSON line on stdout. Libraries write there too, from C (NCCL prints its version line at the first
    collective when the box sets NCCL_DEBUG=VERSION): file descriptor 1 points at stderr while the bench runs, and only
    print()s of this module reach the real stdout."""

    def __enter__(self):
        import builtins
        sys.stdout.flush()
        self.real = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        self.print = builtins.print

        def to_real(*args, **kw):
            if kw.get("file") is None:
                kw["file"] = self.real
            self.print(*args, **kw)
            kw["file"].flush()
        globals()["print"] = to_real
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        self.real.flush()
        os.dup2(self.real.fileno(), 1)
        del globals()["print"]
        return False


def main():
    a = parse_args()
    with JsonOnlyStdout():
        if a.impl == "reference":
            reference_arm(a)
        else:
            ours(a)


if __name__ == "__main__":
    main()
