"""Generate the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref/hexray_ref).

Run in the build container (needs /root/reference for `make -C oracle`):
    make -C oracle && python tests/golden/make_golden.py
Outputs (tests/golden/*.npz, float16/float64, a few MB in total):
    whitted_<scene>.npz   img = reference vfb at reduced resolution (deterministic scenes)
    mc_<scene>.npz        img = reference vfb at HIGH spp (stochastic scenes: the converged target)
    primary_<scene>.npz   rays + raycast records for a coarse pixel grid (TraceContext::raycast)
    visible_<scene>.npz   random segments + visible() answers
Every file records the command line that produced it.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "hexray_ref")
DATA = os.path.join(ROOT, "assets", "data")  # the reference's data/ staged by __graft_entry__.build()

WHITTED = {  # scene -> (W, H)
    "simple": (320, 180), "meshes": (320, 240), "kdtree_test": (320, 240), "heightfield": (320, 240),
    "bumpmap": (320, 240), "Lecture8": (320, 240), "beer": (256, 192),
}
MC = {  # scene -> (W, H, spp for the converged reference)
    "cornell_box": (128, 128, 4096), "smallpt": (128, 96, 4096), "hw12/sphtri": (128, 96, 2048),
    "zaphod": (129, 86, 2048), "hw10/bokeh": (128, 96, 1024), "boxed": (160, 120, 0),
}
PRIMARY = {"kdtree_test": (64, 48), "meshes": (64, 48), "heightfield": (64, 48), "smallpt": (64, 48),
           "simple": (64, 36), "hw10/bokeh": (64, 48), "boxed": (64, 48)}


def run(args):
    p = subprocess.run([REF] + args, cwd=os.path.dirname(DATA), capture_output=True, text=True)
    if p.returncode != 0:
        sys.exit("oracle failed: %s\n%s" % (args, p.stderr))
    return json.loads(p.stdout.strip().splitlines()[-1])


def scene_path(name):
    return os.path.join(DATA, name + ".hexray")


def tag(name):
    return name.replace("/", "_")


def main():
    tmp = "/tmp/hxr_golden.bin"
    for name, (W, H) in WHITTED.items():
        args = ["render", scene_path(name), "--width", str(W), "--height", str(H), "--out", tmp]
        info = run(args)
        img = np.fromfile(tmp, dtype=np.float32).reshape(H, W, 3)
        np.savez_compressed(os.path.join(HERE, "whitted_%s.npz" % tag(name)), img=img.astype(np.float16),
                            cmd=" ".join(["hexray_ref"] + args[:1] + ["data/%s.hexray" % name] + args[2:6]), info=json.dumps(info))
        print("whitted", name, img.shape, float(img.mean()))
    for name, (W, H, spp) in MC.items():
        args = ["render", scene_path(name), "--width", str(W), "--height", str(H), "--out", tmp]
        if spp:
            args += ["--spp", str(spp)]
        # boxed is Whitted with a jittered 8x8 rect light: average several frames instead
        reps = 1 if spp else 16
        acc = np.zeros((H, W, 3), dtype=np.float64)
        for r in range(reps):
            info = run(args)
            acc += np.fromfile(tmp, dtype=np.float32).reshape(H, W, 3)
        img = (acc / reps).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, "mc_%s.npz" % tag(name)), img=img.astype(np.float16), spp=spp, reps=reps,
                            cmd=" ".join(["hexray_ref", "render", "data/%s.hexray" % name] + args[2:]), info=json.dumps(info))
        print("mc", name, img.shape, float(img.mean()))
    rng = np.random.default_rng(1234)
    for name, (W, H) in PRIMARY.items():
        args = ["primary", scene_path(name), "--width", str(W), "--height", str(H), "--out", tmp]
        run(args)
        rec = np.fromfile(tmp, dtype=np.float64).reshape(H * W, 30)
        rays, hits = rec[:, :6], rec[:, 6:]
        np.savez_compressed(os.path.join(HERE, "primary_%s.npz" % tag(name)), rays=rays, hits=hits, W=W, H=H,
                            cmd="hexray_ref primary data/%s.hexray --width %d --height %d" % (name, W, H))
        # visible(): segments between random pairs of primary hit points (+ a little noise), so that
        # many of them graze or cross geometry
        ok = hits[:, 0] == 0
        pts = hits[ok][:, 3:6]
        if len(pts) > 16:
            n = 2000
            a = pts[rng.integers(0, len(pts), n)] + rng.normal(0, 0.5, (n, 3))
            b = pts[rng.integers(0, len(pts), n)] + rng.normal(0, 0.5, (n, 3))
            seg = np.concatenate([a, b], axis=1)
            seg.tofile(tmp + ".in")
            run(["rays", scene_path(name), "--in", tmp + ".in", "--out", tmp, "--mode", "visible"])
            vis = np.fromfile(tmp, dtype=np.float64)
            np.savez_compressed(os.path.join(HERE, "visible_%s.npz" % tag(name)), seg=seg, vis=vis.astype(np.uint8))
        # raytrace() colours of secondary-ish rays: from hit points towards random directions
        print("primary", name, rec.shape, "hits", int(ok.sum()))


def terrain(side=320):
    """C5-style synthetic terrain (the bench workload at a size the reference renders in a minute): explicit rays
    from the bench camera (its x = 0 lies exactly ON a top-level split plane of the grid), from inside the box,
    secondary rays leaving the surface, shadow segments, and a converged GI frame."""
    sys.path.insert(0, ROOT)
    import bench
    import hexray_b200 as hx
    import argparse
    work = "/tmp/hxr_golden_terrain"
    os.makedirs(work, exist_ok=True)
    W, H, spp = 96, 54, 1024
    a = argparse.Namespace(grid_side=side)
    gen = os.path.join(work, "gen.hexray")
    with open(gen, "w") as f:
        f.write(bench.scene_text(a, "synthetic:terrain:%d:0x5EED" % side, W, H, spp))
    sf = hx.SceneFile(gen)  # host front-end only (no GPU needed)
    sf.write_obj(0, os.path.join(work, "terrain.obj"))
    sf.close()
    ref_scene = os.path.join(work, "ref.hexray")
    with open(ref_scene, "w") as f:
        f.write(bench.scene_text(a, "terrain.obj", W, H, spp))

    def oracle(args):
        p = subprocess.run([REF] + args, cwd=work, capture_output=True, text=True)
        if p.returncode != 0:
            sys.exit("oracle failed: %s\n%s" % (args, p.stderr))
        return json.loads(p.stdout.strip().splitlines()[-1])

    rng = np.random.default_rng(77)
    n = 3000
    cam_o = np.tile(np.array([[0.0, 150.0, -600.0]]), (n, 1))
    cam_d = np.stack([rng.uniform(-0.7, 0.7, n), rng.uniform(-0.6, 0.1, n), np.ones(n)], 1)
    in_o = np.stack([rng.uniform(-500, 500, n), rng.uniform(-30, 300, n), rng.uniform(-500, 500, n)], 1)
    in_d = rng.normal(size=(n, 3))
    O, D = np.concatenate([cam_o, in_o]), np.concatenate([cam_d, in_d])
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    tmp = os.path.join(work, "io.bin")

    def raycast(O, D):
        r8 = np.zeros((len(O), 8))
        r8[:, 0:3], r8[:, 3:6] = O, D
        r8.tofile(tmp + ".in")
        oracle(["rays", ref_scene, "--in", tmp + ".in", "--out", tmp])
        return np.fromfile(tmp, dtype=np.float64).reshape(-1, 24)

    h1 = raycast(O, D)
    ok = (h1[:, 0] == 0) & (h1[:, 1] == 0)  # hits on the terrain node
    ip, nn = h1[ok][:, 3:6], h1[ok][:, 6:9]
    d2 = rng.normal(size=ip.shape)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    d2[(d2 * nn).sum(1) < 0] *= -1
    o2 = ip + nn * 1e-6
    h2 = raycast(o2, d2)
    rays = np.concatenate([np.concatenate([O, D], 1), np.concatenate([o2, d2], 1)])
    hits = np.concatenate([h1, h2])
    light = np.array([[0.0, 400.0 - 1e-6, 0.0]]) + rng.uniform(-150, 150, (len(o2), 3)) * np.array([1.0, 0.0, 1.0])
    # shadow segments: light -> surface, and between random pairs of hit points (many cross or graze the terrain)
    pts = hits[hits[:, 0] == 0][:, 3:6]
    m = 4000
    pa = pts[rng.integers(0, len(pts), m)] + rng.normal(0, 0.5, (m, 3))
    pb = pts[rng.integers(0, len(pts), m)] + rng.normal(0, 0.5, (m, 3))
    seg = np.concatenate([np.concatenate([light, o2], axis=1), np.concatenate([pa, pb], axis=1)])
    seg.tofile(tmp + ".in")
    oracle(["rays", ref_scene, "--in", tmp + ".in", "--out", tmp, "--mode", "visible"])
    vis = np.fromfile(tmp, dtype=np.float64)
    info = oracle(["render", ref_scene, "--out", tmp])
    img = np.fromfile(tmp, dtype=np.float32).reshape(H, W, 3)
    np.savez_compressed(os.path.join(HERE, "terrain_%d.npz" % side), rays=rays, hits=hits[:, :20], seg=seg, vis=vis.astype(np.uint8),
                        img=img.astype(np.float16), spp=spp, side=side, W=W, H=H, info=json.dumps(info),
                        cmd="hexray_ref rays|render <bench.py terrain scene, OBJ written by hxr_scene_file_write_obj>")
    print("terrain", side, "rays", len(rays), "terrain hits", int(ok.sum()), "visible", float(vis.mean()), "img mean", float(img.mean()), info["best_ms"])


def stereo():
    """Anaglyph frames (camera.stereoSeparation != 0, src/main.cpp:234-248): no bundled scene uses the feature, so the
    fixtures are bundled scenes with the property added (tests/hxr_testlib.py: stereo_scene)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hxr_testlib as T
    tmp = "/tmp/hxr_golden.bin"
    for name, W, H, spp in (("kdtree_test", 320, 240, 0), ("cornell_box", 96, 96, 4096)):
        variant = T.stereo_scene(name, T.STEREO[name])
        args = ["render", scene_path(variant), "--width", str(W), "--height", str(H), "--out", tmp]
        if spp:
            args += ["--spp", str(spp)]
        info = run(args)
        img = np.fromfile(tmp, dtype=np.float32).reshape(H, W, 3)
        np.savez_compressed(os.path.join(HERE, "stereo_%s.npz" % tag(name)), img=img.astype(np.float16), spp=spp, separation=T.STEREO[name],
                            cmd="hexray_ref render data/%s.hexray + stereoSeparation %g --width %d --height %d" % (name, T.STEREO[name], W, H),
                            info=json.dumps(info))
        print("stereo", name, img.shape, float(img.mean()), img.reshape(-1, 3).mean(0))


def output():
    """What the reference WRITES (a27): takeScreenshot -> Bitmap::saveImage (src/sdl.cpp:103-116, src/bitmap.cpp:202-288) of a
    small frame whose row size needs BMP padding: the float frame (exactly, float32), the BMP file bytes and the EXR file bytes."""
    tmp, bmp, exr = "/tmp/hxr_golden.bin", "/tmp/hxr_golden.bmp", "/tmp/hxr_golden.exr"
    W, H = 203, 77
    args = ["render", scene_path("simple"), "--width", str(W), "--height", str(H), "--out", tmp, "--bmp", bmp, "--exr", exr]
    info = run(args)
    vfb = np.fromfile(tmp, dtype=np.float32).reshape(H, W, 3)
    np.savez_compressed(os.path.join(HERE, "output_simple.npz"), vfb=vfb, bmp=np.frombuffer(open(bmp, "rb").read(), dtype=np.uint8),
                        exr=np.frombuffer(open(exr, "rb").read(), dtype=np.uint8),
                        cmd="hexray_ref render data/simple.hexray --width %d --height %d --out vfb.f32 --bmp out.bmp --exr out.exr" % (W, H), info=json.dumps(info))
    print("output", vfb.shape, float(vfb.mean()), os.path.getsize(bmp), os.path.getsize(exr))


def features():
    """Scenes written for this repo that exercise what no bundled scene does (tests/hxr_testlib.py: FEATURES_*)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hxr_testlib as T
    tmp = "/tmp/hxr_golden.bin"
    name = T.feature_scene("whitted")
    info = run(["render", scene_path(name), "--out", tmp])
    img = np.fromfile(tmp, dtype=np.float32).reshape(240, 320, 3)
    np.savez_compressed(os.path.join(HERE, "features_whitted.npz"), img=img.astype(np.float16), info=json.dumps(info),
                        cmd="hexray_ref render <tests/hxr_testlib.py FEATURES_WHITTED>")
    print("features whitted", float(img.mean()))
    name = T.feature_scene("stochastic")
    reps = 200  # glossy reflection and the DOF lens are stochastic at 16 samples per pixel: average frames
    acc = np.zeros((120, 160, 3), dtype=np.float64)
    for r in range(reps):
        info = run(["render", scene_path(name), "--out", tmp])
        acc += np.fromfile(tmp, dtype=np.float32).reshape(120, 160, 3)
    img = (acc / reps).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "features_stochastic.npz"), img=img.astype(np.float16), reps=reps, info=json.dumps(info),
                        cmd="hexray_ref render <tests/hxr_testlib.py FEATURES_STOCHASTIC> x %d frames averaged" % reps)
    print("features stochastic", float(img.mean()))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "output":
    output()
    sys.exit(0)
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "features":
        features()
    elif len(sys.argv) > 1 and sys.argv[1] == "terrain":
        terrain()
    elif len(sys.argv) > 1 and sys.argv[1] == "stereo":
        stereo()
    else:
        main()
