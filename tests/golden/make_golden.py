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
DATA = os.path.join(ROOT, "oracle", "_ref", "data")

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


if __name__ == "__main__":
    main()
