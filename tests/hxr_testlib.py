"""Shared checks, parameterised by the library under test (product on GPU / emulation on CPU)."""
import json
import os
import subprocess

import numpy as np

import hexray_b200 as hx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE = os.path.join(ROOT, "oracle", "_ref", "hexray_ref")

WHITTED_SCENES = ["simple", "meshes", "kdtree_test", "heightfield", "bumpmap", "Lecture8", "beer"]
MC_SCENES = ["cornell_box", "smallpt", "hw12/sphtri", "zaphod", "hw10/bokeh"]
PRIMARY_SCENES = ["kdtree_test", "meshes", "heightfield", "smallpt", "simple", "hw10/bokeh", "boxed"]

# Parity tolerances (BASELINE.json north_star):
#   deterministic scenes: |clamp(ours) - clamp(reference)| <= 1/255 per channel on >= 99.9 % of pixels
#   stochastic scenes   : with own = RMSE(ours@N seed A, ours@N seed B) / sqrt(2), the noise of ONE N-spp frame, and
#                         Nref the sample count of the converged reference fixture (its noise: own * sqrt(N / Nref)),
#                           RMSE(ours@N, reference@Nref) <= 1.2 * own * sqrt(1 + N / Nref) + 3e-4   (3e-4: the fixture is fp16)
#                           per-channel mean |delta| <= 0.002, and the same RMSE bound on 8x8 box-filtered images
#                         (GPU tier, live oracle: RMSE(ours@N, oracle@N) <= 1.2 * r0, r0 = RMSE of two oracle runs at N)
PIXEL_TOL = 1.0 / 255.0
PIXEL_FRACTION = 0.999


def tag(name):
    return name.replace("/", "_")


def golden(kind, scene):
    return np.load(os.path.join(GOLDEN, "%s_%s.npz" % (kind, tag(scene))))


def scene_path(scene):
    if os.path.isabs(scene):
        return scene
    return os.path.join(hx.data_root(), scene + ".hexray")


_SCRATCH = None


def scratch_scene(name, text):
    """Write a test-made scene where its asset paths resolve WITHOUT touching the data root: a scratch directory whose
    entries are symlinks to the data root's (scene files resolve assets relative to their own directory)."""
    global _SCRATCH
    if _SCRATCH is None:
        import tempfile
        _SCRATCH = tempfile.mkdtemp(prefix="hxr_scratch_")
        root = hx.data_root()
        for e in os.listdir(root):
            os.symlink(os.path.join(root, e), os.path.join(_SCRATCH, e))
    path = os.path.join(_SCRATCH, name + ".hexray")
    with open(path, "w") as f:
        f.write(text)
    return path


def have_oracle():
    return os.path.exists(ORACLE) and os.access(ORACLE, os.X_OK)


def oracle_render(scene, W, H, spp=0, extra=()):
    out = "/tmp/hxr_oracle_%d.f32" % os.getpid()
    args = [ORACLE, "render", scene_path(scene), "--width", str(W), "--height", str(H), "--out", out] + list(extra)
    if spp:
        args += ["--spp", str(spp)]
    p = subprocess.run(args, cwd=os.path.dirname(hx.data_root()), capture_output=True, text=True, check=True)
    info = json.loads(p.stdout.strip().splitlines()[-1])
    return np.fromfile(out, dtype=np.float32).reshape(H, W, 3), info


def stereo_scene(scene, separation):
    """A copy of a bundled scene with `stereoSeparation` set on its camera (anaglyph frames, src/main.cpp:234-248)."""
    text = open(scene_path(scene)).read()
    i = text.index("Camera")
    j = text.index("{", i)
    text = text[:j + 1] + "\n\tstereoSeparation %g" % separation + text[j + 1:]
    return scratch_scene("_stereo_%s" % os.path.basename(scene), text)


STEREO = {"kdtree_test": 12.0, "cornell_box": 40.0}  # scene -> stereoSeparation used by the fixtures


class Session:
    """Caches loaded scenes per library so a test module pays the parse/upload once per scene."""

    def __init__(self, api, queue_capacity=0):
        self.api = api
        self.qc = queue_capacity
        self.cache = {}

    def renderer(self, scene):
        if scene not in self.cache:
            sf = hx.SceneFile(scene_path(scene), api_=self.api)
            r = hx.Renderer(api_=self.api, queue_capacity=self.qc)
            r.load(sf)
            self.cache[scene] = (sf, r)
        return self.cache[scene][1]

    def close(self):
        for sf, r in self.cache.values():
            r.close()
            sf.close()
        self.cache.clear()


def clamp01(x):
    return np.clip(np.asarray(x, dtype=np.float32), 0.0, 1.0)


def pixel_match_fraction(a, b):
    d = np.abs(clamp01(a) - clamp01(b)).max(axis=2)
    return float((d <= PIXEL_TOL + 1e-6).mean()), float(d.max())


def check_whitted(sess, scene):
    g = golden("whitted", scene)
    ref = g["img"].astype(np.float32)
    H, W = ref.shape[:2]
    img, st = sess.renderer(scene).render(width=W, height=H)
    # the fixture is stored as float16: allow its rounding (2^-11 relative) on top of 1/255
    d = np.abs(clamp01(img) - clamp01(ref)).max(axis=2)
    frac = float((d <= PIXEL_TOL + 6e-4).mean())
    assert frac >= PIXEL_FRACTION, "%s: only %.4f%% of pixels within 1/255 (max diff %.4f)" % (scene, frac * 100, d.max())
    assert st["rays_closest"] > 0
    return frac, st


def rmse(a, b):
    return float(np.sqrt(((clamp01(a) - clamp01(b)) ** 2).mean()))


def box_filter(img, k=8):
    H, W = img.shape[:2]
    H2, W2 = H // k * k, W // k * k
    return clamp01(img)[:H2, :W2].reshape(H2 // k, k, W2 // k, k, 3).mean(axis=(1, 3))


MC_REF_SPP = {"cornell_box": 4096, "smallpt": 4096, "hw12/sphtri": 2048, "zaphod": 2048, "hw10/bokeh": 1024}  # tests/golden/make_golden.py


def mc_bounds(a, b, ref, n, n_ref, what):
    """The stochastic-parity assertion (see the header): a, b = two of our N-spp frames, ref = the converged reference."""
    out = {}
    for label, f in (("pixels", clamp01), ("8x8 boxes", box_filter)):
        fa, fb, fr = f(a), f(b), f(ref)
        own = float(np.sqrt(((fa - fb) ** 2).mean())) / np.sqrt(2.0)  # noise of ONE render against the truth
        err = float(np.sqrt(((fa - fr) ** 2).mean()))
        bound = 1.2 * own * np.sqrt(1.0 + n / float(n_ref)) + 3e-4
        assert err <= bound, "%s (%s): rmse vs converged reference %.5f > bound %.5f (own noise %.5f, N %d, Nref %d)" % (what, label, err, bound, own, n, n_ref)
        out[label] = (err, own, bound)
    mean_delta = np.abs(clamp01(a).mean(axis=(0, 1)) - clamp01(ref).mean(axis=(0, 1)))
    assert mean_delta.max() <= 0.002, "%s: per-channel mean delta %s > 0.002" % (what, mean_delta)
    return out


def check_mc(sess, scene, spp):
    g = golden("mc", scene)
    ref = g["img"].astype(np.float32)
    H, W = ref.shape[:2]
    r = sess.renderer(scene)
    a, st = r.render(width=W, height=H, spp=spp, seed=11)
    b, _ = r.render(width=W, height=H, spp=spp, seed=22)
    res = mc_bounds(a, b, ref, spp, MC_REF_SPP[scene], scene)
    err, own, _ = res["pixels"]
    return err, own, st


def check_mc_live(sess, scene, spp, W, H):
    """Against the compiled reference run HERE at the same N: r0 = RMSE between two oracle runs (different thread counts:
    its default-seeded per-thread mt19937 streams land on different buckets), then RMSE(ours@N, oracle@N) <= 1.2 r0 and
    per-channel mean |delta| <= 0.002."""
    # one thread against two: the reference's per-thread streams then cover the 64x64 buckets differently - except the very
    # first bucket, which thread 0 renders from a fresh stream in both runs: it is left out of the comparison
    o1, _ = oracle_render(scene, W, H, spp=spp, extra=("--threads", "1"))
    o2, _ = oracle_render(scene, W, H, spp=spp, extra=("--threads", "2"))
    img, st = sess.renderer(scene).render(width=W, height=H, spp=spp, seed=31)
    m = np.ones((H, W), dtype=bool)
    m[:64, :64] = False

    def rm(a, b):
        return float(np.sqrt(((clamp01(a)[m] - clamp01(b)[m]) ** 2).mean()))

    r0 = rm(o1, o2)
    err = 0.5 * (rm(img, o1) + rm(img, o2))
    mean_delta = np.abs(clamp01(img).mean(axis=(0, 1)) - 0.5 * (clamp01(o1).mean(axis=(0, 1)) + clamp01(o2).mean(axis=(0, 1))))
    assert r0 > 0, "the two oracle runs are identical: no noise floor to compare with"
    assert err <= 1.2 * r0, "%s @%d spp: rmse(ours, oracle) %.5f > 1.2 * r0 (r0 = %.5f)" % (scene, spp, err, r0)
    assert mean_delta.max() <= 0.002, "%s: per-channel mean delta %s > 0.002" % (scene, mean_delta)
    return err, r0, mean_delta


def check_primary(sess, scene):
    g = golden("primary", scene)
    rays, ref = g["rays"], g["hits"]
    hits = sess.renderer(scene).trace_closest(rays)
    status_ref = ref[:, 0].astype(np.int32)
    node_ref = ref[:, 1].astype(np.int32)
    agree = (hits["status"] == status_ref) & (hits["node"] == node_ref)
    assert agree.mean() >= 0.9995, "%s: hit/miss or node differs on %d of %d rays" % (scene, (~agree).sum(), len(rays))
    m = agree & (status_ref == 0)
    rel = np.abs(hits["dist"][m] - ref[m, 2]) / np.maximum(1.0, np.abs(ref[m, 2]))
    # the blurred heightfield is summed in float by a -ffast-math reference build: its heights (and so the
    # hit distances) carry ~1e-6 relative noise; everything else agrees to ~1e-12
    tol = 1e-5 if scene == "heightfield" else 1e-8
    ok = rel < tol
    assert ok.mean() >= 0.999, "%s: dist differs (max rel %.3g)" % (scene, rel.max())
    sel = np.where(m)[0][ok]
    scale = np.maximum(1.0, np.abs(ref[sel, 3:6]).max())
    assert np.abs(hits["ip"][sel] - ref[sel, 3:6]).max() < tol * 100 * scale
    assert np.abs(hits["norm"][sel] - ref[sel, 6:9]).max() < tol * 1000
    uvs = np.maximum(1.0, np.abs(ref[sel, 9:11]).max())
    assert np.abs(hits["u"][sel] - ref[sel, 9]).max() < tol * 100 * uvs and np.abs(hits["v"][sel] - ref[sel, 10]).max() < tol * 100 * uvs
    e = status_ref == 1
    both = e & agree
    if both.any():
        assert np.abs(hits["color"][both] - ref[both, 17:20]).max() < 1e-4
    return float(agree.mean())


def check_visible(sess, scene):
    g = golden("visible", scene)
    vis = sess.renderer(scene).trace_visible(g["seg"])
    agree = (vis == (g["vis"] != 0)).mean()
    assert agree >= 0.999, "%s: visible() differs on %.2f%% of segments" % (scene, (1 - agree) * 100)
    return float(agree)


# ---------------------------------------------------------------------------- synthetic terrain (the bench workload)
def terrain_scene_file(api, side, W=96, H=54, spp=16, tmpdir="/tmp"):
    """The bench.py terrain scene with a `side` x `side` vertex grid (procedural mesh, host generator)."""
    import argparse
    import bench
    path = os.path.join(tmpdir, "hxr_terrain_%d_%d.hexray" % (side, os.getpid()))
    with open(path, "w") as f:
        f.write(bench.scene_text(argparse.Namespace(grid_side=side), "synthetic:terrain:%d:0x5EED" % side, W, H, spp))
    return hx.SceneFile(path, api_=api)


def soup_scene_file(api, n_tris, W=96, H=54, spp=16, tmpdir="/tmp"):
    """bench.py's soup workload: n_tris random triangles in [-500, 500]^3 (overlapping, incoherent: deep stacks, many leaves per ray)."""
    import argparse
    import bench
    path = os.path.join(tmpdir, "hxr_soup_%d_%d.hexray" % (n_tris, os.getpid()))
    with open(path, "w") as f:
        f.write(bench.scene_text(argparse.Namespace(grid_side=0, workload="soup", soup_triangles=n_tris), "synthetic:soup:%d:0x5EEE" % n_tris, W, H, spp))
    return hx.SceneFile(path, api_=api)


def check_soup_walk_equals_brute_force(api, n_tris, n_rays, queue_capacity=1 << 20):
    """The KD walk against testing every triangle in index order (the reference's useKDTree=false path) on the soup: rays
    cross hundreds of leaves, the traversal stack overflows its shared-memory part, triangles overlap and intersect each other."""
    rng = np.random.default_rng(n_tris)
    o = rng.uniform(-480, 480, (n_rays, 3))
    d = rng.normal(size=(n_rays, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.zeros((n_rays, 2))], axis=1)
    seg = np.concatenate([o, o + d * rng.uniform(5, 400, (n_rays, 1))], axis=1)
    sf = soup_scene_file(api, n_tris)
    out = []
    for flags in (0, hx.CFG_BRUTE_FORCE_MESHES):
        r = hx.Renderer(api_=api, queue_capacity=queue_capacity, flags=flags).load(sf)
        out.append((r.trace_closest(rays).copy(), r.trace_visible(seg).copy()))
        r.close()
    sf.close()
    (walk, wvis), (brute, bvis) = out
    assert (walk["node"] == brute["node"]).all() and (walk["status"] == brute["status"]).all()
    m = walk["status"] == 0
    hit_soup = int((walk["node"][m] == 0).sum())  # sparse soups are mostly missed: optical depth ~ n_tris * 1.7 * 500 / 1e9
    assert hit_soup >= min(n_rays // 4, max(3, int(n_rays * n_tris * 1e-7))), hit_soup
    for f in ("dist", "ip", "norm", "u", "v"):
        assert np.array_equal(walk[f][m], brute[f][m]), f
    assert np.array_equal(wvis, bvis)
    assert 0 < wvis.sum() <= n_rays


def check_terrain(api, queue_capacity=1 << 20, spp=64):
    """Explicit rays / shadow segments / a GI frame on the 203k-triangle terrain against the compiled reference."""
    g = np.load(os.path.join(GOLDEN, "terrain_320.npz"))
    side, W, H = int(g["side"]), int(g["W"]), int(g["H"])
    sf = terrain_scene_file(api, side, W, H)
    r = hx.Renderer(api_=api, queue_capacity=queue_capacity).load(sf)
    try:
        rays, ref = g["rays"], g["hits"]
        hits = r.trace_closest(np.concatenate([rays, np.zeros((len(rays), 2))], axis=1))
        agree = (hits["status"] == ref[:, 0].astype(np.int32)) & (hits["node"] == ref[:, 1].astype(np.int32))
        assert agree.all(), "terrain: hit/miss or node differs on %d of %d rays" % ((~agree).sum(), len(rays))
        m = ref[:, 0] == 0
        rel = np.abs(hits["dist"][m] - ref[m, 2]) / np.maximum(1.0, np.abs(ref[m, 2]))
        assert rel.max() < 1e-9, "terrain: dist differs (max rel %.3g)" % rel.max()
        assert np.abs(hits["norm"][m] - ref[m, 6:9]).max() < 1e-7
        vis = r.trace_visible(g["seg"])
        assert (vis == (g["vis"] != 0)).mean() >= 0.9995
        # GI frame against the converged reference frame
        refimg = g["img"].astype(np.float32)
        a, st = r.render(width=W, height=H, spp=spp, seed=3)
        b, _ = r.render(width=W, height=H, spp=spp, seed=4)
        res = mc_bounds(a, b, refimg, spp, 1024, "terrain GI")
        err, own, _ = res["pixels"]
        return err, own, st
    finally:
        r.close()
        sf.close()


def check_layout_switches(api, side=320, queue_capacity=1 << 20):
    """The optional data layouts change what is FETCHED, never the result: the 32-byte packed triangle record (HXR_TRI_PACK,
    filter decisions differ, the exact test decides the same winner) and the full-attribute gather on untextured nodes
    (HXR_FULL_ATTR) must give bit-identical hit records, shadow answers and frames."""
    g = np.load(os.path.join(GOLDEN, "terrain_320.npz"))
    W, H = int(g["W"]), int(g["H"])
    rays = np.concatenate([g["rays"], np.zeros((len(g["rays"]), 2))], axis=1)
    out = []
    for env in ({}, {"HXR_TRI_PACK": "1"}, {"HXR_FULL_ATTR": "1"}):
        os.environ.update(env)
        try:
            sf = terrain_scene_file(api, side, W, H)
            r = hx.Renderer(api_=api, queue_capacity=queue_capacity).load(sf)
        finally:
            for k in env:
                del os.environ[k]
        hits = r.trace_closest(rays).copy()
        vis = r.trace_visible(g["seg"]).copy()
        img, st = r.render(width=W, height=H, spp=4, seed=9)
        out.append((hits, vis, img.copy()))
        r.close()
        sf.close()
    for hits, vis, img in out[1:]:
        for f in ("status", "node", "dist", "ip", "norm", "u", "v"):
            assert np.array_equal(hits[f], out[0][0][f]), f
        assert np.array_equal(vis, out[0][1])
        # (radiance is summed per pixel with atomic float adds: the order, hence the last bits, varies from run to run)
        assert np.allclose(img, out[0][2], rtol=2e-5, atol=1e-6), float(np.abs(img - out[0][2]).max())


def grazing_rays(renderer, n, seed=7):
    """Rays that leave the terrain surface almost tangentially (what deep GI bounces produce): the float bounds of the walk
    decide little on them, so 1-2 % fill their candidate record and are finished by the overflow kernel."""
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(-480, 480, n), np.full(n, 300.0), rng.uniform(-480, 480, n)], axis=1)
    d = np.tile([0.0, -1.0, 0.0], (n, 1))
    h = renderer.trace_closest(np.concatenate([o, d, np.zeros((n, 2))], axis=1))
    m = (h["status"] == 0) & (h["node"] == 0)
    ip, nr = h["ip"][m], h["norm"][m]
    az = rng.uniform(0, 2 * np.pi, len(ip))
    d2 = np.stack([np.cos(az), rng.uniform(-0.02, 0.1, len(ip)), np.sin(az)], axis=1)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    o2 = ip + nr * 1e-6
    rays = np.concatenate([o2, d2, np.zeros((len(ip), 2))], axis=1)
    seg = np.concatenate([o2, o2 + d2 * rng.uniform(20, 900, (len(ip), 1))], axis=1)
    return rays, seg


def check_finish_forms(api, side=708, n=30000, queue_capacity=1 << 20):
    """Rays whose candidate record overflows are finished by one warp each (k_finish_warp: the lanes share a leaf's exact
    tests, the winner comes from a shuffle reduction) or by one thread each (k_finish, HXR_FINISH_WARP=0). Both must return
    what testing every triangle in index order returns (the reference's useKDTree=false path), bit for bit."""
    out = []
    rays = seg = None
    for env, flags in (({}, 0), ({"HXR_FINISH_WARP": "0"}, 0), ({}, hx.CFG_BRUTE_FORCE_MESHES)):
        os.environ.update(env)
        try:
            sf = terrain_scene_file(api, side)
            r = hx.Renderer(api_=api, queue_capacity=queue_capacity, flags=flags).load(sf)
        finally:
            for k in env:
                del os.environ[k]
        if rays is None:
            rays, seg = grazing_rays(r, n)
            assert len(rays) > n // 2
        hits = r.trace_closest(rays).copy()
        vis = r.trace_visible(seg).copy()
        overflow = 0
        if flags == 0:  # a GI frame of the same scene: its stats say that the overflow path is really taken
            _, st = r.render(width=96, height=54, spp=8, seed=5)
            overflow = st["cand_overflow"]
        out.append((hits, vis, overflow))
        r.close()
        sf.close()
    assert out[0][2] > 0 and out[1][2] > 0, "no ray overflowed its candidate record: the test does not reach the finish kernels"
    for hits, vis, _ in out[1:]:
        for f in ("status", "node", "dist", "ip", "norm", "u", "v"):
            assert np.array_equal(hits[f], out[0][0][f]), f
        assert np.array_equal(vis, out[0][1])
    return out[0][2]


def check_stereo(sess, scene, spp=0):
    """Anaglyph frame of a bundled scene with stereoSeparation added, against the compiled reference's frame."""
    g = golden("stereo", scene)
    ref = g["img"].astype(np.float32)
    H, W = ref.shape[:2]
    r = sess.renderer(stereo_scene(scene, STEREO[scene]))
    if spp == 0:
        img, st = r.render(width=W, height=H)
        d = np.abs(clamp01(img) - clamp01(ref)).max(axis=2)
        frac = float((d <= PIXEL_TOL + 6e-4).mean())
        assert frac >= PIXEL_FRACTION, "stereo %s: only %.4f%% of pixels within 1/255 (max diff %.4f)" % (scene, frac * 100, d.max())
        return frac
    a, st = r.render(width=W, height=H, spp=spp, seed=11)
    b, _ = r.render(width=W, height=H, spp=spp, seed=22)
    res = mc_bounds(a, b, ref, spp, 4096, "stereo " + scene)
    return res["pixels"][0]


def many_meshes_scene(n_side=6):
    """n_side^2 instanced little meshes (44-triangle dice, a 10-triangle box) over a floor: many big-mesh nodes per scene."""
    lines = ["GlobalSettings {\n\tframeWidth 160\n\tframeHeight 120\n\tambientLight (0.2, 0.2, 0.2)\n\tmaxTraceDepth 4\n}",
             "Camera camera {\n\tpos (0, 60, -140)\n\taspectRatio 1.33333\n\tpitch -22\n\tfov 90\n}",
             "PointLight l1 {\n\tpos (-60, 160, -80)\n\tcolor (1, 1, 1)\n\tpower 40000\n}",
             "Plane floor {\n\ty 0\n\tlimit 400\n}", "Mesh dice {\n\tfile \"geom/truncated_cube.obj\"\n\tfaceted true\n}",
             "Mesh box {\n\tfile \"cornell/tallblock.obj\"\n\tfaceted true\n}",
             "Lambert grey {\n\tcolor (0.7, 0.7, 0.7)\n}", "Lambert red {\n\tcolor (0.8, 0.3, 0.2)\n}",
             "Reflection mirror {\n\tmultiplier 0.8\n}", "Node nfloor {\n\tgeometry floor\n\tshader grey\n}"]
    k = 0
    for i in range(n_side):
        for j in range(n_side):
            geom, scale = ("dice", 6.0) if (i + j) % 2 == 0 else ("box", 0.05)
            shader = ("red", "mirror", "grey")[k % 3]
            lines.append("Node n%d {\n\tgeometry %s\n\tshader %s\n\tscale (%g, %g, %g)\n\trotate (%d, %d, 0)\n\ttranslate (%g, %g, %g)\n}" % (
                k, geom, shader, scale, scale, scale, 17 * k % 360, 29 * k % 90, (i - n_side / 2) * 28.0, 10.0, (j - n_side / 2) * 28.0))
            k += 1
    return scratch_scene("_many_meshes", "\n".join(lines) + "\n")


def check_many_meshes(api):
    """36 mesh nodes: the walk (with its per-ray task budget) against brute force over every triangle, bit for bit."""
    scene = many_meshes_scene()
    sf = hx.SceneFile(scene_path(scene), api_=api)
    out = []
    for flags in (0, hx.CFG_BRUTE_FORCE_MESHES):
        r = hx.Renderer(api_=api, queue_capacity=1 << 16, flags=flags).load(sf)
        out.append(r.render()[0].copy())
        r.close()
    sf.close()
    assert float(out[0].mean()) > 0.02
    assert np.abs(out[0] - out[1]).max() < 1e-5


def check_screenshot(sess, tmp_path):
    r = sess.renderer("simple")
    img, _ = r.render(width=203, height=77)  # a row size that needs BMP padding
    a, b = str(tmp_path / "device.bmp"), str(tmp_path / "host.bmp")
    r.save_frame_bmp(a)
    hx.save_image(b, img, api_=sess.api)
    da, db = open(a, "rb").read(), open(b, "rb").read()
    assert len(da) == 54 + 77 * 612 and da == db


# ---------------------------------------------------------------------------- features no bundled scene uses
FEATURES_WHITTED = """GlobalSettings {
	frameWidth 320
	frameHeight 240
	ambientLight (0.15, 0.15, 0.15)
	maxTraceDepth 5
	wantAA true
}
Camera camera {
	pos (0, 70, -150)
	aspectRatio 1.33333
	pitch -20
	fov 95
}
PointLight l1 {
	pos (-60, 170, -110)
	color (1, 1, 1)
	power 45000
}
Plane floor {
	y 0
	limit 350
}
CheckerTexture checker {
	color1 (0.9, 0.9, 0.9)
	color2 (0.15, 0.25, 0.55)
	scaling 0.1
}
CheckerTexture smallChecker {
	color1 (0.95, 0.8, 0.2)
	color2 (0.2, 0.6, 0.3)
	scaling 0.125
}
Lambert floorShader {
	color (1, 1, 1)
	texture checker
}
Lambert sphereShader {
	color (1, 1, 1)
	texture smallChecker
}
Phong phong {
	color (0.8, 0.25, 0.2)
	specular (0.7, 0.7, 0.7)
	exponent 40
}
Reflection mirror {
	multiplier 0.85
}
Refraction glass {
	ior 1.45
	multiplier 0.9
}
Const emissive {
	color (0.3, 0.9, 0.4)
}
Sphere uvSphere {
	O (0, 0, 0)
	R 16
	uvscaling 6
}
Cube cubeA {
	O (0, 0, 0)
	side 26
}
Cube cubeB {
	O (9, 9, -9)
	side 22
}
Sphere ball {
	O (0, 0, 0)
	R 17
}
CSGUnion csgUnion {
	left cubeA
	right ball
}
CSGInter csgInter {
	left cubeA
	right ball
}
CSGDiff csgNested {
	left csgUnion
	right cubeB
}
Mesh culledDice {
	file "geom/truncated_cube.obj"
	faceted true
	backfaceCulling true
}
Node nFloor {
	geometry floor
	shader floorShader
}
Node nSphere {
	geometry uvSphere
	shader sphereShader
	translate (-70, 18, 10)
}
Node nUnion {
	geometry csgUnion
	shader phong
	rotate (25, 15, 0)
	translate (-25, 20, 30)
}
Node nInter {
	geometry csgInter
	shader mirror
	rotate (40, 0, 10)
	translate (25, 18, -10)
}
Node nNested {
	geometry csgNested
	shader glass
	rotate (-30, 20, 0)
	translate (70, 20, 35)
}
Node nDice {
	geometry culledDice
	shader emissive
	scale (9, 9, 9)
	rotate (33, 21, 0)
	translate (0, 14, -55)
}
"""

FEATURES_STOCHASTIC = """GlobalSettings {
	frameWidth 160
	frameHeight 120
	ambientLight (0.2, 0.2, 0.2)
	maxTraceDepth 4
	wantAA false
}
Camera camera {
	pos (0, 60, -140)
	aspectRatio 1.33333
	pitch -18
	fov 90
	dof true
	autoFocus true
	fNumber 4
	numSamples 16
}
PointLight l1 {
	pos (-60, 170, -110)
	color (1, 1, 1)
	power 45000
}
Plane floor {
	y 0
	limit 350
}
Lambert grey {
	color (0.75, 0.75, 0.75)
}
Reflection glossy {
	multiplier 0.9
	glossiness 0.78
	numSamples 6
}
Layered floorShader {
	layer grey (1, 1, 1)
	layer glossy (0.35, 0.35, 0.35)
}
Phong phong {
	color (0.2, 0.5, 0.85)
	specular (0.6, 0.6, 0.6)
	exponent 25
}
Sphere ball {
	O (0, 0, 0)
	R 22
}
Cube box {
	O (0, 0, 0)
	side 30
}
Node nFloor {
	geometry floor
	shader floorShader
}
Node nBall {
	geometry ball
	shader phong
	translate (-30, 22, 20)
}
Node nBox {
	geometry box
	shader glossy
	rotate (30, 0, 0)
	translate (35, 15, 0)
}
"""


def feature_scene(kind):
    return scratch_scene("_features_%s" % kind, FEATURES_WHITTED if kind == "whitted" else FEATURES_STOCHASTIC)


def check_features(sess, frames=24):
    """CSG union / intersection / nested CSG, Sphere uvscaling, backface culling, Const, Phong, Refraction (deterministic, per
    pixel) and glossy Reflection + DOF with autoFocus (stochastic, averaged frames) against the compiled reference."""
    g = golden("features", "whitted")
    ref = g["img"].astype(np.float32)
    H, W = ref.shape[:2]
    img, st = sess.renderer(feature_scene("whitted")).render(width=W, height=H)
    d = np.abs(clamp01(img) - clamp01(ref)).max(axis=2)
    frac = float((d <= PIXEL_TOL + 6e-4).mean())
    assert frac >= PIXEL_FRACTION, "features: only %.4f%% of pixels within 1/255 (max diff %.4f)" % (frac * 100, d.max())
    g = golden("features", "stochastic")
    ref = g["img"].astype(np.float32)
    H, W = ref.shape[:2]
    r = sess.renderer(feature_scene("stochastic"))
    acc = np.zeros_like(ref)
    for s in range(frames):
        acc += r.render(width=W, height=H, seed=500 + s)[0]
    acc /= frames
    dd = np.abs(clamp01(acc) - clamp01(ref))
    assert dd.mean() < 0.004 and np.sqrt((dd ** 2).mean()) < 0.012, (float(dd.mean()), float(np.sqrt((dd ** 2).mean())))
    return frac


def check_small_queue(api):
    """A ray queue far smaller than the frame: the driver renders in batches (and redoes the frame in smaller ones when a
    ray tree outgrows the queue); the image must not depend on the queue size."""
    out = []
    for qc in (1 << 20, 4096):
        sf = hx.SceneFile(scene_path("meshes"), api_=api)
        r = hx.Renderer(api_=api, queue_capacity=qc).load(sf)
        out.append(r.render(width=200, height=150)[0].copy())
        r.close()
        sf.close()
    assert np.abs(out[0] - out[1]).max() < 1e-4


def check_device_kd_build(api, kind, size, n_rays):
    """the KD-tree built ON THE DEVICE (level-synchronous SAH build, csrc/device/kdbuild.h) must serve the walk
    exactly like the host-built one: explicit rays through the device-built tree equal testing every triangle in index order (the
    reference's useKDTree=false path, src/mesh.cpp:255-262) bit for bit, and the tree is of the same quality as the host's."""
    rng = np.random.default_rng(size)
    if kind == "terrain":
        sf = terrain_scene_file(api, size)
        o = np.concatenate([np.tile([[0.0, 150.0, -600.0]], (n_rays // 2, 1)),
                            np.stack([rng.uniform(-500, 500, n_rays // 2), rng.uniform(-30, 200, n_rays // 2), rng.uniform(-500, 500, n_rays // 2)], 1)])
        d = np.concatenate([np.stack([rng.uniform(-0.7, 0.7, n_rays // 2), rng.uniform(-0.6, 0.05, n_rays // 2), np.ones(n_rays // 2)], 1),
                            rng.normal(size=(n_rays // 2, 3))])
    else:
        sf = soup_scene_file(api, size)
        o = rng.uniform(-480, 480, (n_rays, 3))
        d = rng.normal(size=(n_rays, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.zeros((n_rays, 2))], axis=1)
    hits, info = {}, {}
    for name, flags in (("host", 0), ("device", hx.CFG_DEVICE_KD_BUILD), ("brute", hx.CFG_BRUTE_FORCE_MESHES)):
        r = hx.Renderer(api_=api, queue_capacity=1 << 20, flags=flags).load(sf)
        hits[name] = r.trace_closest(rays).copy()
        info[name] = r.accel_info(0)
        r.close()
    sf.close()
    assert info["device"]["device_build"] == 1 and info["host"]["device_build"] == 0 and info["device"]["from_cache"] == 0
    assert info["device"]["n_triangles"] == info["host"]["n_triangles"]
    for k in ("tri_refs", "leaves", "nodes"):
        assert abs(info["device"][k] / max(1, info["host"][k]) - 1) < 0.15, (k, info["device"][k], info["host"][k])
    assert info["device"]["max_depth"] <= info["host"]["max_depth"] + 2
    for name in ("device", "host"):
        a, b = hits[name], hits["brute"]
        assert (a["node"] == b["node"]).all() and (a["status"] == b["status"]).all(), name
        m = a["status"] == 0
        assert np.array_equal(a["dist"][m], b["dist"][m]) and np.array_equal(a["ip"][m], b["ip"][m]) and np.array_equal(a["norm"][m], b["norm"][m]), name
    assert (hits["brute"]["node"] >= 0).sum() > n_rays // 8
    return info
