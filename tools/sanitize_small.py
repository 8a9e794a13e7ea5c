"""Smallest run that touches every kernel: for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexray_b200 as hx  # noqa: E402
import hxr_testlib as T  # noqa: E402

api = hx.api()
for scene, kw in (("kdtree_test", dict(width=96, height=72)), ("meshes", dict(width=64, height=48)), ("heightfield", dict(width=64, height=48)),
                  ("cornell_box", dict(width=48, height=48, spp=4)), ("hw10/bokeh", dict(width=48, height=36, spp=2))):
    sf = hx.SceneFile(T.scene_path(scene), api_=api)
    r = hx.Renderer(api_=api, queue_capacity=1 << 16).load(sf)
    img, st = r.render(**kw)
    print(scene, img.shape, float(img.mean()), st["rays_closest"], st["rays_shadow"])
    r.close()
    sf.close()
sf = T.terrain_scene_file(api, 120, 64, 36, 2)
r = hx.Renderer(api_=api, queue_capacity=1 << 16).load(sf)
img, st = r.render(width=64, height=36, spp=2, mode=hx.MODE_MONTECARLO, flags=hx.RENDER_COUNT_TRAVERSAL)
print("terrain", float(img.mean()), st["rays_closest"], st["kd_inner"], st["tri_tests"])
img, st = r.render(width=64, height=36, spp=2, mode=hx.MODE_MONTECARLO)
rays = np.zeros((64, 8)); rays[:, 1] = 150; rays[:, 2] = -600; rays[:, 5] = 1; rays[:, 4] = -0.3
print("trace", r.trace_closest(rays)["node"][:4], r.trace_visible(np.random.default_rng(0).uniform(-400, 400, (64, 6)))[:4])
r.close()
sf.close()
print("sanitize_small done")
