"""N frames of one bundled scene, nothing else (for ncu): python tools/one_frame.py scene W H [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hexray_b200 as hx  # noqa: E402

scene, W, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 2
sf = hx.SceneFile(os.path.join(hx.data_root(), scene + ".hexray"))
r = hx.Renderer().load(sf)
for i in range(frames):
    img, st = r.render(width=W, height=H, seed=i, flags=hx.RENDER_ONE_LANE)
print(scene, st["kernel_launches"], "launches", st["rays_closest"] + st["rays_shadow"], "rays", round(st["render_ms"], 3), "ms")
