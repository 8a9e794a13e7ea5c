"""Per-kernel-class device time of one bundled scene (profiling events on): python tools/scene_prof.py scene [W H] [spp]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hexray_b200 as hx  # noqa: E402

scene = sys.argv[1]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H = int(sys.argv[3]) if len(sys.argv) > 3 else 0
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sf = hx.SceneFile(os.path.join(hx.data_root(), scene + ".hexray"))
r = hx.Renderer().load(sf)
for prof in (False, True):
    r.set_profiling(prof)
    for i in range(4):
        # per-kernel times: every kernel on one stream (the two-lane frame's event pairs overlap)
        img, st = r.render(width=W, height=H, spp=spp, seed=i, flags=hx.RENDER_ONE_LANE if prof else 0)
    rays = st["rays_closest"] + st["rays_shadow"]
    keys = ("render_ms", "walk_ms", "setup_ms", "finish_ms", "shade_ms", "shadow_resolve_ms", "gen_ms", "other_ms", "kernel_launches", "aa_pixels", "cand_overflow")
    print(json.dumps({"scene": scene, "profiling": prof, "rays": rays, "mrays_s": rays / st["render_ms"] / 1e3, **{k: round(st[k], 3) if isinstance(st[k], float) else st[k] for k in keys}}))
