"""Render every bundled scene once on the GPU, compare with the live compiled reference where it
travelled, and print one JSON line per scene (timings, ray counts, parity)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexray_b200 as hx  # noqa: E402
import hxr_testlib as T  # noqa: E402

SCENES = ["simple", "meshes", "kdtree_test", "heightfield", "bumpmap", "Lecture8", "beer", "boxed",
          "cornell_box", "smallpt", "hw12/sphtri", "zaphod", "hw10/bokeh"]


def main():
    only = sys.argv[1:] or SCENES
    for scene in only:
        t0 = time.time()
        sf = hx.SceneFile(T.scene_path(scene))
        r = hx.Renderer()
        r.load(sf)
        t_load = time.time() - t0
        best = None
        for rep in range(3):
            img, st = r.render(seed=rep)
            if best is None or st["render_ms"] < best["render_ms"]:
                best = st
        rec = {"scene": scene, "load_s": round(t_load, 3), "size": list(r.frame_size()),
               "render_ms": round(best["render_ms"], 3), "rays_closest": best["rays_closest"], "rays_shadow": best["rays_shadow"],
               "mrays_s": round((best["rays_closest"] + best["rays_shadow"]) / best["render_ms"] / 1e3, 2),
               "launches": best["kernel_launches"], "spp": best["spp_done"], "aa_pixels": best["aa_pixels"]}
        if T.have_oracle() and "--no-ref" not in sys.argv:
            W, H = r.frame_size()
            t0 = time.time()
            ref, info = T.oracle_render(scene, W, H)
            rec["ref_ms"] = info["best_ms"]
            frac, mx = T.pixel_match_fraction(img, ref)
            rec["within_1_255"] = round(frac, 5)
            rec["rmse"] = round(T.rmse(img, ref), 5)
            rec["mean_delta"] = [round(float(x), 5) for x in np.abs(T.clamp01(img).mean(axis=(0, 1)) - T.clamp01(ref).mean(axis=(0, 1)))]
        print(json.dumps(rec), flush=True)
        r.close()
        sf.close()


if __name__ == "__main__":
    main()
