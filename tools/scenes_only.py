"""The bench's bundled-scene block alone (no terrain step): python tools/scenes_only.py [lib.so ...]; each library is loaded in its
own process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import bench
    import hexray_b200 as hx
    os.environ.setdefault("HEXRAY_DATA", bench.DATA)
    for v in bench.extra_scenes(hx, 0, False):
        print("  ", v["scene"], {k: (round(x["ms_per_frame"], 3), round(x["mrays_per_s"], 1)) for k, x in v.items() if isinstance(x, dict) and "ms_per_frame" in x})
        for k, x in v.items():
            if isinstance(x, dict) and "roofline" in x:
                print("      ", k, "roofline", {a: (round(b, 4) if isinstance(b, float) else b) for a, b in x["roofline"].items() if a not in ("note", "l2_source")})
    sys.exit(0)
lib = os.path.join(ROOT, "hexray_b200", "libhexray_b200.so")
keep = open(lib, "rb").read()
try:
    for v in sys.argv[1:] or ["tree"]:
        open(lib, "wb").write(keep if v == "tree" else open(v, "rb").read())
        print(v, flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], check=False)
finally:
    open(lib, "wb").write(keep)
