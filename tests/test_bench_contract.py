"""CPU tier: the parts of bench.py's contract that need no GPU - the reference arm's JSON line (the compiled reference on the
host cores, nothing of the product loaded), stdout carrying that line alone, and our arm refusing to run without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + list(args), capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "hexray_ref_count")):
        pytest.skip("oracle/_ref is not built (python -c 'import __graft_entry__ as g; g.build()')")
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--grid-side", "64", "--ref-width", "64", "--ref-height", "36", "--ref-spp", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "Mrays/s" and j["unit"] == "Mrays/s" and j["higher_is_better"] is True
    assert j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "terrain" in j["config"]["workload"]
    # the arm must not load the product: it times the reference alone
    assert "libhexray_b200" not in r.stderr


def test_stdout_carries_the_json_line_alone():
    # C libraries write to file descriptor 1 behind Python's back (NCCL's version banner under torchrun): it points at stderr
    # while the bench runs
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "with bench.JsonOnlyStdout():\n"
            "    os.write(1, b'banner from a C library\\n')\n"
            "    bench.print('{\"metric\": \"Mrays/s\"}')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "Mrays/s"}\n'
    assert "banner from a C library" in r.stderr


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "0", "--grid-side", "64", "--spp", "1", "--no-cpu-baseline", "--no-extra")
    assert r.returncode != 0  # no CPU fallback: the arm fails loudly
    assert r.stdout.strip() == ""
