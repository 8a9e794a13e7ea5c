import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs the product library libhexray_b200.so)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on the CPU tier")


@pytest.fixture(scope="session")
def emu_api():
    """tests/emu/libhxr_emu.so: host build of the per-ray functions (test infrastructure, not the product)."""
    from hexray_b200 import capi
    if os.environ.get("HXR_EMU_LIB"):  # e.g. the sanitizer build (tests/emu/build_emu.sh asan)
        return capi.Api(os.path.abspath(os.environ["HXR_EMU_LIB"]))
    so = os.path.join(ROOT, "tests", "emu", "libhxr_emu.so")
    src_dirs = [os.path.join(ROOT, "hexray_b200", "csrc"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "include")]
    newest = 0
    for d in src_dirs:
        for r, _, fs in os.walk(d):
            for f in fs:
                if f.endswith((".h", ".cpp", ".cu", ".sh")):
                    newest = max(newest, os.path.getmtime(os.path.join(r, f)))
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run([os.path.join(ROOT, "tests", "emu", "build_emu.sh")], check=True)
    return capi.Api(so)


@pytest.fixture(scope="session")
def gpu_api():
    """The product library; fails loudly when it is missing."""
    import hexray_b200
    return hexray_b200.api()
