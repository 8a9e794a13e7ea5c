"""Build libhexray_b200.so in-tree: nvcc (sm_100a) for the kernels, g++ for the host front-end.

The library is the product: hexray_b200 loads it with ctypes and refuses to work without it.
Usage: python -m hexray_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libhexray_b200.so")
OBJ = os.path.join(HERE, "build")

HOST_SOURCES = ["abi.cpp", "renderer.cpp", "multi.cpp", "kd_device_build.cpp", "host/scene.cpp", "host/mesh.cpp", "host/flatten.cpp",
                "host/bitmap.cpp", "host/kdtree.cpp", "host/cache.cpp"]
CUDA_SOURCES = ["device/launch_cuda.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _all_inputs():
    inputs = []
    for root, _, files in os.walk(CSRC):
        inputs += [os.path.join(root, f) for f in files]
    inputs.append(os.path.join(HERE, "..", "include", "hxr.h"))
    return inputs


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: A/B builds of the kernels with other compile-time knobs (tools/build_variants.sh), e.g.
    build(out="build_ab/gi3.so", defines=["HXR_SHADE_GI_BLOCKS=3"]); objects of a variant build go to a private directory."""
    global OUT, OBJ
    if out is not None:
        saved = (OUT, OBJ)
        OUT, OBJ = os.path.abspath(out), os.path.abspath(out) + ".obj"
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        try:
            NVCC_FLAGS.extend("-D" + d for d in defines)
            return build(force=True, verbose=verbose)
        finally:
            del NVCC_FLAGS[len(NVCC_FLAGS) - len(defines):]
            OUT, OBJ = saved
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(_all_inputs()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cxx = os.environ.get("CXX", "g++")
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for src in HOST_SOURCES:
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(o)
        procs.append((src, subprocess.Popen([cxx] + CXX_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src in CUDA_SOURCES:
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(o)
        procs.append((src, subprocess.Popen([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (src, out))
        if p.returncode != 0:
            raise RuntimeError("compiling %s failed:\n%s" % (src, out))
    with open(os.path.join(OBJ, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", OUT] + objs + ["-lz", "-ldl", "-lpthread", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return OUT


if __name__ == "__main__":
    if "--out" in sys.argv:  # python -m hexray_b200.build --out build_ab/x.so -DNAME=VALUE ...
        print(build(out=sys.argv[sys.argv.index("--out") + 1], defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
