"""The walk's correctness rests on the FP32 triangle filter being CONSERVATIVE (csrc/device/isect.h: tri_filter):
it may say MAYBE as often as it likes, but MISS must imply that the reference's exact double test (tri_core,
src/mesh.cpp:178-196) rejects the pair, and CERTAIN must imply that it accepts it at a parameter <= ghi.
Millions of random and adversarial (ray, triangle) pairs through the host build of the same function."""
import ctypes as C

import numpy as np
import pytest

from hexray_b200 import capi

MISS, MAYBE, CERTAIN = 0, 1, 2


def run_filter(api, rays, tris, tbest, backface=0, packed=False):
    n = len(rays)
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    tris = np.ascontiguousarray(tris, dtype=np.float64)
    tbest = np.ascontiguousarray(tbest, dtype=np.float64)
    cls = np.zeros(n, dtype=np.int32)
    ghi = np.zeros(n, dtype=np.float32)
    exact = np.zeros(n, dtype=np.int32)
    gamma = np.zeros(n, dtype=np.float64)
    fn = api.lib.hxr_test_tri_filter_packed if packed else api.lib.hxr_test_tri_filter
    st = fn(n, rays.ctypes.data, tris.ctypes.data, tbest.ctypes.data, backface, cls.ctypes.data, ghi.ctypes.data,
            exact.ctypes.data, gamma.ctypes.data)
    assert st == capi.HXR_OK
    return cls, ghi, exact.astype(bool), gamma


def make_pairs(rng, n, world, size, dist, aim, graze_fraction=0.25):
    """triangles of edge ~size around points at |coordinate| ~world; rays from ~dist away aimed at a point of the triangle's
    plane: inside (aim='in'), on an edge or vertex (aim='edge'), just outside (aim='near'), or anywhere (aim='any')"""
    centre = rng.uniform(-world, world, (n, 3))
    A = centre + rng.normal(0, size, (n, 3))
    B = centre + rng.normal(0, size, (n, 3))
    Cc = centre + rng.normal(0, size, (n, 3))
    if aim == "in":
        w = rng.dirichlet([1, 1, 1], n)
    elif aim == "edge":
        w = rng.dirichlet([1, 1, 1], n)
        k = rng.integers(0, 3, n)
        w[np.arange(n), k] = 0.0                      # exactly on an edge
        z = rng.random(n) < 0.3
        w[z, (k[z] + 1) % 3] = 0.0                    # ... or on a vertex
        w /= np.maximum(w.sum(1, keepdims=True), 1e-300)
    elif aim == "near":
        w = rng.dirichlet([1, 1, 1], n)
        k = rng.integers(0, 3, n)
        w[np.arange(n), k] = -10.0 ** rng.uniform(-12, -1, n)   # outside by a relative hair .. a tenth
        w /= w.sum(1, keepdims=True)
    else:
        w = rng.normal(0.33, 1.0, (n, 3))
        w /= w.sum(1, keepdims=True)
    target = w[:, :1] * A + w[:, 1:2] * B + w[:, 2:3] * Cc
    d = rng.normal(size=(n, 3))
    graze = rng.random(n) < graze_fraction           # a quarter of the rays graze the triangle's plane
    nrm = np.cross(B - A, Cc - A)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    d[graze] -= (d[graze] * nrm[graze]).sum(1, keepdims=True) * nrm[graze] * (1 - 10.0 ** rng.uniform(-9, -1, (graze.sum(), 1)))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t = dist * 10.0 ** rng.uniform(-2, 0.3, (n, 1))
    o = target - d * t
    rays = np.concatenate([o, d], axis=1)
    tris = np.concatenate([A, B, Cc], axis=1)
    return rays, tris, t[:, 0]


@pytest.mark.parametrize("packed", [False, True], ids=["f32", "packed32B"])
@pytest.mark.parametrize("world,size,dist", [(1.0, 0.3, 3.0), (500.0, 0.5, 300.0), (500.0, 0.5, 2.0), (1000.0, 1e-3, 50.0), (5.0, 5.0, 1e-3),
                                             (1e4, 30.0, 1e4), (0.0, 1e-4, 1e-2)])
def test_filter_is_conservative(emu_api, world, size, dist, packed):
    # both forms of the record: TriF32 (48 B) and TriPacked (32 B: block-floating-point edges, N rebuilt in the filter)
    rng = np.random.default_rng(int(world * 7 + size * 1e4 + dist * 13) % (2 ** 31))
    n = 120000
    seen = np.zeros(3, dtype=np.int64)
    for aim in ("in", "edge", "near", "any"):
        rays, tris, t = make_pairs(rng, n, world, size, dist, aim)
        for tb in (np.full(n, 1e99), t * (1 + rng.normal(0, 1e-7, n)), t * rng.uniform(0.2, 3.0, n)):
            for backface in (0, 1):
                cls, ghi, exact, gamma = run_filter(emu_api, rays, tris, tb, backface, packed)
                assert not (exact & (cls == MISS)).any(), "filter said MISS for %d pairs the exact test accepts (%s)" % ((exact & (cls == MISS)).sum(), aim)
                c = cls == CERTAIN
                # CERTAIN promises a triangle hit at gamma <= ghi; the exact test may still reject it for lying beyond tbest
                if tb.min() > 1e98:
                    assert exact[c].all(), "filter said CERTAIN for %d pairs the exact test rejects (%s)" % ((~exact[c]).sum(), aim)
                    assert (gamma[c] <= ghi[c].astype(np.float64)).all()
                seen += np.bincount(cls, minlength=3)
    assert seen[MISS] > 0 and seen[CERTAIN] > 0  # the adversarial mix still leaves both verdicts in play


@pytest.mark.parametrize("packed", [False, True], ids=["f32", "packed32B"])
def test_filter_decides_typical_pairs(emu_api, packed):
    # the bench's regime (coordinates ~500, edges ~0.5, rays from a few hundred units): nearly everything is decided in float
    rng = np.random.default_rng(5)
    rays, tris, t = make_pairs(rng, 200000, 500.0, 0.5, 200.0, "any", graze_fraction=0.0)
    cls, ghi, exact, gamma = run_filter(emu_api, rays, tris, np.full(len(rays), 1e99), packed=packed)
    assert (cls == MAYBE).mean() < 0.02
    assert not (exact & (cls == MISS)).any()
