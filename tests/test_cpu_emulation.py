"""CPU tier: the host front-end, the frame driver and the per-ray functions (compiled for the host,
tests/emu) against the committed reference fixtures. The product itself is exercised by test_gpu_*.py."""
import pytest

import hxr_testlib as T


@pytest.fixture(scope="module")
def sess(emu_api):
    s = T.Session(emu_api, queue_capacity=1 << 20)
    yield s
    s.close()


@pytest.mark.parametrize("scene", ["kdtree_test", "heightfield", "smallpt", "boxed", "hw10/bokeh"])
def test_primary_hits(sess, scene):
    T.check_primary(sess, scene)


@pytest.mark.parametrize("scene", ["kdtree_test", "smallpt", "boxed"])
def test_visible(sess, scene):
    T.check_visible(sess, scene)


@pytest.mark.parametrize("scene", ["simple", "meshes", "kdtree_test", "heightfield", "bumpmap"])
def test_whitted_parity(sess, scene):
    T.check_whitted(sess, scene)


@pytest.mark.parametrize("scene,spp", [("hw12/sphtri", 48), ("zaphod", 32)])
def test_montecarlo_parity(sess, scene, spp):
    T.check_mc(sess, scene, spp)


def test_terrain_walk(emu_api):
    # the conservative FP32 block walk (host form) on the bench's terrain; its camera sits ON a split plane
    T.check_terrain(emu_api, spp=8)


def test_soup_walk_equals_brute_force(emu_api):
    T.check_soup_walk_equals_brute_force(emu_api, 100000, 1500)


def test_grazing_rays_equal_brute_force(emu_api):
    # rays leaving the terrain almost tangentially: 1-2 % overflow their candidate record and go through finish_overflowed_ray
    assert T.check_finish_forms(emu_api, side=320, n=1500) > 0


def test_layout_switches_do_not_change_results(emu_api):
    T.check_layout_switches(emu_api)


def test_stereo_anaglyph(sess):
    # src/main.cpp:234-248: two traces per sample mixed into an anaglyph; Whitted + AA (deterministic) and GI
    T.check_stereo(sess, "kdtree_test")
    T.check_stereo(sess, "cornell_box", spp=32)


def test_many_mesh_nodes(emu_api):
    T.check_many_meshes(emu_api)


def test_device_screenshot_equals_host_save(sess, tmp_path):
    # hxr_save_frame_bmp (8-bit conversion on the device) writes the same bytes as hxr_save_image of the downloaded frame
    T.check_screenshot(sess, tmp_path)


def test_features_no_bundled_scene_uses(sess):
    T.check_features(sess, frames=12)


def test_small_queue(emu_api):
    T.check_small_queue(emu_api)
