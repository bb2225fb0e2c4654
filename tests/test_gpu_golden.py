"""The scenes of tests/test_golden_images.py (restated from the reference's test-renderers suite) through the CUDA path.

The CUDA path is held to the oracle on other scenes (tests/test_gpu_parity.py); these cases add the reference's own
expected images on top, and the same bit-exact comparison against the oracle (f64-rounded-once powf / expf)."""
import os

import numpy as np
import pytest

import aicb200
import orc
from aicb200 import Camera, GraphicsOptions, RtRenderer, Space, SpaceRaytracer, Viewport
from test_golden_images import (build_fog_universe, build_light_spread_universe, build_tone_mapping_universe,
                                check_threshold, golden)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _oracle_rounds_once():
    prev = orc.get_libm()
    orc.set_libm(orc.LIBM_CR)
    yield
    orc.set_libm(prev)


def gpu_and_oracle(space, cam, opts):
    r = RtRenderer(cam)
    r.update(space)
    img = r.draw().data
    ref = orc.OracleScene(space).render(cam, opts)
    h, w = img.shape[:2]
    assert np.array_equal(img.reshape(-1, 4), ref["srgb8"])
    aux = r.draw_colorbuf()
    assert np.array_equal(aux["hit"], ref["hit"]) and np.array_equal(aux["steps"], ref["steps"])
    assert orc.max_ulp_diff(aux["colorbuf"], ref["colorbuf"]) == 0
    return img


def test_debug_pixel_cost_image_on_gpu():
    """The reference's debug_pixel_cost-ray.png encodes per-pixel cubes_traced; the oracle reproduces it exactly."""
    space = build_fog_universe()
    opts = GraphicsOptions.unaltered_colors()
    opts.debug_pixel_cost = True
    cam = Camera(opts, Viewport((128.0, 96.0), (128, 96)))
    cam.look_at_y_up((0.0, 10.0, 0.0), (0.4, 10.0, -1.0))
    img = gpu_and_oracle(space, cam, opts)
    exp = golden("debug_pixel_cost-ray")
    assert np.array_equal(img[..., :2], exp[..., :2])          # red / green: the step counts
    assert np.abs(img.astype(int) - exp.astype(int)).max() <= 1  # blue: a fifth of the luminance


@pytest.mark.parametrize("name,fog", [("Abrupt", aicb200.FOG_ABRUPT), ("Compromise", aicb200.FOG_COMPROMISE),
                                      ("Physical", aicb200.FOG_PHYSICAL)])
def test_fog_on_gpu(name, fog):
    space = build_fog_universe()
    opts = GraphicsOptions.unaltered_colors()
    opts.lighting_display = aicb200.LIGHT_LINEAR
    opts.view_distance = 50.0
    opts.fog = fog
    cam = Camera(opts, Viewport((128.0, 96.0), (128, 96)))
    cam.look_at_y_up((0.0, 10.0, 0.0), (0.4, 10.0, -1.0))
    img = gpu_and_oracle(space, cam, opts)
    check_threshold(img, golden(f"fog-{name}-all"), [(2, 2200), (15, 100)])


@pytest.mark.parametrize("name,lighting", [("None", aicb200.LIGHT_NONE), ("Flat", aicb200.LIGHT_FLAT),
                                           ("Coarse", aicb200.LIGHT_COARSE), ("Linear", aicb200.LIGHT_LINEAR),
                                           ("Smoothstep", aicb200.LIGHT_SMOOTHSTEP)])
def test_light_spread_on_gpu(name, lighting):
    space = build_light_spread_universe()
    opts = GraphicsOptions.unaltered_colors()
    opts.lighting_display = lighting
    opts.fov_y = 45.0
    cam = Camera(opts, Viewport((128.0, 96.0), (128, 96)))
    cam.set_view_transform((0.0, 0.0, 0.0, 1.0), (0.0, 0.0, 8.0))
    img = gpu_and_oracle(space, cam, opts)
    check_threshold(img, golden(f"light_spread-{name}-all"), [(8, 128 * 96)])   # the reference's 7 + the 1 code of GPU vs oracle


@pytest.mark.parametrize("tmo,max_intensity,exposure", [("Clamp", 1.0, 0.5), ("Reinhard", 0.5, 0.5), ("Reinhard", 1.0, 2.0)])
def test_tone_map_on_gpu(tmo, max_intensity, exposure):
    space, eye = build_tone_mapping_universe()
    opts = GraphicsOptions.unaltered_colors()
    opts.lighting_display = aicb200.LIGHT_FLAT
    opts.fov_y = 45.0
    opts.maximum_intensity = max_intensity
    opts.exposure = exposure
    opts.tone_mapping = aicb200.TONE_CLAMP if tmo == "Clamp" else aicb200.TONE_REINHARD
    cam = Camera(opts, Viewport((256.0, 320.0), (256, 320)))
    cam.set_view_transform((0.0, 0.0, 0.0, 1.0), eye)
    img = gpu_and_oracle(space, cam, opts)
    check_threshold(img, golden(f"tone_map-{tmo}-{max_intensity}-{exposure}-all"), [(2, 256 * 320), (4, 700), (11, 500)])


def test_white_furnace_with_gpu_light():
    """Space::set x3 + evaluate_light(0) on the GPU, then the frame: white blocks under a uniform sky stay invisible."""
    from aicb200 import Block
    white = Block(color=(1.0, 1.0, 1.0, 1.0))
    ids = np.zeros((3, 3, 3), dtype=np.uint16)
    light = np.zeros((3, 3, 3, 4), dtype=np.uint8)
    light[..., 3] = 1
    sky = [(0.75, 0.75, 0.75)]
    space = Space((-1, -1, -1), ids, [Block.air(), white], light=light, sky_colors=sky, light_max_distance=30)
    opts = GraphicsOptions(fov_y=45.0, view_distance=10.0, fog=aicb200.FOG_NONE)
    cam = Camera(opts, Viewport((128.0, 96.0), (128, 96)))
    cam.look_at_y_up((-3.0, 4.0, 4.0), (-2.0, 3.0, 3.0))
    r = RtRenderer(cam)
    r.update(space)
    r.rt.light_edit_and_propagate([(-1, -1, 1), (1, -1, 0), (-1, 1, -1)], [1, 1, 1], 0)
    img = r.draw().data
    check_threshold(img, golden("furnace-Clear-Opaque-all"), [(2, 128 * 96)])
    assert int(img[..., :3].min()) >= 223 and int(img[..., :3].max()) <= 228


def test_one_cube_cases_on_gpu():
    """viewport_prime (a 101 x 37 viewport) and no_update (NO_WORLD_TO_SHOW before the first update, the cube after it):
    cases/src/lib.rs:1215-1229, 988-1005 — through the CUDA path, bit-identical to the oracle and within the
    reference's thresholds of its expected images."""
    from test_golden_images import common_camera, no_world_to_show, one_cube_space
    opts = GraphicsOptions.unaltered_colors()
    cam = Camera(opts, Viewport((101.0, 37.0), (101, 37)))
    cam.set_view_transform((0.0, 0.0, 0.0, 1.0), (0.5, 0.5, 2.0))
    img = gpu_and_oracle(one_cube_space(), cam, opts)
    check_threshold(img, golden("viewport_prime-all"), [(2, 101 * 37)])
    cam = common_camera(opts)
    img = gpu_and_oracle(one_cube_space(), cam, opts)
    check_threshold(img, golden("no_update-2-all"), [(5, 128 * 96)])
    nothing = Space((0, 0, 0), np.zeros((1, 1, 1), dtype=np.uint16), [aicb200.Block.air()], sky_colors=[(0.5, 0.5, 0.5)])
    first = aicb200.render_layers(None, (SpaceRaytracer(nothing, opts), cam, opts), no_world=no_world_to_show())
    assert np.array_equal(first.data.reshape(96, 128, 4), golden("no_update-all"))


@pytest.mark.parametrize("name", ["layers_all-all", "layers_hidden_ui-all", "layers_ui_only-all", "layers_none_but_text-all"])
def test_layers_on_gpu(name):
    """RtScene::trace_ray_through_layers (renderer.rs:454-478) through aicb_render_layers_srgb8 against the reference's
    layers_* expectations: byte-identical outside the box where the reference draws its info text."""
    from test_golden_images import check_layers_image, layer_cameras, layer_cases, no_world_to_show, one_cube_space, ui_space
    has_world, has_ui, opts = layer_cases()[name]
    wcam, ucam = layer_cameras(opts)
    world = (SpaceRaytracer(one_cube_space(), opts), wcam, opts) if has_world else None
    ui = (SpaceRaytracer(ui_space(), opts), ucam, opts) if has_ui else None
    if not world and not ui:
        ui = (SpaceRaytracer(Space((0, 0, 0), np.zeros((1, 1, 1), dtype=np.uint16), [aicb200.Block.air()]), opts), ucam, opts)
    img = aicb200.render_layers(world, ui, no_world=no_world_to_show()).data.reshape(96, 128, 4)
    check_layers_image(img, name)
    oworld = (orc.OracleScene(one_cube_space()), wcam, opts) if has_world else None
    oui = (orc.OracleScene(ui_space()), ucam, opts) if has_ui else None
    if oworld or oui:
        assert np.array_equal(img.reshape(-1, 4), orc.render_layers(oworld, oui, no_world=no_world_to_show())["srgb8"])


@pytest.mark.parametrize("name,lighting", [("None", aicb200.LIGHT_NONE), ("Flat", aicb200.LIGHT_FLAT),
                                           ("Coarse", aicb200.LIGHT_COARSE), ("Linear", aicb200.LIGHT_LINEAR),
                                           ("Smoothstep", aicb200.LIGHT_SMOOTHSTEP)])
def test_light_on_slab_on_gpu(name, lighting):
    """Light on surfaces inside their cubes (partial voxel bounds): the reference's light_on_slab expectations through
    the CUDA path, bit-identical to the oracle."""
    from test_golden_images import build_light_on_slab_universe, light_on_slab_camera
    space = _slab_universe()
    opts = GraphicsOptions.unaltered_colors()
    opts.lighting_display = lighting
    opts.fov_y = 45.0
    img = gpu_and_oracle(space, light_on_slab_camera(opts), opts)
    check_threshold(img, golden(f"light_on_slab-{name}-all"), [(8, 128 * 96)])


_SLAB = []


def _slab_universe():
    if not _SLAB:
        from test_golden_images import build_light_on_slab_universe
        _SLAB.append(build_light_on_slab_universe())
    return _SLAB[0]
