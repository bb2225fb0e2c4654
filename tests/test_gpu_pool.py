"""GPU tests of the per-CTA ray-pool scheduler (csrc/pool_kernel.cuh, AICB_SCHED=pool): it must produce the very frame
the warp scheduler (trace_kernel) produces — same sRGB8 bytes, same cubes_traced — on every option combination, and
stay within 1 sRGB8 code of the oracle like the warp scheduler does."""
import os

import numpy as np
import pytest

import orc
from aicb200 import (FOG_ABRUPT, FOG_NONE, FOG_PHYSICAL, LIGHT_FLAT, LIGHT_LINEAR, LIGHT_NONE, LIGHT_SMOOTHSTEP,
                     TRANSPARENCY_SURFACE, TRANSPARENCY_THRESHOLD, TRANSPARENCY_VOLUMETRIC, GraphicsOptions, RtRenderer,
                     scenes)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


def draw_with(sched, renderer, **kw):
    old = os.environ.get("AICB_SCHED")
    if sched:
        os.environ["AICB_SCHED"] = sched
    else:
        os.environ.pop("AICB_SCHED", None)
    try:
        return renderer.draw(**kw)
    finally:
        if old is None:
            os.environ.pop("AICB_SCHED", None)
        else:
            os.environ["AICB_SCHED"] = old


def both(space, opts, w, h, direction=(1.0, 0.6, 1.0), shard=None):
    cam = scenes.standard_camera(space, opts, w, h, direction=direction)
    r = RtRenderer(cam)
    r.update(space)
    a = draw_with("warp", r, shard=shard)
    b = draw_with("pool", r, shard=shard)
    return cam, a, b


@pytest.mark.parametrize("transparency", [TRANSPARENCY_SURFACE, TRANSPARENCY_VOLUMETRIC, TRANSPARENCY_THRESHOLD])
@pytest.mark.parametrize("lighting", [LIGHT_NONE, LIGHT_FLAT, LIGHT_LINEAR, LIGHT_SMOOTHSTEP])
@pytest.mark.parametrize("fog", [FOG_NONE, FOG_ABRUPT, FOG_PHYSICAL])
def test_pool_equals_warp_option_matrix(transparency, lighting, fog):
    space = scenes.small_mixed_scene(n=12, seed=7)
    opts = GraphicsOptions(fog=fog, lighting_display=lighting, transparency=transparency, view_distance=40.0,
                           transparency_threshold=0.3)
    cam, a, b = both(space, opts, 96, 64)
    assert np.array_equal(a.data, b.data)
    assert a.info.cubes_traced == b.info.cubes_traced


def test_pool_equals_warp_and_oracle_on_reduced_bench_scenes():
    for space, vd in ((scenes.config_c2(n=48, n_voxel_blocks=8, with_light=True), 192.0),
                      (scenes.config_c1(n=32, n_voxel_blocks=16, with_light=True), 128.0)):
        for opts in (GraphicsOptions(view_distance=vd), GraphicsOptions.unaltered_colors()):
            opts.view_distance = vd
            cam, a, b = both(space, opts, 240, 135)
            assert np.array_equal(a.data, b.data)
            assert a.info.cubes_traced == b.info.cubes_traced
            ref = orc.OracleScene(space).render(cam, opts)
            assert np.abs(b.data.reshape(-1, 4).astype(int) - ref["srgb8"].astype(int)).max() <= 1
            assert b.info.cubes_traced == ref["cubes_traced"]


def test_pool_antialiasing_inside_camera_and_shards():
    space = scenes.small_mixed_scene(n=12, seed=7)
    opts = GraphicsOptions(view_distance=40.0, antialiasing_always=True)
    cam, a, b = both(space, opts, 64, 48)
    assert np.array_equal(a.data, b.data)
    opts = GraphicsOptions(view_distance=40.0)
    cam, a, b = both(space, opts, 64, 48, direction=(0.0, 0.0, 1.0))
    assert np.array_equal(a.data, b.data)
    for index in range(3):
        cam, a, b = both(space, opts, 64, 50, shard=(4, index, 3))
        assert np.array_equal(a.data, b.data)


def test_pool_full_size_frame_is_deterministic_and_equal():
    space = scenes.config_c2(n=96, n_voxel_blocks=16, with_light=True)
    opts = GraphicsOptions(view_distance=384.0)
    cam, a, b = both(space, opts, 1920, 1080)
    assert np.array_equal(a.data, b.data)
    assert a.info.cubes_traced == b.info.cubes_traced
    r = RtRenderer(cam)
    r.update(space)
    again = draw_with("pool", r)
    assert np.array_equal(again.data, b.data)
