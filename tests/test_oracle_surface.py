"""Pins the oracle's SurfaceIter / DepthIter / accumulators against the reference's known-answer
tests: all-is-cubes-render/src/raytracer/surface.rs:541-835, accum.rs:386-496,
all-is-cubes/src/raytracer_components.rs:269-331."""
import math

import numpy as np

import aicb200
import orc
from aicb200 import Block, Space

NAN = float("nan")
ES, INV, EB = 0, 1, 2  # record kinds


def rec_tuple(r):
    kind = int(r[0])
    if kind == 0:
        return ("surface", r[1], None if math.isnan(r[2]) else r[2], tuple(r[3:6]), tuple(int(v) for v in r[6:9]),
                tuple(int(v) for v in r[9:12]), int(r[12]), orc.FACES[int(r[13])], tuple(np.float32(r[14:18])))
    if kind == 1:
        return ("invisible", None if math.isnan(r[1]) else r[1])
    return ("enter_block", r[1])


def slab_block(resolution, color, height=1, full_bounds=False):
    """content::make_slab / voxels_fn slab: voxels with y < height are `color`; bounds shrink to them."""
    r = resolution
    if full_bounds:
        idx = np.zeros((r, r, r), dtype=np.uint16)
        idx[:, :height, :] = 1
        lower = (0, 0, 0)
    else:
        idx = np.ones((r, height, r), dtype=np.uint16)
        lower = (0, 0, 0)
    pal = np.zeros((2, 8), dtype=np.float32)
    pal[1, :4] = color
    return Block(resolution=r, voxel_lower=lower, indices=idx, palette=pal)


RED = (1.0, 0.0, 0.0, 1.0)
YELLOW = (1.0, 1.0, 0.0, 1.0)


# surface.rs:541-679
def test_surface_and_depth_iter_basic():
    # slab_with_extra_space: R4, AIR iff y >= 2 && x != 0 (full 4^3 bounds because of the x == 0 column)
    idx = np.ones((4, 4, 4), dtype=np.uint16)
    idx[1:, 2:, :] = 0
    pal = np.zeros((2, 8), dtype=np.float32)
    pal[1, :4] = YELLOW
    slab = Block(resolution=4, voxel_lower=(0, 0, 0), indices=idx, palette=pal)
    ids = np.array([0, 1, 2], dtype=np.uint16).reshape(1, 3, 1)
    space = Space((0, 0, 0), ids, [Block.air(), Block(color=RED), slab])
    o = orc.OracleScene(space)
    ray = [0.25, -0.5, 0.25, 0.0, 1.0, 0.0]

    steps = [rec_tuple(r) for r in o.surface_steps(ray)]
    f = np.float32
    assert steps == [
        ("invisible", 0.5),
        ("surface", 1.5, None, (0.25, 1.0, 0.25), (0, 1, 0), (0, 0, 0), 1, "NY", (f(1), f(0), f(0), f(1))),
        ("enter_block", 2.5),
        ("surface", 2.5, None, (0.25, 2.0, 0.25), (0, 2, 0), (1, 0, 1), 4, "NY", (f(1), f(1), f(0), f(1))),
        ("surface", 2.75, None, (0.25, 2.25, 0.25), (0, 2, 0), (1, 1, 1), 4, "NY", (f(1), f(1), f(0), f(1))),
        ("invisible", 3.0),
        ("invisible", 3.25),
        ("invisible", 3.5),
        ("invisible", 3.5),
    ]

    dsteps = [rec_tuple(r) for r in o.surface_steps(ray, depth_iter=True)]
    assert dsteps == [
        ("invisible", None),
        ("invisible", None),
        ("surface", 1.5, 2.5, (0.25, 1.0, 0.25), (0, 1, 0), (0, 0, 0), 1, "NY", (f(1), f(0), f(0), f(1))),
        ("enter_block", 2.5),
        ("invisible", None),
        ("surface", 2.5, 2.75, (0.25, 2.0, 0.25), (0, 2, 0), (1, 0, 1), 4, "NY", (f(1), f(1), f(0), f(1))),
        ("surface", 2.75, 3.0, (0.25, 2.25, 0.25), (0, 2, 0), (1, 1, 1), 4, "NY", (f(1), f(1), f(0), f(1))),
        ("invisible", None),
        ("invisible", None),
        ("invisible", None),
    ]


# surface.rs:682-709
def test_surface_iter_exit_block_at_end_of_space():
    space = Space((0, 0, 0), np.zeros((1, 1, 1), dtype=np.uint16), [Block(color=RED)])
    steps = [rec_tuple(r) for r in orc.OracleScene(space).surface_steps([-0.5, 0.5, 0.5, 1.0, 0.0, 0.0])]
    f = np.float32
    assert steps == [
        ("surface", 0.5, None, (0.0, 0.5, 0.5), (0, 0, 0), (0, 0, 0), 1, "NX", (f(1), f(0), f(0), f(1))),
        ("invisible", 1.5),
    ]


# surface.rs:713-749
def test_ray_misses_voxels():
    space = Space((0, 0, 0), np.zeros((1, 1, 1), dtype=np.uint16), [slab_block(2, (0.5, 0.5, 0.5, 1.0))])
    o = orc.OracleScene(space)
    ray = [-0.5, 0.75, 0.25, 1.0, 0.0, 0.0]
    assert [rec_tuple(r) for r in o.surface_steps(ray)] == [("enter_block", 0.5), ("invisible", 1.5)]
    assert [rec_tuple(r) for r in o.surface_steps(ray, depth_iter=True)] == [
        ("invisible", None), ("enter_block", 0.5), ("invisible", None)]


# surface.rs:755-835
def test_depth_iter_exiting_block_volume_before_cube():
    color = (1.0, 1.0, 0.0, 0.5)
    ids = np.array([0, 1, 0], dtype=np.uint16).reshape(1, 3, 1)
    space = Space((0, 0, 0), ids, [Block.air(), slab_block(2, color)])
    o = orc.OracleScene(space)
    ray = [0.25, -0.5, 0.25, 0.0, 1.0, 0.0]
    f = np.float32
    surf = lambda exit_t: ("surface", 1.5, exit_t, (0.25, 1.0, 0.25), (0, 1, 0), (0, 0, 0), 2, "NY",
                           (f(1), f(1), f(0), f(0.5)))
    assert [rec_tuple(r) for r in o.surface_steps(ray)] == [
        ("invisible", 0.5), ("enter_block", 1.5), surf(None), ("invisible", 2.0), ("invisible", 2.5), ("invisible", 3.5)]
    assert [rec_tuple(r) for r in o.surface_steps(ray, depth_iter=True)] == [
        ("invisible", None), ("invisible", None), ("enter_block", 1.5), ("invisible", None), surf(2.0),
        ("invisible", None), ("invisible", None)]


# accum.rs:442-496
def test_depth_buf():
    space = Space((0, 0, 0), np.zeros((1, 1, 1), dtype=np.uint16), [slab_block(2, (0.5, 0.5, 0.5, 1.0))])
    o = orc.OracleScene(space)
    opts = aicb200.GraphicsOptions()  # GraphicsOptions::default()
    d = 0.25
    rays = [
        ([0.25, 0.25, 0.0, 0.0, 0.0, 1.0], 0.0),
        ([0.25, 0.25, -d, 0.0, 0.0, 1.0], d),
        ([0.25, 0.25, -d, 0.0, 0.0, 4.0], d / 4.0),
        ([0.5, 5.25, 0.5, 0.0, -1.0, 0.0], 5.25 - 0.5),
        ([0.5, 0.75, -0.5, 0.0, 0.0, 1.0], math.inf),
    ]
    for ray, expected in rays:
        got = o.trace_rays([ray], opts, include_sky=False, accum_mode=2)["depth"][0]
        assert got == expected, (ray, got, expected)
        # the passive depth observer of the ColorBuf accumulator agrees (first hit is the same)
        got0 = o.trace_rays([ray], opts, include_sky=False, accum_mode=0)["depth"][0]
        assert got0 == expected


# raytracer_components.rs:269-331
def test_apply_transmittance():
    color = (1.0, 0.5, 0.0, 0.5)
    assert orc.apply_transmittance(color, 1.0) == (color, 1.0)
    assert orc.apply_transmittance(color, -0.125) == ((0.0, 0.0, 0.0, 0.0), 0.0)
    assert orc.apply_transmittance(color, 0.0) == ((0.0, 0.0, 0.0, 0.0), 0.0)
    opaque = (1.0, 0.5, 0.0, 1.0)
    assert orc.apply_transmittance(opaque, -0.125) == (opaque, 1.0)
    assert orc.apply_transmittance(opaque, 0.0) == (opaque, 1.0)
    # apply_transmittance_equivalence: count layers of thickness 1/count == one unit layer
    for count in (1, 2, 8):
        mod, _ = orc.apply_transmittance(color, float(np.float32(1.0) / np.float32(count)))
        light = np.zeros(3, dtype=np.float32)
        t = np.float32(1.0)
        a = np.float32(mod[3])
        for _ in range(count):
            light = light + np.array(mod[:3], dtype=np.float32) * a * t
            t = t * (np.float32(1.0) - a)
        alpha = np.float32(1.0) - t
        actual = list(light / alpha) + [alpha]
        assert sum(float(x) - y for x, y in zip(actual, color)) < 0.00001


# accum.rs:386-440 (ColorBuf opacity transitions) through single-voxel blocks
def test_color_buf_accumulation():
    ids = np.array([1, 2, 3], dtype=np.uint16).reshape(3, 1, 1)
    blocks = [Block.air(), Block(color=(1, 0, 0, 0.75)), Block(color=(0, 1, 0, 0.5)), Block(color=(0, 0, 1, 1.0))]
    space = Space((0, 0, 0), ids, blocks)
    o = orc.OracleScene(space)
    opts = aicb200.GraphicsOptions.unaltered_colors()
    opts.transparency = aicb200.TRANSPARENCY_SURFACE
    r = o.trace_rays([[-0.5, 0.5, 0.5, 1.0, 0.0, 0.0]], opts, include_sky=False)
    cb = r["colorbuf"][0]
    # premultiplied accumulation: 0.75 red, then 0.25*0.5 green, then 0.125 blue; transmittance 0
    assert list(cb) == [0.75, 0.125, 0.125, 0.0]
    assert r["steps"][0] == 4  # three surfaces + the count that sees opacity (sr.rs:639-655)
    # the first surface alone
    space1 = Space((0, 0, 0), ids[:1], blocks)
    cb1 = orc.OracleScene(space1).trace_rays([[-0.5, 0.5, 0.5, 1.0, 0.0, 0.0]], opts, include_sky=False)["colorbuf"][0]
    assert list(cb1) == [0.75, 0.0, 0.0, 0.25]


# cases/src/lib.rs:1140-1142: one transparent block (alpha 0.5, rgb (1,0,0)) in front of a sky of 0.5
# is a 50 % mix = [0.75,0.25,0.25] = #E18989.  On an axis-aligned ray the Volumetric thickness is
# exactly 1, so apply_transmittance is the identity and both modes agree.  (The full 128x96 images are
# checked in test_golden_images.py.)
def test_transparent_one_colors():
    for transparency, expect in ((aicb200.TRANSPARENCY_SURFACE, (225, 137, 137, 255)),
                                 (aicb200.TRANSPARENCY_VOLUMETRIC, (225, 137, 137, 255))):
        ids = np.ones((1, 1, 1), dtype=np.uint16)
        space = Space((0, 0, 0), ids, [Block.air(), Block(color=(1.0, 0.0, 0.0, 0.5))], sky_colors=[(0.5, 0.5, 0.5)])
        opts = aicb200.GraphicsOptions.unaltered_colors()
        opts.transparency = transparency
        cb = orc.OracleScene(space).trace_rays([[0.5, 0.5, 2.0, 0.0, 0.0, -1.0]], opts)["colorbuf"][0]
        assert orc.to_srgb8(cb) == expect, (transparency, cb, orc.to_srgb8(cb))
        assert orc.to_srgb8([0.5, 0.5, 0.5, 0.0]) == (188, 188, 188, 255)
