"""Pins the host Camera code (all-is-cubes_b200/host/camera.cpp, restating euclid 0.22.14) and the
oracle's whole-image path against the reference's tests:
camera/tests.rs:78-109 (exact frustum corners), :198-222 (project_ndc_into_world),
raytracer/text.rs:196-258 and :265-341 (80x40 ASCII hit-identity images)."""
import json
import math
import os

import numpy as np

import aicb200
import orc
from aicb200 import Block, Camera, GraphicsOptions, Space, Viewport

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unproject(cam, x, y, z):
    m = cam.inverse_projection_view
    h = np.array([x, y, z, 1.0]) @ m  # row-vector convention
    # evaluate exactly as euclid does (left-to-right sums) to compare bit patterns
    hx = x * m[0, 0] + y * m[1, 0] + z * m[2, 0] + m[3, 0]
    hy = x * m[0, 1] + y * m[1, 1] + z * m[2, 1] + m[3, 1]
    hz = x * m[0, 2] + y * m[1, 2] + z * m[2, 2] + m[3, 2]
    hw = x * m[0, 3] + y * m[1, 3] + z * m[2, 3] + m[3, 3]
    assert np.allclose(h, [hx, hy, hz, hw])
    return (hx / hw, hy / hw, hz / hw)


# camera/tests.rs:78-109 — exact f64 equality in the reference
def test_view_frustum_exact():
    opts = GraphicsOptions(view_distance=100.0, fov_y=90.0)
    cam = Camera(opts, Viewport.with_scale(1.0, (10, 5)))
    x_near, y_near, z_near = 0.062499999999999986, 0.031249999999999993, -0.03125
    x_far, y_far, z_far = 200.00000000003973, 100.00000000001987, -100.0000000000199
    assert unproject(cam, -1, -1, 0) == (-x_near, -y_near, z_near)
    assert unproject(cam, -1, 1, 0) == (-x_near, y_near, z_near)
    assert unproject(cam, 1, -1, 0) == (x_near, -y_near, z_near)
    assert unproject(cam, 1, 1, 0) == (x_near, y_near, z_near)
    assert unproject(cam, -1, -1, 1) == (-x_far, -y_far, z_far)
    assert unproject(cam, -1, 1, 1) == (-x_far, y_far, z_far)
    assert unproject(cam, 1, -1, 1) == (x_far, -y_far, z_far)
    assert unproject(cam, 1, 1, 1) == (x_far, y_far, z_far)


# camera/tests.rs:198-222
def test_project_ndc_into_world():
    cam = Camera(GraphicsOptions(), Viewport((2.0, 2.0), (2, 2)))
    near = 1.0 / 32.0
    ray = cam.project_ndc_into_world(0.0, 0.0)
    assert tuple(ray[:3]) == (0.0, 0.0, -near)
    assert np.allclose(ray[3:], [0.0, 0.0, -(200.0 - near)], atol=1e-6)
    # Rotation3D::around_y(frac_pi_2), translation (0,100,0)
    h = math.pi / 2 / 2
    cam.set_view_transform((0.0, math.sin(h), 0.0, math.cos(h)), (0.0, 100.0, 0.0))
    ray = cam.project_ndc_into_world(0.0, 0.0)
    assert np.allclose(ray[:3], [-near, 100.0, 0.0], atol=1e-6)
    assert np.allclose(ray[3:], [-(200.0 - near), 0.0, 0.0], atol=1e-6)


# camera/tests.rs:224-234
def test_project_ndc_edge_cases():
    cam = Camera(GraphicsOptions(), Viewport((2.0, 2.0), (2, 2)))
    for bad in (math.nan, math.inf):
        ray = cam.project_ndc_into_world(bad, 0.0)
        assert math.isnan(ray[0]) and math.isnan(ray[3])


def test_oracle_pixel_ray_matches_host_camera():
    """renderer.rs:424-451 patch centre == project_ndc_into_world at the same NDC point."""
    cam = Camera(GraphicsOptions(), Viewport.with_scale(1.0, (16, 12)))
    cam.look_at_y_up((5.0, 7.0, 9.0), (0.5, 0.5, 0.5))
    for (x, y) in [(0, 0), (15, 11), (7, 3)]:
        x0, x1 = x / 16 * 2.0 - 1.0, (x + 1) / 16 * 2.0 - 1.0
        y0, y1 = -(y / 12 * 2.0 - 1.0), -((y + 1) / 12 * 2.0 - 1.0)
        want = cam.project_ndc_into_world((x0 + x1) / 2.0, (y0 + y1) / 2.0)
        got = orc.pixel_ray(cam, x, y)
        assert np.array_equal(want, got)


# ---- text.rs ASCII images ------------------------------------------------------------------------
def color_for_make_blocks(i, n):
    """content.rs color_for_make_blocks: a grey ramp (make_some_blocks_1/2 tests, content.rs:295-331)."""
    if n <= 1:
        return (0.5, 0.5, 0.5, 1.0)
    v = i / (n - 1)
    return (v, v, v, 1.0)


def print_space(space, direction, chars):
    """PrintSpace::fmt (text.rs:158-180): Camera::new(default, Viewport{nominal 40x40, fb 80x40}),
    look_at_y_up(eye_for_look_at(bounds, direction), bounds.center()), to_text::<CharacterBuf>."""
    opts = GraphicsOptions()
    cam = Camera(opts, Viewport((40.0, 40.0), (80, 40)))
    center = [space.lower[a] + space.size[a] / 2.0 for a in range(3)]
    cam.look_at_y_up(aicb200.eye_for_look_at(space.lower, space.size, direction), center)
    rays = []
    for ych in range(40):
        y = -((ych + 0.5) / 40.0 * 2.0 - 1.0)  # normalize_fb_y (viewport.rs:96-100)
        for xch in range(80):
            x = (xch + 0.5) / 80.0 * 2.0 - 1.0  # normalize_fb_x (viewport.rs:89-92)
            rays.append(cam.project_ndc_into_world(x, y))
    text = orc.OracleScene(space).trace_rays(np.array(rays), opts, include_sky=True, accum_mode=1)["text"]
    rows = []
    for ych in range(40):
        row = ""
        for xch in range(80):
            t = int(text[ych * 80 + xch])
            row += "." if t == -2 else " " if t in (-1, -4) else "X" if t == -3 else chars[t]
        rows.append(row)
    return rows


def test_print_space_golden():
    golden = json.load(open(os.path.join(GOLDEN, "text_images.json")))["print_space_test"]
    ids = np.array([1, 2, 3], dtype=np.uint16).reshape(3, 1, 1)
    blocks = [Block.air()] + [Block(color=color_for_make_blocks(i, 3)) for i in range(3)]
    space = Space((0, 0, 0), ids, blocks)
    rows = print_space(space, (1.0, 1.0, 1.0), {1: "0", 2: "1", 3: "2"})
    diffs = sum(a != b for ra, rb in zip(rows, golden) for a, b in zip(ra, rb))
    assert diffs == 0, "\n".join(rows)


def test_partial_voxels_golden():
    golden = json.load(open(os.path.join(GOLDEN, "text_images.json")))["partial_voxels"]
    # R4 block whose voxel data is a 4x2x4 white slab (partial voxel_bounds)
    idx = np.zeros((4, 2, 4), dtype=np.uint16)
    pal = np.zeros((1, 8), dtype=np.float32)
    pal[0, :4] = (1.0, 1.0, 1.0, 1.0)
    partial = Block(resolution=4, voxel_lower=(0, 0, 0), indices=idx, palette=pal)
    ids = np.array([1, 2], dtype=np.uint16).reshape(2, 1, 1)
    space = Space((0, 0, 0), ids, [Block.air(), Block(color=color_for_make_blocks(0, 1)), partial])
    rows = print_space(space, (1.0, 1.0, 1.0), {1: "0", 2: "P"})
    diffs = sum(a != b for ra, rb in zip(rows, golden) for a, b in zip(ra, rb))
    assert diffs == 0, "\n".join(rows)
