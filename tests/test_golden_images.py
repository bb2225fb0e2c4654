"""Image-level pins: the oracle rendering restated test-renderers scenes must reproduce the
reference's expected PNGs (test-renderers/expected/renderers/*.png, committed as
tests/golden/png_*.npy by make_goldens.py) within the thresholds the reference itself uses
(test-renderers/cases/src/lib.rs: transparent_one :1138-1164 threshold COLOR_ROUNDING_MAX_DIFF=2
:1237; emission :297-348 threshold 1; emission_only/semi :351-418 histogram thresholds)."""
import os

import numpy as np
import pytest

import aicb200
import orc
from aicb200 import Block, Camera, GraphicsOptions, Space, Viewport, srgb8_to_linear

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, f"png_{name}.npy"))


def common_camera(options):
    """StandardCameras::from_constant_for_test with COMMON_VIEWPORT 128x96 (types/src/render.rs:135) and
    looking_at_one_cube_spawn: eye (0.5,0.5,2), look (0,0,-1) (cases/src/lib.rs:1250-1257)."""
    cam = Camera(options, Viewport((128.0, 96.0), (128, 96)))
    cam.set_view_transform((0.0, 0.0, 0.0, 1.0), (0.5, 0.5, 2.0))
    return cam


def diff_histogram(img, expected):
    """rendiff-style comparison: per pixel, the smallest max-channel difference against the 3x3
    neighbourhood of the expected image (rendiff tolerates one-pixel position error)."""
    h, w, _ = expected.shape
    a = img.astype(np.int32)
    pad = np.pad(expected.astype(np.int32), ((1, 1), (1, 1), (0, 0)), mode="edge")
    best = np.full((h, w), 255, dtype=np.int32)
    for dy in range(3):
        for dx in range(3):
            d = np.abs(a - pad[dy:dy + h, dx:dx + w]).max(axis=2)
            best = np.minimum(best, d)
    return best


def check_threshold(img, expected, levels):
    d = diff_histogram(img, expected)
    exact = np.abs(img.astype(np.int32) - expected.astype(np.int32)).max(axis=2)
    prev = 0
    for limit, count in levels:
        n = int(((d > prev) & (d <= limit)).sum())
        assert n <= count, f"{n} pixels differ by ({prev},{limit}] (allowed {count}); exact max diff {exact.max()}"
        prev = limit
    assert int((d > prev).sum()) == 0, f"pixels differ by more than {prev}: max {d.max()}"
    return int(exact.max()), int((exact > 0).sum())


@pytest.mark.parametrize("mode,name", [(aicb200.TRANSPARENCY_SURFACE, "transparent_one-surf-all"),
                                       (aicb200.TRANSPARENCY_VOLUMETRIC, "transparent_one-vol-all")])
def test_transparent_one(mode, name):
    space = Space((0, 0, 0), np.ones((1, 1, 1), dtype=np.uint16), [Block.air(), Block(color=(1.0, 0.0, 0.0, 0.5))],
                  sky_colors=[(0.5, 0.5, 0.5)])
    opts = GraphicsOptions.unaltered_colors()
    opts.transparency = mode
    img = orc.OracleScene(space).render(common_camera(opts), opts)["srgb8"].reshape(96, 128, 4)
    exp = golden(name)
    assert tuple(exp[48, 64]) == ((225, 137, 137, 255) if mode == aicb200.TRANSPARENCY_SURFACE else (225, 136, 136, 255))
    if mode == aicb200.TRANSPARENCY_SURFACE:
        assert tuple(img[48, 64]) == tuple(exp[48, 64])  # #E18989 exactly (cases/src/lib.rs:1140-1142)
    # The shared "-all" expectation was not produced by the CPU raytracer alone; the reference accepts
    # any renderer within COLOR_ROUNDING_MAX_DIFF = 2 (cases/src/lib.rs:1237), and so do we.
    check_threshold(img, exp, [(2, 128 * 96)])


def test_emission():
    c200 = srgb8_to_linear((200, 0, 0))[0]
    pal = np.zeros((3, 8), dtype=np.float32)
    pal[0, :4] = (1, 1, 1, 1)                       # white
    pal[1, :4] = (c200, 0, 0, 1); pal[1, 4:7] = (0, c200, 0)   # E: colour + emission
    pal[2, :4] = (0, 0, 0, 1); pal[2, 4:7] = (0, c200, 0)      # e: emission only (black)
    idx = np.zeros((4, 4, 4), dtype=np.uint16)
    idx[1, 2, :] = 1   # 'E' at text row 1, col 1 -> y = 3 - 1 = 2, x = 1
    idx[2, 1, :] = 2   # 'e' at text row 2, col 2 -> y = 1, x = 2
    space = Space((0, 0, 0), np.ones((1, 1, 1), dtype=np.uint16),
                  [Block.air(), Block(resolution=4, indices=idx, palette=pal)], sky_colors=[(0.5, 0.5, 0.5)])
    opts = GraphicsOptions.unaltered_colors()
    img = orc.OracleScene(space).render(common_camera(opts), opts)["srgb8"].reshape(96, 128, 4)
    check_threshold(img, golden("emission-all"), [(1, 128 * 96)])


@pytest.mark.parametrize("kind", ["only", "semi"])
@pytest.mark.parametrize("mode,tag", [(aicb200.TRANSPARENCY_SURFACE, "surf"), (aicb200.TRANSPARENCY_VOLUMETRIC, "vol")])
def test_emission_voxel_shapes(kind, mode, tag):
    g = srgb8_to_linear((0, 200, 0))[1]
    if kind == "only":
        atom = dict(color=(0.0, 0.0, 0.0, 0.0), emission=(0.0, g, 0.0))
    else:
        atom = dict(color=(0.0, 0.0, 0.0, 1.0 - 2.0 ** -3), emission=(0.0, g, 0.0))
    pal = np.zeros((2, 8), dtype=np.float32)
    pal[1, :4] = atom["color"]
    pal[1, 4:7] = atom["emission"]
    idx = np.zeros((2, 2, 2), dtype=np.uint16)
    for x in range(2):
        for y in range(2):
            for z in range(2):
                if x == 0 or y == 0 or z == 0:
                    idx[x, y, z] = 1
    ids = np.zeros((4, 1, 1), dtype=np.uint16)
    ids[0, 0, 0] = 1   # cube (-1,0,0): the atom
    ids[2, 0, 0] = 2   # cube (1,0,0): the voxel block
    space = Space((-1, 0, 0), ids, [Block.air(), Block(**atom), Block(resolution=2, indices=idx, palette=pal)],
                  sky_colors=[srgb8_to_linear((0, 0, 127))])
    opts = GraphicsOptions.unaltered_colors()
    opts.transparency = mode
    img = orc.OracleScene(space).render(common_camera(opts), opts)["srgb8"].reshape(96, 128, 4)
    check_threshold(img, golden(f"emission_{kind}-{tag}-all"), [(2, 1000), (5, 200), (15, 80)])
