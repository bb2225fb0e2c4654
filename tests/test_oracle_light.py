"""Pins the light-propagation oracle (oracle/aic_light.cpp) against the reference's tests in
all-is-cubes/src/space/light/tests.rs (line numbers per test) and the chart statistics."""
import numpy as np
import pytest

import aicb200
import orc
from aicb200 import Block, Space

OPAQUE, NO_RAYS, UNINIT, VISIBLE = 128, 1, 0, 255
WHITE = (1.0, 1.0, 1.0, 1.0)


def some(rgb):
    L = orc.lib()
    return tuple(L.orc_packed_light_scalar_in(float(np.float32(v))) for v in rgb) + (VISIBLE,)


def value(texel):
    L = orc.lib()
    return tuple(np.float32(L.orc_packed_light_lut(int(v))) for v in texel[:3])


def make(size, blocks, sets=(), sky=((0.9, 0.9, 0.9),), max_distance=30, initial=NO_RAYS):
    ids = np.zeros(size, dtype=np.uint16)
    light = np.zeros(size + (4,), dtype=np.uint8)
    light[..., 3] = initial
    sp = Space((0, 0, 0), ids, [Block.air()] + list(blocks), light=light, sky_colors=list(sky), light_max_distance=max_distance)
    ol = orc.OracleLight(sp)
    for cube, bid in sets:
        ol.set_cubes([cube], [bid])
    return sp, ol


def test_chart_statistics():
    """generator.rs:49-113: 602 rays on the 11^3-9^3 shell; SURVEY 8(a) L1: 114 779 tree nodes."""
    w, ch = orc.light_chart()
    assert w.shape[0] == 114779
    assert (ch > 0).sum() == w.shape[0] - 1          # every node but the root has exactly one parent
    # root weights: sum over all rays of max(0, cos) per face; symmetric in sign
    assert np.allclose(w[0, :3], w[0, 3:], rtol=1e-6)
    # a node's weight >= the sum of its children's (rays may end at a node, never appear below it)
    kids = np.where(ch[0] > 0)[0]
    assert w[ch[0][kids]].sum(axis=0).max() <= w[0].max() * 1.0001 * 6


# tests.rs:18-31
def test_initial_values():
    sp, ol = make((1, 1, 1), [Block(color=WHITE)])
    assert tuple(ol.field()[0, 0, 0]) == (0, 0, 0, NO_RAYS)
    ol.set_cubes([(0, 0, 0)], [1])
    assert tuple(ol.field()[0, 0, 0]) == (0, 0, 0, OPAQUE)          # set_cube_opaque_notification, tests.rs:176-200


# tests.rs:77-109
@pytest.mark.parametrize("opacity", [0.0, 0.5, 1.0])
def test_out_of_bounds_light_is_sky(opacity):
    sky = [(2.0,) * 3, (3.0,) * 3, (5.0,) * 3, (7.0,) * 3, (11.0,) * 3, (13.0,) * 3, (17.0,) * 3, (19.0,) * 3]
    sp, ol = make((1, 1, 1), [Block(color=(1.0, 0.0, 0.0, opacity))], sets=[((0, 0, 0), 1)], sky=sky)
    # BlockSky::for_blocks (sky.rs:54-82): mean of the four octants on each face's side
    def face_mean(face):  # face index NX..PZ
        axis, positive = face % 3, face >= 3
        vals = [sky[k][0] for k in range(8) if bool((k >> (2 - axis)) & 1) == positive]
        return float(np.float32(sum(np.float32(v) for v in vals)) * np.float32(0.25))
    for x in range(-2, 3):
        for y in range(-2, 3):
            for z in range(-2, 3):
                if (x, y, z) == (0, 0, 0):
                    continue
                got = ol.get((x, y, z))
                nonzero = [(a, v) for a, v in enumerate((x, y, z)) if v != 0]
                if len(nonzero) == 1 and abs(nonzero[0][1]) == 1:
                    face = nonzero[0][0] + (3 if nonzero[0][1] > 0 else 0)
                    assert got == some((face_mean(face),) * 3), (x, y, z)
                else:
                    assert got == (0, 0, 0, NO_RAYS), (x, y, z)


# tests.rs:111-160 (`step`): one update, the cube next to the new block receives the sky colour
def test_step_single_update():
    color = (1.0, 0.0, 0.0)
    sp, ol = make((3, 1, 1), [Block(color=WHITE)], sky=[color])
    ol.set_cubes([(0, 0, 0)], [1])
    f = ol.field()
    assert tuple(f[0, 0, 0]) == (0, 0, 0, OPAQUE) and tuple(f[1, 0, 0]) == (0, 0, 0, NO_RAYS) and tuple(f[2, 0, 0]) == (0, 0, 0, NO_RAYS)
    n, max_diff = ol.evaluate(0)
    assert n == 1
    f = ol.field()
    assert tuple(f[0, 0, 0]) == (0, 0, 0, OPAQUE)
    assert tuple(f[1, 0, 0]) == some(color)
    assert tuple(f[2, 0, 0]) == (0, 0, 0, NO_RAYS)
    # max_update_difference == sky_light.difference_priority(NO_RAYS) (tests.rs:145)
    s = some(color)
    assert max_diff == min(255, max(s[0], s[1], s[2]) + 63)
    assert ol.queue_len() == 0


# tests.rs:162-174 (`evaluate_light`): 0, then 2 updates after setting the middle block, then 0
def test_evaluate_light_counts():
    sp, ol = make((3, 1, 1), [Block(color=WHITE)])
    assert ol.evaluate(0)[0] == 0
    ol.set_cubes([(1, 0, 0)], [1])
    assert ol.evaluate(0)[0] == 2
    assert ol.evaluate(0)[0] == 0


# tests.rs:219-231
def test_light_source_self_illumination_transparent():
    light = (0.5, 1.0, 2.0)
    sp, ol = make((3, 3, 3), [Block(color=(1.0, 0.0, 0.0, 0.125), emission=light)], sets=[((1, 1, 1), 1)], sky=[(0.0, 0.0, 0.0)])
    ol.evaluate(0)
    assert tuple(ol.field()[1, 1, 1]) == some(light)


# tests.rs:233-261 — exact neighbour values around an opaque emitter
def test_light_source_self_illumination_opaque():
    light = (0.5, 1.0, 2.0)
    sp, ol = make((3, 3, 3), [Block(color=WHITE, emission=light)], sets=[((1, 1, 1), 1)], sky=[(0.0, 0.0, 0.0)])
    ol.evaluate(0)
    f = ol.field()
    assert tuple(f[1, 1, 1]) == some(light)
    f32 = np.float32
    expect = {
        "nx": (f32(0.13397168), f32(0.26794338), f32(0.53588676)),
        "ny": (f32(0.1649385), f32(0.32987696), f32(0.6597539)),
        "nz": (f32(0.21763763), f32(0.43527526), f32(0.8705506)),
    }
    got = {
        "nx": value(f[0, 1, 1]), "px": value(f[2, 1, 1]), "ny": value(f[1, 0, 1]), "py": value(f[1, 2, 1]),
        "nz": value(f[1, 1, 0]), "pz": value(f[1, 1, 2]),
    }
    assert got["nx"] == expect["nx"] and got["px"] == expect["nx"]
    assert got["ny"] == expect["ny"] and got["py"] == expect["ny"]
    assert got["nz"] == expect["nz"] and got["pz"] == expect["nz"]


# tests.rs:263-299 (without the animation-hint case, which is block evaluation)
def test_visible_statuses():
    for block, want in ((None, (NO_RAYS, NO_RAYS)), (Block(color=(1.0, 1.0, 1.0, 0.5)), (VISIBLE, VISIBLE))):
        sp, ol = make((3, 3, 3), [block] if block else [Block(color=WHITE)], sets=[((1, 1, 1), 1)] if block else [])
        ol.evaluate(0)
        f = ol.field()
        assert (f[1, 1, 1, 3], f[0, 1, 1, 3]) == want


# tests.rs:301-319
def test_reflectance_is_clamped():
    over = Block(color=(16.0, 1.0, 0.0, 1.0))
    sky = (0.5, 0.5, 0.5)
    sp, ol = make((5, 3, 3), [over], sets=[((1, 1, 1), 1), ((3, 1, 1), 1)], sky=[sky])
    ol.evaluate(0)
    red = value(ol.field()[2, 1, 1])[0]
    assert red <= np.float32(0.5)


# tests.rs:33-75 (fast_evaluate_light): sky above an obstacle, uninitialized below it
def test_fast_evaluate_light():
    ids = np.zeros((3, 3, 3), dtype=np.uint16)
    ids[1, 1, 1] = 1
    sp = Space((0, 0, 0), ids, [Block.air(), Block(color=(1.0, 0.0, 0.0, 1.0))], light_max_distance=10)
    ol = orc.OracleLight(sp)
    ol.fast_evaluate()
    f = ol.field()
    sky_py = ol.get((1, 3, 1))          # block_sky.in_direction(PY)
    assert tuple(f[1, 2, 1]) == sky_py and sky_py[3] == VISIBLE
    assert tuple(f[1, 0, 1]) == (0, 0, 0, UNINIT)
    assert tuple(f[1, 1, 1]) == (0, 0, 0, OPAQUE)
    assert tuple(f[0, 0, 0]) == (0, 0, 0, NO_RAYS)   # nothing visible nearby
    assert ol.queue_len() > 0
    ol.evaluate(0)
    assert ol.queue_len() == 0
    assert ol.field()[1, 0, 1, 3] == VISIBLE


def test_product_chart_equals_oracle_chart():
    """The product's host chart generator (csrc/light.cu) against the oracle's (generator.rs restatement)."""
    w0, c0 = orc.light_chart()
    w1, c1 = aicb200.light_chart()
    assert w0.shape == w1.shape and np.array_equal(c0, c1)
    assert np.array_equal(w0, w1)


def test_order_dependence_of_the_reference_algorithm():
    """queue.rs:226-246 pops an ARBITRARY element of the highest priority; apply_light_update drops 1-unit
    differences (updater.rs:348-360).  Two legal orders therefore converge to fixed points that differ by a
    few units (statuses never differ).  This documents the tolerance of the GPU-vs-oracle contract (L4)."""
    from test_gpu_light import light_scene
    sp = light_scene()
    fields = []
    for order in (0, 1):
        ol = orc.OracleLight(sp)
        ol.set_pop_order(order)
        ol.fast_evaluate()
        ol.evaluate(0)
        fields.append(ol.field())
    assert np.array_equal(fields[0][..., 3], fields[1][..., 3])
    d = np.abs(fields[0][..., :3].astype(int) - fields[1][..., :3].astype(int)).max(axis=-1)
    assert 1 <= d.max() <= 8
    assert (d <= 2).mean() > 0.97


def test_threaded_update_is_deterministic_and_within_the_contract_of_the_sequential_one():
    """update_light_from_queue with `auto-threads` (updater.rs:211-252: 32 cubes popped, computed in parallel from the
    same stored light, applied in pop order) — the variant the CPU arm of bench.py --workload c4 times.  Its result
    must not depend on the thread count, and it must land where the one-at-a-time variant lands up to the order
    dependence the reference itself has (same statuses, values within a few units)."""
    from aicb200 import scenes
    space = scenes.config_c4(20)
    fields, counts = {}, {}
    for mode in ("sequential", 1, 3, 8):
        ol = orc.OracleLight(space)
        ol.fast_evaluate()
        counts[mode], _ = ol.evaluate(1) if mode == "sequential" else ol.evaluate_threaded(1, mode)
        fields[mode] = ol.field()
        assert ol.queue_len() == 0 or ol.queue_peek() <= 1
    assert np.array_equal(fields[1], fields[3]) and np.array_equal(fields[1], fields[8])
    assert counts[1] == counts[3] == counts[8]
    assert np.array_equal(fields["sequential"][..., 3], fields[1][..., 3])
    d = np.abs(fields["sequential"][..., :3].astype(int) - fields[1][..., :3].astype(int))
    assert d.max() <= 8 and (d > 2).mean() < 0.03
