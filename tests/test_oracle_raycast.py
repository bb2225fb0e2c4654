"""Pins the oracle's Amanatides–Woo DDA against the reference's own known-answer tests:
all-is-cubes-base/src/raycast/tests.rs (line numbers cited per test) and the doc-tests in
raycast.rs:185-192, 340-345, 362-368, 399-406."""
import math

import numpy as np
import pytest

import orc

I32_MAX = 2**31 - 1
I32_MIN = -(2**31)


def prefix(origin, direction, expected, **kw):
    got = orc.raycast(origin, direction, max_steps=len(expected) + 2, **kw)
    assert len(got) >= len(expected), got
    for g, e in zip(got, expected):
        assert g[0] == tuple(e[0:3]) and g[1] == e[3] and g[2] == e[4], (g, e)


def exactly(origin, direction, expected, **kw):
    got = orc.raycast(origin, direction, max_steps=len(expected) + 4, **kw)
    assert [(g[0], g[1], g[2]) for g in got] == [(tuple(e[0:3]), e[3], e[4]) for e in expected], got


# raycast/tests.rs:95-146
@pytest.mark.parametrize("d,cubes,face", [
    ((0.01, 0.0001, 0.0001), [(11, 20, 30), (12, 20, 30)], "NX"),
    ((-0.01, 0.0001, 0.0001), [(9, 20, 30), (8, 20, 30)], "PX"),
    ((0.0001, 0.01, 0.0001), [(10, 21, 30), (10, 22, 30)], "NY"),
    ((0.0001, -0.01, 0.0001), [(10, 19, 30), (10, 18, 30)], "PY"),
    ((0.0001, 0.0001, 0.01), [(10, 20, 31), (10, 20, 32)], "NZ"),
    ((0.0001, 0.0001, -0.01), [(10, 20, 29), (10, 20, 28)], "PZ"),
])
def test_simple_almost_1d(d, cubes, face):
    prefix((10.5, 20.5, 30.5), d, [(10, 20, 30, "Within", 0.0), cubes[0] + (face, 50.0), cubes[1] + (face, 150.0)])


# raycast/tests.rs:148-167
def test_simple_exactly_1d():
    prefix((10.5, 20.5, 30.5), (0.01, 0.0, 0.0), [(10, 20, 30, "Within", 0.0), (11, 20, 30, "NX", 50.0), (12, 20, 30, "NX", 150.0)])
    prefix((10.5, 20.5, 30.5), (-0.01, 0.0, 0.0), [(10, 20, 30, "Within", 0.0), (9, 20, 30, "PX", 50.0), (8, 20, 30, "PX", 150.0)])


# raycast/tests.rs:169-194
@pytest.mark.parametrize("d", [(0.0, 0.0, 0.0), (-0.0, -0.0, -0.0), (1.0, 2.0, math.nan)])
def test_direction_degenerate_produces_origin_cube_only(d):
    exactly((10.5, 20.5, 30.5), d, [(10, 20, 30, "Within", 0.0)])


# raycast/tests.rs:198-236
def test_start_on_cube_edge_parallel():
    prefix((10.0, 20.5, 30.5), (2.0, 0.1, 0.1), [(10, 20, 30, "Within", 0.0), (11, 20, 30, "NX", 0.5), (12, 20, 30, "NX", 1.0)])
    prefix((10.0, 20.5, 30.5), (-2.0, 0.1, 0.1), [(10, 20, 30, "Within", 0.0), (9, 20, 30, "PX", 0.5), (8, 20, 30, "PX", 1.0)])
    prefix((-10.0, 20.5, 30.5), (2.0, 0.1, 0.1), [(-10, 20, 30, "Within", 0.0), (-9, 20, 30, "NX", 0.5), (-8, 20, 30, "NX", 1.0)])
    prefix((-10.0, 20.5, 30.5), (-2.0, 0.1, 0.1), [(-10, 20, 30, "Within", 0.0), (-11, 20, 30, "PX", 0.5), (-12, 20, 30, "PX", 1.0)])


# raycast/tests.rs:240-278
def test_start_on_cube_edge_perpendicular():
    prefix((10.0, 20.5, 30.5), (0.125, 1.0, 0.0), [(10, 20, 30, "Within", 0.0), (10, 21, 30, "NY", 0.5), (10, 22, 30, "NY", 1.5)])
    prefix((10.0, 20.5, 30.5), (-0.125, -1.0, 0.0), [(10, 20, 30, "Within", 0.0), (10, 19, 30, "PY", 0.5), (10, 18, 30, "PY", 1.5)])
    prefix((-10.0, -20.5, 30.5), (0.125, 1.0, 0.0), [(-10, -21, 30, "Within", 0.0), (-10, -20, 30, "NY", 0.5), (-10, -19, 30, "NY", 1.5)])
    prefix((-10.0, -20.5, 30.5), (-0.125, -1.0, 0.0), [(-10, -21, 30, "Within", 0.0), (-10, -22, 30, "PY", 0.5), (-10, -23, 30, "PY", 1.5)])


# raycast/tests.rs:280-285
@pytest.mark.parametrize("include_exit", [False, True])
def test_start_just_past_bounds(include_exit):
    assert orc.raycast((1.5, 0.5, 0.5), (1.0, 0.0, 0.0), bounds=((0, 0, 0), (1, 1, 1)), include_exit=include_exit) == []


# raycast/tests.rs:287-305
def test_start_outside_of_integer_range():
    assert orc.raycast((0.5, 0.5, I32_MAX + 1.5), (0.0, 0.0, -1.0)) == []
    assert orc.raycast((0.5, 0.5, I32_MAX + 2.5), (0.0, 0.0, -1.0)) == []
    assert orc.raycast((0.5, 0.5, I32_MIN - 0.5), (0.0, 0.0, 1.0)) == []
    assert orc.raycast((0.5, 0.5, I32_MIN - 1.5), (0.0, 0.0, 1.0)) == []


# raycast/tests.rs:309-315
@pytest.mark.parametrize("include_exit", [False, True])
def test_start_outside_of_integer_range_with_bounds(include_exit):
    assert orc.raycast((0.0, 1e303, 0.0), (0.0, -1e303, 0.0), bounds=((0, 0, 0), (10, 10, 10)), include_exit=include_exit) == []


# raycast/tests.rs:319-352
def test_exiting_integer_limits():
    highest = I32_MAX - 1
    exactly((0.5, 0.5, float(highest) - 0.5), (0.0, 0.0, 1.0), [(0, 0, highest - 1, "Within", 0.0), (0, 0, highest, "NZ", 0.5)])
    lowest = I32_MIN
    exactly((0.5, 0.5, float(lowest) + 1.5), (0.0, 0.0, -1.0), [(0, 0, lowest + 1, "Within", 0.0), (0, 0, lowest, "PZ", 0.5)])


# raycast/tests.rs:354-380
@pytest.mark.parametrize("include_exit", [False, True])
def test_within_bounds(include_exit):
    expected = [(2, 1, 1, "NX", 2.0), (2, 2, 1, "NY", 2.25), (2, 2, 2, "NZ", 2.5), (3, 2, 2, "NX", 3.0),
                (3, 3, 2, "NY", 3.25), (3, 3, 3, "NZ", 3.5)]
    if include_exit:
        expected.append((4, 3, 3, "NX", 4.0))
    exactly((0.0, -0.25, -0.5), (1.0, 1.0, 1.0), expected, bounds=((2, -10, -10), (4, 10, 10)), include_exit=include_exit)


# raycast/tests.rs:383-396
def test_regression_1():
    prefix((4.833333333333334, 4.666666666666666, -3.0), (0.0, 0.0, 10.0),
           [(4, 4, -3, "Within", 0.0), (4, 4, -2, "NZ", 0.1), (4, 4, -1, "NZ", 0.2)])


# raycast/tests.rs:400-411
@pytest.mark.parametrize("include_exit", [False, True])
def test_regression_2(include_exit):
    assert orc.raycast((18.166666666666668, 4.666666666666666, -3.0), (0.0, 0.0, 16.0), bounds=((0, 0, 0), (10, 10, 10)),
                       include_exit=include_exit) == []


# raycast/tests.rs:417-434 — exact double 0.010000000000000002 after fast_forward
def test_regression_long_distance_fast_forward():
    prefix((6.749300603672869e-67, 6.750109954921438e-67, -85891558.96000093), (1.1036366354256313e-305, 0.0, 8589152896.000092),
           [(0, 0, -30, "NZ", 0.010000000000000002)], bounds=((-10, -20, -30), (10, 20, 30)))


# raycast/tests.rs:437-449
def test_regression_invalid_position_from_beginning():
    assert orc.raycast((10.0, 1.1319598848574732e-72, 2.848094540588472e-306), (-3.39850991e-315, 3.53100099615357e-310, 0.0),
                       bounds=((-10, -20, -30), (10, 20, 30)), include_exit=False) == []


# raycast/tests.rs:451-460 and raycast.rs:399-406
def test_intersection_point_faces():
    got = orc.raycast((0.5, 0.5, 0.5), (-1.0, 0.0, 0.0), max_steps=3)
    assert [g[3] for g in got] == [(0.5, 0.5, 0.5), (0.0, 0.5, 0.5), (-1.0, 0.5, 0.5)]
    got = orc.raycast((0.5, 0.5, 0.5), (1.0, 0.0, 0.0), max_steps=3)
    assert [g[3] for g in got] == [(0.5, 0.5, 0.5), (1.0, 0.5, 0.5), (2.0, 0.5, 0.5)]
    assert [g[1] for g in got] == ["Within", "NX", "NX"]  # raycast.rs:340-345


# raycast.rs:185-192 doc-test
def test_doc_cube_sequence():
    got = orc.raycast((0.5, 0.5, 0.5), (1.0, 0.5, 0.0), max_steps=4)
    assert [g[0] for g in got] == [(0, 0, 0), (1, 0, 0), (1, 1, 0), (2, 1, 0)]


# raycast/tests.rs:463-505 (property test; our own seeded rays since the Rust RNG stream is not reproducible here)
def test_intersection_point_random():
    rng = np.random.default_rng(0)
    n_hits = 0
    for case in range(1000):
        o = rng.uniform(-1.0, 2.0, 3)
        d = rng.uniform(-1.0, 1.0, 3)
        steps = orc.raycast(o, d, bounds=((0, 0, 0), (1, 1, 1)), include_exit=True, max_steps=8)
        assert len(steps) in (0, 2), (case, o, d, steps)
        for s in steps:
            surfaces = sum(1 for v in s[3] if v == 0.0 or v == 1.0)
            interiors = sum(1 for v in s[3] if 0.0 < v < 1.0)
            assert surfaces + interiors == 3 and (surfaces > 0 or s[1] == "Within"), (case, o, d, s)
        n_hits += len(steps) == 2
    assert n_hits > 100


# raycast/tests.rs:507-530
def test_recursive_simple():
    outer = orc.raycast((-1.0, 10.125, 0.125), (1.0, 0.0, 0.0), max_steps=2)
    assert outer[1][0] == (0, 10, 0)
    sub, steps = orc.recursive_raycast((-1.0, 10.125, 0.125), (1.0, 0.0, 0.0), 1, 4, ((0, 0, 0), (4, 4, 4)))
    assert list(sub) == [-4.0, 0.5, 0.5, 1.0, 0.0, 0.0]
    assert steps == [((0, 0, 0), "NX", 4.0), ((1, 0, 0), "NX", 5.0), ((2, 0, 0), "NX", 6.0), ((3, 0, 0), "NX", 7.0),
                     ((4, 0, 0), "NX", 8.0)]


# raycast/tests.rs:532-571
def test_scale_to_integer_step():
    f = orc.scale_to_integer_step
    assert f(1.25, 0.25) == 3.0 and f(1.25, -0.25) == 1.0 and f(-1.25, 0.25) == 1.0 and f(-1.25, -0.25) == 3.0
    for s, ds in [(1.5, 0.0), (1.5, -0.0), (0.0, 0.0), (0.0, -0.0), (-0.0, 0.0)]:
        assert f(s, ds) == math.inf
    assert f(3.0, 0.5) == 2.0 and f(3.0, -0.5) == 2.0 and f(-3.0, 0.5) == 2.0 and f(-3.0, -0.5) == 2.0
    assert math.isnan(f(1.5, math.nan)) and math.isnan(f(math.nan, 1.0)) and math.isnan(f(math.nan, 0.0))
    assert f(-1.9656826074480345e-251, 0.0) == math.inf


# fuzz/fuzz_targets/fuzz_raycast.rs properties: t monotone, each step moves to a face-adjacent cube
def test_fuzz_properties():
    rng = np.random.default_rng(1)
    for case in range(300):
        o = rng.uniform(-20.0, 20.0, 3)
        d = rng.normal(size=3) * 10.0 ** rng.uniform(-3, 3)
        steps = orc.raycast(o, d, bounds=((-8, -8, -8), (8, 8, 8)), max_steps=60)
        last_t = -1.0
        for i, s in enumerate(steps):
            assert s[2] >= last_t
            last_t = s[2]
            if i > 0:
                diff = np.abs(np.array(s[0]) - np.array(steps[i - 1][0]))
                assert diff.sum() == 1, (case, steps[i - 1], s)
