"""Pins the oracle's colour arithmetic: PackedLight LUT/quantiser (space/light/data.rs:301-354,
:357-496), sRGB encode (math/color.rs:1038-1054 + doc-tests), tone mapping
(graphics_options.rs:352-368)."""
import json
import math
import os

import numpy as np

import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_packed_light_lut_matches_reference_table():
    table = json.load(open(os.path.join(GOLDEN, "packed_light_lut.json")))
    got = [orc.lib().orc_packed_light_lut(i) for i in range(256)]
    assert [np.float32(v) for v in table] == [np.float32(v) for v in got]
    assert got[0] == 0.0 and got[144] == 1.0 and got[154] == 2.0


# light/data.rs tests: round trip through the quantiser is the identity on table entries, ONE = 144
def test_packed_light_quantiser_round_trip():
    L = orc.lib()
    assert L.orc_packed_light_scalar_in(1.0) == 144
    assert L.orc_packed_light_scalar_in(0.0) == 0          # log2(0) = -inf saturates to 0
    assert L.orc_packed_light_scalar_in(1e30) == 255
    for i in range(1, 256):
        assert L.orc_packed_light_scalar_in(L.orc_packed_light_lut(i)) == i


def test_srgb_encode_known_values():
    # transmittance 0 => alpha 255; linear 0, 1, 0.5, and the linear segment
    assert orc.to_srgb8([0.0, 1.0, 0.5, 0.0]) == (0, 255, 188, 255)
    assert orc.to_srgb8([0.0031308, 0.002, 0.0001, 0.0]) == (10, 7, 0, 255)
    # out-of-range saturates (`as u8`), fully transparent buffer is Rgba::TRANSPARENT
    assert orc.to_srgb8([7.0, 2.0, 1.5, 0.0]) == (255, 255, 255, 255)
    assert orc.to_srgb8([0.3, 0.3, 0.3, 1.0]) == (0, 0, 0, 0)
    # un-premultiply: light 0.25 at alpha 0.5 is colour 0.5
    assert orc.to_srgb8([0.25, 0.25, 0.25, 0.5]) == (188, 188, 188, 128)
    # negative / NaN light becomes pure red (raytracer_components.rs:142-143)
    assert orc.to_srgb8([-1.0, 0.5, 0.5, 0.0]) == (255, 0, 0, 255)
    assert orc.to_srgb8([math.nan, 0.5, 0.5, 0.0]) == (255, 0, 0, 255)


def test_srgb8_round_trip_all_levels():
    """from_srgb8 -> to_srgb8 is the identity for all 256 levels (color.rs tests)."""
    from aicb200 import srgb8_to_linear
    for v in range(256):
        lin = srgb8_to_linear((v, v, v))
        assert orc.to_srgb8([lin[0], lin[1], lin[2], 0.0])[:3] == (v, v, v)


def test_tone_mapping():
    # Clamp with finite maximum; infinite maximum disables (graphics_options.rs:352-357)
    assert orc.to_srgb8([4.0, 0.5, 0.0, 0.0], tone_mapping=0, maximum_intensity=1.0) == (255, 188, 0, 255)
    assert orc.to_srgb8([0.5, 0.5, 0.5, 0.0], exposure=2.0) == (255, 255, 255, 255)
    # Reinhard: c / (1 + lum/max); grey 1.0 at max 1 -> 0.5
    assert orc.to_srgb8([1.0, 1.0, 1.0, 0.0], tone_mapping=1, maximum_intensity=1.0) == (188, 188, 188, 255)
    assert orc.to_srgb8([1.0, 1.0, 1.0, 0.0], tone_mapping=1) == (255, 255, 255, 255)
