"""ingestion of the reference's native save format (SpaceV1 / UniverseV1 JSON) without Rust,
pinned on the serialized example of the reference's own test (all-is-cubes/src/save/tests.rs:653-746)."""
import numpy as np
import pytest

from aicb200 import ingest

NL, IN, BS = [0, 0, 0, 1], [0, 0, 0, 2], [144, 144, 144, 3]   # LightSerV1 of NO_RAYS, OPAQUE, Rgb::ONE (tests.rs:676-678)


def space_success_json():
    contents = [0, 0, 1, 0, 0, 0, 0, 0, 0] + [0] * 18
    light = [NL, BS, IN, NL, NL, BS, NL, NL, NL, NL, NL, BS] + [NL] * 15
    return {
        "type": "SpaceV1",
        "bounds": {"lower": [1, 2, 3], "upper": [4, 5, 6]},
        "physics": {"gravity": [0.0, 0.25, 1.0], "sky": {"type": "UniformV1", "color": [1.0, 1.0, 1.0]},
                    "light": {"type": "RaysV1", "maximum_distance": 123}},
        "spawn": {"type": "SpawnV1", "bounds": {"lower": [1, 2, 6], "upper": [4, 5, 46]}, "eye_position": None,
                  "inventory": [], "look_direction": [0.0, 0.0, -1.0]},
        "blocks": [
            {"type": "BlockV1", "primitive": {"type": "AirV1"}},
            {"type": "BlockV1", "primitive": {"type": "AtomV1", "color": [0.5, 0.5, 0.5, 1.0]},
             "modifiers": [{"type": "DisplayNameV1", "display_name": "0"}]},
        ],
        "contents": ingest.gz_encode(np.array(contents, dtype="<u2").tobytes()),
        "light": ingest.gz_encode(np.array(light, dtype=np.uint8).tobytes()),
    }


def test_space_v1_example_of_the_reference():
    s = ingest.space_from_value(space_success_json())
    assert tuple(s.lower) == (1, 2, 3) and tuple(s.size) == (3, 3, 3)
    assert s.light_max_distance == 123
    assert len(s.blocks) == 2 and s.blocks[0].is_air
    assert tuple(s.blocks[1].palette[0, :4]) == (0.5, 0.5, 0.5, 1.0)
    # m.set([1, 2, 5], block): cube (0, 0, 2) relative to the lower corner, Z-major (vol.rs:1013-1018)
    assert s.block_ids[0, 0, 2] == 1 and int(s.block_ids.sum()) == 1
    # status bytes: LightSerV1 -> PackedLight texel (NoRays 1 -> 1, Opaque 2 -> 128, Visible 3 -> 255)
    assert list(s.light[0, 0, 0]) == [0, 0, 0, 1]
    assert list(s.light[0, 0, 1]) == [144, 144, 144, 255]
    assert list(s.light[0, 0, 2]) == [0, 0, 0, 128]
    assert list(s.light[0, 1, 2]) == [144, 144, 144, 255]
    assert list(s.light[1, 0, 2]) == [144, 144, 144, 255]
    assert [tuple(c) for c in s.sky_colors] == [(1.0, 1.0, 1.0)]


def test_gzserde_round_trip_and_base64_without_padding():
    for n in (1, 2, 3, 100):
        data = bytes(range(n % 256)) * 3 + b"x" * n
        enc = ingest.gz_encode(data)
        assert "=" not in enc["Base64Gzip"]                      # STANDARD_NO_PAD (compress.rs:103-104)
        assert ingest.gz_decode(enc) == data


def test_invalid_index_is_rejected():
    v = space_success_json()
    v["contents"] = ingest.gz_encode(np.array([0, 999, 0] + [0] * 24, dtype="<u2").tobytes())
    with pytest.raises(ValueError):
        ingest.space_from_value(v)                               # save/tests.rs:749-785


def test_universe_with_a_recursive_block():
    voxels = {
        "type": "SpaceV1", "bounds": {"lower": [0, 0, 0], "upper": [4, 2, 4]},
        "physics": {"gravity": [0, 0, 0], "sky": {"type": "UniformV1", "color": [0, 0, 0]}, "light": {"type": "NoneV1"}},
        "blocks": [{"type": "BlockV1", "primitive": {"type": "AirV1"}},
                   {"type": "BlockV1", "primitive": {"type": "AtomV1", "color": [1.0, 0.0, 0.0, 1.0], "light_emission": [0.0, 2.0, 0.0]}}],
        "contents": ingest.gz_encode(np.array([1] * 32, dtype="<u2").tobytes()), "light": None,
    }
    world = {
        "type": "SpaceV1", "bounds": {"lower": [0, 0, 0], "upper": [2, 1, 1]},
        "physics": {"gravity": [0, 0, 0], "sky": {"type": "OctantsV1", "colors": [[0.1 * i, 0.0, 0.0] for i in range(8)]},
                    "light": {"type": "NoneV1"}},
        "blocks": [{"type": "BlockV1", "primitive": {"type": "AirV1"}},
                   {"type": "BlockV1", "primitive": {"type": "RecurV1", "space": {"type": "HandleV1", "Specific": "vox"}, "resolution": 4}},
                   {"type": "BlockV1", "primitive": {"type": "AtomV1", "color": [0, 0, 1, 1]}, "modifiers": [{"type": "RotateV1", "rotation": "RXyZ"}]}],
        "contents": ingest.gz_encode(np.array([1, 0], dtype="<u2").tobytes()), "light": None,
    }
    u = {"type": "UniverseV1", "members": [{"name": {"Specific": "vox"}, "member_type": "Space", "value": voxels},
                                          {"name": {"Specific": "world"}, "member_type": "Space", "value": world}]}
    spaces = ingest.spaces_from_universe(u)
    assert ingest.name_key({"Specific": "vox"}) in spaces
    assert ingest.name_key({"Specific": "world"}) not in spaces      # RotateV1 needs the block evaluator: left out
    world["blocks"].pop()
    spaces = ingest.spaces_from_universe(u)
    w = spaces[ingest.name_key({"Specific": "world"})]
    b = w.blocks[1]
    assert b.resolution == 4 and b.voxel_size == (4, 2, 4) and b.voxel_lower == (0, 0, 0)   # partial voxel bounds
    assert tuple(b.palette[1, :7]) == (1.0, 0.0, 0.0, 1.0, 0.0, 2.0, 0.0)
    assert len(w.sky_colors) == 8 and w.light is None and w.light_max_distance == 0
