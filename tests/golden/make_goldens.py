"""Generates the committed golden fixtures from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_goldens.py

Outputs (all small, committed):
  packed_light_lut.json   the 256-entry PackedLight decode table (all-is-cubes/src/space/light/data.rs:301-354)
  text_images.json        the two 80x40 ASCII renderings (all-is-cubes-render/src/raytracer/text.rs:196-258, 265-341)
  png_*.npy               expected renderer images (test-renderers/expected/renderers/*.png) as uint8 [H,W,4]
"""
import json
import os
import re

import numpy as np
from PIL import Image

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def lut():
    src = open(f"{REF}/all-is-cubes/src/space/light/data.rs").read()
    i = src.index("static PACKED_LIGHT_SCALAR_LOOKUP_TABLE")
    j = src.index("];", i)
    vals = [float(np.float32(x)) for x in re.findall(r"ps32\(([^)]+)\)", src[i:j])]
    assert len(vals) == 256
    json.dump(vals, open(f"{OUT}/packed_light_lut.json", "w"))


def text_images():
    src = open(f"{REF}/all-is-cubes-render/src/raytracer/text.rs").read()
    images = []
    for m in re.finditer(r'assert_eq!\(\s*output,\s*"\\\n(.*?)\s*"\s*\);', src, re.S):
        rows = re.findall(r"^\s*(.{80})\\n\\$", m.group(1), re.M)
        assert len(rows) == 40, len(rows)
        images.append(rows)
    assert len(images) == 2
    json.dump({"print_space_test": images[0], "partial_voxels": images[1]}, open(f"{OUT}/text_images.json", "w"), indent=0)


def pngs():
    for name in ["transparent_one-surf-all", "transparent_one-vol-all", "emission-all", "emission_only-surf-all",
                 "emission_only-vol-all", "emission_semi-surf-all", "emission_semi-vol-all", "debug_pixel_cost-ray", "color_srgb_ramp-all", "furnace-Clear-Opaque-all",
                 "furnace-Clear-Transparent-all", "furnace-Foggy-Opaque-all", "furnace-Foggy-Transparent-all",
                 "fog-Abrupt-all", "fog-Compromise-all", "fog-Physical-all", "light_spread-None-all", "light_spread-Flat-all",
                 "light_spread-Coarse-all", "light_spread-Linear-all", "light_spread-Smoothstep-all", "tone_map-Clamp-1.0-0.5-all",
                 "tone_map-Clamp-1.0-2.0-all", "tone_map-Reinhard-0.5-0.5-all", "tone_map-Reinhard-1.0-0.5-all",
                 "tone_map-Reinhard-1.0-2.0-all", "viewport_prime-all", "no_update-all", "no_update-2-all", "layers_all-all", "layers_hidden_ui-all", "layers_ui_only-all",
                 "layers_none_but_text-all", "light_on_slab-None-all", "light_on_slab-Flat-all", "light_on_slab-Coarse-all",
                 "light_on_slab-Linear-all", "light_on_slab-Smoothstep-all"]:
        im = np.array(Image.open(f"{REF}/test-renderers/expected/renderers/{name}.png").convert("RGBA"))
        np.save(f"{OUT}/png_{name}.npy", im)


if __name__ == "__main__":
    lut()
    text_images()
    pngs()
    print("goldens written to", OUT)
