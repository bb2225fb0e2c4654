"""Renders a few frames of one bench workload (for `ncu` captures: keeps the profiled command short).

  python tools/prof_frame.py c2 [frames]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-is-cubes_b200"))
import bench  # noqa: E402
import aicb200  # noqa: E402
from aicb200 import scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
space, opts, w, h, desc = bench.make_workload(name)
cam = scenes.standard_camera(space, opts, w, h)
r = aicb200.RtRenderer(cam)
r.update(space)
for i in range(frames):
    img = r.draw()
print(desc, "frame_ms", img.info.kernel_ms, "stages", img.info.stage_ms)
