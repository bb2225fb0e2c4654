"""Frame time against resolution on a bench workload: separates the throughput term from the tail (GPU only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-is-cubes_b200"))
import bench  # noqa: E402
import aicb200  # noqa: E402
from aicb200 import scenes  # noqa: E402


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2"]
    for name in names:
        space, opts, w, h, desc = bench.make_workload(name)
        for (rw, rh) in [(240, 135), (480, 270), (960, 540), (1920, 1080), (3840, 2160), (7680, 4320)]:
            cam = scenes.standard_camera(space, opts, rw, rh)
            r = aicb200.RtRenderer(cam)
            r.update(space)
            ms = []
            for i in range(6):
                img = r.draw()
                if i >= 2:
                    ms.append(img.info.kernel_ms)
            print(f"{name} {rw}x{rh} rays={rw * rh} frame_ms={np.mean(ms):.3f} (min {np.min(ms):.3f}) Mrays/s={rw * rh / np.mean(ms) / 1e3:.0f}", flush=True)


if __name__ == "__main__":
    main()
