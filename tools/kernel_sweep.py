"""Measures trace-kernel time on the bench workloads for a sweep of runtime knobs (GPU only)."""
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
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c1"]
    thresholds = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4", "8", "16", "32"])]
    for name in names:
        space, opts, w, h, desc = bench.make_workload(name)
        cam = scenes.standard_camera(space, opts, w, h)
        r = aicb200.RtRenderer(cam)
        r.update(space)
        events = [int(v) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["8"])]
        for ev in events:
            for th in thresholds:
                os.environ["AICB_REFILL_THRESHOLD"] = str(th)
                os.environ["AICB_EVENT_THRESHOLD"] = str(ev)
                ms, stages = [], []
                for i in range(6):
                    img = r.draw()
                    if i >= 2:
                        ms.append(img.info.kernel_ms)
                        stages.append([float(v) for v in img.info.stage_ms])
                st = np.mean(np.array(stages), axis=0)
                print(f"{name} event_thr={ev:2d} refill_thr={th:2d} frame_ms={np.mean(ms):.3f} (min {np.min(ms):.3f})  "
                      f"gen/march/shade/encode {st[0]:.3f}/{st[1]:.3f}/{st[2]:.3f}/{st[3]:.3f}  Mrays/s={w * h / np.mean(ms) / 1e3:.0f}", flush=True)


if __name__ == "__main__":
    main()
