"""Per-rank kernel times of one shard of the frame for N = 1, 2, 4, 8 ranks, on ONE GPU (what each rank of an N-GPU
run computes; no delivery): separates the serial floor of the marching kernel from the multi-GPU plumbing.

  python tools/shard_sweep.py c2[,c3]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-is-cubes_b200"))
import bench  # noqa: E402
import aicb200  # noqa: E402
from aicb200 import scenes  # noqa: E402

for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2"]):
    space, opts, w, h, desc = bench.make_workload(name)
    cam = scenes.standard_camera(space, opts, w, h)
    r = aicb200.RtRenderer(cam)
    r.update(space)
    for count in (1, 2, 4, 8):
        worst = None
        for index in range(count):
            ms, st = [], []
            for i in range(5):
                img = r.draw(shard=(16, index, count))
                if i >= 2:
                    ms.append(img.info.kernel_ms)
                    st.append(img.info.stage_ms)
            m = float(np.mean(ms))
            if worst is None or m > worst[0]:
                worst = (m, np.mean(np.array(st), axis=0), index)
        m, s, idx = worst
        print(f"{name} ranks={count} slowest shard {idx}: frame {m:.3f} ms  gen/march/shade/encode {s[0]:.3f}/{s[1]:.3f}/{s[2]:.3f}/{s[3]:.3f}", flush=True)
