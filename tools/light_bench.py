"""BASELINE configs[4]: N^3 Space of res-1 blocks, LightPhysics::Rays{30}; converge
(fast_evaluate_light + evaluate_light(1)), apply K random edits, propagate to epsilon 1, re-render.
Prints one JSON line with cube-updates/s and chart-node-visits/s for the GPU and (on a bounded N) the oracle."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-is-cubes_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aicb200  # noqa: E402
from aicb200 import Block, GraphicsOptions, Space, SpaceRaytracer, scenes  # noqa: E402


def make_space(n, seed=4):
    return scenes.config_c4(n, seed)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_edits = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    oracle_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    space = make_space(n)
    rng = np.random.default_rng(7)
    cubes = np.stack([rng.integers(0, n, n_edits), rng.integers(n // 4 - 2, n, n_edits), rng.integers(0, n, n_edits)], axis=1).astype(np.int32)
    new_ids = rng.integers(0, len(space.blocks), n_edits).astype(np.uint16)
    opts = GraphicsOptions(view_distance=4.0 * n)
    cam = scenes.standard_camera(space, opts, 1920, 1080)
    rt = SpaceRaytracer(space, opts)
    t0 = time.perf_counter()
    rt.light_fast_evaluate()
    upd0, md0, nv0 = rt.light_evaluate(1)
    t_conv = time.perf_counter() - t0
    t0 = time.perf_counter()
    upd1, md1 = rt.light_edit_and_propagate(cubes, new_ids, 1)
    t_edit = time.perf_counter() - t0
    r = aicb200.RtRenderer(cam)
    r.rt = rt
    img = r.draw()
    out = {"workload": f"C4: {n}^3 res-1 Space, LightPhysics::Rays{{30}}, octant sky; converge, {n_edits} random edits, propagate (eps 1), re-render 1080p",
           "initial_convergence": {"cube_updates": upd0, "seconds": t_conv, "cube_updates_per_s": upd0 / t_conv, "chart_node_visits": nv0,
                                   "node_visits_per_s": nv0 / t_conv},
           "after_edits": {"cube_updates": upd1, "seconds": t_edit, "cube_updates_per_s": upd1 / max(t_edit, 1e-9), "max_difference": md1},
           "rerender_kernel_ms": img.info.kernel_ms}
    if oracle_n:
        import orc
        sp2 = make_space(oracle_n)
        ol = orc.OracleLight(sp2)
        t0 = time.perf_counter()
        ol.fast_evaluate()
        nup, _ = ol.evaluate(1)
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"n": oracle_n, "cube_updates": nup, "seconds": dt, "cube_updates_per_s": nup / dt, "threads": 1,
                             "note": "sequential port of update_light_from_queue (non-threaded variant)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
