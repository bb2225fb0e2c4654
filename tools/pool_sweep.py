"""Times the pool scheduler variants (GPU only): python tools/pool_sweep.py c2 name:blocks ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[2:]:
    name, _, blocks = spec.partition(":")
    env = dict(os.environ, AICB_SCHED="pool", AICB_PROFILE_KERNELS="1")
    if blocks:
        env["AICB_BLOCKS_PER_SM"] = blocks
    if name != "product":
        env["AICB200_LIB"] = os.path.join(ROOT, "build_variants", f"lib_{name}.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_sweep.py"), sys.argv[1], "16", "32"], env=env,
                       capture_output=True, text=True, timeout=200)
    lines = [l for l in (r.stdout + r.stderr).splitlines() if "gen " in l or "march warps" in l]
    print(f"{spec:14s}", lines[-1][10:60] if lines else r.stderr[-300:], "|", lines[-2][10:] if len(lines) > 1 else "", flush=True)
