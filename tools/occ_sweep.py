"""Marching-kernel time against the number of resident blocks per SM (GPU only)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for b in sys.argv[2].split(","):
    env = dict(os.environ, AICB_BLOCKS_PER_SM=b, AICB_PROFILE_KERNELS="1")
    if len(sys.argv) > 3:
        env["AICB200_LIB"] = os.path.join(ROOT, "build_variants", f"lib_{sys.argv[3]}.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_sweep.py"), sys.argv[1], "16", "32"], env=env,
                       capture_output=True, text=True, timeout=300)
    lines = [l for l in (r.stdout + r.stderr).splitlines() if "gen " in l or "event_thr" in l]
    print(f"blocks/SM={b} {sys.argv[3] if len(sys.argv) > 3 else 'product'}:", lines[-2][10:] if len(lines) > 1 else "", "|", lines[-1] if lines else r.stderr[-300:], flush=True)
