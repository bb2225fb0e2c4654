"""Builds experimental variants of libaicb200.so (compile-time knobs) and times each on the bench workloads.

  python tools/variants.py build  name:DEF=V,DEF=V ...     (here, no GPU)
  python tools/variants.py run c2,c1 name name ...         (on the GPU box)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build_variants")


def main():
    if sys.argv[1] == "build":
        import __graft_entry__ as g
        os.makedirs(OUT, exist_ok=True)
        for spec in sys.argv[2:]:
            name, _, defs = spec.partition(":")
            g.build_library(defines=[d for d in defs.split(",") if d], out=os.path.join(OUT, f"lib_{name}.so"))
            print("built", name)
    else:
        workloads = sys.argv[2]
        for name in sys.argv[3:]:
            env = dict(os.environ, AICB200_LIB=os.path.join(OUT, f"lib_{name}.so"))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_sweep.py"), workloads, "1", os.environ.get("SWEEP_EV", "12")],
                               env=env, capture_output=True, text=True, timeout=300)
            for line in r.stdout.splitlines():
                print(f"[{name}] {line}", flush=True)
            if r.returncode:
                print(f"[{name}] FAILED: {r.stderr[-500:]}")


if __name__ == "__main__":
    main()
