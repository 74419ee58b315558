"""A/B of the rough post-physics kernel with the terrain windows staged by 2-D TMA (default) vs per-point gathers (B200GYM_HEIGHT_TMA=0);
one process per setting (the knob is read once).  python tools/ab_rough_tma.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if "--one" in sys.argv:
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as B
    n = int(sys.argv[sys.argv.index("--one") + 1])
    r = B.rough_lstm(num_envs=n, steps=30)
    print(json.dumps(dict(num_envs=n, tma=os.environ.get("B200GYM_HEIGHT_TMA", "1"), ms_per_step=r["ms_per_step"],
                          post_physics_ms=r["post_physics_rough"]["avg_launch_ms"], post_physics_frac=r["post_physics_rough"]["frac"])))
    sys.exit(0)
for n in (16384, 262144, 1048576):
    for tma in ("1", "0"):
        out = subprocess.run([sys.executable, __file__, "--one", str(n)], env=dict(os.environ, B200GYM_HEIGHT_TMA=tma), capture_output=True, text=True)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else "ERR " + out.stderr[-400:], flush=True)
