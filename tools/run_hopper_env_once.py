"""ncu driver: a few HopperTrajectory env steps at 1 M envs (tools/profile_r2.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as B   # noqa: E402

print(B.hopper_env_step(num_envs=1 << 20, steps=2))
