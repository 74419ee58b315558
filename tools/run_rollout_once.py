"""ncu driver: one PPO training iteration at 4096 envs (act / store kernels, external reset) — tools/profile_r2.sh."""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from legged_gym_dev_b200 import synthetic as S                                   # noqa: E402
from legged_gym_dev_b200.physics import ReplayPhysics                            # noqa: E402
from legged_gym_dev_b200.task_registry import task_registry                      # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tape = S.make_state_tape(N, frames=4, seed=10, device="cuda")
args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
runner.learn(num_learning_iterations=2)
env.reset_idx(torch.arange(0, N, 2, device="cuda"))
torch.cuda.synchronize()
print("ok")
