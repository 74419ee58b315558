"""ncu driver: the Hopper kernels once each at 1M envs.  Usage: ncu ... python tools/run_hopper_once.py [num_envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_dev_b200.hopper import HopperActuation   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = HopperActuation(N, device="cuda", torque_limits=[9000.0, 80.0, 80.0, 80.0])
g = torch.Generator(device="cuda").manual_seed(1)
for t in (env.dof_state, env.root_states, env.base_ang_vel, env.base_lin_vel, env.contact_forces, env.commands, env.last_dof_vel):
    t.normal_(generator=g)
act = torch.randn(N, 4, device="cuda", generator=g)
env.actions.copy_(act)
for _ in range(2):
    env._compute_torques(act)
    env.compute_observations()
    env._reward_terms()
torch.cuda.synchronize()
print("ok", float(env.obs_buf.abs().mean()))
