"""Synthetic / replayed physics state for the step pipeline.

PhysX is closed source (reference legged_robot.py:92-96,111-112 call into it), so the pipeline is
driven by a *replay tape*: K frames, each holding what `gym.simulate` + `refresh_*` would have left in
the aliased state tensors for one env step — 4 `dof_state` sub-frames (one per decimation sub-step),
one `root_states`, one `contact_forces` — plus the policy actions for that step.  Distributions follow
SURVEY.md §8d (cfg 2 / cfg 3).
"""
import math
from types import SimpleNamespace

import torch

NUM_DOF = 12
NUM_BODIES = 17
# body order of the collapsed ANYmal-C asset: base, then (HIP, THIGH, SHANK, FOOT) per leg LF, LH, RF, RH
FEET_INDICES = (4, 8, 12, 16)
PENALISED_INDICES = (3, 7, 11, 15, 2, 6, 10, 14)   # SHANK*, then THIGH* (anymal_c_rough_config.py:76)
TERMINATION_INDICES = (0,)                         # base (anymal_c_rough_config.py:77)
DOF_NAMES = tuple(f"{leg}_{j}" for leg in ("LF", "LH", "RF", "RH") for j in ("HAA", "HFE", "KFE"))
DEFAULT_DOF_POS = (0.0, 0.4, -0.8, 0.0, -0.4, 0.8, 0.0, 0.4, -0.8, 0.0, -0.4, 0.8)


def anymal_dof_limits():
    """Soft joint limits / velocity / effort limits of the collapsed ANYmal-C asset (what LeggedRobot reads from the URDF through
    `gym.get_asset_dof_properties`, legged_robot.py:264-283): synthetic stand-in used wherever no Isaac Gym asset is loaded."""
    lo = torch.tensor([-0.72, -1.2, -1.8] * 4, dtype=torch.float)
    hi = torch.tensor([0.49, 1.2, 1.8] * 4, dtype=torch.float)
    lo[3:6], hi[3:6] = torch.tensor([-0.49, -1.2, -1.8]), torch.tensor([0.72, 1.2, 1.8])
    return dict(dof_pos_limits=torch.stack([lo, hi], dim=1), dof_vel_limits=torch.full((12,), 20.0),
                torque_limits=torch.full((12,), 80.0))


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def make_state_tape(num_envs, frames=32, seed=0, rough=False, device="cpu", decimation=4,
                    base_contact_prob=0.002, dtype=torch.float32):
    """Returns a namespace of stacked frames:
       root [K,N,13], dof [K,decimation,N*12,2], contact [K,N,17,3], actions [K,N,12]."""
    N, K = int(num_envs), int(frames)
    g = _gen(seed, device)
    rn = lambda *s: torch.randn(*s, generator=g, device=device, dtype=dtype)
    ru = lambda *s: torch.rand(*s, generator=g, device=device, dtype=dtype)

    root = torch.empty(K, N, 13, device=device, dtype=dtype)
    if rough:
        root[..., 0] = ru(K, N) * 80.0
        root[..., 1] = ru(K, N) * 160.0
    else:
        root[..., 0:2] = ru(K, N, 2) * 50.0
    root[..., 2] = 0.55 + 0.03 * rn(K, N)
    q = torch.cat([0.1 * rn(K, N, 3), torch.ones(K, N, 1, device=device, dtype=dtype)], dim=-1)
    root[..., 3:7] = q / q.norm(dim=-1, keepdim=True)
    root[..., 7:10] = 0.5 * rn(K, N, 3)
    root[..., 10:13] = 0.7 * rn(K, N, 3)

    q0 = torch.tensor(DEFAULT_DOF_POS, device=device, dtype=dtype)
    dof = torch.empty(K, decimation, N, NUM_DOF, 2, device=device, dtype=dtype)
    dof[..., 0] = q0 + 0.2 * rn(K, decimation, N, NUM_DOF)
    dof[..., 1] = 2.0 * rn(K, decimation, N, NUM_DOF)
    dof = dof.reshape(K, decimation, N * NUM_DOF, 2)

    contact = torch.zeros(K, N, NUM_BODIES, 3, device=device, dtype=dtype)
    feet = list(FEET_INDICES)
    fz = torch.clamp(120.0 + 80.0 * rn(K, N, 4), min=0.0) * (ru(K, N, 4) > 0.4)
    contact[:, :, feet, 2] = fz
    contact[:, :, feet, 0:2] = 15.0 * rn(K, N, 4, 2) * (fz > 0).unsqueeze(-1)
    pen = list(PENALISED_INDICES)
    contact[:, :, pen, :] = 5.0 * rn(K, N, 8, 3) * (ru(K, N, 8, 1) < 0.02)
    contact[:, :, 0, :] = 20.0 * rn(K, N, 3) * (ru(K, N, 1) < base_contact_prob)

    actions = rn(K, N, NUM_DOF)
    return SimpleNamespace(root=root, dof=dof, contact=contact, actions=actions,
                           num_envs=N, frames=K, decimation=decimation)


def make_episode_lengths(num_envs, max_episode_length=1000, seed=0, device="cpu"):
    """init_at_random_ep_len-style episode counters (rsl_rl OnPolicyRunner.learn)."""
    g = _gen(seed + 7919, device)
    return torch.randint(0, int(max_episode_length) + 1, (num_envs,), generator=g, device=device, dtype=torch.int64)


def make_heightfield(rows=1300, cols=2100, seed=0, vertical_scale=0.005, device="cpu"):
    """int16 terrain (cfg 3): three low-frequency sinusoids (<=0.4 m total) + 20 % of 8 m x 8 m tiles
    carrying 0.1-0.2 m steps, quantised at `vertical_scale`."""
    g = _gen(seed + 104729, "cpu")
    x = torch.arange(rows, dtype=torch.float32).unsqueeze(1) * 0.1
    y = torch.arange(cols, dtype=torch.float32).unsqueeze(0) * 0.1
    h = torch.zeros(rows, cols)
    for _ in range(3):
        fx, fy = (torch.rand(2, generator=g) * 0.25 + 0.02).tolist()
        ph = (torch.rand(1, generator=g) * 2 * math.pi).item()
        amp = (torch.rand(1, generator=g) * 0.1 + 0.03).item()
        h += amp * torch.sin(fx * x + fy * y + ph)
    tr, tc = (rows + 79) // 80, (cols + 79) // 80
    stepped = torch.rand(tr, tc, generator=g) < 0.2
    step_h = torch.rand(tr, tc, generator=g) * 0.1 + 0.1
    cell = (torch.arange(rows).unsqueeze(1) // 5 + torch.arange(cols).unsqueeze(0) // 5) % 2
    tile_mask = stepped.repeat_interleave(80, 0)[:rows].repeat_interleave(80, 1)[:, :cols]
    tile_step = step_h.repeat_interleave(80, 0)[:rows].repeat_interleave(80, 1)[:, :cols]
    h = h + tile_mask * tile_step * cell
    return torch.round(h / vertical_scale).to(torch.int16).to(device)


def make_terrain_origins(num_rows=10, num_cols=20, terrain_length=8.0, terrain_width=8.0, border=25.0, seed=0):
    """Per-tile platform origins [num_rows, num_cols, 3] like Terrain.env_origins (terrain.py:147-187),
    expressed in world coordinates (border already subtracted by the reference's mesh transform)."""
    g = _gen(seed + 15485863, "cpu")
    i = torch.arange(num_rows, dtype=torch.float32).unsqueeze(1).expand(num_rows, num_cols)
    j = torch.arange(num_cols, dtype=torch.float32).unsqueeze(0).expand(num_rows, num_cols)
    o = torch.empty(num_rows, num_cols, 3)
    o[..., 0] = (i + 0.5) * terrain_length
    o[..., 1] = (j + 0.5) * terrain_width
    o[..., 2] = torch.rand(num_rows, num_cols, generator=g) * 0.3
    return o
