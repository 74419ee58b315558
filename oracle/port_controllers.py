"""CPU restatement of deep_tube_learning/controllers.py:4-81 (RaibertHeuristic).  TEST INFRASTRUCTURE.

The reference file is pure torch and runs unchanged in the build container; tests/test_oracle_cpu.py checks this restatement
against it bit for bit and tests/golden/raibert_reference.npz (written by oracle/make_golden_controllers.py from the reference
itself) travels to the GPU box.  Every line follows the reference's operation order.
"""
import torch


def quat_to_yaw(quat):                                                    # controllers.py:75-81
    x, y, z, w = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
    return torch.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))


def omega_to_quat(pitch, roll, yaw):                                      # controllers.py:23-36
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    return torch.stack((cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp * cy), dim=-1)


def raibert_policy(obs, Kp, Kv, K_ff, clip_pos, clip_vel, clip_ang):      # controllers.py:38-73
    pex, pey, evx, evy, dvx, dvy = obs[:, 0], -obs[:, 1], -obs[:, 2], obs[:, 3], obs[:, 4], -obs[:, 5]
    pitch_pos = torch.clamp(-Kp * pex, -clip_pos, clip_pos)
    roll_pos = torch.clamp(-Kp * pey, -clip_pos, clip_pos)
    vx = torch.clamp(-Kv * evx + K_ff * dvx, -clip_vel, clip_vel)
    vy = torch.clamp(-Kv * evy + K_ff * dvy, -clip_vel, clip_vel)
    pitch = torch.clamp(pitch_pos + vx, -clip_ang, clip_ang)
    roll = torch.clamp(roll_pos + vy, -clip_ang, clip_ang)
    return omega_to_quat(pitch, roll, quat_to_yaw(obs[:, 6:10]))


def sample_obs(n, seed=0):
    """Inputs that exercise every clamp: errors of O(1), velocities of O(2), random unit quaternions (xyzw)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 12, generator=g)
    o[:, 0:2] *= 2.0
    o[:, 2:6] *= 1.5
    q = torch.randn(n, 4, generator=g)
    o[:, 6:10] = q / q.norm(dim=-1, keepdim=True)
    return o


GAINS = dict(Kp=-0.3, Kv=-0.9, K_ff=0.0, clip_pos=0.5, clip_vel=1.0, clip_ang=0.2)   # hopper_trajectory_config.py rewards.raibert
GAINS_B = dict(Kp=0.8, Kv=0.25, K_ff=0.4, clip_pos=0.3, clip_vel=0.4, clip_ang=0.6)
