"""Restatement of the six `isaacgym.torch_utils` helpers the hot path calls.  TEST INFRASTRUCTURE.

Isaac Gym (Preview 3, reference README.md:29) is closed source and absent from /root/reference, so
these follow its published Python helper file.  Call sites in the reference:
  quat_rotate_inverse  legged_gym/envs/base/legged_robot.py:119-121,579-581
  quat_apply           legged_robot.py:352, legged_gym/utils/math.py:42
  normalize            legged_gym/utils/math.py:41
  torch_rand_float     legged_robot.py:277,371-384,423,443,449,459,740 (restated by the reference itself at
                       legged_gym/utils/helpers.py:129-130)
  to_torch / get_axis_params  legged_robot.py:558,560,727
The quaternion maths was checked against scipy.spatial.transform.Rotation (tests/test_oracle_cpu.py).
Parity of the fp32 *operation order* with the closed-source originals is unpinned.
"""
import numpy as np
import torch


def to_torch(x, dtype=torch.float, device="cpu", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    zs = np.zeros((n_dims,))
    assert axis_idx < n_dims
    zs[axis_idx] = 1.0
    params = np.where(zs == 1.0, value, zs)
    params[0] = x_value
    return list(params.astype(dtype))


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = xyz.cross(b, dim=-1) * 2
    return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_rotate_inverse(q, v):
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower
