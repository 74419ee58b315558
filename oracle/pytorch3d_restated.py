"""Restatement of the pytorch3d.transforms functions the Hopper classes import (legged_gym/envs/hopper/hopper.py:38, hopper_trajectory.py:39-40):
quaternion_invert, quaternion_multiply, quaternion_to_matrix, so3_log_map, Rotate(...).transform_points (torque law), euler_angles_to_matrix,
matrix_to_quaternion (yaw randomisation of the reset).  TEST INFRASTRUCTURE.

pytorch3d is a third-party dependency that is absent from this image and from /root/reference, and the reference pins no version of it
(no requirements entry, README silent).  What follows restates the published algorithm of pytorch3d v0.7.x
(pytorch3d/transforms/rotation_conversions.py, so3.py, math.py, transform3d.py), real-first (w, x, y, z) quaternions, in the same
operation order.  oracle/ref_harness.py injects this module as `pytorch3d.transforms`, so the UNMODIFIED Hopper._compute_torques runs on
it; tests/test_hopper_cpu.py checks each function against scipy's Rotation (fp64) — that pins the mathematics, not pytorch3d's rounding.
"""
import math

import torch


def quaternion_invert(quaternion):
    return quaternion * quaternion.new_tensor([1, -1, -1, -1])


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def standardize_quaternion(quaternions):
    return torch.where(quaternions[..., 0:1] < 0, -quaternions, quaternions)


def quaternion_multiply(a, b):
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def quaternion_to_matrix(quaternions):
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def acos_linear_extrapolation(x, bounds):
    """math.py: arccos inside (lower, upper), first-order Taylor continuation outside."""
    lower, upper = bounds
    if lower > upper:
        raise ValueError("lower bound has to be smaller or equal to upper bound.")
    if lower <= -1.0 or upper >= 1.0:
        raise ValueError("Both lower bound and upper bound have to be within (-1, 1).")
    acos_extrap = torch.empty_like(x)
    x_upper, x_lower = x >= upper, x <= lower
    x_mid = (~x_upper) & (~x_lower)
    acos_extrap[x_mid] = torch.acos(x[x_mid])
    acos_extrap[x_upper] = _acos_linear_approximation(x[x_upper], upper)
    acos_extrap[x_lower] = _acos_linear_approximation(x[x_lower], lower)
    return acos_extrap


def _acos_linear_approximation(x, x0):
    return (x - x0) * (-1.0 / math.sqrt(1.0 - x0 * x0)) + math.acos(x0)


def so3_rotation_angle(R, eps=1e-4, cos_angle=False, cos_bound=1e-4):
    N, dim1, dim2 = R.shape
    if dim1 != 3 or dim2 != 3:
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    rot_trace = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    if ((rot_trace < -1.0 - eps) + (rot_trace > 3.0 + eps)).any():
        raise ValueError("A matrix has trace outside valid range [-1-eps,3+eps].")
    phi_cos = (rot_trace - 1.0) * 0.5
    if cos_angle:
        return phi_cos
    if cos_bound > 0.0:
        return acos_linear_extrapolation(phi_cos, (-1.0 + cos_bound, 1.0 - cos_bound))
    return torch.acos(phi_cos)


def hat_inv(h):
    N, dim1, dim2 = h.shape
    if dim1 != 3 or dim2 != 3:
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    ss_diff = torch.abs(h + h.permute(0, 2, 1)).max()
    if float(ss_diff) > 1e-5:
        raise ValueError("One of input matrices is not skew-symmetric.")
    return torch.stack((h[:, 2, 1], h[:, 0, 2], h[:, 1, 0]), dim=1)


def so3_log_map(R, eps=0.0001, cos_bound=1e-4):
    N, dim1, dim2 = R.shape
    if dim1 != 3 or dim2 != 3:
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    phi = so3_rotation_angle(R, cos_bound=cos_bound, eps=eps)
    phi_sin = torch.sin(phi)
    phi_factor = torch.empty_like(phi)
    ok_denom = phi_sin.abs() > (0.5 * eps)
    phi_factor[~ok_denom] = 0.5 + (phi[~ok_denom] ** 2) * (1.0 / 12)
    phi_factor[ok_denom] = phi[ok_denom] / (2.0 * phi_sin[ok_denom])
    log_rot_hat = phi_factor[:, None, None] * (R - R.permute(0, 2, 1))
    return hat_inv(log_rot_hat)


class Rotate:
    """transform3d.py: Rotate(R) holds the 4x4 matrix with R in its upper-left block; transform_points multiplies ROW vectors from the
    left (points_h @ M) and divides by the homogeneous coordinate."""

    def __init__(self, R, dtype=torch.float32, device=None, orthogonal_tol=1e-5):
        R = torch.as_tensor(R, dtype=dtype, device=device)
        if R.dim() == 2:
            R = R[None]
        if R.shape[-2:] != (3, 3):
            raise ValueError("R must have shape (3, 3) or (N, 3, 3)")
        mat = torch.eye(4, dtype=dtype, device=device).view(1, 4, 4).repeat(R.shape[0], 1, 1)
        mat[:, :3, :3] = R
        self._matrix = mat

    def get_matrix(self):
        return self._matrix

    def transform_points(self, points, eps=None):
        points_batch = points.clone()
        if points_batch.dim() == 2:
            points_batch = points_batch[None]
        if points_batch.dim() != 3:
            raise ValueError("Expected points to have dim = 2 or dim = 3")
        N, P, _3 = points_batch.shape
        ones = torch.ones(N, P, 1, dtype=points.dtype, device=points.device)
        points_batch = torch.cat([points_batch, ones], dim=2)
        points_out = torch.bmm(points_batch, self._matrix.expand(N, 4, 4))
        denom = points_out[..., 3:]
        if eps is not None:
            denom = denom.sign() * torch.clamp(denom.abs(), eps)
        points_out = points_out[..., :3] / denom
        if points_out.shape[0] == 1 and points.dim() == 2:
            points_out = points_out.reshape(points.shape)
        return points_out


def _axis_angle_rotation(axis, angle):
    cos, sin = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, cos, -sin, zero, sin, cos)
    elif axis == "Y":
        flat = (cos, zero, sin, zero, one, zero, -sin, zero, cos)
    elif axis == "Z":
        flat = (cos, -sin, zero, sin, cos, zero, zero, zero, one)
    else:
        raise ValueError("letter must be either X, Y or Z.")
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    """rotation_conversions.py: product of the three axis rotations in the order of `convention` (used by the Hopper's yaw randomisation,
    hopper_trajectory.py:344)."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    matrices = [_axis_angle_rotation(c, e) for c, e in zip(convention, torch.unbind(euler_angles, -1))]
    return torch.matmul(torch.matmul(matrices[0], matrices[1]), matrices[2])


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    positive_mask = x > 0
    ret[positive_mask] = torch.sqrt(x[positive_mask])
    return ret


def matrix_to_quaternion(matrix):
    """rotation_conversions.py (v0.7.x): four candidate quaternions, the one with the largest denominator wins; real part first.  Later
    releases standardise the sign (w >= 0), earlier ones do not — irrelevant where the reference uses it: the result is only ever the right
    factor of a quaternion_multiply, which standardises its own product."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    out = quat_candidates[torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch_dim + (4,))
    return standardize_quaternion(out)


def __getattr__(name):
    """Anything else a module of the reference may import from pytorch3d.transforms: importable, loud when used."""
    if name.startswith("__"):
        raise AttributeError(name)

    def _absent(*a, **k):
        raise NotImplementedError(f"pytorch3d.transforms.{name} is not restated (not on the Hopper torque path)")
    return _absent
