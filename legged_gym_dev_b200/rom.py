"""Drop-in mirrors of the reference's ROM classes whose per-step work runs in the fused kernels (csrc/rom.cu).

Reference surface kept (SURVEY.md §1, §8b iii):
  RomDynamics.{f, proj_z, des_pose_vel, clip_v, clip_v_z, compute_state_dependent_input_bounds}  trajopt/rom_dynamics.py:74-118
  SingleInt2D :182, DoubleInt2D :214
  TrajectoryGenerator.{reset, reset_idx, step, step_idx, get_trajectory, get_v_trajectory} + attributes
      k, t, v, trajectory, v_trajectory, weights, t_final, ...                              rom_dynamics.py:441-615
  CustomSim.{step, reset, reset_idx, get_observations, get_state} + root_states, trajectory, traj_gen, rom, model
                                                                                              deep_tube_learning/custom_sim.py:5-103
  DoubleSingleTracking(Kp, Kd, state_dependent_input_bound)(obs)                             deep_tube_learning/controllers.py:80-92
plus `collect_epoch`, the loop body of deep_tube_learning/data_collection_trajectory.py:104-149 as one launch.
The small algebraic helpers (f / proj_z / clip_v_z ...) of the two integrator classes stay torch expressions on the device — they
are API, not the hot loop; the hot loop (`step`, `reset`, the controller, the epoch rollout) goes through the C ABI.
The unicycle family (Unicycle :263, LateralUnicycle :307, ExtendedUnicycle :336, ExtendedLateralUnicycle :397 — SURVEY 8f row 4)
runs entirely through the generic kernels of csrc/rom_family.cu: its algebra (one launch per call, incl. proj_z, which the
reference evaluates with scipy on the host) and a stand-alone TrajectoryGenerator over any rom class
(`reset(z)`, `reset_idx(idx, z)`, `step()`, `step_idx(idx)`, `get_input_t(t, z)`).
"""
import ctypes as C
import math

import os

import torch

from . import _lib

SINGLE_INT_2D, DOUBLE_INT_2D, UNICYCLE, LATERAL_UNICYCLE, EXTENDED_UNICYCLE, EXTENDED_LATERAL_UNICYCLE = range(6)


class RomDynamics:
    n: int
    m: int

    def __init__(self, dt, z_min, z_max, v_min, v_max, n_robots=1, backend="torch", device="cuda"):
        if backend != "torch":
            raise ValueError("the b200gym ROM classes implement the reference's torch backend only")
        self.dt, self.n_robots, self.device = dt, n_robots, device
        t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=device)
        self.z_min, self.z_max, self.v_min, self.v_max = t(z_min), t(z_max), t(v_min), t(v_max)
        self.vel_inds = None

    def clip_v(self, v):
        return torch.max(torch.min(v, self.v_max), self.v_min)

    def compute_state_dependent_input_bounds(self, z):
        return (torch.repeat_interleave(self.v_min[None, :], z.shape[0], dim=0),
                torch.repeat_interleave(self.v_max[None, :], z.shape[0], dim=0))

    def get_weighting_vector(self, reward_weighting):
        raise NotImplementedError


class SingleInt2D(RomDynamics):
    n, m, kind = 2, 2, SINGLE_INT_2D

    def __init__(self, dt, z_min, z_max, v_min, v_max, n_robots=1, backend="torch", device="cuda"):
        super().__init__(dt, z_min, z_max, v_min, v_max, n_robots, backend, device)
        self.A = torch.tensor([[1.0, 0], [0, 1.0]], device=device)
        self.B = torch.tensor([[dt, 0], [0, dt]], device=device)
        self.vel_inds = torch.tensor([False, False], device=device)

    def f(self, x, u):
        return (self.A @ x.T).T + (self.B @ u.T).T

    def proj_z(self, x):
        return x[..., :2]

    def des_pose_vel(self, z, v):
        return (torch.hstack((z, torch.arctan2(v[:, 1], v[:, 0])[:, None])),
                torch.hstack((v, torch.zeros((v.shape[0], 1), device=self.device))))

    def clip_v_z(self, z, v):
        return v

    def get_weighting_vector(self, rw):
        return torch.tensor([rw.position, rw.position], dtype=torch.float32, device=self.device)


class DoubleInt2D(RomDynamics):
    n, m, kind = 4, 2, DOUBLE_INT_2D

    def __init__(self, dt, z_min, z_max, v_min, v_max, n_robots=1, backend="torch", device="cuda"):
        super().__init__(dt, z_min, z_max, v_min, v_max, n_robots, backend, device)
        self.A = torch.tensor([[1.0, 0, dt, 0], [0, 1.0, 0, dt], [0, 0, 1.0, 0], [0, 0, 0, 1.0]], device=device)
        self.B = torch.tensor([[0, 0], [0, 0], [dt, 0], [0, dt]], device=device)
        self.vel_inds = torch.tensor([False, False, True, True], device=device)

    def f(self, x, u):
        return (self.A @ x.T).T + (self.B @ u.T).T

    def proj_z(self, x):
        return torch.hstack((x[..., :2], x[..., 7:9]))

    def des_pose_vel(self, z, v):
        return (torch.hstack((z[:, :2], torch.arctan2(z[:, 3], z[:, 2])[:, None])),
                torch.hstack((z[:, 2:], torch.zeros((v.shape[0], 1), device=self.device))))

    def compute_state_dependent_input_bounds(self, z):
        v_max_z = torch.min(self.v_max, (self.z_max[2:] - z[:, 2:]) / self.dt)
        v_min_z = torch.max(self.v_min, (self.z_min[2:] - z[:, 2:]) / self.dt)
        return v_min_z, v_max_z

    def clip_v_z(self, z, v):
        lo, hi = self.compute_state_dependent_input_bounds(z)
        return torch.max(torch.min(v, hi), lo)

    def get_weighting_vector(self, rw):
        return torch.tensor([rw.position, rw.position, rw.velocity, rw.velocity], dtype=torch.float32, device=self.device)


def _pad(v, n):
    v = [float(x) for x in v]
    return v + [0.0] * (n - len(v))


def _family_pod(rom):
    p = _lib.RomFamilyParamsPOD()
    p.num_envs, p.rom_type, p.window, p.dN = rom.n_robots, rom.kind, 2, 1
    p.rom_dt, p.dt_loop = rom.dt, rom.dt
    p.z_min[:], p.z_max[:] = _pad(rom.z_min.tolist(), 8), _pad(rom.z_max.tolist(), 8)
    p.v_min[:], p.v_max[:] = _pad(rom.v_min.tolist(), 4), _pad(rom.v_max.tolist(), 4)
    return p


class Unicycle(RomDynamics):
    """rom_dynamics.py:263-305; f / des_pose_vel / proj_z / clip_v_z are single launches of the b200gym_romfam_* entry points.
    Unlike the reference (whose f() fills a [n_robots, n] buffer, :270) any number of rows is accepted."""
    n, m, kind = 3, 2, UNICYCLE
    _vel_inds = (False, False, False)

    def __init__(self, dt, z_min, z_max, v_min, v_max, n_robots=1, backend="torch", device="cuda"):
        super().__init__(dt, z_min, z_max, v_min, v_max, n_robots, backend, device)
        if torch.device(device).type != "cuda":
            raise RuntimeError("the b200gym unicycle-family ROMs run on CUDA devices only (no CPU fallback)")
        if len(self.z_min) != self.n or len(self.z_max) != self.n or len(self.v_min) != self.m or len(self.v_max) != self.m:
            raise ValueError(f"{type(self).__name__}: z bounds need {self.n} entries, v bounds {self.m}")
        self.vel_inds = torch.tensor(self._vel_inds, device=device)
        self._fp = _family_pod(self)
        self.lib = _lib.lib()

    def _rows(self, t, width, name):
        _lib.require_cuda(t, name)
        if t.dim() != 2 or t.shape[1] != width or t.dtype != torch.float32:
            raise ValueError(f"{type(self).__name__}: `{name}` must be a float32 [rows, {width}] tensor (got {tuple(t.shape)}, {t.dtype})")
        return t.contiguous()

    def f(self, x, u):
        x, u = self._rows(x, self.n, "x"), self._rows(u, self.m, "u")
        out = torch.empty_like(x)
        if x.shape[0]:
            _lib.check(self.lib.b200gym_romfam_f(self.kind, self.dt, _lib.ptr(x), _lib.ptr(u), _lib.ptr(out), x.shape[0],
                                                 _lib.stream_ptr(x.device)), "romfam_f")
        return out

    def des_pose_vel(self, z, v):
        z, v = self._rows(z, self.n, "z"), self._rows(v, self.m, "v")
        pose, vel = torch.empty(z.shape[0], 3, device=z.device), torch.empty(z.shape[0], 3, device=z.device)
        if z.shape[0]:
            _lib.check(self.lib.b200gym_romfam_des_pose_vel(self.kind, _lib.ptr(z), _lib.ptr(v), _lib.ptr(pose), _lib.ptr(vel), z.shape[0],
                                                            _lib.stream_ptr(z.device)), "romfam_des_pose_vel")
        return pose, vel

    def proj_z(self, x):
        x = self._rows(x, 13, "x")
        z = torch.empty(x.shape[0], self.n, device=x.device)
        if x.shape[0]:
            _lib.check(self.lib.b200gym_romfam_proj_z(self.kind, _lib.ptr(x), _lib.ptr(z), x.shape[0], _lib.stream_ptr(x.device)), "romfam_proj_z")
        return z

    def compute_state_dependent_input_bounds(self, z):
        z = self._rows(z, self.n, "z")
        lo, hi = torch.empty(z.shape[0], self.m, device=z.device), torch.empty(z.shape[0], self.m, device=z.device)
        if z.shape[0]:
            _lib.check(self.lib.b200gym_romfam_input_bounds(self._fp, _lib.ptr(z), None, _lib.ptr(lo), _lib.ptr(hi), None, z.shape[0],
                                                            _lib.stream_ptr(z.device)), "romfam_input_bounds")
        return lo, hi

    def clip_v_z(self, z, v):
        z, v = self._rows(z, self.n, "z"), self._rows(v, self.m, "v")
        out = torch.empty_like(v)
        if z.shape[0]:
            _lib.check(self.lib.b200gym_romfam_input_bounds(self._fp, _lib.ptr(z), _lib.ptr(v), None, None, _lib.ptr(out), z.shape[0],
                                                            _lib.stream_ptr(z.device)), "romfam_input_bounds")
        return out

    def get_weighting_vector(self, rw):                                   # :302-304
        return torch.tensor([rw.position, rw.position, rw.orientation], dtype=torch.float32, device=self.device)


class LateralUnicycle(Unicycle):
    """rom_dynamics.py:307-333 (des_pose_vel keeps the reference's om = v[:, 1], :321)."""
    n, m, kind = 3, 3, LATERAL_UNICYCLE

    def get_weighting_vector(self, rw):                                   # :329-333
        return torch.tensor([rw.position, rw.position, rw.orientation, rw.velocity, rw.velocity, rw.angular_velocity], dtype=torch.float32,
                            device=self.device)


class ExtendedUnicycle(Unicycle):
    """rom_dynamics.py:336-394."""
    n, m, kind = 5, 2, EXTENDED_UNICYCLE
    _vel_inds = (False, False, False, True, True)

    def get_weighting_vector(self, rw):                                   # :390-394
        return torch.tensor([rw.position, rw.position, rw.orientation, rw.velocity, rw.angular_velocity], dtype=torch.float32,
                            device=self.device)


class ExtendedLateralUnicycle(ExtendedUnicycle):
    """rom_dynamics.py:397-438.  proj_z: the reference's (:422-427) raises on both of its array backends (torch.squeeze of a numpy
    array); built as ExtendedUnicycle's with both local velocity components, the evident intent."""
    n, m, kind = 6, 3, EXTENDED_LATERAL_UNICYCLE
    _vel_inds = (False, False, False, True, True, True)

    def get_weighting_vector(self, rw):                                   # :434-438
        return torch.tensor([rw.position, rw.position, rw.orientation, rw.velocity, rw.velocity, rw.angular_velocity], dtype=torch.float32,
                            device=self.device)


ROM_CLASSES = {c.__name__: c for c in (SingleInt2D, DoubleInt2D, Unicycle, LateralUnicycle, ExtendedUnicycle, ExtendedLateralUnicycle)}


class UniformSampleHoldDT:
    """utils.py:27-43: only the bounds matter to the fused generator (the draw itself happens in-kernel)."""

    def __init__(self, t_low, t_high, seed=42, backend="torch", device="cuda"):
        self.t_low, self.t_high = t_low, t_high


class UniformWeightSampler:            # utils.py:46-54
    zero_col = -1

    def __init__(self, dim=4, seed=42, device="cuda"):
        self.dim = dim


class UniformWeightSamplerNoExtreme(UniformWeightSampler):    # utils.py:57-67
    zero_col = 2


class UniformWeightSamplerNoRamp(UniformWeightSampler):       # utils.py:70-79
    zero_col = 1


def _pad4(v):
    v = [float(x) for x in v]
    return v + [0.0] * (4 - len(v))


class TrajectoryGenerator:
    """rom_dynamics.py:441-615 with the torch backend; state tensors have the reference's names and shapes."""
    kind = 0   # B200GYM_GEN_RANDOM; the deterministic subclasses below override it

    def __init__(self, rom, t_sampler, weight_sampler, dt_loop=0.02, N=4, freq_low=0.01, freq_high=10, seed=42,
                 backend="torch", device="cuda", prob_stationary=.01, dN=1, env_id_offset=0, model=None, generic_kernels=None):
        if backend != "torch":
            raise ValueError("backend must be 'torch'")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the b200gym TrajectoryGenerator runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.rom, self.dt_loop, self.N, self.dN = rom, dt_loop, N, dN
        self.t_sampler, self.weight_sampler = t_sampler, weight_sampler
        self.freq_low, self.freq_high, self.prob_stationary, self.seed = freq_low, freq_high, prob_stationary, int(seed)
        self.env_id_offset = int(env_id_offset)
        n, W, dev = rom.n_robots, N * dN, self.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.weights, self.t_final, self.t, self.k = z(n, 4), z(n), z(n), z(n)
        self.sample_hold_input, self.extreme_input = z(n, rom.m), z(n, rom.m)
        self.ramp_t_start, self.ramp_v_start, self.ramp_v_end = z(n), z(n, rom.m), z(n, rom.m)
        self.sin_mag, self.sin_freq, self.sin_off, self.sin_mean = z(n, rom.m), z(n, rom.m), z(n, rom.m), z(n, rom.m)
        self.trajectory = z(n, W + 1, rom.n)
        self.v_trajectory = z(n, W, rom.m)
        self.v = z(n, rom.m)
        self.stationary_inds = torch.zeros(n, dtype=torch.bool, device=dev)
        self.rng_ctr = torch.zeros(n, dtype=torch.int32, device=dev)
        if self.kind != 0 and rom.kind != SINGLE_INT_2D:
            raise ValueError("Only SingleInt2D is fused for the Zero / Square / Circle generators")
        self.center = z(n, 2) if self.kind == 3 else None
        self._model = model
        self._sim = None
        # the unicycle-family classes are served only by the generic kernels (csrc/rom_family.cu).  For the integrator classes both kernel
        # sets exist and agree bit for bit (tests/test_rom_family_gpu.py); `generic_kernels` picks the one behind step / step_idx (reset(z) /
        # get_input_t always use the generic ones): None = the generic, window-staging kernels for a stand-alone random generator (measured
        # 1.4-1.6x the register-resident ones at 1 M envs, profiles/r1_rom_family_hopper.md), the register-resident ones of csrc/rom.cu as soon
        # as a model / CustomSim / trajectory env is attached (they also refresh the env's interpolated trajectory and observation views).
        if generic_kernels is None:
            generic_kernels = self.kind == 0 and model is None
        self._family = rom.kind > DOUBLE_INT_2D or bool(generic_kernels)
        if self._family and (self.kind != 0 or model is not None):
            raise ValueError("the generic rom-family kernels serve the stand-alone random TrajectoryGenerator only")
        self._build_family_pod()
        if rom.kind > DOUBLE_INT_2D:
            self._p = None
            _lib.check(self.lib.b200gym_romfam_gen_init(self._fp, self._s, self.env_id_offset, _lib.stream_ptr(dev)), "romfam_gen_init")
            return
        self._build_pod()      # the integrator classes always carry both parameter blocks; `_family` only selects the step kernels
        _lib.check(self.lib.b200gym_rom_init(self._p, self._s, self.env_id_offset, _lib.stream_ptr(dev)), "rom_init")

    # ---- POD structs ----------------------------------------------------------------------------
    def _state_pod(self):
        s = _lib.RomStatePOD()
        for name in ("trajectory", "v_trajectory", "v", "t", "k", "t_final", "weights", "sample_hold_input", "extreme_input",
                     "ramp_v_start", "ramp_v_end", "ramp_t_start", "sin_mag", "sin_freq", "sin_off", "sin_mean",
                     "stationary_inds", "rng_ctr"):
            t = getattr(self, name)
            _lib.require_cuda(t, name)
            setattr(s, name, t.data_ptr())
        return s

    def _build_family_pod(self):
        """B200RomFamilyParams of the generic generator kernels (any rom class; the stand-alone reset(z) / get_input_t path)."""
        rom = self.rom
        p = _family_pod(rom)
        p.window, p.dN, p.dt_loop = self.N * self.dN, self.dN, self.dt_loop
        p.t_low, p.t_span = self.t_sampler.t_low, self.t_sampler.t_high - self.t_sampler.t_low
        p.freq_low, p.freq_high, p.prob_stationary = self.freq_low, self.freq_high, self.prob_stationary
        p.weight_zero_col = getattr(self.weight_sampler, "zero_col", -1)
        p.seed_lo, p.seed_hi = self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF
        self._fp = p
        if rom.kind > DOUBLE_INT_2D:
            self._s = self._state_pod()

    def _build_pod(self, sim=None):
        rom, model = self.rom, self._model
        if sim is not None and rom.kind <= DOUBLE_INT_2D:
            self._family = False      # an attached env needs its views refreshed by rom_step
        p = _lib.RomParamsPOD()
        p.num_envs, p.rom_type = rom.n_robots, rom.kind
        p.model_type = model.kind if model is not None else rom.kind
        p.window, p.dN, p.horizon = self.N * self.dN, self.dN, self.N
        p.rom_dt, p.dt_loop = rom.dt, self.dt_loop
        p.model_dt = model.dt if model is not None else self.dt_loop
        p.rom_z_min[:], p.rom_z_max[:] = _pad4(rom.z_min.tolist()), _pad4(rom.z_max.tolist())
        p.rom_v_min[:], p.rom_v_max[:] = rom.v_min.tolist(), rom.v_max.tolist()
        if model is not None:
            p.model_z_min[:], p.model_z_max[:] = _pad4(model.z_min.tolist()), _pad4(model.z_max.tolist())
            p.model_v_min[:], p.model_v_max[:] = model.v_min.tolist(), model.v_max.tolist()
        p.t_low, p.t_span = self.t_sampler.t_low, self.t_sampler.t_high - self.t_sampler.t_low
        p.freq_low, p.freq_high, p.prob_stationary = self.freq_low, self.freq_high, self.prob_stationary
        p.weight_zero_col = getattr(self.weight_sampler, "zero_col", -1)
        p.seed_lo, p.seed_hi = self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF
        p.gen_kind = self.kind
        if self.kind == 2:      # rom_dynamics.py:633-640, evaluated in fp32 tensors exactly as written there
            vmax, vmin = rom.v_max.cpu(), rom.v_min.cpu()
            c1 = 2 / vmax[1]
            c2 = c1 + 1 / vmax[0]
            c3 = c2 + 2 / abs(vmin[1])
            c4 = c3 + 1 / abs(vmin[0])
            p.gen_c[:] = [float(c1), float(c2), float(c3), float(c4)]
            p.gen_v[:] = [float(vmax[1] / 2), float(vmax[0]), float(vmin[1] / 2), float(vmin[1])]
        elif self.kind == 3:    # :692
            p.gen_v[0] = float(torch.min(torch.minimum(rom.v_max, torch.abs(rom.v_min))))
        if sim is not None:
            p.randomize_rom_distance = int(sim.randomize_rom_distance)
            p.max_rom_distance[:] = _pad4(sim.max_rom_distance.tolist())
            p.zero_rom_dist_llh = sim.zero_rom_dist_llh
            p.noise_lower[:], p.noise_upper[:] = _pad4(sim.root_state_noise_lower.tolist()), _pad4(sim.root_state_noise_upper.tolist())
            p.Kp, p.Kd = sim.Kp, sim.Kd
        s = self._state_pod()
        if self.center is not None:
            s.center = self.center.data_ptr()
        if sim is not None:
            s.root_states, s.env_trajectory, s.obs = sim.root_states.data_ptr(), sim.trajectory.data_ptr(), sim.obs_buf.data_ptr()
        self._p, self._s, self._sim = p, s, sim

    def _mask_ptr(self, idx):
        """idx -> uint8 mask (NULL when idx covers every env, the reference's arange)."""
        n = self.rom.n_robots
        if idx is None:
            return None, None
        if idx.dtype == torch.bool:
            m = idx.to(torch.uint8)
        else:
            if idx.numel() == n:
                return None, None
            m = torch.zeros(n, dtype=torch.uint8, device=self.device)
            m[idx] = 1
        return m, C.c_void_p(m.data_ptr())

    # ---- reference API --------------------------------------------------------------------------
    def step(self):
        self.step_idx(None)

    def step_idx(self, idx):
        m, mp = self._mask_ptr(idx)
        # a CustomSim's observation view is refreshed by b200gym_rom_step only; a trajectory env's interpolated window (env_trajectory) is
        # written by the window-staging kernel as well (one bulk store per 128 envs)
        if self._family and not self._s.obs:
            _lib.check(self.lib.b200gym_romfam_gen_step(self._fp, self._s, mp, self.env_id_offset, _lib.stream_ptr(self.device)),
                       "romfam_gen_step")
            return
        _lib.check(self.lib.b200gym_rom_step(self._p, self._s, None, mp, self.env_id_offset, _lib.stream_ptr(self.device)), "rom_step")

    def env_step(self, stream):
        """The generator step of an attached trajectory env (legged_robot_trajectory.py:409-410): `traj_gen.step()` plus the refresh of the
        env's interpolated window.  The random generator goes through the window-staging kernel (csrc/rom_family.cu: the state window
        and the env's view move as bulk tiles, bit-identical to b200gym_rom_step); the Zero / Square / Circle generators and
        B200GYM_ENV_GEN_TILE=0 through b200gym_rom_step.  Returns the C-ABI status."""
        if self.kind == 0 and not self._s.obs and os.environ.get("B200GYM_ENV_GEN_TILE", "1") != "0":
            return self.lib.b200gym_romfam_gen_step(self._fp, self._s, None, self.env_id_offset, stream)
        return self.lib.b200gym_rom_step(self._p, self._s, None, None, self.env_id_offset, stream)

    def get_trajectory(self):                                             # rom_dynamics.py:607-612
        t0, t1 = self.trajectory[:, :-1, :], self.trajectory[:, 1:, :]
        interp = t0 + (t1 - t0) * (self.t - (self.k - 1) * self.rom.dt)[:, None, None] / self.rom.dt
        return interp[:, ::self.dN, :]

    def get_v_trajectory(self):
        return self.v_trajectory[:, ::self.dN, :]

    def _z_rows(self, z):
        _lib.require_cuda(z, "z")
        if tuple(z.shape) != (self.rom.n_robots, self.rom.n) or z.dtype != torch.float32:
            raise ValueError(f"z must be a float32 [{self.rom.n_robots}, {self.rom.n}] tensor (got {tuple(z.shape)}, {z.dtype})")
        return z.contiguous()

    def reset(self, z):                                                   # rom_dynamics.py:592-593
        self.reset_idx(None, z)

    def reset_idx(self, idx, z):
        """rom_dynamics.py:595-605 for the random generator over any rom class (one launch: the W warm-up knots are written in place).
        The Zero / Square / Circle generators are reset through the trajectory env / CustomSim paths only."""
        if self.kind != 0:
            raise NotImplementedError("stand-alone reset of the Zero / Square / Circle generators is driven through the env's reset path")
        z = self._z_rows(z)
        m, mp = self._mask_ptr(idx)
        _lib.check(self.lib.b200gym_romfam_gen_reset(self._fp, self._s, _lib.ptr(z), mp, self.env_id_offset, _lib.stream_ptr(self.device)),
                   "romfam_gen_reset")

    def get_input_t(self, t, z):
        """rom_dynamics.py:560-566 with caller-supplied clock(s) `t` (scalar or [n_robots]) and states z: resamples the envs with
        t > t_final, returns the weighted, clipped input [n_robots, m] (the open-loop use of trajopt/trajectory_gen.py:35-41)."""
        if self.kind != 0:
            raise NotImplementedError("get_input_t is fused for the random TrajectoryGenerator")
        z = self._z_rows(z)
        n = self.rom.n_robots
        tt = torch.as_tensor(t, dtype=torch.float32, device=self.device).expand(n).contiguous()
        v = torch.empty(n, self.rom.m, device=self.device)
        _lib.check(self.lib.b200gym_romfam_gen_input(self._fp, self._s, _lib.ptr(tt), _lib.ptr(z), _lib.ptr(v), self.env_id_offset,
                                                     _lib.stream_ptr(self.device)), "romfam_gen_input")
        return v


class ZeroTrajectoryGenerator(TrajectoryGenerator):
    """rom_dynamics.py:618-624: zero input, reset envs become stationary."""
    kind = 1


class SquareTrajectoryGenerator(TrajectoryGenerator):
    """rom_dynamics.py:627-675 (SingleInt2D legs :630-640; the leg boundaries are computed here in fp32, as the reference does)."""
    kind = 2


class CircleTrajectoryGenerator(TrajectoryGenerator):
    """rom_dynamics.py:678-698 (SingleInt2D :686-692); `center` is re-taken from z for EVERY env whenever any env resets (:680-683)."""
    kind = 3


class CustomSim:
    """deep_tube_learning/custom_sim.py:5-103.  cfg has the reference's shape (env.model / rom / trajectory_generator /
    domain_rand / init_state); `cfg.controller.{Kp,Kd}` (optional) parameterises the fused tracking controller."""

    def __init__(self, cfg, device=None, env_id_offset=0):
        self.cfg = cfg
        self.dt = cfg.env.model.dt
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("the b200gym CustomSim runs on CUDA devices only (no CPU fallback)")
        self.num_envs = cfg.env.num_envs
        self.env_id_offset = int(env_id_offset)
        dev = self.device
        classes = {"SingleInt2D": SingleInt2D, "DoubleInt2D": DoubleInt2D}
        mc, rc = cfg.env.model, cfg.rom
        self.model = classes[mc.cls](dt=self.dt, z_min=mc.z_min, z_max=mc.z_max, v_min=mc.v_min, v_max=mc.v_max,
                                     n_robots=self.num_envs, backend="torch", device=dev)
        self.rom = classes[rc.cls](dt=rc.dt, z_min=rc.z_min, z_max=rc.z_max, v_min=rc.v_min, v_max=rc.v_max,
                                   n_robots=self.num_envs, backend="torch", device=dev)
        tc = cfg.trajectory_generator
        samplers = {"UniformWeightSampler": UniformWeightSampler, "UniformWeightSamplerNoExtreme": UniformWeightSamplerNoExtreme,
                    "UniformWeightSamplerNoRamp": UniformWeightSamplerNoRamp}
        if tc.cls != "TrajectoryGenerator":
            raise NotImplementedError(f"{tc.cls}: only the random TrajectoryGenerator is fused (SURVEY.md §8f-4)")
        self.traj_gen = TrajectoryGenerator(self.rom, UniformSampleHoldDT(tc.t_low, tc.t_high), samplers[tc.weight_samp_cls](),
                                            dt_loop=self.dt, N=tc.N, freq_low=tc.freq_low, freq_high=tc.freq_high, seed=tc.seed,
                                            backend="torch", device=dev, prob_stationary=tc.prob_stationary, dN=tc.dN,
                                            env_id_offset=env_id_offset, model=self.model)
        self.root_states = torch.zeros((self.num_envs, self.model.n), device=dev)
        self.trajectory = torch.zeros(self.num_envs, tc.N, self.rom.n, dtype=torch.float, device=dev)
        self.obs_buf = torch.zeros(self.num_envs, self.model.n + self.rom.n + self.rom.m, device=dev)
        self._dones = torch.zeros(self.num_envs, dtype=torch.bool, device=dev)
        self.max_rom_distance = torch.tensor(cfg.domain_rand.max_rom_distance, device=dev, dtype=torch.float32)
        self.zero_rom_dist_llh = cfg.domain_rand.zero_rom_dist_llh
        self.randomize_rom_distance = cfg.domain_rand.randomize_rom_distance
        self.root_state_noise_lower = torch.tensor(cfg.init_state.default_noise_lower, device=dev, dtype=torch.float32)
        self.root_state_noise_upper = torch.tensor(cfg.init_state.default_noise_upper, device=dev, dtype=torch.float32)
        ctl = getattr(cfg, "controller", None)
        self.Kp, self.Kd = (float(ctl.Kp), float(ctl.Kd)) if ctl is not None else (0.0, 0.0)
        self.traj_gen._build_pod(sim=self)
        self.lib = self.traj_gen.lib

    def step(self, action):                                               # custom_sim.py:71-75
        _lib.require_cuda(action, "action")
        g = self.traj_gen
        _lib.check(self.lib.b200gym_rom_step(g._p, g._s, _lib.ptr(action), None, self.env_id_offset, _lib.stream_ptr(self.device)),
                   "rom_step")
        return self.obs_buf, None, None, self._dones, None

    def reset(self):
        self.reset_idx(None)

    def reset_idx(self, idx):                                             # custom_sim.py:87-93
        g = self.traj_gen
        m, mp = g._mask_ptr(idx)
        _lib.check(self.lib.b200gym_rom_reset(g._p, g._s, mp, self.env_id_offset, _lib.stream_ptr(self.device)), "rom_reset")

    def get_observations(self):                                           # custom_sim.py:95-100
        return self.obs_buf

    def get_state(self):
        return torch.clone(self.root_states.detach())

    def collect_epoch(self, obs, T, save_debugging_data=False):
        """One epoch of data_collection_trajectory.py:102-149: env.reset() then T ROM steps of
        {actions = policy(obs); obs = env.step(actions)}, logging z, v, pz_x, done (and x).  `obs` is the observation
        the first action is computed from (the reference never refreshes it after reset, :94,:111); it is updated in
        place.  Returns the epoch_data dict of device tensors."""
        g, N, dev = self.traj_gen, self.num_envs, self.device
        _lib.require_cuda(obs, "obs")
        z = torch.empty(N, T + 1, self.rom.n, device=dev)
        pz_x = torch.empty(N, T + 1, self.rom.n, device=dev)
        v = torch.empty(N, T, self.rom.m, device=dev)
        done = torch.empty(N, T, dtype=torch.bool, device=dev)
        x = torch.empty(N, T + 1, self.model.n, device=dev) if save_debugging_data else None
        _lib.check(self.lib.b200gym_rom_rollout(g._p, g._s, _lib.ptr(obs), T, _lib.ptr(x), _lib.ptr(z), _lib.ptr(pz_x), _lib.ptr(v),
                                                _lib.ptr(done), self.env_id_offset, _lib.stream_ptr(dev)), "rom_rollout")
        out = dict(z=z, v=v, pz_x=pz_x, done=done)
        if x is not None:
            out["x"] = x
        return out


class DoubleSingleTracking:
    """controllers.py:80-92.  `state_dependent_input_bound` is the bound method `env.model.clip_v_z`
    (data_collection_trajectory.py:90); the fused controller reads the model's bounds from it."""

    def __init__(self, Kp, Kd, state_dependent_input_bound):
        self.K_p, self.K_d = Kp, Kd
        self.state_dependent_input_bound = state_dependent_input_bound
        model = getattr(state_dependent_input_bound, "__self__", None)
        if not isinstance(model, DoubleInt2D):
            raise TypeError("DoubleSingleTracking needs DoubleInt2D.clip_v_z as its input bound")
        self.model = model
        p = _lib.RomParamsPOD()
        p.num_envs, p.model_type, p.model_dt = model.n_robots, DOUBLE_INT_2D, model.dt
        p.model_z_min[:], p.model_z_max[:] = _pad4(model.z_min.tolist()), _pad4(model.z_max.tolist())
        p.model_v_min[:], p.model_v_max[:] = model.v_min.tolist(), model.v_max.tolist()
        p.Kp, p.Kd = float(Kp), float(Kd)
        self._p = p
        self._action = torch.empty(model.n_robots, 2, device=model.device)
        self.lib = _lib.lib()

    def __call__(self, obs):
        _lib.require_cuda(obs, "obs")
        _lib.check(self.lib.b200gym_rom_tracking_policy(self._p, _lib.ptr(obs), _lib.ptr(self._action),
                                                        _lib.stream_ptr(obs.device)), "rom_tracking_policy")
        return self._action


class RaibertHeuristic:
    """deep_tube_learning/controllers.py:4-81 (the hopper's tracking controller): same constructor (`cfg.controller.{K_p, K_v, K_ff,
    clip_value_pos, clip_value_vel, clip_value_total}`), `get_inference_policy(device)` and static `raibert_policy(...)`; the policy
    is one launch of `b200gym_raibert_policy`."""

    def __init__(self, cfg):
        self.cfg = cfg
        c = cfg.controller
        self.K_p, self.K_v, self.K_ff = c.K_p, c.K_v, c.K_ff
        self.clip_value_pos, self.clip_value_vel, self.clip_value_total = c.clip_value_pos, c.clip_value_vel, c.clip_value_total

    def get_inference_policy(self, device):
        def policy(obs):
            return RaibertHeuristic.raibert_policy(obs, self.K_p, self.K_v, self.K_ff, self.clip_value_pos, self.clip_value_vel,
                                                   self.clip_value_total)
        return policy

    @staticmethod
    def raibert_policy(obs, Kp, Kv, K_ff, clip_pos, clip_vel, clip_ang):
        _lib.require_cuda(obs, "obs")
        if obs.dim() != 2 or obs.shape[1] < 10 or obs.dtype != torch.float32:
            raise ValueError("raibert_policy: obs must be a float32 [n, >= 10] tensor")
        out = torch.empty(obs.shape[0], 4, device=obs.device)
        _lib.check(_lib.lib().b200gym_raibert_policy(_lib.ptr(obs), obs.shape[1], obs.shape[0], float(Kp), float(Kv), float(K_ff),
                                                     float(clip_pos), float(clip_vel), float(clip_ang), _lib.ptr(out),
                                                     _lib.stream_ptr(obs.device)), "raibert_policy")
        return out
