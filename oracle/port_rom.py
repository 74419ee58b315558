"""CPU restatement (torch, fp32) of the reference's ROM rollout path.  TEST INFRASTRUCTURE.

Follows trajopt/rom_dynamics.py (SingleInt2D :182-211, DoubleInt2D :214-260, TrajectoryGenerator :441-615),
deep_tube_learning/custom_sim.py:5-103, deep_tube_learning/controllers.py:80-92 and the loop body of
deep_tube_learning/data_collection_trajectory.py:104-149.  Pinned by tests/test_oracle_cpu.py against the
unmodified reference (oracle/ref_harness.py) and by tests/golden/rom_*.npz.

Randomness: `rng="philox"` uses the per-env draw-event stream of oracle/philox.py (event counter `ctr`);
`rng="torch"` calls torch.rand like the reference (CPU-baseline timing mode).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import philox as P


def rom_params(num_envs, **over):
    """Plain numbers of configs/data_generation/double_single_int.yaml:29-87 (+ default_custom.yaml:24-25)."""
    d = dict(num_envs=num_envs, model_cls="DoubleInt2D", rom_cls="SingleInt2D", model_dt=0.05, rom_dt=0.1, pos_max=1e9,
             vel_max=0.3, acc_max=0.5, vel_max_rom=0.2, N=10, dN=1, t_low=1.0, t_high=2.0, freq_low=0.01, freq_high=2.0,
             prob_stationary=0.0005, weight_sampler="UniformWeightSamplerNoRamp", randomize_rom_distance=True,
             max_rom_distance=[1.0, 1.0], zero_rom_dist_llh=0.25, noise_lower=[0.0, 0.0, -0.1, -0.1],
             noise_upper=[0.0, 0.0, 0.1, 0.1], Kp=10.0, Kd=10.0, seed=0, episode_length_s=20, generator="TrajectoryGenerator")
    d.update(over)
    return SimpleNamespace(**d)


# the unicycle family, rom_dynamics.py:263-438: class -> (n, m, first velocity state)
FAMILY = {"Unicycle": (3, 2, 3), "LateralUnicycle": (3, 3, 3), "ExtendedUnicycle": (5, 2, 3), "ExtendedLateralUnicycle": (6, 3, 3)}


class Rom:
    """The RomDynamics classes with the torch backend (rom_dynamics.py:182-438)."""

    def __init__(self, cls, dt, z_min, z_max, v_min, v_max):
        t = lambda v: torch.tensor(v, dtype=torch.float32)
        self.cls, self.dt = cls, dt
        self.z_min, self.z_max, self.v_min, self.v_max = t(z_min), t(z_max), t(v_min), t(v_max)
        self.m = 2
        if cls in FAMILY:
            self.n, self.m, vs = FAMILY[cls]
            self.vs = vs
            self.vel_inds = torch.tensor([c >= vs for c in range(self.n)])
        elif cls == "SingleInt2D":
            self.n = 2
            self.A = t([[1.0, 0], [0, 1.0]])
            self.B = t([[dt, 0], [0, dt]])
            self.vel_inds = torch.tensor([False, False])
        elif cls == "DoubleInt2D":
            self.n = 4
            self.A = t([[1.0, 0, dt, 0], [0, 1.0, 0, dt], [0, 0, 1.0, 0], [0, 0, 0, 1.0]])
            self.B = t([[0, 0], [0, 0], [dt, 0], [0, dt]])
            self.vel_inds = torch.tensor([False, False, True, True])
        else:
            raise ValueError(cls)

    def f(self, x, u):
        if self.cls in FAMILY:
            # Unicycle.f :273-278, LateralUnicycle.f :311-316, ExtendedUnicycle.f :345-352, ExtendedLateralUnicycle.f :406-414
            gu = torch.zeros(x.shape[0], self.n)
            c, s = torch.cos(x[:, 2]), torch.sin(x[:, 2])
            if self.cls == "Unicycle":
                gu[:, 0], gu[:, 1], gu[:, 2] = u[:, 0] * c, u[:, 0] * s, u[:, 1]
            elif self.cls == "LateralUnicycle":
                gu[:, 0] = u[:, 0] * c - u[:, 1] * s
                gu[:, 1] = u[:, 0] * s + u[:, 1] * c
                gu[:, 2] = u[:, 2]
            elif self.cls == "ExtendedUnicycle":
                gu[:, 0], gu[:, 1], gu[:, 2] = x[:, 3] * c, x[:, 3] * s, x[:, 4]
                gu[:, 3], gu[:, 4] = u[:, 0], u[:, 1]
            else:
                gu[:, 0] = x[:, 3] * c - x[:, 4] * s
                gu[:, 1] = x[:, 3] * s + x[:, 4] * c
                gu[:, 2] = x[:, 5]
                gu[:, 3], gu[:, 4], gu[:, 5] = u[:, 0], u[:, 1], u[:, 2]
            return x + self.dt * gu
        return (self.A @ x.T).T + (self.B @ u.T).T

    def des_pose_vel(self, z, v):
        """:198-199, :230-232, :286-290, :318-322 (om = v[:, 1] as written there), :354-358, :416-420."""
        zero = torch.zeros((v.shape[0], 1))
        if self.cls == "SingleInt2D":
            return torch.hstack((z, torch.arctan2(v[:, 1], v[:, 0])[:, None])), torch.hstack((v, zero))
        if self.cls == "DoubleInt2D":
            return torch.hstack((z[:, :2], torch.arctan2(z[:, 3], z[:, 2])[:, None])), torch.hstack((z[:, 2:], zero))
        c, s = torch.cos(z[:, 2]), torch.sin(z[:, 2])
        if self.cls == "Unicycle":
            vx, vy, om = v[:, 0] * c, v[:, 0] * s, v[:, 1]
        elif self.cls == "LateralUnicycle":
            vx, vy, om = v[:, 0] * c - v[:, 1] * s, v[:, 0] * s + v[:, 1] * c, v[:, 1]
        elif self.cls == "ExtendedUnicycle":
            vx, vy, om = z[:, 3] * c, z[:, 3] * s, z[:, 4]
        else:
            vx, vy, om = z[:, 3] * c - z[:, 4] * s, z[:, 3] * s + z[:, 4] * c, z[:, 5]
        return z[:, :3], torch.hstack((vx[:, None], vy[:, None], om[:, None]))

    def proj_z(self, x):
        if self.cls == "SingleInt2D":
            return x[..., :2]
        if self.cls == "DoubleInt2D":
            return torch.hstack((x[..., :2], x[..., 7:9]))
        # :280-284, :360-365, :422-427 — the reference runs these on numpy arrays (scipy Rotation, fp64 euler angles)
        from scipy.spatial.transform import Rotation
        xn = x.numpy() if torch.is_tensor(x) else np.asarray(x)
        yaw = Rotation.from_quat(xn[:, 3:7]).as_euler("xyz", degrees=False)[..., -1]
        cols = [xn[..., :2], yaw[:, None]]
        if self.cls in ("ExtendedUnicycle", "ExtendedLateralUnicycle"):
            y2r = np.zeros((yaw.shape[0], 2, 2))                          # yaw2rot, deep_tube_learning/utils.py:88-96
            y2r[:, 0, 0] = y2r[:, 1, 1] = np.cos(yaw)
            y2r[:, 0, 1], y2r[:, 1, 0] = np.sin(yaw), -np.sin(yaw)
            vl = np.squeeze(y2r @ xn[:, 7:9][:, :, None], axis=-1)
            cols += [vl[:, :1] if self.cls == "ExtendedUnicycle" else vl, xn[:, -1][:, None]]
        return np.hstack(cols)

    def bounds(self, z):
        if self.cls in FAMILY and self.vs < self.n:                       # :367-379
            v_max_z = torch.min(self.v_max, (self.z_max[self.vs:] - z[:, self.vs:]) / self.dt)
            v_min_z = torch.max(self.v_min, (self.z_min[self.vs:] - z[:, self.vs:]) / self.dt)
            return v_min_z, v_max_z
        if self.cls == "SingleInt2D" or self.cls in FAMILY:
            n = z.shape[0]
            return (torch.repeat_interleave(self.v_min[None, :], n, dim=0),
                    torch.repeat_interleave(self.v_max[None, :], n, dim=0))
        v_max_z = torch.min(self.v_max, (self.z_max[2:] - z[:, 2:]) / self.dt)
        v_min_z = torch.max(self.v_min, (self.z_min[2:] - z[:, 2:]) / self.dt)
        return v_min_z, v_max_z

    def clip_v_z(self, z, v):
        if self.cls == "SingleInt2D" or (self.cls in FAMILY and self.vs == self.n):   # :201, :292
            return v
        lo, hi = self.bounds(z)
        return torch.max(torch.min(v, hi), lo)


def make_roms(p):
    pm, vm, am, vr = p.pos_max, p.vel_max, p.acc_max, p.vel_max_rom
    def mk(cls, dt, vmax_state, vmax_in):
        if cls == "SingleInt2D":
            return Rom(cls, dt, [-pm, -pm], [pm, pm], [-vmax_in] * 2, [vmax_in] * 2)
        return Rom(cls, dt, [-pm, -pm, -vmax_state, -vmax_state], [pm, pm, vmax_state, vmax_state], [-vmax_in] * 2, [vmax_in] * 2)
    model = mk(p.model_cls, p.model_dt, vm, am if p.model_cls == "DoubleInt2D" else vm)
    rom = mk(p.rom_cls, p.rom_dt, vm, vr if p.rom_cls == "SingleInt2D" else am)
    return model, rom


class RomPort:
    """CustomSim + TrajectoryGenerator + DoubleSingleTracking in one object; attribute names as in the reference."""

    def __init__(self, p, rng="philox", env_id_offset=0):
        self.p, self.rng, self.off = p, rng, env_id_offset
        N = p.num_envs
        self.N = N
        self.model, self.rom = make_roms(p)
        rom = self.rom
        self.W = p.N * p.dN
        self.ctr = np.zeros(N, dtype=np.int64)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32)
        # TrajectoryGenerator.__init__ (rom_dynamics.py:485-508)
        self.weights, self.t_final, self.t, self.k = z(N, 4), z(N), z(N), z(N)
        self.sample_hold_input, self.extreme_input = z(N, rom.m), z(N, rom.m)
        self.ramp_t_start, self.ramp_v_start = z(N), z(N, rom.m)
        ids = torch.arange(N)
        self.ramp_v_end = self._uniform(rom.v_min, rom.v_max, P.SITE_ROM_INIT, ids, self._events(ids), rom.m)
        self.sin_mag, self.sin_freq, self.sin_off, self.sin_mean = z(N, rom.m), z(N, rom.m), z(N, rom.m), z(N, rom.m)
        self.traj = z(N, self.W + 1, rom.n)
        self.v_traj = z(N, self.W, rom.m)
        self.v = z(N, rom.m)
        self.stationary = torch.zeros(N, dtype=torch.bool)
        # CustomSim.__init__ (custom_sim.py:29-35)
        self.root_states = z(N, self.model.n)
        self.trajectory = z(N, p.N, rom.n)
        self.max_rom_distance = torch.tensor(p.max_rom_distance, dtype=torch.float32)
        self.noise_lower = torch.tensor(p.noise_lower, dtype=torch.float32)[: self.model.n]
        self.noise_upper = torch.tensor(p.noise_upper, dtype=torch.float32)[: self.model.n]

    # ---- randomness ---------------------------------------------------------------------------
    def _events(self, ids):
        ids = ids.numpy()
        ev = self.ctr[ids].copy()
        self.ctr[ids] += 1
        return ev

    def _u(self, site, ids, ev, ncols):
        if self.rng == "torch":
            return torch.rand(len(ids), ncols)
        return torch.from_numpy(P.uniform01(self.p.seed, ids.numpy() + self.off, ev, site, ncols))

    def _uniform(self, lo, hi, site, ids, ev, ncols):
        return (hi - lo) * self._u(site, ids, ev, ncols) + lo

    # ---- TrajectoryGenerator ------------------------------------------------------------------
    def resample(self, idx, z):                                           # rom_dynamics.py:510-545
        kind = getattr(self.p, "generator", "TrajectoryGenerator")
        if kind == "ZeroTrajectoryGenerator":                             # :619-620
            self.stationary[idx] = True
            return
        if kind == "SquareTrajectoryGenerator":                           # :629-630
            return
        if kind == "CircleTrajectoryGenerator":                           # :680-683: every env is re-centred from z
            self.center = z.detach().clone()[:, :2]
            self.center[:, 0] -= 0.5
            return
        if len(idx) == 0:
            return
        p, rom = self.p, self.rom
        ev = self._events(idx)
        v_min, v_max = rom.bounds(z[idx, :])
        self.sample_hold_input[idx, :] = self._uniform(v_min, v_max, P.SITE_ROM_CONST, idx, ev, rom.m)
        self.ramp_v_start[idx, :] = rom.clip_v_z(z[idx, :], self.ramp_v_end[idx, :])
        self.ramp_v_end[idx, :] = self._uniform(v_min, v_max, P.SITE_ROM_RAMP, idx, ev, rom.m)
        self.ramp_t_start[idx] = self.t_final[idx]
        if self.rng == "torch":
            choice = torch.randint(0, 3, (len(idx), rom.m, 1))
        else:
            choice = torch.from_numpy(P.randint(p.seed, idx.numpy() + self.off, ev, P.SITE_ROM_EXTREME, rom.m, 3)).unsqueeze(-1)
        arr = torch.cat((v_min[:, :, None], torch.zeros_like(v_min)[:, :, None], v_max[:, :, None]), -1)
        mask = torch.arange(3)[None, None, :] == choice
        self.extreme_input[idx, :] = arr[mask].reshape(v_min.shape)
        self.sin_mag[idx, :] = self._uniform(torch.zeros_like(v_max), (v_max - v_min) / 2, P.SITE_ROM_SIN_MAG, idx, ev, rom.m)
        self.sin_mean[idx, :] = self._uniform(v_min + self.sin_mag[idx, :], v_max - self.sin_mag[idx, :],
                                              P.SITE_ROM_SIN_MEAN, idx, ev, rom.m)
        fl, fh = torch.tensor([p.freq_low]), torch.tensor([p.freq_high])
        self.sin_freq[idx, :] = self._uniform(fl, fh, P.SITE_ROM_SIN_FREQ, idx, ev, rom.m)
        pi = torch.tensor([math.pi])
        self.sin_off[idx, :] = self._uniform(-pi, pi, P.SITE_ROM_SIN_OFF, idx, ev, rom.m)
        self.t_final[idx] += ((p.t_high - p.t_low) * self._u(P.SITE_ROM_TFINAL, idx, ev, 1) + p.t_low).squeeze(1)
        w = self._u(P.SITE_ROM_WEIGHTS, idx, ev, 4).clone()
        if p.weight_sampler == "UniformWeightSamplerNoRamp":
            w[:, 1] = 0
        elif p.weight_sampler == "UniformWeightSamplerNoExtreme":
            w[:, 2] = 0
        self.weights[idx, :] = w / torch.sum(w, axis=-1, keepdims=True)
        self.stationary[idx] = self._u(P.SITE_ROM_STATIONARY, idx, ev, 1).squeeze(1) < p.prob_stationary

    def get_input_t(self, t, z):                                          # rom_dynamics.py:550-566
        rom = self.rom
        kind = getattr(self.p, "generator", "TrajectoryGenerator")
        if kind != "TrajectoryGenerator":                                 # the subclasses override get_input_t: no resample, no clip
            v = torch.zeros(self.N, rom.m)
            if rom.cls != "SingleInt2D":
                raise ValueError("port: Zero / Square / Circle generators restated for SingleInt2D")
            if kind == "SquareTrajectoryGenerator":                       # :630-640
                c1 = 2 / rom.v_max[1]
                c2 = c1 + 1 / rom.v_max[0]
                c3 = c2 + 2 / abs(rom.v_min[1])
                c4 = c3 + 1 / abs(rom.v_min[0])
                v[(0 <= t) & (t < c1), 1] = rom.v_max[1] / 2
                v[(c1 <= t) & (t < c2), 0] = rom.v_max[0]
                v[(c2 <= t) & (t < c3), 1] = rom.v_min[1] / 2
                v[(c3 <= t) & (t < c4), 0] = rom.v_min[1]
            elif kind == "CircleTrajectoryGenerator":                     # :686-692
                e = z - self.center
                v[:, 0] = -e[:, 1]
                v[:, 1] = e[:, 0]
                v += -(e - 0.5 * e / torch.linalg.norm(v, dim=-1, keepdim=True))
                v = v / torch.linalg.norm(v, dim=-1, keepdim=True) * torch.min(torch.minimum(rom.v_max, torch.abs(rom.v_min)))
            return v
        idx = torch.nonzero(t > self.t_final).reshape((-1,))
        self.resample(idx, z)
        ramp = self.ramp_v_start + (self.ramp_v_end - self.ramp_v_start) * \
            ((t - self.ramp_t_start) / (self.t_final - self.ramp_t_start))[:, None]
        sinus = self.sin_mag * torch.sin(self.sin_freq * t[:, None] + self.sin_off) + self.sin_mean
        return self.weights[:, 0][:, None] * rom.clip_v_z(z, self.sample_hold_input) + \
            self.weights[:, 1][:, None] * rom.clip_v_z(z, ramp) + \
            self.weights[:, 2][:, None] * rom.clip_v_z(z, self.extreme_input) + \
            self.weights[:, 3][:, None] * rom.clip_v_z(z, sinus)

    def step_rom_idx(self, idx, increment_rom_time=False):                # rom_dynamics.py:577-590
        rom = self.rom
        self.v = self.get_input_t(self.t, self.traj[:, -1, :])
        self.v[self.stationary, :] = 0
        z_next = rom.f(self.traj[idx, -1, :], self.v[idx, :])
        mask = self.stationary[:, None] & rom.vel_inds
        z_next[mask[idx, :]] = 0
        self.traj[idx, :-1, :] = self.traj[idx, 1:, :].clone()
        self.traj[idx, -1, :] = z_next
        self.v_traj[idx, :-1, :] = self.v_traj[idx, 1:, :].clone()
        self.v_traj[idx, -1, :] = self.v[idx]
        self.k[idx] += 1
        if increment_rom_time:
            self.t[idx] += rom.dt

    def gen_step_idx(self, idx):                                          # rom_dynamics.py:571-575
        masked = idx[self.t[idx] >= self.k[idx] * self.rom.dt - 1e-5]
        self.step_rom_idx(masked)
        self.t[idx] += self.p.model_dt

    def gen_reset_idx(self, idx, z):                                      # rom_dynamics.py:595-605
        rom, W = self.rom, self.W
        if getattr(self.p, "generator", "") == "SquareTrajectoryGenerator":   # :673-675
            z[:, rom.vel_inds] = 0
        self.traj[idx, :, :] = 0.0
        self.v_traj[idx, :, :] = 0.0
        self.traj[idx, -1, :] = z[idx, :]
        self.k[idx] = -W
        self.t[idx] = self.k[idx] * rom.dt
        self.t_final[idx] = self.k[idx] * rom.dt
        self.resample(idx, z)
        for _ in range(W):
            self.step_rom_idx(idx, increment_rom_time=True)

    def get_trajectory(self):                                             # rom_dynamics.py:607-612
        t0, t1 = self.traj[:, :-1, :], self.traj[:, 1:, :]
        interp = t0 + (t1 - t0) * (self.t - (self.k - 1) * self.rom.dt)[:, None, None] / self.rom.dt
        return interp[:, ::self.p.dN, :]

    # ---- CustomSim ----------------------------------------------------------------------------
    def step(self, action):                                               # custom_sim.py:71-75
        self.root_states = self.model.f(self.root_states, action)
        self.gen_step_idx(torch.arange(self.N))
        self.trajectory = self.get_trajectory().clone()
        return self.get_observations(), torch.zeros(self.N, dtype=torch.bool)

    def get_observations(self):                                           # custom_sim.py:95-100
        return torch.cat((self.root_states.clone(), self.trajectory[:, 0, :], self.v_traj[:, 1, :].clone()), dim=1)

    def reset_idx(self, idx):                                             # custom_sim.py:80-93
        p = self.p
        ev = self._events(idx)
        self.root_states[idx, :] = self._uniform(self.noise_lower, self.noise_upper, P.SITE_ROM_ROOT, idx, ev, self.model.n)
        self.reset_traj(idx, self.rom.proj_z(self.root_states.clone()))
        return self.step(torch.zeros(self.N, self.model.m))

    def reset_traj(self, idx, p_zx):
        """reset_traj: custom_sim.py:80-85 == legged_robot_trajectory.py:248-253 (p_zx = proj_z of the root states)."""
        p = self.p
        if p.randomize_rom_distance:
            ev2 = self._events(idx)
            m = self._u(P.SITE_ROM_DIST_MASK, idx, ev2, 1).squeeze(1) > p.zero_rom_dist_llh
            sel = idx[m]
            d = self.max_rom_distance
            off = self._uniform(-d, d, P.SITE_ROM_DIST, sel, ev2[m.numpy()], self.rom.n)
            p_zx = p_zx.clone()
            p_zx[sel, :] += off
        elif self.rng == "philox":
            self._events(idx)     # the harness takes the event whether or not the branch draws
        self.gen_reset_idx(idx, p_zx)

    def reset(self):
        return self.reset_idx(torch.arange(self.N))

    # ---- DoubleSingleTracking -----------------------------------------------------------------
    def policy(self, obs):                                                # controllers.py:87-92
        xt, zt, vt = obs[:, :4], obs[:, 4:6], obs[:, 6:]
        u = self.p.Kp * (zt - xt[:, :2]) + self.p.Kd * (vt - xt[:, 2:])
        return self.model.clip_v_z(xt, u)

    # ---- data_collection_trajectory.py:104-149, one epoch -------------------------------------
    def collect_epoch(self, obs, T):
        N, rom, model = self.N, self.rom, self.model
        x = torch.zeros(N, T + 1, model.n)
        z = torch.zeros(N, T + 1, rom.n)
        pz_x = torch.zeros(N, T + 1, rom.n)
        v = torch.zeros(N, T, rom.m)
        done = torch.zeros(N, T, dtype=torch.bool)
        self.reset()
        x[:, 0, :] = self.root_states
        pz_x[:, 0, :] = rom.proj_z(self.root_states.clone())
        z[:, 0, :] = self.traj[:, 0, :]
        for t in range(T):
            k = self.k.clone()
            while torch.any(self.k == k):
                obs, dones = self.step(self.policy(obs))
            proj = rom.proj_z(self.root_states.clone())
            done[:, t] = dones
            v[:, t, :] = self.v
            x[:, t + 1, :] = self.root_states
            z[:, t + 1, :] = self.get_trajectory()[:, 0, :]
            z[done[:, t], t + 1, :] = proj[done[:, t], :]
            pz_x[:, t + 1, :] = proj
        return dict(x=x, z=z, pz_x=pz_x, v=v, done=done), obs


def gen_params(num_envs, rom_cls, **over):
    """A stand-alone TrajectoryGenerator over any rom class (the shape of trajopt/trajectory_gen.py:13-19,72-98)."""
    n, m, _ = FAMILY.get(rom_cls, (2 if rom_cls == "SingleInt2D" else 4, 2, 0))
    vel = {"SingleInt2D": [1.0, 1.0], "DoubleInt2D": [1.0, 1.0], "Unicycle": [1.0, 2.0], "LateralUnicycle": [1.0, 0.5, 2.0],
           "ExtendedUnicycle": [1.0, 4.0], "ExtendedLateralUnicycle": [1.0, 0.6, 4.0]}[rom_cls]
    zmax = {"SingleInt2D": [1e9, 1e9], "DoubleInt2D": [1e9, 1e9, 0.6, 0.6], "Unicycle": [1e9, 1e9, 1e9], "LateralUnicycle": [1e9, 1e9, 1e9],
            "ExtendedUnicycle": [1e9, 1e9, 1e9, 1.0, 2.0], "ExtendedLateralUnicycle": [1e9, 1e9, 1e9, 1.0, 0.5, 2.0]}[rom_cls]
    d = dict(num_envs=num_envs, rom_cls=rom_cls, rom_dt=0.1, dt_loop=0.02, z_min=[-a for a in zmax], z_max=zmax, v_min=[-a for a in vel],
             v_max=vel, N=6, dN=1, t_low=0.3, t_high=1.2, freq_low=0.01, freq_high=3.0, prob_stationary=0.05,
             weight_sampler="UniformWeightSampler", seed=0, generator="TrajectoryGenerator")
    d.update(over)
    d["model_dt"] = d["dt_loop"]
    return SimpleNamespace(**d)


class GenPort(RomPort):
    """TrajectoryGenerator alone (rom_dynamics.py:441-615) over any rom class: reset(z) / reset_idx / step / step_idx /
    get_input_t / get_trajectory, with RomPort's restatement of the generator underneath."""

    def __init__(self, p, rng="philox", env_id_offset=0):
        self.p, self.rng, self.off = p, rng, env_id_offset
        N = self.N = p.num_envs
        rom = self.rom = Rom(p.rom_cls, p.rom_dt, p.z_min, p.z_max, p.v_min, p.v_max)
        self.W = p.N * p.dN
        self.ctr = np.zeros(N, dtype=np.int64)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32)
        self.weights, self.t_final, self.t, self.k = z(N, 4), z(N), z(N), z(N)
        self.sample_hold_input, self.extreme_input = z(N, rom.m), z(N, rom.m)
        self.ramp_t_start, self.ramp_v_start = z(N), z(N, rom.m)
        ids = torch.arange(N)
        self.ramp_v_end = self._uniform(rom.v_min, rom.v_max, P.SITE_ROM_INIT, ids, self._events(ids), rom.m)
        self.sin_mag, self.sin_freq, self.sin_off, self.sin_mean = z(N, rom.m), z(N, rom.m), z(N, rom.m), z(N, rom.m)
        self.traj = z(N, self.W + 1, rom.n)
        self.v_traj = z(N, self.W, rom.m)
        self.v = z(N, rom.m)
        self.stationary = torch.zeros(N, dtype=torch.bool)

    def reset(self, z):
        self.gen_reset_idx(torch.arange(self.N), z)

    def step(self):
        self.gen_step_idx(torch.arange(self.N))
