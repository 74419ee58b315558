"""Executes the UNMODIFIED reference (/root/reference) as the pinning oracle.  TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box; there the byte-identical copy staged into the git-ignored oracle/_ref/ by
oracle/stage_reference.py (run by __graft_entry__.build() in the container) is imported instead.  Users: oracle/make_golden.py
(writes tests/golden/*.npz), tests/test_oracle_cpu.py (validates the travelling restatements in oracle/port_*.py; `-m "not gpu"`,
skipped when no tree is present) and bench.py's CPU arm (`--impl reference`, `cpu_baseline`), which times the reference's own
LeggedRobot.step.  No `-m gpu` test, smoke() or product module imports it.

How: the absent heavy dependencies are replaced in sys.modules by mocks, except
`isaacgym.torch_utils`, which is the restatement in oracle/isaacgym_restated.py; `Anymal.__init__` is
bypassed and the attributes the asset loader would have produced are set by hand; `self.gym` is a stub
whose simulate()/refresh_*() copy frames of a replay tape into the aliased state tensors.  The
reference's own methods (`step`, `post_physics_step`, `_compute_torques`, `CustomSim.step`, ...) then
run verbatim.  Randomness goes through oracle/rng_shim.py.
"""
import os
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

from . import philox as P
from . import rng_shim, isaacgym_restated

# the tree itself in the build container; on the GPU box the byte-identical copy staged by oracle/stage_reference.py (git-ignored)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference/legged_gym") else _STAGED


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "legged_gym"))


_imported = {}


def import_reference():
    """Stub-inject and import the reference packages once; returns a namespace of modules."""
    if _imported:
        return SimpleNamespace(**_imported)
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for name in ("casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "omegaconf", "hydra",
                 "rsl_rl", "rsl_rl.env", "rsl_rl.runners", "pytorch3d", "pytorch3d.transforms", "wandb",
                 "isaacgym.gymapi", "isaacgym.gymtorch", "isaacgym.gymutil", "isaacgym.terrain_utils"):
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if isinstance(sys.modules["pytorch3d.transforms"], MagicMock):
        # the Hopper's torque law really calls these (hopper.py:38,213-221): restated, not mocked
        from . import pytorch3d_restated
        sys.modules["pytorch3d.transforms"] = pytorch3d_restated
        sys.modules["pytorch3d"].transforms = pytorch3d_restated
    if "isaacgym" not in sys.modules or isinstance(sys.modules["isaacgym"], MagicMock):
        pkg = types.ModuleType("isaacgym")
        pkg.__path__ = []
        tu = types.ModuleType("isaacgym.torch_utils")
        for k in ("to_torch", "get_axis_params", "normalize", "quat_apply", "quat_rotate_inverse", "torch_rand_float"):
            setattr(tu, k, getattr(isaacgym_restated, k))
        tu.torch = torch
        tu.np = np
        tu.__all__ = ["to_torch", "get_axis_params", "normalize", "quat_apply", "quat_rotate_inverse",
                      "torch_rand_float", "torch", "np"]
        pkg.torch_utils = tu
        for sub in ("gymapi", "gymtorch", "gymutil", "terrain_utils"):
            setattr(pkg, sub, sys.modules[f"isaacgym.{sub}"])
        sys.modules["isaacgym"] = pkg
        sys.modules["isaacgym.torch_utils"] = tu
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import legged_gym.envs as envs                                        # noqa: E402
    import legged_gym.envs.base.legged_robot as legged_robot              # noqa: E402
    import legged_gym.envs.anymal_c.anymal as anymal                      # noqa: E402
    import trajopt.rom_dynamics as rom_dynamics                           # noqa: E402
    import deep_tube_learning.utils as dtl_utils                          # noqa: E402
    import deep_tube_learning.controllers as controllers                  # noqa: E402
    import deep_tube_learning.custom_sim as custom_sim                    # noqa: E402
    import legged_gym.envs.hopper.hopper as hopper                        # noqa: E402
    _imported["hopper"] = hopper
    _imported.update(envs=envs, legged_robot=legged_robot, anymal=anymal, rom_dynamics=rom_dynamics,
                     dtl_utils=dtl_utils, controllers=controllers, custom_sim=custom_sim)
    return SimpleNamespace(**_imported)


# --------------------------------------------------------------------------------------------
# Group R: Anymal / LeggedRobot
# --------------------------------------------------------------------------------------------
UPSTREAM_REWARD_SCALES = dict(  # the commented-out upstream defaults, legged_robot_config.py:155-168
    tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-2.0, ang_vel_xy=-0.05, torques=-1e-5,
    dof_acc=-2.5e-7, feet_air_time=1.0, collision=-1.0, action_rate=-0.01, termination=-0.0)
UPSTREAM_COMMAND_RANGES = dict(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1],
                               heading=[-3.14, 3.14])     # legged_robot_config.py:76-79
ALL_REWARD_SCALES = dict(       # every _reward_* term switched on (legged_robot.py:918-1015)
    action_rate=-0.01, ang_vel_xy=-0.05, base_height=-1.0, collision=-1.0, dof_acc=-2.5e-7,
    dof_pos_limits=-10.0, dof_vel=-1e-4, dof_vel_limits=-0.5, feet_air_time=1.0,
    feet_contact_forces=-0.01, lin_vel_z=-2.0, orientation=-5.0, stand_still=-0.1, stumble=-0.3,
    termination=-3.0, torque_limits=-0.02, torques=-1e-5, tracking_ang_vel=0.5, tracking_lin_vel=1.0)


class _ReplayGym:
    """Stands in for `self.gym`: simulate()/refresh_*() replay a tape into the aliased tensors."""

    def __init__(self, env, tape):
        self.env, self.tape = env, tape
        self.frame = 0
        self.sub = 0

    def simulate(self, sim):
        t = self.tape
        self.env.dof_state.copy_(t.dof[self.frame % t.frames, self.sub])
        self.sub += 1

    def refresh_actor_root_state_tensor(self, sim):
        t = self.tape
        self.env.root_states.copy_(t.root[self.frame % t.frames])

    def refresh_net_contact_force_tensor(self, sim):
        t = self.tape
        self.env.contact_forces.copy_(t.contact[self.frame % t.frames].reshape(self.env.contact_forces.shape))
        self.frame += 1
        self.sub = 0

    def __getattr__(self, name):     # every other gym call is a no-op
        return lambda *a, **k: None


def make_reference_anymal(task, num_envs, tape, seed=0, reward_scales=None, command_ranges=None,
                          use_actuator_network=None, heightfield=None, terrain_origins=None,
                          episode_lengths=None, overrides=None):
    """Build the reference's Anymal env for `task` in {"anymal_c_flat","anymal_c_rough"} without Isaac Gym."""
    ref = import_reference()
    from legged_gym_dev_b200 import synthetic as S
    envs = ref.envs
    cfg = envs.AnymalCFlatCfg() if task == "anymal_c_flat" else envs.AnymalCRoughCfg()
    cfg.curriculum.use_curriculum = False          # annotation-only in the fork (legged_robot_config.py:179)
    cfg.domain_rand.max_push_vel = 1.0             # read at legged_robot.py:827 but never defined
    cfg.env.num_envs = num_envs
    if reward_scales is not None:
        for k, v in reward_scales.items():
            setattr(cfg.rewards.scales, k, v)
    if command_ranges is not None:
        for k, v in command_ranges.items():
            setattr(cfg.commands.ranges, k, list(v))
    if use_actuator_network is not None:
        cfg.control.use_actuator_network = use_actuator_network
    for path, v in (overrides or {}).items():
        obj = cfg
        parts = path.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p)
        setattr(obj, parts[-1], v)

    Anymal = ref.anymal.Anymal
    env = Anymal.__new__(Anymal)
    env.cfg = cfg
    env.sim_params = SimpleNamespace(dt=cfg.sim.dt, use_gpu_pipeline=False)
    env.height_samples = None
    env.debug_viz = False
    env.init_done = False
    env._parse_cfg(cfg)                                                   # reference code
    # --- BaseTask.__init__ buffers (base_task.py:59-79) ---
    env.device = "cpu"
    env.headless = True
    env.viewer = None
    env.enable_viewer_sync = False
    env.sim = None
    env.num_envs, env.num_obs = cfg.env.num_envs, cfg.env.num_observations
    env.num_privileged_obs, env.num_actions = cfg.env.num_privileged_obs, cfg.env.num_actions
    env.obs_buf = torch.zeros(env.num_envs, env.num_obs, dtype=torch.float)
    env.rew_buf = torch.zeros(env.num_envs, dtype=torch.float)
    env.reset_buf = torch.ones(env.num_envs, dtype=torch.long)
    env.episode_length_buf = torch.zeros(env.num_envs, dtype=torch.long)
    env.time_out_buf = torch.zeros(env.num_envs, dtype=torch.bool)
    env.privileged_obs_buf = None
    env.extras = {}
    # --- what create_sim/_create_envs would have produced (legged_robot.py:679-772) ---
    env.up_axis_idx = 2
    env.num_dof = env.num_dofs = S.NUM_DOF
    env.num_bodies = S.NUM_BODIES
    env.dof_names = list(S.DOF_NAMES)
    env.feet_indices = torch.tensor(S.FEET_INDICES, dtype=torch.long)
    env.penalised_contact_indices = torch.tensor(S.PENALISED_INDICES, dtype=torch.long)
    env.termination_contact_indices = torch.tensor(S.TERMINATION_INDICES, dtype=torch.long)
    lim = synthetic_dof_limits()
    env.dof_pos_limits = lim["dof_pos_limits"].clone()
    env.dof_vel_limits = lim["dof_vel_limits"].clone()
    env.torque_limits = lim["torque_limits"].clone()
    isl = cfg.init_state
    env.base_init_state = torch.tensor(isl.pos + isl.rot + isl.lin_vel + isl.ang_vel, dtype=torch.float)
    if cfg.terrain.mesh_type in ("heightfield", "trimesh"):
        env.terrain = SimpleNamespace(cfg=cfg.terrain, env_length=cfg.terrain.terrain_length,
                                      env_width=cfg.terrain.terrain_width,
                                      env_origins=terrain_origins.numpy())
        env.height_samples = heightfield
        g = torch.Generator().manual_seed(seed + 31)
        state = torch.get_rng_state()
        torch.manual_seed(seed + 31)
        env._get_env_origins()                                            # reference code (:790-817)
        torch.set_rng_state(state)
    else:
        env._get_env_origins()
    # --- aliased physics tensors (legged_robot.py:545-551): wrap_tensor returns ours, in order ---
    root = tape.root[0].clone()
    dof = tape.dof[0, 0].clone()
    contact = tape.contact[0].reshape(num_envs * S.NUM_BODIES, 3).clone()
    handed = iter([root, dof, contact])
    ref.legged_robot.gymtorch.wrap_tensor = lambda _t: next(handed)
    env.gym = MagicMock()
    env._init_buffers()                                                   # reference code
    env._prepare_reward_function()                                        # reference code
    env.init_done = True
    if cfg.control.use_actuator_network:
        path = cfg.control.actuator_net_file.format(LEGGED_GYM_ROOT_DIR=REF_ROOT)
        env.actuator_network = torch.jit.load(path)
    env.gym = _ReplayGym(env, tape)
    if episode_lengths is not None:
        env.episode_length_buf[:] = episode_lengths
    env._shim_seed = seed
    _install_legged_wrappers(ref)
    return env


# --------------------------------------------------------------------------------------------
# AnymalTrajectory / LeggedRobotTrajectory
# --------------------------------------------------------------------------------------------
TRAJECTORY_ALL_REWARD_SCALES = dict(   # every _reward_* of legged_robot_trajectory.py:1000-1110 except stand_still (reads self.commands)
    action_rate=-0.01, ang_vel_xy=-0.05, base_height=-1.0, collision=-1.0, differential_error=2.0, dof_acc=-2.5e-7,
    dof_pos_limits=-10.0, dof_vel=-1e-4, dof_vel_limits=-0.5, feet_air_time=1.0, feet_contact_forces=-0.01, lin_vel_z=-2.0,
    orientation=-5.0, stumble=-0.3, termination=-3.0, torque_limits=-0.02, torques=-1e-5, tracking_rom=6.0)


def complete_trajectory_cfg(cfg):
    """The fork's anymal *_trajectory configs lack attributes LeggedRobotTrajectory reads (they live on in the hopper
    configs, hopper_trajectory_config.py:113-260).  Filled in here — and identically in legged_gym_dev_b200/configs.py."""
    cfg.rewards.tracking_sigma = 0.25                                  # read at legged_robot_trajectory.py:892
    cfg.domain_rand.randomize_rom_distance = True                      # :250
    cfg.domain_rand.max_rom_dist = [1.0, 1.0]                          # :887
    cfg.domain_rand.zero_rom_distance_likelihood = 0.25                # :888
    cfg.curriculum = SimpleNamespace(use_curriculum=False, curriculum_steps=[2500, 5000])   # :82, :414
    tg = cfg.trajectory_generator
    tg.weight_samp_cls = "UniformWeightSampler"                        # 'WeightSamplerSampleAndHold' does not exist
    tg.dN = 1                                                          # :122 reads dN, the cfg defines DN
    tg.prob_stationary = cfg.rom.prob_stationary                       # :121 reads it from the generator cfg
    cfg.terrain.curriculum = False                                     # _update_terrain_curriculum needs self.commands (:508)
    return cfg


def make_reference_anymal_trajectory(task, num_envs, tape, seed=0, reward_scales=None, use_actuator_network=None,
                                     heightfield=None, terrain_origins=None, episode_lengths=None,
                                     time_until_next_push=None, overrides=None):
    """The reference's AnymalTrajectory for `task` in {"anymal_c_flat_trajectory","anymal_c_rough_trajectory"}."""
    ref = import_reference()
    from legged_gym_dev_b200 import synthetic as S
    envs = ref.envs
    import legged_gym.envs.base.legged_robot_trajectory as ltmod
    cfg = envs.AnymalCFlatTrajectoryCfg() if task == "anymal_c_flat_trajectory" else envs.AnymalCRoughTrajectoryCfg()
    complete_trajectory_cfg(cfg)
    cfg.env.num_envs = num_envs
    if task != "anymal_c_flat_trajectory":
        cfg.env.num_observations = 65 + 187
    if reward_scales is not None:
        for k, v in reward_scales.items():
            setattr(cfg.rewards.scales, k, v)
    if use_actuator_network is not None:
        cfg.control.use_actuator_network = use_actuator_network
    for path, v in (overrides or {}).items():
        obj = cfg
        parts = path.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p)
        setattr(obj, parts[-1], v)

    AT = envs.AnymalTrajectory
    env = AT.__new__(AT)
    env.cfg = cfg
    env.sim_params = SimpleNamespace(dt=cfg.sim.dt, use_gpu_pipeline=False)
    env.height_samples = None
    env.debug_viz = False
    env.init_done = False
    env._parse_cfg(cfg)                                                   # reference code (:877-902)
    env.device = "cpu"
    env.headless = True
    env.viewer = None
    env.enable_viewer_sync = False
    env.sim = None
    env.num_envs, env.num_obs = cfg.env.num_envs, cfg.env.num_observations
    env.num_privileged_obs, env.num_actions = cfg.env.num_privileged_obs, cfg.env.num_actions
    env.obs_buf = torch.zeros(env.num_envs, env.num_obs, dtype=torch.float)
    env.rew_buf = torch.zeros(env.num_envs, dtype=torch.float)
    env.reset_buf = torch.ones(env.num_envs, dtype=torch.long)
    env.episode_length_buf = torch.zeros(env.num_envs, dtype=torch.long)
    env.time_out_buf = torch.zeros(env.num_envs, dtype=torch.bool)
    env.privileged_obs_buf = None
    env.extras = {}
    env.up_axis_idx = 2
    env.num_dof = env.num_dofs = S.NUM_DOF
    env.num_bodies = S.NUM_BODIES
    env.dof_names = list(S.DOF_NAMES)
    env.feet_indices = torch.tensor(S.FEET_INDICES, dtype=torch.long)
    env.penalised_contact_indices = torch.tensor(S.PENALISED_INDICES, dtype=torch.long)
    env.termination_contact_indices = torch.tensor(S.TERMINATION_INDICES, dtype=torch.long)
    lim = synthetic_dof_limits()
    env.dof_pos_limits = lim["dof_pos_limits"].clone()
    env.dof_vel_limits = lim["dof_vel_limits"].clone()
    env.torque_limits = lim["torque_limits"].clone()
    isl = cfg.init_state
    env.base_init_state = torch.tensor(isl.pos + isl.rot + isl.lin_vel + isl.ang_vel, dtype=torch.float)
    if cfg.terrain.mesh_type in ("heightfield", "trimesh"):
        env.terrain = SimpleNamespace(cfg=cfg.terrain, env_length=cfg.terrain.terrain_length,
                                      env_width=cfg.terrain.terrain_width, env_origins=terrain_origins.numpy())
        env.height_samples = heightfield
        state = torch.get_rng_state()
        torch.manual_seed(seed + 31)
        env._get_env_origins()                                            # reference code
        torch.set_rng_state(state)
    else:
        env._get_env_origins()
    # --- the tail of LeggedRobotTrajectory.__init__ (:72-88), in its order ---
    env.max_rom_distance = torch.tensor(env.nominal_max_rom_distance)
    env.zero_rom_dist_llh = env.nominal_zero_rom_distance_likelihood
    # (UniformWeightSampler.sample hard-codes device='cuda' in torch.rand; under the shim that call returns CPU numbers)
    for name in ("UniformWeightSamplerNoRamp", "UniformWeightSamplerNoExtreme"):
        base = getattr(ref.dtl_utils, name)     # default device='cuda' and the env passes no args (:109)
        cpu_cls = type(name, (base,), {"__init__": (lambda b: lambda self, dim=4, seed=42, device="cpu":
                                                    b.__init__(self, dim=dim, seed=seed, device="cpu"))(base)})
        setattr(ltmod, name, cpu_cls)
    _install_rom_wrappers(ref)
    holder = SimpleNamespace(seed=int(cfg.trajectory_generator.seed), ctr=np.zeros(num_envs, dtype=np.int64))
    env._init_rom()                                                       # reference code
    ref.rom_dynamics.TrajectoryGenerator._pending_holder = holder
    try:
        env._init_trajectory_generator()                                  # reference code
    finally:
        ref.rom_dynamics.TrajectoryGenerator._pending_holder = None
    env.traj_gen._shim = holder
    env._traj_shim = holder
    root = tape.root[0].clone()
    dof = tape.dof[0, 0].clone()
    contact = tape.contact[0].reshape(num_envs * S.NUM_BODIES, 3).clone()
    handed = iter([root, dof, contact])
    ltmod.gymtorch.wrap_tensor = lambda _t: next(handed)
    env.gym = MagicMock()
    env._init_buffers()                                                   # reference code
    env._prepare_reward_function()                                        # reference code
    env.init_done = True
    if cfg.control.use_actuator_network:
        path = cfg.control.actuator_net_file.format(LEGGED_GYM_ROOT_DIR=REF_ROOT)
        env.actuator_network = torch.jit.load(path)
    if time_until_next_push is not None:                                  # the torch_rand_float draw of :85-88 as an input
        env.time_until_next_push = time_until_next_push.clone().reshape(num_envs, 1)
    env.gym = _ReplayGym(env, tape)
    if episode_lengths is not None:
        env.episode_length_buf[:] = episode_lengths
    env._shim_seed = seed
    _install_trajectory_wrappers(ref, ltmod)
    return env


def _install_trajectory_wrappers(ref, ltmod):
    LT = ltmod.LeggedRobotTrajectory
    rng_shim.install()
    all_ids = lambda s: np.arange(s.num_envs)
    # draw event: common_step_counter inside a step; an EXTERNAL reset_idx keys its draws with (external reset count << 40) | step, as for
    # the base env (oracle/port_legged.py reset_idx)
    ev = lambda s: getattr(s, "_shim_event", None) or s.common_step_counter
    _wrap(LT, "_push_robots", lambda s, push_idx: (push_idx.nonzero().flatten(), ev(s), [(P.SITE_PUSH, 0)]))
    _wrap(LT, "_reset_dofs", lambda s, env_ids: (env_ids, ev(s), [(P.SITE_RESET_DOF, 0)]))
    _wrap(LT, "_reset_root_states",
          lambda s, env_ids: (env_ids, ev(s),
                              [(P.SITE_RESET_XY, 0), (P.SITE_RESET_VEL, 0)] if s.custom_origins
                              else [(P.SITE_RESET_VEL, 0)]))
    _wrap(LT, "compute_observations", lambda s: (all_ids(s), ev(s), [(P.SITE_OBS_NOISE, 0)]))

    # the push-timer redraw is inline in post_physics_step (:175-178): an outer context resolved at draw time
    def pps_ctx(self):
        ids = lambda call, hist: (self.time_until_next_push <= 0).reshape(-1).nonzero().flatten().numpy()
        ev = lambda call, hist: self.common_step_counter
        return (ids, ev, [(P.SITE_PUSH_TIMER, 0)])
    _wrap(LT, "post_physics_step", pps_ctx)
    if (LT, "reset_idx") not in _wrapped:
        _wrapped.add((LT, "reset_idx"))
        orig_reset, orig_pps = LT.reset_idx, LT.post_physics_step

        def post_physics_step(self):
            self._in_pps = True
            try:
                return orig_pps(self)
            finally:
                self._in_pps = False

        def reset_idx(self, env_ids):
            external = not getattr(self, "_in_pps", False) and hasattr(self, "_shim_seed") and len(env_ids) > 0
            if external:
                self._ext_resets = getattr(self, "_ext_resets", 0) + 1
                self._shim_event = (self._ext_resets << 40) | int(self.common_step_counter)
            try:
                return orig_reset(self, env_ids)
            finally:
                self._shim_event = None
        LT.reset_idx = reset_idx
        LT.post_physics_step = post_physics_step

    if (LT, "reset_traj") not in _wrapped:
        _wrapped.add((LT, "reset_traj"))
        orig = LT.reset_traj

        def reset_traj(self, env_ids):
            holder = getattr(self, "_traj_shim", None)
            if holder is None:
                return orig(self, env_ids)
            ids = env_ids.detach().cpu().numpy()
            ev = holder.ctr[ids].copy()
            holder.ctr[ids] += 1
            llh = self.zero_rom_dist_llh
            lazy_ids = lambda call, hist: ids if call == 0 else ids[(hist[0] > llh).numpy().reshape(-1)]
            lazy_ev = lambda call, hist: ev if call == 0 else ev[(hist[0] > llh).numpy().reshape(-1)]
            with rng_shim.draws(holder.seed, lazy_ids, lazy_ev, [(P.SITE_ROM_DIST_MASK, 0), (P.SITE_ROM_DIST, 0)]):
                return orig(self, env_ids)
        LT.reset_traj = reset_traj



# --------------------------------------------------------------------------------------------
# HopperTrajectory as a whole env (hopper_trajectory.py), the only Hopper class of the fork that can run a full step
# --------------------------------------------------------------------------------------------
class _HopperReplayGym:
    """`self.gym` of HopperTrajectory.step: simulate() + the three refresh calls of one sub-step replay a per-sub-step tape."""

    def __init__(self, env, tape):
        self.env, self.tape, self.frame, self.sub = env, tape, 0, 0

    def simulate(self, sim):
        t, f = self.tape, self.frame % self.tape.frames
        self.env.dof_state.copy_(t.dof[f, self.sub].reshape(self.env.dof_state.shape))
        self.env.root_states.copy_(t.root[f, self.sub])
        self.env.contact_forces.copy_(t.contact[f, self.sub])
        self.sub += 1
        if self.sub == t.decimation:
            self.sub, self.frame = 0, self.frame + 1

    def __getattr__(self, name):
        return lambda *a, **k: None


def make_reference_hopper_trajectory(hp, dr, tape, time_until_next_push, episode_lengths=None):
    """The reference's HopperTrajectory (hopper_trajectory.py) without Isaac Gym, configured from the port's parameter namespace `hp`
    (oracle/port_hopper_env.hopper_env_params): the constructor is bypassed as for the ANYmal classes, its body is replayed by hand
    with the reference's own methods (`_parse_cfg`, `_init_rom`, `_init_trajectory_generator`, `_init_buffers`,
    `_prepare_reward_function`), the asset loader's outputs and the construction-time random multipliers (`_update_envs`) are inputs."""
    ref = import_reference()
    import legged_gym.envs.base.legged_robot_trajectory as ltmod
    import legged_gym.envs.hopper.hopper_trajectory as ht
    from legged_gym.envs.hopper.flat_trajectory.hopper_trajectory_config import HopperRoughTrajectoryCfg
    from .port_hopper_env import apply_params_to_cfg
    N = hp.num_envs
    cfg = HopperRoughTrajectoryCfg()
    apply_params_to_cfg(cfg, hp)
    rw, d, tg, isl = cfg.rewards, cfg.domain_rand, cfg.trajectory_generator, cfg.init_state
    dof_names = ["foot_slide", "wheel1_rotation", "wheel2_rotation", "wheel3_rotation"]

    HT = ht.HopperTrajectory
    env = HT.__new__(HT)
    # --- head of HopperTrajectory.__init__ (:46-58) ---
    env.nominal_spring_stiffness, env.nominal_spring_damping = cfg.asset.spring_stiffness, cfg.asset.spring_damping
    env.nominal_spring_setpoint = cfg.control.foot_pos_des
    env.zero_action = torch.repeat_interleave(torch.tensor(cfg.control.zero_action).reshape((1, -1)), N, 0).float()
    # --- LeggedRobotTrajectory.__init__ (:65-88) with BaseTask.__init__ / create_sim replaced by their outputs ---
    env.cfg = cfg
    env.sim_params = SimpleNamespace(dt=hp.sim_dt, use_gpu_pipeline=False)
    env.height_samples, env.debug_viz, env.init_done = None, False, False
    env._parse_cfg(cfg)                                                   # reference code
    env.device, env.headless, env.viewer, env.enable_viewer_sync, env.sim = "cpu", True, None, False, None
    env.num_envs, env.num_obs = N, cfg.env.num_observations
    env.num_privileged_obs, env.num_actions = None, 4
    env.obs_buf = torch.zeros(N, env.num_obs, dtype=torch.float)
    env.rew_buf = torch.zeros(N, dtype=torch.float)
    env.reset_buf = torch.ones(N, dtype=torch.long)
    env.episode_length_buf = torch.zeros(N, dtype=torch.long)
    env.time_out_buf = torch.zeros(N, dtype=torch.bool)
    env.privileged_obs_buf, env.extras, env.up_axis_idx = None, {}, 2
    env.num_dof = env.num_dofs = 4
    env.num_bodies, env.dof_names = hp.num_bodies, dof_names
    env.feet_indices = torch.tensor([hp.foot_body], dtype=torch.long)
    env.penalised_contact_indices = torch.tensor(hp.penalised_bodies, dtype=torch.long)
    env.termination_contact_indices = torch.tensor(hp.termination_bodies, dtype=torch.long)
    env.dof_pos_limits = torch.tensor(hp.dof_pos_limits, dtype=torch.float)
    env.dof_vel_limits, env.torque_limits = torch.tensor(hp.dof_vel_limits), torch.tensor(hp.torque_limits)
    env.base_init_state = torch.tensor(isl.pos + isl.rot + isl.lin_vel + isl.ang_vel, dtype=torch.float)
    env._get_env_origins()                                                # reference code (plane: a grid)
    env.max_rom_distance = torch.tensor(env.nominal_max_rom_distance)
    env.zero_rom_dist_llh = env.nominal_zero_rom_distance_likelihood
    for name in ("UniformWeightSamplerNoRamp", "UniformWeightSamplerNoExtreme"):
        base = getattr(ref.dtl_utils, name)
        cpu_cls = type(name, (base,), {"__init__": (lambda b: lambda self, dim=4, seed=42, device="cpu":
                                                    b.__init__(self, dim=dim, seed=seed, device="cpu"))(base)})
        setattr(ltmod, name, cpu_cls)
    _install_rom_wrappers(ref)
    holder = SimpleNamespace(seed=int(tg.seed), ctr=np.zeros(N, dtype=np.int64))
    env._init_rom()                                                       # reference code
    ref.rom_dynamics.TrajectoryGenerator._pending_holder = holder
    try:
        env._init_trajectory_generator()                                  # reference code
    finally:
        ref.rom_dynamics.TrajectoryGenerator._pending_holder = None
    env.traj_gen._shim = holder
    env._traj_shim = holder
    handed = iter([tape.root[0, 0].clone(), tape.dof[0, 0].reshape(N * 4, 2).clone(), tape.contact[0, 0].reshape(N * hp.num_bodies, 3).clone()])
    ltmod.gymtorch.wrap_tensor = lambda _t: next(handed)
    env.gym = MagicMock()
    env._init_buffers()                                                   # reference code (HopperTrajectory's, :415-434)
    env._prepare_reward_function()                                        # reference code
    env.init_done = True
    env.time_until_next_push = time_until_next_push.clone().reshape(N, 1)   # the torch_rand_float draw of :85-88 as an input
    # --- tail of HopperTrajectory.__init__ (:59-100) ---
    env.raibert_Kp, env.raibert_Kv, env.raibert_Kff = rw.raibert.Kp, rw.raibert.Kv, rw.raibert.Kff
    env.raibert_clip_pos, env.raibert_clip_vel, env.raibert_clip_ang = rw.raibert.clip_pos, rw.raibert.clip_vel, rw.raibert.clip_ang
    env.use_raibert = False
    env.foot_joint_index = torch.tensor([0])
    env.wxyz_quat_inds = torch.tensor([6, 3, 4, 5])
    env.wheel_joint_indices = torch.tensor([1, 2, 3])
    env.actuator_transform = sys.modules["pytorch3d.transforms"].Rotate(torch.tensor(cfg.asset.rot_actuator), device="cpu")
    env.torque_speed_bound_ratio = cfg.asset.torque_speed_bound_ratio
    t = lambda v: torch.tensor(v, dtype=torch.float)
    env.default_dof_pos_noise_lower, env.default_dof_pos_noise_upper = t(isl.default_dof_pos_noise_lower), t(isl.default_dof_pos_noise_upper)
    env.default_dof_vel_noise_lower, env.default_dof_vel_noise_upper = t(isl.default_dof_vel_noise_lower), t(isl.default_dof_vel_noise_upper)
    env.default_root_pos_noise_lower, env.default_root_pos_noise_upper = t(isl.default_root_pos_noise_lower), t(isl.default_root_pos_noise_upper)
    env.default_root_vel_noise_lower, env.default_root_vel_noise_upper = t(isl.default_root_vel_noise_lower), t(isl.default_root_vel_noise_upper)
    env.max_vel = t(d.max_push_vel)
    for k, v in dr.items():                                               # what _update_envs (:372-413) draws at construction
        setattr(env, k, v.clone())
    env.gym = _HopperReplayGym(env, tape)
    if episode_lengths is not None:
        env.episode_length_buf[:] = episode_lengths
    env._shim_seed = hp.seed
    _install_trajectory_wrappers(ref, ltmod)
    _install_hopper_wrappers(ht)
    return env


def _install_hopper_wrappers(ht):
    HT = ht.HopperTrajectory
    rng_shim.install()
    all_ids = lambda s: np.arange(s.num_envs)
    _wrap(HT, "_push_robots", lambda s, push_idx: (push_idx, s.common_step_counter, [(P.SITE_HOP_PUSH, 0)]))
    _wrap(HT, "_reset_dofs", lambda s, env_ids: (env_ids, s.common_step_counter, [(P.SITE_HOP_DOF_POS, 0), (P.SITE_HOP_DOF_VEL, 0)]))
    _wrap(HT, "_reset_root_states",
          lambda s, env_ids: (env_ids, s.common_step_counter,
                              [(P.SITE_HOP_ROOT_POS, 0)] + ([(P.SITE_HOP_YAW, 0)] if s.cfg.init_state.randomize_yaw else []) + [(P.SITE_HOP_ROOT_VEL, 0)]))
    _wrap(HT, "compute_observations", lambda s: (all_ids(s), s.common_step_counter, [(P.SITE_OBS_NOISE, 0)]))

    def pps_ctx(self):   # the push-timer redraw is inline in post_physics_step (:156-159): rows = the envs whose timer ran out, in order
        ids = lambda call, hist: (self.time_until_next_push <= 0).reshape(-1).nonzero().flatten().numpy()
        ev = lambda call, hist: self.common_step_counter
        return (ids, ev, [(P.SITE_PUSH_TIMER, 0)])
    _wrap(HT, "post_physics_step", pps_ctx)


def synthetic_dof_limits():
    """The URDF carries effort/velocity limits (80 N m, 20 rad/s: anymal_c.urdf:540) but no position
    limits; synthetic +-(HAA 0.72, HFE/KFE 9.42*0.1...) style limits make the limit rewards non-trivial."""
    lo = torch.tensor([-0.72, -1.2, -1.8] * 4, dtype=torch.float)
    hi = torch.tensor([0.49, 1.2, 1.8] * 4, dtype=torch.float)
    lo[3:6], hi[3:6] = torch.tensor([-0.49, -1.2, -1.8]), torch.tensor([0.72, 1.2, 1.8])
    return dict(dof_pos_limits=torch.stack([lo, hi], dim=1),
                dof_vel_limits=torch.full((12,), 20.0), torque_limits=torch.full((12,), 80.0))


_wrapped = set()


def _wrap(cls, name, make_ctx):
    key = (cls, name)
    if key in _wrapped:
        return
    _wrapped.add(key)
    orig = getattr(cls, name)

    def wrapper(self, *a, **k):
        if not hasattr(self, "_shim_seed"):
            return orig(self, *a, **k)
        ctx = make_ctx(self, *a, **k)
        if ctx is None:
            return orig(self, *a, **k)
        with rng_shim.draws(self._shim_seed, *ctx):
            return orig(self, *a, **k)

    wrapper.__wrapped__ = orig
    setattr(cls, name, wrapper)


def _install_legged_wrappers(ref):
    LR = ref.legged_robot.LeggedRobot
    rng_shim.install()
    all_ids = lambda s: np.arange(s.num_envs)

    # draw event: common_step_counter inside a step; an EXTERNAL reset_idx (BaseTask.reset, user code) keys its draws with
    # (external reset count << 40) | common_step_counter so that they never coincide with a step's (oracle/port_legged.py reset_idx)
    ev = lambda s: getattr(s, "_shim_event", None) or s.common_step_counter

    def cmd_ctx(self, env_ids):
        site = P.SITE_CMD_RESET if getattr(self, "_in_reset", False) else P.SITE_CMD_PERIODIC
        return (env_ids, ev(self), [(site, 0), (site, 1), (site, 2)])

    _wrap(LR, "_resample_commands", cmd_ctx)
    _wrap(LR, "_push_robots", lambda s: (all_ids(s), ev(s), [(P.SITE_PUSH, 0)]))
    _wrap(LR, "_update_terrain_curriculum",
          lambda s, env_ids: (env_ids, ev(s), [(P.SITE_TERRAIN, 0)]))
    _wrap(LR, "_reset_dofs", lambda s, env_ids: (env_ids, ev(s), [(P.SITE_RESET_DOF, 0)]))
    _wrap(LR, "_reset_root_states",
          lambda s, env_ids: (env_ids, ev(s),
                              [(P.SITE_RESET_XY, 0), (P.SITE_RESET_VEL, 0)] if s.custom_origins
                              else [(P.SITE_RESET_VEL, 0)]))
    _wrap(LR, "compute_observations", lambda s: (all_ids(s), ev(s), [(P.SITE_OBS_NOISE, 0)]))
    if (LR, "reset_idx") not in _wrapped:
        _wrapped.add((LR, "reset_idx"))
        orig = LR.reset_idx
        orig_pps = LR.post_physics_step

        def post_physics_step(self):
            self._in_pps = True
            try:
                return orig_pps(self)
            finally:
                self._in_pps = False

        def reset_idx(self, env_ids):
            external = not getattr(self, "_in_pps", False) and hasattr(self, "_shim_seed") and len(env_ids) > 0
            if external:
                self._ext_resets = getattr(self, "_ext_resets", 0) + 1
                self._shim_event = (self._ext_resets << 40) | int(self.common_step_counter)
            self._in_reset = True
            try:
                return orig(self, env_ids)
            finally:
                self._in_reset = False
                self._shim_event = None
        LR.reset_idx = reset_idx
        LR.post_physics_step = post_physics_step


# --------------------------------------------------------------------------------------------
# Group M: CustomSim / TrajectoryGenerator / DoubleSingleTracking
# --------------------------------------------------------------------------------------------
def rom_config(num_envs, **over):
    """Plain-namespace mirror of configs/data_generation/double_single_int.yaml:29-87 (+ default_custom.yaml)."""
    pos = 1e9
    d = dict(acc=0.5, vel=0.3, vel_rom=0.2, model_dt=0.05, rom_dt=0.1, N=10, dN=1, t_low=1, t_high=2,
             freq_low=0.01, freq_high=2, prob_stationary=0.0005, max_rom_distance=[1.0, 1.0],
             zero_rom_dist_llh=0.25, noise_lower=[0.0, 0.0, -0.1, -0.1], noise_upper=[0.0, 0.0, 0.1, 0.1],
             weight_samp_cls="UniformWeightSamplerNoRamp", model_cls="DoubleInt2D", rom_cls="SingleInt2D",
             randomize_rom_distance=True, Kp=10, Kd=10)
    d.update(over)
    ns = SimpleNamespace
    cfg = ns(
        env=ns(num_envs=num_envs, episode_length_s=20,
               model=ns(dt=d["model_dt"], cls=d["model_cls"], z_min=[-pos, -pos, -d["vel"], -d["vel"]],
                        z_max=[pos, pos, d["vel"], d["vel"]], v_min=[-d["acc"]] * 2, v_max=[d["acc"]] * 2)),
        rom=ns(cls=d["rom_cls"], dt=d["rom_dt"], z_min=[-pos, -pos], z_max=[pos, pos],
               v_min=[-d["vel_rom"]] * 2, v_max=[d["vel_rom"]] * 2),
        trajectory_generator=ns(cls="TrajectoryGenerator", t_samp_cls="UniformSampleHoldDT",
                                weight_samp_cls=d["weight_samp_cls"], N=d["N"], t_low=d["t_low"],
                                t_high=d["t_high"], freq_low=d["freq_low"], freq_high=d["freq_high"], seed=0,
                                prob_stationary=d["prob_stationary"], dN=d["dN"]),
        noise=ns(add_noise=False),
        domain_rand=ns(randomize_rom_distance=d["randomize_rom_distance"],
                       max_rom_distance=d["max_rom_distance"], zero_rom_dist_llh=d["zero_rom_dist_llh"]),
        init_state=ns(default_noise_lower=d["noise_lower"], default_noise_upper=d["noise_upper"]),
        controller=ns(Kp=d["Kp"], Kd=d["Kd"]))
    return cfg


def make_reference_custom_sim(num_envs, seed=0, **over):
    """The reference's CustomSim + DoubleSingleTracking on CPU, RNG through the shim."""
    ref = import_reference()
    cs, rd = ref.custom_sim, ref.rom_dynamics
    cfg = rom_config(num_envs, **over)
    # weight samplers default to device='cuda' and CustomSim passes no args (custom_sim.py:55, utils.py:52-79)
    for name in ("UniformWeightSamplerNoRamp", "UniformWeightSamplerNoExtreme"):
        base = getattr(ref.dtl_utils, name)
        if not getattr(base, "_cpu_patched", False):
            cpu_cls = type(name, (base,), {"__init__": (lambda b: lambda self, dim=4, seed=42, device="cpu":
                                                        b.__init__(self, dim=dim, seed=seed, device="cpu"))(base),
                                           "_cpu_patched": True})
            setattr(cs, name, cpu_cls)
    _install_rom_wrappers(ref)
    holder = SimpleNamespace(seed=seed, ctr=np.zeros(num_envs, dtype=np.int64))
    rd.TrajectoryGenerator._pending_holder = holder
    try:
        env = cs.CustomSim(cfg)
    finally:
        rd.TrajectoryGenerator._pending_holder = None
    env._shim = holder
    env.traj_gen._shim = holder
    policy = ref.controllers.DoubleSingleTracking(cfg.controller.Kp, cfg.controller.Kd, env.model.clip_v_z)
    return env, policy, cfg


def make_reference_generator(p, seed=0):
    """The reference's rom class `p.rom_cls` (torch backend on CPU) under a stand-alone TrajectoryGenerator (rom_dynamics.py:441-615),
    RNG through the shim.  `p` = oracle.port_rom.gen_params(...)."""
    ref = import_reference()
    rd, du = ref.rom_dynamics, ref.dtl_utils
    _install_rom_wrappers(ref)
    t = lambda v: torch.tensor(v, dtype=torch.float32)
    rom = getattr(rd, p.rom_cls)(p.rom_dt, t(p.z_min), t(p.z_max), t(p.v_min), t(p.v_max), n_robots=p.num_envs, backend="torch", device="cpu")
    t_samp = du.UniformSampleHoldDT(p.t_low, p.t_high, backend="torch", device="cpu")
    w_samp = getattr(du, p.weight_sampler)()
    holder = SimpleNamespace(seed=seed, ctr=np.zeros(p.num_envs, dtype=np.int64))
    rd.TrajectoryGenerator._pending_holder = holder
    try:
        tg = rd.TrajectoryGenerator(rom, t_samp, w_samp, dt_loop=p.dt_loop, N=p.N, freq_low=p.freq_low, freq_high=p.freq_high, seed=seed,
                                    backend="torch", device="cpu", prob_stationary=p.prob_stationary, dN=p.dN)
    finally:
        rd.TrajectoryGenerator._pending_holder = None
    tg._shim = holder
    return tg, rom


def _install_rom_wrappers(ref):
    rd, cs = ref.rom_dynamics, ref.custom_sim
    TG, CS = rd.TrajectoryGenerator, cs.CustomSim
    rng_shim.install()
    if (TG, "__init__") in _wrapped:
        return
    _wrapped.add((TG, "__init__"))

    def take_events(holder, ids):
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        ev = holder.ctr[ids].copy()
        holder.ctr[ids] += 1
        return ev

    tg_init = TG.__init__

    def init(self, rom, *a, **k):
        holder = getattr(TG, "_pending_holder", None)
        if holder is None:
            return tg_init(self, rom, *a, **k)
        ids = np.arange(rom.n_robots)
        with rng_shim.draws(holder.seed, ids, take_events(holder, ids), [(P.SITE_ROM_INIT, 0)]):
            return tg_init(self, rom, *a, **k)
    TG.__init__ = init

    tg_resample = TG.resample

    def resample(self, idx, z):
        holder = getattr(self, "_shim", None)
        if holder is None or len(idx) == 0:
            return tg_resample(self, idx, z)
        ids = idx.detach().cpu().numpy()
        plan = [(s, 0) for s in (P.SITE_ROM_CONST, P.SITE_ROM_RAMP, P.SITE_ROM_EXTREME, P.SITE_ROM_SIN_MAG,
                                 P.SITE_ROM_SIN_MEAN, P.SITE_ROM_SIN_FREQ, P.SITE_ROM_SIN_OFF,
                                 P.SITE_ROM_TFINAL, P.SITE_ROM_WEIGHTS, P.SITE_ROM_STATIONARY)]
        with rng_shim.draws(holder.seed, ids, take_events(holder, ids), plan):
            return tg_resample(self, idx, z)
    TG.resample = resample

    cs_reset_idx = CS.reset_idx

    def reset_idx(self, idx):
        holder = getattr(self, "_shim", None)
        if holder is None:
            return cs_reset_idx(self, idx)
        self._root_event = take_events(holder, idx.detach().cpu().numpy())
        return cs_reset_idx(self, idx)
    CS.reset_idx = reset_idx

    # root draw happens inside reset_idx before reset_traj: give it its own one-call context by wrapping
    # torch_rand_vec_float as seen from custom_sim's namespace.
    vec = cs.torch_rand_vec_float

    def torch_rand_vec_float(lower, upper, shape, device):
        env = _current_sim[-1] if _current_sim else None
        if env is None or getattr(env, "_root_event", None) is None:
            return vec(lower, upper, shape, device)
        ev, env._root_event = env._root_event, None
        with rng_shim.draws(env._shim.seed, env._reset_ids, ev, [(P.SITE_ROM_ROOT, 0)]):
            return vec(lower, upper, shape, device)
    cs.torch_rand_vec_float = torch_rand_vec_float

    inner_reset_idx = CS.reset_idx

    def reset_idx_outer(self, idx):
        self._reset_ids = idx.detach().cpu().numpy()
        _current_sim.append(self)
        try:
            return inner_reset_idx(self, idx)
        finally:
            _current_sim.pop()
    CS.reset_idx = reset_idx_outer

    cs_reset_traj = CS.reset_traj

    def reset_traj(self, env_ids):
        holder = getattr(self, "_shim", None)
        if holder is None:
            return cs_reset_traj(self, env_ids)
        ids = env_ids.detach().cpu().numpy()
        ev = take_events(holder, ids)
        llh = self.zero_rom_dist_llh

        def lazy_ids(call, hist):
            if call == 0:
                return ids
            return ids[(hist[0] > llh).numpy().reshape(-1)]

        def lazy_ev(call, hist):
            if call == 0:
                return ev
            return ev[(hist[0] > llh).numpy().reshape(-1)]

        prev, self._root_event = getattr(self, "_root_event", None), None
        with rng_shim.draws(holder.seed, lazy_ids, lazy_ev, [(P.SITE_ROM_DIST_MASK, 0), (P.SITE_ROM_DIST, 0)]):
            return cs_reset_traj(self, env_ids)
    CS.reset_traj = reset_traj


_current_sim = []


# --------------------------------------------------------------------------------------------
# Hopper._compute_torques
# --------------------------------------------------------------------------------------------
def reference_hopper_torques(case, actions, cls="Hopper"):
    """Runs the UNMODIFIED Hopper._compute_torques (legged_gym/envs/hopper/hopper.py:168-237) on a stub `self` that carries exactly the
    attributes the method reads, built from `case` (oracle.port_hopper.hopper_case).  pytorch3d.transforms = oracle/pytorch3d_restated."""
    ref = import_reference()
    N = case["num_envs"]
    t = lambda k: case[k].clone()
    dof_state = t("dof_state")                                   # [N, 4, 2], as gymtorch hands it out (hopper uses .view(N, num_dof, 2))
    stub = SimpleNamespace(
        cfg=SimpleNamespace(control=SimpleNamespace(control_type=case["control_type"], action_scale=case["action_scale"])),
        num_envs=N, device="cpu", dof_pos=dof_state[..., 0], dof_vel=dof_state[..., 1], contact_forces=t("contact_forces"),
        feet_indices=torch.tensor([case["foot_body"]]), foot_joint_index=torch.tensor([0]), wheel_joint_indices=torch.tensor([1, 2, 3]),
        wxyz_quat_inds=torch.tensor([6, 3, 4, 5]), root_states=t("root_states"), base_ang_vel=t("base_ang_vel"),
        p_gains=t("p_gains"), d_gains=t("d_gains"), p_gain_random=t("p_gain_random"), d_gain_random=t("d_gain_random"),
        kd_spindown=t("kd_spindown"), spring_stiffness=t("spring_stiffness"), spring_damping=t("spring_damping"), foot_pos_des=t("foot_pos_des"),
        torque_speed_bound_ratio=case["torque_speed_bound_ratio"], torque_speed_bound_ratio_random=t("torque_speed_bound_ratio_random"),
        torque_limits=t("torque_limits"), torque_limit_random=t("torque_limit_random"), wheel_speed_limits=t("wheel_speed_limits"),
        wheel_limit_random=t("wheel_limit_random"), torques=torch.zeros(N, 4), last_dof_vel=torch.zeros(N, 4), sim_params=SimpleNamespace(dt=0.005),
        actuator_transform=sys.modules["pytorch3d.transforms"].Rotate(torch.tensor(case["rot_actuator"]), device="cpu"))
    if cls == "HopperTrajectory":       # the class of this fork that can run a full step (hopper_trajectory.py:184-253)
        import legged_gym.envs.hopper.hopper_trajectory as ht
        out = ht.HopperTrajectory._compute_torques(stub, actions.clone())
    else:
        out = ref.hopper.Hopper._compute_torques(stub, actions.clone())
    return out, stub.torques


def reference_hopper_observations(case, cfg, seed=0, event=1):
    """The UNMODIFIED Hopper._get_noise_scale_vec + Hopper.compute_observations (hopper.py:407-430, 239-258) + the clip of step (:116-117)
    on a stub `self`; torch.rand_like goes through the shim (site OBS_NOISE, event = common_step_counter)."""
    ref = import_reference()
    rng_shim.install()
    N = case["num_envs"]
    ns = SimpleNamespace
    t = lambda k: case[k].clone()
    dof_state = t("dof_state")
    hcfg = ns(noise=ns(add_noise=cfg["add_noise"], noise_level=cfg["noise_level"], noise_scales=ns(**cfg["noise_scales"])),
              terrain=ns(measure_heights=False), normalization=ns(clip_observations=cfg["clip_observations"]))
    obs_scales = ns(z_pos=cfg["z_pos"], lin_vel=cfg["lin_vel"], ang_vel=cfg["ang_vel"], dof_vel=cfg["dof_vel"], height_measurements=5.0)
    stub = ns(cfg=hcfg, obs_scales=obs_scales, obs_buf=torch.zeros(N, 21), actions=t("actions"), root_states=t("root_states"),
              base_quat=t("root_states")[:, 3:7], base_lin_vel=t("base_lin_vel"), base_ang_vel=t("base_ang_vel"), dof_vel=dof_state[..., 1],
              wheel_joint_indices=torch.tensor([1, 2, 3]), commands=t("commands"),
              commands_scale=torch.tensor([cfg["lin_vel"], cfg["lin_vel"], cfg["ang_vel"]]), last_dof_vel=t("last_dof_vel"), torques=t("torques"),
              dt=case.get("dt", 0.02))
    H = ref.hopper.Hopper
    stub.noise_scale_vec = H._get_noise_scale_vec(stub, hcfg)
    with rng_shim.draws(seed, np.arange(N), event, [(P.SITE_OBS_NOISE, 0)]):
        H.compute_observations(stub)
    obs = torch.clip(stub.obs_buf, -cfg["clip_observations"], cfg["clip_observations"])
    terms = torch.stack((H._reward_torque_limits(stub), H._reward_dof_acc(stub), H._reward_unit_quat(stub)), dim=1)
    return obs, stub.noise_scale_vec, terms


def reference_hopper_trajectory_observations(case, cfg, trajectory, trajectory_scale, gains, desired_velocity, seed=0, event=1):
    """The UNMODIFIED HopperTrajectory._get_noise_scale_vec / compute_observations / _reward_raibert (hopper_trajectory.py:439-468, 255-282,
    482-505) on a stub `self` with a SingleInt2D rom; `traj_gen.get_trajectory()` returns `trajectory`, `traj_gen.v` = desired_velocity."""
    ref = import_reference()
    rng_shim.install()
    import legged_gym.envs.hopper.hopper_trajectory as ht
    N = case["num_envs"]
    ns = SimpleNamespace
    t = lambda k: case[k].clone()
    rom = ref.rom_dynamics.SingleInt2D(0.1, torch.tensor([-1e9, -1e9]), torch.tensor([1e9, 1e9]), torch.tensor([-1.0, -1.0]), torch.tensor([1.0, 1.0]),
                                       n_robots=N, backend="torch", device="cpu")
    hcfg = ns(noise=ns(add_noise=cfg["add_noise"], noise_level=cfg["noise_level"], noise_scales=ns(**cfg["noise_scales"])),
              terrain=ns(measure_heights=False))
    obs_scales = ns(z_pos=cfg["z_pos"], lin_vel=cfg["lin_vel"], ang_vel=cfg["ang_vel"], dof_vel=cfg["dof_vel"], height_measurements=5.0)
    W, n = trajectory.shape[1], trajectory.shape[2]
    traj_gen = ns(N=W, rom=rom, v=desired_velocity.clone(), get_trajectory=lambda: trajectory.clone())
    stub = ns(cfg=hcfg, obs_scales=obs_scales, obs_buf=torch.zeros(N, 14 + W * n + 4), actions=t("actions"), root_states=t("root_states"),
              base_quat=t("root_states")[:, 3:7], base_lin_vel=t("base_lin_vel"), base_ang_vel=t("base_ang_vel"), dof_vel=t("dof_state")[..., 1],
              wheel_joint_indices=torch.tensor([1, 2, 3]), trajectory=trajectory.clone(), trajectory_scale=trajectory_scale.clone(), rom=rom,
              traj_gen=traj_gen, num_envs=N, raibert_Kp=gains["Kp"], raibert_Kv=gains["Kv"], raibert_Kff=gains["K_ff"],
              raibert_clip_pos=gains["clip_pos"], raibert_clip_vel=gains["clip_vel"], raibert_clip_ang=gains["clip_ang"])
    H = ht.HopperTrajectory
    stub.noise_scale_vec = H._get_noise_scale_vec(stub, hcfg)
    with rng_shim.draws(seed, np.arange(N), event, [(P.SITE_OBS_NOISE, 0)]):
        H.compute_observations(stub)
    obs = torch.clip(stub.obs_buf, -cfg["clip_observations"], cfg["clip_observations"])
    return obs, stub.noise_scale_vec, H._reward_raibert(stub)


def reference_hopper_trajectory_reset(state, env_ids, env_origins, cfg, seed=0, event=1, push_idx=None):
    """The UNMODIFIED HopperTrajectory._reset_dofs / _reset_root_states / _push_robots (hopper_trajectory.py:298-372) on a stub `self` whose
    gym calls are no-ops; draws through the shim at the HOP_* sites.  `state` tensors are modified in place."""
    import_reference()
    rng_shim.install()
    import legged_gym.envs.hopper.hopper_trajectory as ht
    ns = SimpleNamespace
    t = lambda v: torch.tensor(v, dtype=torch.float32)
    N = state["root_states"].shape[0]
    ds = state["dof_state"]
    stub = ns(dof_pos=ds[..., 0], dof_vel=ds[..., 1], dof_state=ds, root_states=state["root_states"], actions=state["actions"], num_dof=4, device="cpu",
              default_dof_pos=t(cfg["default_dof_pos"])[None, :], zero_action=t(cfg["zero_action"])[None, :].repeat(N, 1),
              default_dof_pos_noise_lower=t(cfg["dof_pos_noise"][0]), default_dof_pos_noise_upper=t(cfg["dof_pos_noise"][1]),
              default_dof_vel_noise_lower=t(cfg["dof_vel_noise"][0]), default_dof_vel_noise_upper=t(cfg["dof_vel_noise"][1]),
              default_root_pos_noise_lower=t(cfg["root_pos_noise"][0]), default_root_pos_noise_upper=t(cfg["root_pos_noise"][1]),
              default_root_vel_noise_lower=t(cfg["root_vel_noise"][0]), default_root_vel_noise_upper=t(cfg["root_vel_noise"][1]),
              custom_origins=False, base_init_state=t(cfg["base_init_state"]), env_origins=env_origins, wxyz_quat_inds=torch.tensor([6, 3, 4, 5]),
              cfg=ns(init_state=ns(randomize_yaw=cfg["randomize_yaw"])), max_vel=t(cfg["max_push_vel"]), gym=MagicMock(), sim=None)
    ids = env_ids.numpy()
    H = ht.HopperTrajectory
    with rng_shim.draws(seed, ids, event, [(P.SITE_HOP_DOF_POS, 0), (P.SITE_HOP_DOF_VEL, 0)]):
        H._reset_dofs(stub, env_ids)
    plan = [(P.SITE_HOP_ROOT_POS, 0)] + ([(P.SITE_HOP_YAW, 0)] if cfg["randomize_yaw"] else []) + [(P.SITE_HOP_ROOT_VEL, 0)]
    with rng_shim.draws(seed, ids, event, plan):
        H._reset_root_states(stub, env_ids)
    if push_idx is not None:
        with rng_shim.draws(seed, push_idx.numpy(), event, [(P.SITE_HOP_PUSH, 0)]):
            H._push_robots(stub, push_idx)
