"""Name -> (env class, env cfg, train cfg) registry with the reference's API (legged_gym/utils/task_registry.py:45-158):
`register`, `get_task_class`, `get_cfgs`, `make_env`, `make_alg_runner`.  The fused classes are registered under the
reference's task names with a `_b200` suffix (SURVEY.md §8b i) and take the same cfg objects."""
import copy
from types import SimpleNamespace

from . import configs


def _to_dict(obj):
    if hasattr(obj, "__dict__"):
        return {k: _to_dict(v) for k, v in vars(obj).items() if not k.startswith("_")}
    return obj


class TaskRegistry:
    def __init__(self):
        self.task_classes, self.env_cfgs, self.train_cfgs = {}, {}, {}

    def register(self, name, task_class, env_cfg, train_cfg):
        self.task_classes[name], self.env_cfgs[name], self.train_cfgs[name] = task_class, env_cfg, train_cfg

    def get_task_class(self, name):
        return self.task_classes[name]

    def get_cfgs(self, name):
        train_cfg, env_cfg = self.train_cfgs[name], self.env_cfgs[name]
        env_cfg.seed = train_cfg.seed                                   # task_registry.py:62-63
        return env_cfg, train_cfg

    def make_env(self, name, args=None, env_cfg=None, physics=None, **env_kwargs):
        """task_registry.py:66-103.  `physics` replaces the Isaac Gym simulation the reference creates here."""
        if name not in self.task_classes:
            raise ValueError(f"Task with name: {name} was not registered")
        task_class = self.get_task_class(name)
        if env_cfg is None:
            env_cfg, _ = self.get_cfgs(name)
            env_cfg = copy.deepcopy(env_cfg)
        if args is not None and getattr(args, "num_envs", None) is not None:
            env_cfg.env.num_envs = args.num_envs                        # helpers.py:208-231 (update_cfg_from_args)
        device = getattr(args, "sim_device", None) or env_kwargs.pop("sim_device", "cuda")
        # task_registry.py:89 set_seed(env_cfg.seed): the configured seed keys the env's Philox streams and torch's generator
        # (policy initialisation; PPO.act's sample stream is keyed by torch.initial_seed())
        seed = getattr(env_cfg, "seed", None)
        if seed is not None and "seed" not in env_kwargs:
            env_kwargs["seed"] = int(seed)
            import torch
            torch.manual_seed(int(seed))
        sim_params = SimpleNamespace(dt=env_cfg.sim.dt, use_gpu_pipeline=True)
        env = task_class(cfg=env_cfg, sim_params=sim_params, physics_engine=getattr(args, "physics_engine", None),
                         sim_device=device, headless=getattr(args, "headless", True), physics=physics, **env_kwargs)
        return env, env_cfg

    def make_alg_runner(self, env, name=None, args=None, train_cfg=None, log_root="default", wandb_callback=None):
        """task_registry.py:105-158 (log-dir / resume path handling reduced to explicit arguments)."""
        from .ppo import OnPolicyRunner
        if train_cfg is None:
            if name is None:
                raise ValueError("Either 'name' or 'train_cfg' must be not None")
            _, train_cfg = self.get_cfgs(name)
        log_dir = None if log_root in (None, "default") else log_root
        runner = OnPolicyRunner(env, _to_dict(train_cfg), log_dir, device=env.device, wandb_callback=wandb_callback)
        resume_path = getattr(getattr(train_cfg, "runner", None), "resume_path", None)
        if getattr(getattr(train_cfg, "runner", None), "resume", False) and resume_path:
            runner.load(resume_path)
        return runner, train_cfg


task_registry = TaskRegistry()


def _register_defaults():
    from .legged_robot import Anymal
    task_registry.register("anymal_c_rough_b200", Anymal, configs.anymal_c_rough_cfg(), configs.anymal_c_rough_cfg_ppo())
    task_registry.register("anymal_c_flat_b200", Anymal, configs.anymal_c_flat_cfg(), configs.anymal_c_flat_cfg_ppo())
    from .legged_robot_trajectory import AnymalTrajectory                # legged_gym/envs/__init__.py registrations
    task_registry.register("anymal_c_rough_trajectory_b200", AnymalTrajectory, configs.anymal_c_rough_trajectory_cfg(),
                           configs.anymal_c_rough_trajectory_cfg_ppo())
    task_registry.register("anymal_c_flat_trajectory_b200", AnymalTrajectory, configs.anymal_c_flat_trajectory_cfg(),
                           configs.anymal_c_flat_trajectory_cfg_ppo())
    from .hopper_trajectory import HopperTrajectory                      # legged_gym/envs/__init__.py: "hopper_flat_trajectory"
    task_registry.register("hopper_flat_trajectory_b200", HopperTrajectory, configs.hopper_flat_trajectory_cfg(),
                           configs.hopper_flat_trajectory_cfg_ppo())


_register_defaults()
