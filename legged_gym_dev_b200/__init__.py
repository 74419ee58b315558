"""b200gym: B200-native (sm_100a) implementation of legged_gym_dev's per-step tensor pipeline.

Public surface (mirrors the reference):
    legged_gym_dev_b200.task_registry.task_registry       register / make_env / make_alg_runner
    legged_gym_dev_b200.legged_robot.{LeggedRobot, Anymal}
    legged_gym_dev_b200.rom.{SingleInt2D, DoubleInt2D, TrajectoryGenerator, CustomSim, DoubleSingleTracking}
    legged_gym_dev_b200.ppo.{RolloutStorage, ActorCritic, PPO, OnPolicyRunner}
The compute path is the C-ABI library libb200gym.so (include/b200gym.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
