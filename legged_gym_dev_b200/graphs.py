"""CUDA-graph replay of whole env steps.

The step pipeline is 5 launches (+1 finaliser) per env step; below ~256K envs per GPU the launch latency and the
Python around it are comparable to the kernels' run time.  With a replayed physics tape every pointer of a tape cycle
is known in advance, so one cycle of F env steps (F = tape frames) is captured ONCE and replayed: the launch cost per
env step drops to a fraction of a graph launch.  The step number the kernels use (RNG event, push schedule) lives in
device memory and is advanced by the kernels themselves, so replays draw fresh random numbers every time."""
import torch


class GraphedReplay:
    def __init__(self, env, actions):
        """env: LeggedRobot driven by ReplayPhysics(copy=False); actions: list of F device tensors [N, 12], one per frame."""
        ph = env.physics
        if getattr(ph, "copy", True):
            raise ValueError("GraphedReplay needs a zero-copy ReplayPhysics (copy=False)")
        if len(actions) != ph.frames:
            raise ValueError("one action tensor per tape frame")
        self.env, self.F = env, ph.frames
        while ph.frame % self.F:          # align to the start of a tape cycle
            env.step(actions[ph.frame % self.F])
        env.use_device_step_counter()
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side):     # warm-up outside capture (function attributes, lazy module loading)
            for f in range(self.F):
                env.step(actions[f])
        torch.cuda.current_stream(env.device).wait_stream(side)
        torch.cuda.synchronize(env.device)
        self.graph = torch.cuda.CUDAGraph()
        c0, f0 = env.common_step_counter, ph.frame
        with torch.cuda.graph(self.graph):
            for f in range(self.F):
                env.step(actions[f])
        env.common_step_counter, ph.frame, ph.sub = c0, f0, 0   # capture executed nothing
        env._stream = None

    def replay(self):
        """Runs F env steps (one tape cycle)."""
        self.graph.replay()
        self.env.common_step_counter += self.F
        self.env.physics.frame += self.F


class GraphedRollout:
    """One PPO rollout — T x [PPO.act, env.step, PPO.process_env_step] of an OnPolicyRunner — captured ONCE as a CUDA graph and replayed per
    training iteration (rsl_rl OnPolicyRunner.learn's inner loop; ~10 launches per env step become one graph launch per rollout).  Needs an
    env whose step is graph-capturable: here a LeggedRobot on a replayed tape whose length divides T.  Both counters the kernels key their
    random draws with (env step, act event) live in device memory and are advanced by the kernels, so every replay draws fresh numbers."""

    def __init__(self, runner, rollout_fn):
        env, alg, T = runner.env, runner.alg, runner.num_steps_per_env
        ph = env.physics
        F = getattr(ph, "frames", None)
        if F is None or T % F or ph.frame % F:
            raise ValueError("GraphedRollout needs a replayed tape whose length divides the rollout length, at the start of a tape cycle")
        if alg.storage.step != 0:
            raise ValueError("GraphedRollout: capture at the start of a rollout")
        self.runner, self.T = runner, T
        env.use_device_step_counter()
        alg.use_device_act_counter()
        env._stream = None
        state = (env.common_step_counter, ph.frame, ph.sub, alg._act_event)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            rollout_fn()
        env.common_step_counter, ph.frame, ph.sub, alg._act_event = state     # capture executed nothing
        alg.storage.step = 0
        env._stream = None

    def replay(self):
        env, alg = self.runner.env, self.runner.alg
        self.graph.replay()
        env.common_step_counter += self.T
        env.physics.frame += self.T
        alg._act_event += self.T
        alg.storage.step = self.T
