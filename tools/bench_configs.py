"""Timings of the other BASELINE.json configs (1, 3, 4, 5) — the headline bench (config 2) lives in bench.py.
Each function returns a small dict; bench.py attaches them under "extra_configs".  CUDA-event timing after warm-up."""
import os
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PEAK = None


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def rough_lstm(num_envs=16384, steps=50, warmup=5, device="cuda", peak=6535.7):
    """cfg 3: anymal_c_rough, LSTM actuator net x4 + 187-point height scan + 235 observations."""
    from legged_gym_dev_b200 import configs, synthetic as S
    cfg = configs.anymal_c_rough_cfg()                # shipped reward table, trimesh terrain with curriculum, heading commands
    cfg.control.use_actuator_network = True
    cfg.env.num_envs = num_envs
    F = 4
    tape = S.make_state_tape(num_envs, frames=F, seed=7, rough=True, device=device)
    g = torch.Generator().manual_seed(3)
    levels = torch.randint(0, cfg.terrain.max_init_terrain_level + 1, (num_envs,), generator=g)
    types = torch.div(torch.arange(num_envs), (num_envs / cfg.terrain.num_cols), rounding_mode="floor").to(torch.long)
    to = S.make_terrain_origins(seed=1)
    terrain = dict(height_samples=S.make_heightfield(seed=1), terrain_origins=to, terrain_levels=levels, terrain_types=types,
                   env_origins=to[levels, types].clone())
    from legged_gym_dev_b200.legged_robot import Anymal
    from legged_gym_dev_b200.physics import ReplayPhysics
    env = Anymal(cfg, SimpleNamespace(dt=cfg.sim.dt), None, device, True, physics=ReplayPhysics(tape, device=device, copy=False),
                 asset=S.anymal_dof_limits(), seed=0, terrain=terrain)
    env.episode_length_buf = S.make_episode_lengths(num_envs, seed=1, device=device)
    acts = [tape.actions[f] for f in range(F)]
    for s in range(warmup):
        env.step(acts[s % F])
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for s in range(steps):
        env.step(acts[s % F])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    env._timing = []
    for s in range(20):
        env.step(acts[s % F])
    torch.cuda.synchronize()
    pp = [x.elapsed_time(y) for k, x, y in env._timing if k == "post_physics"]
    tq = [x.elapsed_time(y) for k, x, y in env._timing if k == "torques"]
    t_pp, t_tq = sum(pp) / len(pp), sum(tq) / len(tq)
    K = len(env.params.active_terms)
    pp_bytes = 986 + 8 * K + 748 + 748
    graph_ms = None
    try:   # the 6 launches of a step (4 x LSTM torques, post-physics, finaliser; PDL between them) replayed from ONE CUDA graph per tape cycle
        from legged_gym_dev_b200.graphs import GraphedReplay
        env._timing = None
        g = GraphedReplay(env, acts)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a.record()
        reps = max(1, steps // F)
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        graph_ms = a.elapsed_time(b) / (reps * F)
    except Exception as e:   # noqa: BLE001
        graph_ms = f"{type(e).__name__}: {e}"
    step_bytes = 4 * (3264 + 48) + pp_bytes
    return dict(config="anymal_c_rough: 4x LSTM actuator torques + post_physics with 187-pt height scan, 235 obs",
                num_envs=num_envs, ms_per_step=ms, env_steps_per_s=num_envs / (ms * 1e-3), graph_ms_per_step=graph_ms,
                graph_env_steps_per_s=(num_envs / (graph_ms * 1e-3) if isinstance(graph_ms, float) else None),
                graph_step_frac=(step_bytes * num_envs / (graph_ms * 1e-3) / 1e9 / peak if isinstance(graph_ms, float) else None),
                lstm_torques=dict(avg_launch_ms=t_tq, algorithmic_bytes_per_env=3264 + 48,
                                  achieved_gbs=(3264 + 48) * num_envs / (t_tq * 1e-3) / 1e9,
                                  frac=(3264 + 48) * num_envs / (t_tq * 1e-3) / 1e9 / peak),
                post_physics_rough=dict(avg_launch_ms=t_pp, algorithmic_bytes_per_env=pp_bytes,
                                        achieved_gbs=pp_bytes * num_envs / (t_pp * 1e-3) / 1e9,
                                        frac=pp_bytes * num_envs / (t_pp * 1e-3) / 1e9 / peak))


def trajectory_env(num_envs=4096, steps=50, warmup=5, device="cuda", peak=6535.7):
    """anymal_c_flat_trajectory env step = 4x PD torques + generator step + fused post-physics (traj_mode, 65 obs,
    all reward terms of the class) + generator reset.  Algorithmic bytes of the fused kernel: the flat figure with 65 obs columns
    + trajectory [10,2] read + prev_error / push timer read+write."""
    from legged_gym_dev_b200 import configs, synthetic as S
    from legged_gym_dev_b200.legged_robot_trajectory import AnymalTrajectory
    from legged_gym_dev_b200.physics import ReplayPhysics
    cfg = configs.anymal_c_flat_trajectory_cfg()
    for k, v in configs.TRAJECTORY_ALL_REWARD_SCALES.items():
        setattr(cfg.rewards.scales, k, v)
    cfg.control.use_actuator_network = False
    cfg.env.num_envs = num_envs
    F = 4
    tape = S.make_state_tape(num_envs, frames=F, seed=7, device=device)
    env = AnymalTrajectory(cfg, SimpleNamespace(dt=cfg.sim.dt), None, device, True, physics=ReplayPhysics(tape, device=device, copy=False),
                           asset=S.anymal_dof_limits(), seed=0)
    env.episode_length_buf = S.make_episode_lengths(num_envs, seed=1, device=device)
    env.reset_traj(torch.arange(num_envs, device=device))
    acts = [tape.actions[f] for f in range(F)]
    for s in range(warmup):
        env.step(acts[s % F])
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for s in range(steps):
        env.step(acts[s % F])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    env._timing = []
    for s in range(20):
        env.step(acts[s % F])
    torch.cuda.synchronize()
    pp = [x.elapsed_time(y) for k, x, y in env._timing if k == "post_physics"]
    t_pp = sum(pp) / len(pp)
    K = len(env.params.active_terms)
    pp_bytes = 986 + 8 * K + (65 - 48) * 4 + 80 + 2 * 8 + 2 * 4          # fused kernel only
    gen_bytes = 2 * (88 + 80 + 4 * 4 + 16 + 9 * 8 + 4 + 1 + 4) + 80         # generator step: state read + written, window out
    finite = bool(torch.isfinite(env.rew_buf).all()) and bool(torch.isfinite(env.obs_buf).all())
    graph_ms = None
    try:   # the same loop as CUDA-graph replays of whole tape cycles (no host branch in the trajectory step)
        from legged_gym_dev_b200.graphs import GraphedReplay
        env._timing = None
        g = GraphedReplay(env, acts)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = _events()
        a.record()
        reps = max(1, steps // F)
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        graph_ms = a.elapsed_time(b) / (reps * F)
    except Exception as e:
        graph_ms = f"{type(e).__name__}: {e}"
    return dict(graph_ms_per_step=graph_ms,
                graph_env_steps_per_s=(num_envs / (graph_ms * 1e-3) if isinstance(graph_ms, float) else None),
                config="anymal_c_flat_trajectory (AnymalTrajectory): 4x PD torques + generator step + fused post-physics (65 obs, 18 reward "
                       "terms) + generator reset",
                num_envs=num_envs, ms_per_step=ms, env_steps_per_s=num_envs / (ms * 1e-3), finite=finite,
                generator_plus_post_physics=dict(avg_launch_ms=t_pp, algorithmic_bytes_per_env=pp_bytes + gen_bytes,
                                                 achieved_gbs=(pp_bytes + gen_bytes) * num_envs / (t_pp * 1e-3) / 1e9,
                                                 frac=(pp_bytes + gen_bytes) * num_envs / (t_pp * 1e-3) / 1e9 / peak))


def rom_per_call(num_envs=4096, loop_steps=1000, device="cuda", cpu=True):
    """cfg 1: CustomSim.step + DoubleSingleTracking, 4096 envs x 1000 loop steps, call-per-step API; CPU port beside it."""
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.rom import CustomSim, DoubleSingleTracking
    env = CustomSim(configs.double_single_int_cfg(num_envs, seed=0), device=device)
    pol = DoubleSingleTracking(10, 10, env.model.clip_v_z)
    env.reset()
    obs = env.get_observations()
    for _ in range(20):
        obs, _, _, _, _ = env.step(pol(obs))
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(loop_steps):
        obs, _, _, _, _ = env.step(pol(obs))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    out = dict(config="DoubleInt2D model / SingleInt2D rom / DoubleSingleTracking, call-per-step API", num_envs=num_envs,
               loop_steps=loop_steps, ms_total=ms, env_steps_per_s=num_envs * loop_steps / (ms * 1e-3),
               note="2 launches per loop step: launch-latency bound at this size (SURVEY H5)")
    # the same 4096 x 1000 loop steps as ONE launch of the persistent rollout kernel (collect_epoch: reset + 500 ROM steps
    # = 1000 loop steps + 11 warm-up steps), and the call-per-step loop replayed from a CUDA graph (100 steps per replay)
    env2 = CustomSim(configs.double_single_int_cfg(num_envs, seed=0), device=device)
    o2 = torch.zeros(num_envs, 8, device=device)
    env2.collect_epoch(o2, loop_steps // 2)
    torch.cuda.synchronize()
    a.record()
    env2.collect_epoch(o2, loop_steps // 2)
    b.record()
    torch.cuda.synchronize()
    ms_p = a.elapsed_time(b)
    out["persistent_kernel"] = dict(ms_total=ms_p, env_steps_per_s=num_envs * loop_steps / (ms_p * 1e-3))
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                obs, _, _, _, _ = env.step(pol(obs))
        torch.cuda.current_stream(device).wait_stream(side)
        ob_static = obs.clone()
        with torch.cuda.graph(g):
            o = ob_static
            for _ in range(100):
                o, _, _, _, _ = env.step(pol(o))
            ob_static.copy_(o)
        g.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(loop_steps // 100):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        ms_g = a.elapsed_time(b)
        out["call_per_step_graph"] = dict(ms_total=ms_g, env_steps_per_s=num_envs * loop_steps / (ms_g * 1e-3))
    except Exception as e:
        out["call_per_step_graph"] = dict(error=f"{type(e).__name__}: {e}")
    if cpu:
        from oracle.port_rom import RomPort, rom_params
        torch.set_num_threads(os.cpu_count() or 1)
        port = RomPort(rom_params(num_envs, seed=0), rng="torch")
        o, _ = port.reset()
        t0 = time.perf_counter()
        for _ in range(loop_steps):
            o, _ = port.step(port.policy(o))
        dt = time.perf_counter() - t0
        out["cpu_port"] = dict(env_steps_per_s=num_envs * loop_steps / dt, seconds=dt, cores=torch.get_num_threads())
    return out


def rom_rollout(num_envs=1 << 20, T=200, device="cuda", peak=6535.7, debug=False):
    """cfg 4: one data-collection epoch (reset + T ROM steps = 2T loop steps) as ONE persistent launch."""
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.rom import CustomSim
    env = CustomSim(configs.double_single_int_cfg(num_envs, seed=0), device=device)
    obs = torch.zeros(num_envs, 8, device=device)
    env.collect_epoch(obs, 8)
    warm = env.collect_epoch(obs, T, save_debugging_data=debug)   # same log sizes: the timed call reuses the allocator's cached blocks
    del warm                                                        # (a fresh 5 GB cudaMalloc inside the timed region costs ~90 ms)
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    log = env.collect_epoch(obs, T, save_debugging_data=debug)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    loop_steps = 2 * T + 11
    log_bytes = num_envs * ((T + 1) * 16 + T * 9 + ((T + 1) * 16 if debug else 0))
    return dict(config="ROM tube data collection epoch, persistent rollout kernel", num_envs=num_envs, rom_steps=T,
                ms_epoch=ms, env_loop_steps_per_s=num_envs * loop_steps / (ms * 1e-3), log_gb=log_bytes / 1e9,
                log_gbs=log_bytes / (ms * 1e-3) / 1e9, frac_hbm=log_bytes / (ms * 1e-3) / 1e9 / peak,
                note="ALU-bound (sinf, IEEE division, Philox); HBM traffic is only the materialised logs")


def tube_dataset(num_envs=262144, T=200, N=10, device="cuda", peak=6535.7):
    """rollout logs -> tube error + N-deep sliding window, on the device (datasets.py:60-71)."""
    from legged_gym_dev_b200 import _lib
    L = _lib.lib()
    z = torch.randn(num_envs, T + 1, 2, device=device)
    pz = torch.randn(num_envs, T + 1, 2, device=device)
    v = torch.randn(num_envs, T, 2, device=device)
    w = torch.empty(num_envs, T, device=device)
    data = torch.cat((w[:, :, None], v), dim=-1).contiguous()
    out = torch.empty(num_envs, T, N * 3, device=device)
    st = _lib.stream_ptr(z.device)

    def run():
        _lib.check(L.b200gym_tube_error(_lib.ptr(z), _lib.ptr(pz), _lib.ptr(w), num_envs, T, T + 1, 2, st))
        _lib.check(L.b200gym_sliding_window(_lib.ptr(data), _lib.ptr(out), num_envs, T, 3, N, 1, 2, st))
    run()
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(5):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    byt = num_envs * (2 * (T + 1) * 8 + T * 4 + T * 12 * (1 + N))
    return dict(config="tube error + sliding window (N=10, dN=1) from rollout logs", num_envs=num_envs, T=T, ms=ms,
                gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / peak)


def gae_update(num_envs=4096, T=24, device="cuda", peak=6535.7):
    """cfg 5: GAE + advantage normalisation, and one PPO update iteration (5 epochs x 4 minibatches, flat nets)."""
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    ac = ActorCritic(48, 48, 12, actor_hidden_dims=[128, 64, 32], critic_hidden_dims=[128, 64, 32], init_noise_std=1.0)
    alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device=device)
    alg.init_storage(num_envs, T, [48], [None], [12])
    st = alg.storage
    g = torch.Generator(device=device).manual_seed(1)
    st.observations.normal_(generator=g)
    st.actions.normal_(generator=g)
    st.mu.normal_(generator=g)
    st.sigma.fill_(1.0)
    st.rewards.normal_(0.02, 0.05, generator=g)
    st.values.normal_(0.5, 0.3, generator=g)
    st.actions_log_prob.normal_(-17.0, 2.0, generator=g)
    st.dones.copy_(torch.rand(T, num_envs, 1, device=device, generator=g) < 0.005)
    last = torch.randn(num_envs, 1, device=device, generator=g)
    for _ in range(3):
        st.compute_returns(last, 0.99, 0.95)
    torch.cuda.synchronize()
    a, b = _events()
    # device time of the storage pass (zero stats + GAE scan + normalisation): replayed from a CUDA graph so that the host-side
    # cost of three Python-issued launches does not hide it
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st.compute_returns(last, 0.99, 0.95)
    g.replay()
    torch.cuda.synchronize()
    a.record()
    reps = 50
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    ms_gae = a.elapsed_time(b) / reps
    gae_sizes = {}
    for n2 in (65536, 1048576):
        st2 = type(st)(n2, T, [1], [None], [1], device)
        st2.rewards.normal_(0.02, 0.05)
        st2.values.normal_(0.5, 0.3)
        l2 = torch.randn(n2, 1, device=device)
        st2.compute_returns(l2, 0.99, 0.95)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            st2.compute_returns(l2, 0.99, 0.95)
        g2.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            g2.replay()
        b.record()
        torch.cuda.synchronize()
        m2 = a.elapsed_time(b) / 20
        gae_sizes[str(n2)] = dict(ms=m2, gbs=604 * n2 / (m2 * 1e-3) / 1e9, frac=604 * n2 / (m2 * 1e-3) / 1e9 / peak)
        del st2, g2
    alg.update()
    st.step = T
    torch.cuda.synchronize()
    a.record()
    alg.update()
    b.record()
    torch.cuda.synchronize()
    ms_upd = a.elapsed_time(b)
    # policy forward: tcgen05 fused kernel vs torch (cuBLAS) on the rollout batch and on a minibatch-sized batch
    mlp = {}
    for B_ in (num_envs, T * num_envs // 4):
        xb = torch.randn(B_, 48, device=device)
        with torch.no_grad():
            for _ in range(5):
                ac._fused_actor(xb)
                ac.actor(xb)
            torch.cuda.synchronize()
            a.record()
            for _ in range(50):
                ac._fused_actor(xb)
            b.record()
            torch.cuda.synchronize()
            t_f = a.elapsed_time(b) / 50
            a.record()
            for _ in range(50):
                ac.actor(xb)
            b.record()
            torch.cuda.synchronize()
            t_t = a.elapsed_time(b) / 50
        flops = 2.0 * B_ * (48 * 128 + 128 * 64 + 64 * 32 + 32 * 12)
        mlp[str(B_)] = dict(fused_tcgen05_ms=t_f, torch_ms=t_t, fused_tflops=flops / (t_f * 1e-3) / 1e12)
    return dict(config="rollout storage GAE + normalisation; PPO update 5 epochs x 4 minibatches (nets 48-128-64-32)", num_envs=num_envs,
                actor_forward=mlp, gae_sizes=gae_sizes,
                T=T, gae_ms=ms_gae, gae_env_steps_per_s=num_envs * T / (ms_gae * 1e-3),
                gae_gbs=604 * num_envs / (ms_gae * 1e-3) / 1e9, gae_frac=604 * num_envs / (ms_gae * 1e-3) / 1e9 / peak,
                update_ms=ms_upd, update_samples_per_s=5 * T * num_envs / (ms_upd * 1e-3),
                note="GAE at 4096 envs moves 2.5 MB: launch-latency bound; update = ONE replay of a captured graph of all 20 minibatch steps "
                     "(row gather -> chained tcgen05 forward / loss / dgrad -> grouped wgrad GEMM -> fused clip + Adam + fp16 repack)")


def cpu_port_baselines(quick=False):
    """The reference's torch CPU path beside cfg 3 and the trajectory env (measured N only, no extrapolation): the oracle
    ports (torch.rand like the reference) on all host cores.  The only function of this file, with rom_per_call, that touches oracle/."""
    import legged_case as LC
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    for key, case_name, n, steps in (("cfg3_rough_lstm_cpu_port", "rough_lstm_allterms", 4096 if quick else 16384, 5),
                                     ("next1_trajectory_env_cpu_port", "traj_flat_allterms", 4096, 20)):
        case = LC.build_case(case_name, n, frames=4)
        port, phys = LC.make_port(case, rng="torch")
        if case.traj:
            port.gen.reset_traj(torch.arange(n), port.proj_z())
        for s in range(2):
            port.step(case.tape.actions[s % 4], phys)
        t0 = time.perf_counter()
        for s in range(steps):
            port.step(case.tape.actions[s % 4], phys)
        dt = time.perf_counter() - t0
        out[key] = dict(num_envs=n, steps=steps, ms_per_step=1e3 * dt / steps, env_steps_per_s=n * steps / dt, cores=torch.get_num_threads(),
                        kind="port", sample=f"oracle port of case {case_name}, {n} envs x {steps} steps, {dt:.2f} s")
    # the generator step over the 6-state rom class and the Hopper torque law, torch on the host cores
    from oracle.port_rom import GenPort, gen_params
    from oracle.port_hopper import hopper_case, hopper_torques as port_hopper_torques
    n, steps = 16384 if quick else 65536, 50
    gen = GenPort(gen_params(n, "ExtendedLateralUnicycle", N=10, t_low=1.0, t_high=2.0, freq_high=2.0, prob_stationary=0.0005), rng="torch")
    gen.reset(torch.randn(n, 6) * 0.3)
    t0 = time.perf_counter()
    for _ in range(steps):
        gen.step()
    dt = time.perf_counter() - t0
    out["next4_rom_family_step_cpu_port"] = dict(num_envs=n, steps=steps, ms_per_step=1e3 * dt / steps, env_steps_per_s=n * steps / dt,
                                                 cores=torch.get_num_threads(), kind="port",
                                                 sample=f"oracle GenPort over ExtendedLateralUnicycle, {n} envs x {steps} steps, {dt:.2f} s")
    case, act = hopper_case(n, seed=1, torque_limits=[9000.0, 80.0, 80.0, 80.0])
    port_hopper_torques(case, act)
    t0 = time.perf_counter()
    for _ in range(20):
        port_hopper_torques(case, act)
    dt = time.perf_counter() - t0
    out["next3_hopper_torques_cpu_port"] = dict(num_envs=n, calls=20, ms_per_call=1e3 * dt / 20, env_calls_per_s=n * 20 / dt, cores=torch.get_num_threads(),
                                                kind="port", sample=f"oracle port of Hopper._compute_torques, {n} envs x 20 calls, {dt:.2f} s")
    return out


def rom_family_step(num_envs=1 << 20, cls="ExtendedLateralUnicycle", steps=100, device="cuda", peak=6535.7):
    """stand-alone TrajectoryGenerator.step over the 6-state / 3-input rom class through the generic kernels (W = 10 knots,
    dt_loop 0.02 / rom.dt 0.1: one call in five appends a knot).  Algorithmic bytes per env and call: 4 (P + n + 1 + m) + 8 ((W+1) n + W m) / 5."""
    from legged_gym_dev_b200 import rom as R
    zmax, vmax = {"ExtendedLateralUnicycle": ([1e9, 1e9, 1e9, 1.0, 0.5, 2.0], [1.0, 0.6, 4.0]), "Unicycle": ([1e9] * 3, [1.0, 2.0])}[cls]
    rom = R.ROM_CLASSES[cls](0.1, [-a for a in zmax], zmax, [-a for a in vmax], vmax, n_robots=num_envs, device=device)
    gen = R.TrajectoryGenerator(rom, R.UniformSampleHoldDT(1.0, 2.0), R.UniformWeightSampler(), dt_loop=0.02, N=10, freq_low=0.01, freq_high=2.0, seed=1,
                                device=device, prob_stationary=0.0005)
    z0 = torch.randn(num_envs, rom.n, device=device) * 0.3
    for _ in range(3):
        gen.reset(z0)
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(5):
        gen.reset(z0)
    b.record()
    torch.cuda.synchronize()
    reset_ms = a.elapsed_time(b) / 5
    for _ in range(10):
        gen.step()
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        gen.step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    n, m, W = rom.n, rom.m, 10
    per = 4 * (9 * m + 6 + n + 1 + m) + 8 * ((W + 1) * n + W * m) / 5
    return dict(workload=f"TrajectoryGenerator.step over {cls}, {num_envs} envs", ms_per_step=ms, env_steps_per_s=num_envs / ms * 1e3,
                algorithmic_bytes_per_env=per, achieved_gbs=per * num_envs / ms / 1e6, frac=per * num_envs / ms / 1e6 / peak,
                reset_ms=reset_ms, reset_bytes_per_env=4 * (n + 2 * (9 * m + 6) + (W + 1) * n + W * m),
                reset_frac=4 * (n + 2 * (9 * m + 6) + (W + 1) * n + W * m) * num_envs / reset_ms / 1e6 / peak)


def hopper_torques(num_envs=1 << 20, steps=50, device="cuda", peak=6535.7):
    """Hopper._compute_torques, 188 algorithmic bytes per env (csrc/hopper.cu)."""
    from legged_gym_dev_b200.hopper import HopperActuation
    env = HopperActuation(num_envs, device=device, torque_limits=[9000.0, 80.0, 80.0, 80.0])
    g = torch.Generator(device=device).manual_seed(1)
    for t in (env.dof_state, env.root_states, env.base_ang_vel, env.contact_forces):
        t.normal_(generator=g)
    act = torch.randn(num_envs, 4, device=device, generator=g)
    for _ in range(5):
        env._compute_torques(act)
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(steps):
        env._compute_torques(act)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return dict(workload=f"Hopper._compute_torques, {num_envs} envs", ms_per_call=ms, env_calls_per_s=num_envs / ms * 1e3,
                algorithmic_bytes_per_env=188, achieved_gbs=188 * num_envs / ms / 1e6, frac=188 * num_envs / ms / 1e6 / peak)


def hopper_env_step(num_envs=1 << 20, steps=32, device="cuda", peak=6535.7):
    """HopperTrajectory.step = 4 x torque law + generator step + prologue / fused post-physics / finaliser + generator reset,
    yaml reward table (11 terms).  Synthetic per-sub-step state generated on the device.  Algorithmic bytes per env: 4 x 188 (torque law)
    + 682 (post-physics, csrc/hopper_env.cu header) + generator step."""
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.hopper_trajectory import HopperTrajectory
    cfg = configs.hopper_flat_trajectory_cfg()
    cfg.env.num_envs = num_envs
    for k, v in configs.HOPPER_YAML_REWARD_SCALES.items():
        setattr(cfg.rewards.scales, k, v)

    class DevicePhysics:   # fixed random state tensors: the kernels' traffic does not depend on the values
        def __init__(self):
            g = torch.Generator(device=device).manual_seed(0)
            self.root_states = torch.randn(num_envs, 13, device=device, generator=g)
            self.root_states[:, 3:7] = torch.nn.functional.normalize(self.root_states[:, 3:7] + torch.tensor([0, 0, 0, 3.0], device=device), dim=-1)
            self.dof_state = torch.randn(num_envs, 4, 2, device=device, generator=g)
            self.contact_forces = torch.zeros(num_envs, 5, 3, device=device)
            self.contact_forces[:, 4, 2] = 100.0 * (torch.rand(num_envs, device=device, generator=g) > 0.5)

        def simulate(self, torques):
            pass

        def commit_resets(self, reset_buf):
            pass
    env = HopperTrajectory(cfg, SimpleNamespace(dt=0.005), None, device, True, physics=DevicePhysics(), seed=0)
    env.episode_length_buf.copy_(torch.randint(0, 1000, (num_envs,), device=device))
    env.reset_traj_all()
    act = torch.randn(num_envs, 4, device=device) * 0.2 + torch.tensor([1.0, 0, 0, 0], device=device)
    for _ in range(4):
        env.step(act)
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(steps):
        env.step(act)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    a.record()
    for _ in range(steps):
        env.post_physics_step()
    b.record()
    torch.cuda.synchronize()
    ms_pp = a.elapsed_time(b) / steps
    per_pp = 682 + 2 * (88 + 80 + 4 * 4 + 16 + 9 * 8 + 4 + 1 + 4) + 80
    return dict(workload=f"HopperTrajectory.step, yaml reward table, {num_envs} envs", ms_per_step=ms, env_steps_per_s=num_envs / ms * 1e3,
                post_physics_ms=ms_pp, post_physics_bytes_per_env=per_pp, post_physics_frac=per_pp * num_envs / ms_pp / 1e6 / peak,
                step_bytes_per_env=4 * 188 + per_pp, step_frac=(4 * 188 + per_pp) * num_envs / ms / 1e6 / peak,
                finite=bool(torch.isfinite(env.obs_buf).all()) and bool(torch.isfinite(env.rew_buf).all()))


def run_all(device="cuda", peak=6535.7, quick=False):
    out = {}
    for name, fn, kw in (("cfg1_rom_per_call", rom_per_call, dict(device=device, loop_steps=200 if quick else 1000)),
                         ("cfg3_rough_lstm", rough_lstm, dict(device=device, peak=peak)),
                         ("cfg3_rough_lstm_262144", rough_lstm, dict(device=device, peak=peak, num_envs=16384 if quick else 262144, steps=20)),
                         ("next1_trajectory_env_4096", trajectory_env, dict(device=device, peak=peak)),
                         ("next1_trajectory_env_262144", trajectory_env, dict(device=device, peak=peak, num_envs=16384 if quick else 262144, steps=20)),
                         ("next1_trajectory_env_1048576", trajectory_env, dict(device=device, peak=peak, num_envs=16384 if quick else (1 << 20), steps=20)),
                         ("cfg4_rom_rollout", rom_rollout, dict(device=device, peak=peak, num_envs=(1 << 17) if quick else (1 << 20))),
                         ("cfg4b_tube_dataset", tube_dataset, dict(device=device, peak=peak, num_envs=16384 if quick else 262144)),
                         ("cfg5_gae_update", gae_update, dict(device=device, peak=peak)),
                         ("next4_rom_family_step", rom_family_step, dict(device=device, peak=peak, num_envs=(1 << 17) if quick else (1 << 20))),
                         ("next3_hopper_torques", hopper_torques, dict(device=device, peak=peak, num_envs=(1 << 17) if quick else (1 << 20))),
                         ("next3_hopper_env_step", hopper_env_step, dict(device=device, peak=peak, num_envs=(1 << 17) if quick else (1 << 20)))):
        try:
            out[name] = fn(**kw)
        except Exception as e:   # an extra must never take the headline line down
            out[name] = dict(error=f"{type(e).__name__}: {e}")
        torch.cuda.empty_cache()
    try:
        out.update(cpu_port_baselines(quick=quick))
    except Exception as e:
        out["cpu_port_baselines"] = dict(error=f"{type(e).__name__}: {e}")
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run_all(quick="--quick" in sys.argv), indent=1))
