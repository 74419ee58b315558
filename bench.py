#!/usr/bin/env python
"""Throughput bench of the fused step pipeline (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): anymal_c_flat env step —
4 x PD torques + the fused post-physics pass — upstream reward table (SURVEY.md §8d cfg 2b), on a replayed
synthetic state tape, `--envs` environments per GPU (default 1 048 576: the 1M-env end of the metric's range on ONE
B200; the `sweep` key carries 4096 ... 1M including 131 072 = the per-GPU shard of 1M envs on 8 GPUs).  One "step" = one
`env.step(actions)` for every env.  `value` = env-steps/s with the tape resident in HBM; `e2e` = the same
step driven through the public API from PINNED HOST buffers (H2D of the step's physics state + actions, D2H
of obs/rew/reset inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "env-steps/s of step pipeline (anymal_c_flat: 4x PD torques + fused post_physics_step)"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=1048576, help="environments per GPU")
    ap.add_argument("--frames", type=int, default=8, help="frames of the replay tape")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-envs", type=int, default=16384, help="envs of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager Python API instead of CUDA-graph replays of tape cycles")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 4096..1M envs sweep (extra key, not the headline)")
    ap.add_argument("--no-extra", action="store_true", help="skip the timings of configs 1, 3, 4, 5 (extra key)")
    ap.add_argument("--no-multi", action="store_true", help="skip the sharded configs 4 / 2 (1M envs total) / 5 (multi_gpu key, every N)")
    return ap.parse_args()


def workload_cfg(num_envs):
    from legged_gym_dev_b200 import configs
    cfg = configs.with_upstream_rewards(configs.anymal_c_flat_cfg(), pd_control=True)
    cfg.env.num_envs = num_envs
    return cfg


def algorithmic_bytes(num_sum_rows, num_bodies=17):
    """every API-visible tensor counted once per read and once per write, per env."""
    K = num_sum_rows
    rd = 52 + 96 + 12 * num_bodies + 48 + 48 + 48 + 48 + 16 + 16 + 4 + 8 + 4 * K
    wr = 36 + 16 + 16 + 4 + 8 + 2 + 4 + 4 * K + 192 + 48 + 48 + 24
    post = rd + wr
    pd = 48 + 96 + 48
    return dict(post_physics=post, pd_torques=pd, step=96 + 4 * pd + post)


def measured_traffic(num_envs):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, scaled per env (None if absent)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)["post_physics_kernel<32,false>"]
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["envs"] * num_envs
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def build_env(num_envs, frames, device, rank, copy=False, host=False):
    from types import SimpleNamespace
    from legged_gym_dev_b200 import synthetic as S
    from legged_gym_dev_b200.legged_robot import Anymal
    from legged_gym_dev_b200.physics import ReplayPhysics, HostReplayPhysics
    cfg = workload_cfg(num_envs)
    tape = S.make_state_tape(num_envs, frames=frames, seed=100 + rank, device=device)
    if host:   # generated on the device (fast), then moved to the host: the e2e arm replays it from pinned host memory
        for k in ("root", "dof", "contact", "actions"):
            setattr(tape, k, getattr(tape, k).cpu())
        torch.cuda.empty_cache()
    phys = HostReplayPhysics(tape, device=device) if host else ReplayPhysics(tape, device=device, copy=copy)
    lim = S.anymal_dof_limits()
    env = Anymal(cfg, SimpleNamespace(dt=cfg.sim.dt), None, device, True, physics=phys, asset=lim, seed=0,
                 env_id_offset=rank * num_envs)
    env.episode_length_buf.copy_(S.make_episode_lengths(num_envs, seed=rank, device=device))
    return env, tape


def time_steps(env, actions, steps, warmup, world, device, graphed=None):
    """Times EXACTLY `steps` env steps.  graphed: GraphedReplay — whole tape cycles are replayed from one CUDA graph,
    the remainder (steps % F) runs through the eager API at the end."""
    import torch.distributed as dist
    F = len(actions)
    if graphed is not None:
        for _ in range((warmup + F - 1) // F):
            graphed.replay()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps // F):
            graphed.replay()
        for s in range(steps % F):
            env.step(actions[s])
        t1.record()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    for s in range(warmup):
        env.step(actions[s % F])
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(steps):
        env.step(actions[(warmup + s) % F])
    t1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def kernel_time(env, actions, steps, device):
    """Average duration of the dominant kernel (post_physics) and of pd_torques, CUDA events on the launch stream."""
    F = len(actions)
    env._timing = []
    for s in range(steps):
        env.step(actions[s % F])
    torch.cuda.synchronize(device)
    pp = [a.elapsed_time(b) for (k, a, b) in env._timing if k == "post_physics"]
    pd = [a.elapsed_time(b) for (k, a, b) in env._timing if k == "torques"]
    env._timing = None
    return sum(pp) / len(pp), sum(pd) / len(pd)


def graph_kernel_times(env, actions, reps=20):
    """Average duration of pd_torques and of the fused post-physics pass over BACK-TO-BACK launches replayed from a CUDA graph
    (4 F torque launches, resp. F post-physics passes over the tape frames): the per-launch figure without the ~5 us of host
    launch gap that event-bracketing a single eager launch includes — what matters below ~256 K envs per GPU."""
    dev, F = env.device, len(actions)
    ph = env.physics
    env.use_device_step_counter()
    out = {}
    for name in ("pd_torques", "post_physics"):
        def body():
            for f in range(F):
                if name == "pd_torques":
                    for i in range(env.params.decimation):
                        env._compute_torques(actions[f], write_clipped=(i == 0))
                        ph.simulate(env.torques)
                    ph.refresh()
                else:
                    env.post_physics_step()
        c0, f0 = env.common_step_counter, ph.frame
        env._stream = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        env.common_step_counter, ph.frame, ph.sub = c0, f0, 0
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize(dev)
        n_launch = reps * F * (env.params.decimation if name == "pd_torques" else 1)
        out[name + "_ms"] = a.elapsed_time(b) / n_launch
    return out


def e2e_run(num_envs, frames, steps, warmup, device, rank, world):
    """Public-API step with HOST-resident inputs: every sub-step's dof_state, the root/contact frame and the
    actions come from pinned host memory; obs/rew/reset are read back to pinned host memory every step."""
    import torch.distributed as dist
    env, tape = build_env(num_envs, frames, device, rank, host=True)
    acts = [tape.actions[f].pin_memory() for f in range(frames)]
    obs_h = torch.empty(num_envs, env.num_obs, pin_memory=True)
    rew_h = torch.empty(num_envs, pin_memory=True)
    rst_h = torch.empty(num_envs, dtype=torch.bool, pin_memory=True)
    a_dev = torch.empty(num_envs, 12, device=device)

    out_stream = torch.cuda.Stream(device=device)
    done_ev, read_ev = torch.cuda.Event(), torch.cuda.Event()

    def one(s):
        cur = torch.cuda.current_stream(device)
        a_dev.copy_(acts[s % frames], non_blocking=True)
        cur.wait_event(read_ev)                     # previous results have left the device buffers
        obs, _, rew, rst, _ = env.step(a_dev)
        done_ev.record(cur)
        with torch.cuda.stream(out_stream):         # D2H of the results overlaps the next step's H2D (PCIe is full duplex)
            out_stream.wait_event(done_ev)
            obs_h.copy_(obs, non_blocking=True)
            rew_h.copy_(rew, non_blocking=True)
            rst_h.copy_(rst, non_blocking=True)
            read_ev.record(out_stream)

    for s in range(warmup):
        one(s)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        one(warmup + s)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    h2d = num_envs * (4 * 96 + 52 + 204 + 48)
    d2h = num_envs * (env.num_obs * 4 + 4 + 1 + 4 * 48)     # obs, rew, reset + the torques of the 4 sub-steps (what a host-side physics needs back)
    return dt, h2d, d2h


def cpu_baseline(num_envs, steps, warmup=3, port_too=True):
    """The reference's own LeggedRobot.step (unmodified sources staged in oracle/_ref by __graft_entry__.build(), driven through
    oracle/ref_harness.py with the replay tape standing in for Isaac Gym and its stock torch.rand draws) on the host cores:
    kind "reference".  The oracle port (oracle/port_legged.py) is timed beside it (`port_value`), and alone when no reference tree
    is present (kind "port")."""
    import legged_case as LC
    torch.set_num_threads(os.cpu_count() or 1)
    case = LC.build_case("flat_pd_upstream", num_envs, frames=4, base_contact_prob=0.002)

    def timed(step):
        for s in range(warmup):
            step(case.tape.actions[s % 4])
        t0 = time.perf_counter()
        n = 0
        while n < steps:
            step(case.tape.actions[n % 4])
            n += 1
            if time.perf_counter() - t0 > 45.0:
                break
        return n, time.perf_counter() - t0

    out, ref_err = None, None
    try:
        from oracle import ref_harness as H
        if H.reference_available():
            task, rs, cr, lstm, over = LC.CASES["flat_pd_upstream"]
            env = H.make_reference_anymal(task, num_envs, case.tape, seed=case.seed, reward_scales=rs, command_ranges=cr,
                                          use_actuator_network=lstm, episode_lengths=case.ep, overrides=over)
            del env._shim_seed                      # stock code path: the reference draws from torch.rand itself
            n, dt = timed(lambda a: env.step(a.clone()))
            out = dict(value=num_envs * n / dt, unit=UNIT, cores=torch.get_num_threads(), kind="reference",
                       sample=f"the reference's own LeggedRobot.step (legged_gym/envs/base/legged_robot.py:80-104, staged unmodified in "
                              f"oracle/_ref), anymal_c_flat PD + upstream rewards, {num_envs} envs x {n} steps, {dt:.2f} s",
                       ms_per_step=1e3 * dt / n)
    except Exception as e:   # noqa: BLE001 — the port below still gives a CPU figure
        ref_err = f"{type(e).__name__}: {e}"
    if out is None or port_too:
        port, phys = LC.make_port(case, rng="torch")
        n, dt = timed(lambda a: port.step(a, phys))
        pv = dict(value=num_envs * n / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                  sample=f"oracle/port_legged.py (torch CPU restatement of the reference), anymal_c_flat PD + upstream rewards, "
                         f"{num_envs} envs x {n} steps, {dt:.2f} s", ms_per_step=1e3 * dt / n)
        if out is None:
            out = pv
            if ref_err:
                out["reference_error"] = ref_err
        else:
            out["port_value"] = pv["value"]
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_baseline(args.cpu_envs, max(1, min(args.steps, 200)), warmup=max(3, min(args.warmup, 10)), port_too=False)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"anymal_c_flat env.step (4x PD torques + post_physics_step), upstream reward table, "
                                       f"replayed synthetic state; reference's torch CPU path ({cb['kind']}) on a bounded sample of {args.cpu_envs} envs"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "reference_error") if k in cb},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the b200gym path has no CPU fallback)")
    import torch.distributed as dist
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from legged_gym_dev_b200 import _lib
    _lib.lib()

    N = args.envs
    env, tape = build_env(N, args.frames, device, rank)
    actions = [tape.actions[f].to(device) for f in range(args.frames)]
    graphed = None
    if not args.no_graph:
        from legged_gym_dev_b200.graphs import GraphedReplay
        graphed = GraphedReplay(env, actions)
    with ClockSampler(local_rank) as clk:   # sampled over every timed loop of the same workload (eager, per-kernel, graph replay)
        ms_eager = time_steps(env, actions, args.steps, args.warmup, world, device)
        t_pp, t_pd = kernel_time(env, actions, min(args.steps, 50), device)
        ms = time_steps(env, actions, args.steps, args.warmup, world, device, graphed=graphed)
        if len(clk.samples) < 3:            # short runs: keep the same load up until a few samples exist
            t_end = time.perf_counter() + 0.6
            while time.perf_counter() < t_end and len(clk.samples) < 3:
                time_steps(env, actions, 8 * len(actions), 0, 1, device, graphed=graphed)
    value = world * N * args.steps / (ms * 1e-3)
    ab = algorithmic_bytes(len(env.params.active_terms))
    peak, peak_src = measured_peak()
    ach = ab["post_physics"] * N / (t_pp * 1e-3) / 1e9
    ach_pd = ab["pd_torques"] * N / (t_pd * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        env = tape = graphed = None
        torch.cuda.empty_cache()
        esteps = max(5, min(args.steps, 30 if N <= 262144 else 12))
        dt, h2d, d2h = e2e_run(N, 2 if N > 262144 else min(args.frames, 4), esteps, 3, device, rank, world)
        e2e = {"value": world * N * esteps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": esteps, "note": "LeggedRobot.step through the C ABI with pinned-host physics frames + actions in; obs / rew / reset and the 4 sub-steps' torques out"}

    multi = None
    if not args.no_multi:   # every rank: the sharded configurations of BASELINE.json (strong scaling), N = 1 included
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_multigpu
        env = tape = graphed = None
        torch.cuda.empty_cache()
        multi = bench_multigpu.run_all(rank, world, device)

    sweep = None
    extra = None
    if world == 1 and not args.no_sweep:
        sweep = {}
        for n in (4096, 16384, 65536, 131072, 262144, 1048576):
            e2, tp2 = build_env(n, 4, device, 0)
            a2 = [tp2.actions[f].to(device) for f in range(4)]
            m2 = time_steps(e2, a2, 100, 10, 1, device)
            tpp, tpd = kernel_time(e2, a2, 30, device)
            from legged_gym_dev_b200.graphs import GraphedReplay
            m3 = time_steps(e2, a2, 96, 8, 1, device, graphed=GraphedReplay(e2, a2))
            gk = graph_kernel_times(e2, a2)   # back-to-back launches replayed from a graph: no host launch gap in the figure
            sweep[str(n)] = {"env_steps_per_s": n * 100 / (m2 * 1e-3), "ms_per_step": m2 / 100, "graph_ms_per_step": m3 / 96,
                             "graph_env_steps_per_s": n * 96 / (m3 * 1e-3), "post_physics_ms": tpp,
                             "post_physics_frac": ab["post_physics"] * n / (tpp * 1e-3) / 1e9 / peak,
                             "pd_torques_ms": tpd, "pd_torques_frac": ab["pd_torques"] * n / (tpd * 1e-3) / 1e9 / peak,
                             "post_physics_graph_ms": gk["post_physics_ms"], "pd_torques_graph_ms": gk["pd_torques_ms"],
                             "post_physics_graph_frac": ab["post_physics"] * n / (gk["post_physics_ms"] * 1e-3) / 1e9 / peak,
                             "pd_torques_graph_frac": ab["pd_torques"] * n / (gk["pd_torques_ms"] * 1e-3) / 1e9 / peak,
                             "step_frac": ab["step"] * n * 96 / (m3 * 1e-3) / 1e9 / peak}
            del e2, tp2, a2
            torch.cuda.empty_cache()
    if world == 1 and not args.no_extra:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        extra = bench_configs.run_all(device=device, peak=peak)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args.cpu_envs, args.cpu_steps)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"anymal_c_flat env.step (4x PD torques + fused post_physics_step), upstream reward table "
                                       f"(SURVEY 8d cfg 2b), {N} envs/GPU, {args.frames}-frame replay tape resident in HBM",
                           "envs_per_gpu": N, "total_envs": world * N, "parallelism": f"env-sharded x{world}, no data-path collective",
                           "launch": ("eager Python API (5 launches + finaliser per env step)" if graphed is None else
                                      f"CUDA graph of one {args.frames}-step tape cycle replayed; step counter / RNG event in device memory"),
                           "eager_ms_per_step": ms_eager / args.steps,
                           "cache": f"inputs larger than L2: {ab['step'] * N / 1e6:.0f} MB touched per step, tape cycles {args.frames} frames"},
                "gpu_launches": args.steps * (env_launches()),
                "roofline": {"bound": "hbm", "kernel": "post_physics_kernel<32,false> (+ extras_finalize_kernel)", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "traffic": measured_traffic(N),
                             "traffic_note": "ncu dram read+write bytes of one launch at 1 048 576 envs, scaled per env (profiles/r1_traffic.json)",
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_env": ab["post_physics"], "avg_launch_ms": t_pp,
                             "pd_torques": {"achieved": ach_pd, "frac": ach_pd / peak, "avg_launch_ms": t_pd,
                                            "algorithmic_bytes_per_env": ab["pd_torques"]},
                             "step_frac": ab["step"] * N * args.steps / (ms * 1e-3) / 1e9 / peak},
                "clocks": clk.summary()}
        if e2e:
            line["e2e"] = e2e
        if cb:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "port_value", "reference_error") if k in cb}
            if sweep and str(args.cpu_envs) in sweep:   # the GPU arm at the CPU arm's own size: a same-config ratio
                g = sweep[str(args.cpu_envs)]
                line["cpu_baseline"]["gpu_value_at_same_envs"] = g["graph_env_steps_per_s"]
                line["cpu_baseline"]["same_config_ratio"] = g["graph_env_steps_per_s"] / cb["value"]
        if sweep:
            line["sweep"] = sweep
        if extra:
            line["extra_configs"] = extra
        if multi:
            line["multi_gpu"] = multi
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def env_launches():
    return 6   # 4 torque launches + fused post-physics + its 1-warp extras finaliser per env step


if __name__ == "__main__":
    main()
