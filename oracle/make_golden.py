"""Writes tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/ref_harness.py).

Build-container only (needs /root/reference).  Usage:  python -m oracle.make_golden
Each file holds the inputs of one small case (replay tape, counters, terrain) and, for every step,
the reference's outputs.  tests/test_golden.py replays them through the oracle port (CPU) and the fused
CUDA path (GPU).  The reference's own tests hold no golden vectors for this path, so
these reference-generated fixtures are what pins the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as H          # noqa: E402
import legged_case as LC                     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def legged_golden(name, num_envs=32, frames=4, steps=8):
    case = LC.build_case(name, num_envs, frames=frames)
    task, rs, cr, lstm, over = LC.CASES[name]
    if case.rough:   # small field so the fixture stays small; a few robots outside it exercise the clip path
        hf = case.terrain["height_samples"][:420, :520].contiguous()
        case.terrain["height_samples"] = hf
        case.tape.root[..., 0] = case.tape.root[..., 0] / 80.0 * 20.0 - 26.0
        case.tape.root[..., 1] = case.tape.root[..., 1] / 160.0 * 30.0 - 26.0
        if case.traj:   # robots stay next to their origins: move the origins the same way
            case.terrain["env_origins"][:, 0] = case.terrain["env_origins"][:, 0] / 80.0 * 20.0 - 26.0
            case.terrain["env_origins"][:, 1] = case.terrain["env_origins"][:, 1] / 160.0 * 30.0 - 26.0
    if case.traj:
        env = H.make_reference_anymal_trajectory(task, num_envs, case.tape, seed=case.seed, reward_scales=rs, use_actuator_network=lstm,
                                                 heightfield=case.terrain["height_samples"] if case.rough else None,
                                                 terrain_origins=case.terrain["terrain_origins"] if case.rough else None,
                                                 episode_lengths=case.ep, time_until_next_push=case.tpush, overrides=over)
    else:
        env = H.make_reference_anymal(task, num_envs, case.tape, seed=case.seed, reward_scales=rs, command_ranges=cr,
                                      use_actuator_network=lstm, heightfield=case.terrain["height_samples"] if case.rough else None,
                                      terrain_origins=case.terrain["terrain_origins"] if case.rough else None,
                                      episode_lengths=case.ep, overrides=over)
    if case.rough:
        if not case.traj:
            env.terrain_levels[:] = case.terrain["terrain_levels"]
            env.terrain_types[:] = case.terrain["terrain_types"]
        env.env_origins[:] = case.terrain["env_origins"]
    if case.traj:
        env.reset_traj(torch.arange(num_envs))     # as reset() does (see tests/test_oracle_cpu.py)
    out = dict(tape_root=case.tape.root.numpy(), tape_dof=case.tape.dof.numpy(), tape_contact=case.tape.contact.numpy(),
               tape_actions=case.tape.actions.numpy(), ep=case.ep.numpy(), seed=np.int64(case.seed),
               num_envs=np.int64(num_envs), frames=np.int64(frames), steps=np.int64(steps))
    if case.rough:
        for k, v in case.terrain.items():
            out["terrain_" + k] = v.numpy()
    if case.traj:
        out["tpush"] = case.tpush.numpy()
    for s in range(steps):
        a = case.tape.actions[s % frames] * (150.0 if s == 3 else 1.0)
        env.step(a.clone())
        snap = dict(obs=env.obs_buf, rew=env.rew_buf, reset=env.reset_buf, time_out=env.time_out_buf, torques=env.torques,
                    ep_len=env.episode_length_buf, feet_air_time=env.feet_air_time,
                    last_contacts=env.last_contacts, root=env.root_states, dof=env.dof_state,
                    last_dof_vel=env.last_dof_vel, last_root_vel=env.last_root_vel, base_lin_vel=env.base_lin_vel)
        if case.traj:
            tg = env.traj_gen
            snap.update(trajectory=env.trajectory, prev_error=env.prev_error, time_until_next_push=env.time_until_next_push,
                        gen_trajectory=tg.trajectory, gen_v_trajectory=tg.v_trajectory, gen_t=tg.t, gen_k=tg.k,
                        gen_t_final=tg.t_final, gen_weights=tg.weights, gen_stationary=tg.stationary_inds)
        else:
            snap["commands"] = env.commands
        for k, v in env.episode_sums.items():
            snap["sum_" + k] = v
        if case.rough:
            snap["heights"] = env.measured_heights
            if not case.traj:
                snap["terrain_levels"] = env.terrain_levels
        if lstm:
            snap["lstm_h"], snap["lstm_c"] = env.sea_hidden_state, env.sea_cell_state
        for k, v in env.extras.get("episode", {}).items():
            snap["extras_" + k] = torch.as_tensor(v)
        for k, v in snap.items():
            out[f"s{s}_{k}"] = v.detach().cpu().numpy().copy()
    path = os.path.join(GOLD, f"legged_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; resets:",
          sum(int(out[f"s{s}_reset"].sum()) for s in range(steps)))


def main():
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    for name in ("flat_pd_upstream", "flat_lstm_shipped", "flat_allterms_v", "rough_lstm_allterms",
                 "traj_flat_allterms", "traj_flat_lstm_shipped", "traj_rough_lstm_allterms"):
        if not only or name in only:
            legged_golden(name, steps=12 if name.startswith("traj") else 8)
    if only:
        return
    try:
        from oracle.make_golden_rom import main as rom_main
        rom_main()
    except ImportError:
        pass


if __name__ == "__main__":
    main()
