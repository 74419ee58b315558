"""Counter-based RNG shared by the oracle and the CUDA kernels (TEST INFRASTRUCTURE side).

The reference draws every random number from torch's global generator with
batch-composition-dependent sizes (legged_gym/envs/base/legged_robot.py:371-384,423,443,449,459,226;
trajopt/rom_dynamics.py:480,520,529-545; deep_tube_learning/custom_sim.py:83-91), so "same inputs"
parity needs random numbers to be *inputs*.  Both sides therefore use Philox4x32-10 keyed as

    key     = (seed_lo, seed_hi)
    counter = (global_env_id, event_lo, event_hi, (site << 16) | block)

where `event` is the env-step counter (LeggedRobot: common_step_counter) or the per-env draw-event
counter (ROM generator), `site` names the draw site and `block` selects 4 consecutive columns of the
draw.  Column j of a draw is word j%4 of block j//4.  Uniforms are (x >> 8) * 2^-24 in [0, 1) (the
same 24-bit grid torch.rand uses for fp32); bounded ints are mulhi(x, bound).

This numpy implementation is the specification; legged_gym_dev_b200/csrc/philox.cuh must agree with
it bit for bit (tests/test_philox.py checks the known-answer vectors of the Random123 distribution).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
SH32 = np.uint64(32)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R. Inputs broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> SH32, p0 & MASK
        hi1, lo1 = p1 >> SH32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _words(seed, env_ids, event, site, ncols):
    """uint32 words [n, ncols] for the draw (env, event, site)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    event = np.broadcast_to(np.asarray(event, dtype=np.uint64).reshape(-1), env_ids.shape)
    nblk = (ncols + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    c3 = (np.uint64(site) << np.uint64(16)) | blk
    w = philox4x32(env_ids[:, None], (event & MASK)[:, None], (event >> SH32)[:, None], c3[None, :],
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.stack(w, axis=-1).reshape(env_ids.shape[0], nblk * 4)
    return out[:, :ncols]


def uniform01(seed, env_ids, event, site, ncols):
    """fp32 uniforms in [0,1), shape [len(env_ids), ncols]."""
    w = _words(seed, env_ids, event, site, ncols)
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def randint(seed, env_ids, event, site, ncols, bound):
    """int64 in [0, bound), shape [len(env_ids), ncols]: mulhi(x, bound)."""
    w = _words(seed, env_ids, event, site, ncols).astype(np.uint64)
    return ((w * np.uint64(bound)) >> SH32).astype(np.int64)


# ---- draw-site ids (must match legged_gym_dev_b200/csrc/philox.cuh) ----
# LeggedRobot step pipeline, event = common_step_counter (after the += 1 of legged_robot.py:115)
SITE_CMD_PERIODIC = 1     # cols 0,1,2 = lin_vel_x, lin_vel_y, heading-or-yaw   (legged_robot.py:371-384 via :350)
SITE_PUSH = 2             # cols 0,1                                           (legged_robot.py:459)
SITE_TERRAIN = 3          # col 0, bounded int                                 (legged_robot.py:482)
SITE_RESET_DOF = 4        # cols 0..num_dof-1                                  (legged_robot.py:423)
SITE_RESET_XY = 5         # cols 0,1                                           (legged_robot.py:443)
SITE_RESET_VEL = 6        # cols 0..5                                          (legged_robot.py:449)
SITE_CMD_RESET = 7        # cols 0,1,2                                         (legged_robot.py:371-384 via :166)
SITE_OBS_NOISE = 8        # cols 0..num_obs-1                                  (legged_robot.py:226)
SITE_PUSH_TIMER = 9       # col 0: time_until_next_push redraw                  (legged_robot_trajectory.py:175-178)
# ROM generator, event = per-env draw-event counter
SITE_ROM_INIT = 16        # ramp_v_end at construction                         (rom_dynamics.py:495)
SITE_ROM_ROOT = 17        # CustomSim.reset_idx root state                     (custom_sim.py:88-91)
SITE_ROM_DIST_MASK = 18   # reset_traj: rand(n) > zero_rom_dist_llh            (custom_sim.py:83)
SITE_ROM_DIST = 19        # reset_traj: offset U(-d, d)                        (custom_sim.py:84)
SITE_ROM_CONST = 20       # resample sub-sites, in reference draw order        (rom_dynamics.py:513-520)
SITE_ROM_RAMP = 21
SITE_ROM_EXTREME = 22
SITE_ROM_SIN_MAG = 23
SITE_ROM_SIN_MEAN = 24
SITE_ROM_SIN_FREQ = 25
SITE_ROM_SIN_OFF = 26
SITE_ROM_TFINAL = 27
SITE_ROM_WEIGHTS = 28
SITE_ROM_STATIONARY = 29
# HopperTrajectory reset / push, event = common_step_counter (oracle groundwork; no kernel consumes these yet)
SITE_HOP_DOF_POS = 32     # cols 0..3                                          (hopper_trajectory.py:306-309)
SITE_HOP_DOF_VEL = 33     # cols 0..3                                          (hopper_trajectory.py:310-313)
SITE_HOP_ROOT_POS = 34    # cols 0..4 = z, qx, qy, qz, qw offsets              (hopper_trajectory.py:338-341)
SITE_HOP_YAW = 35         # col 0                                              (hopper_trajectory.py:343)
SITE_HOP_ROOT_VEL = 36    # cols 0..5                                          (hopper_trajectory.py:351-354)
SITE_HOP_PUSH = 37        # cols 0..5                                          (hopper_trajectory.py:366-367)
SITE_POLICY_SAMPLE = 48   # cols 0..num_actions-1: Box-Muller pairs, event = act counter  (rsl_rl PPO.act -> Normal.sample)
