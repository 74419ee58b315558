"""tests/golden/romfam_reference.npz from the UNMODIFIED reference rom classes (trajopt/rom_dynamics.py:182-438) and its
TrajectoryGenerator (:441-615) over each of them (oracle/ref_harness.make_reference_generator).  Build-container only.
Usage: python -m oracle.make_golden_rom_family

Per class: the algebra (f, des_pose_vel, compute_state_dependent_input_bounds, clip_v_z with the torch backend; proj_z with the
numpy backend — the reference's unicycle proj_z mixes scipy arrays into torch.hstack and only runs on numpy) on seeded inputs, and
a generator run: reset(z0) then `steps` calls of step(), snapshots at the steps in `keep`.  The unicycle classes' f() builds a
[n_robots, n] buffer (:270), so the reference can only advance them with every env due on every call: those runs use
dt_loop == rom.dt (the lock-step use of trajopt/trajectory_gen.py); the integrator classes run with the shipped dt_loop < rom.dt."""
import os

import numpy as np
import torch

from oracle import ref_harness as H
from oracle.port_rom import FAMILY, gen_params

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CLASSES = ("SingleInt2D", "DoubleInt2D", "Unicycle", "LateralUnicycle", "ExtendedUnicycle", "ExtendedLateralUnicycle")
KEEP = (0, 1, 7, 30, 79)


def case_params(cls, num_envs=24, seed=7):
    return gen_params(num_envs, cls, seed=seed, **({} if cls not in FAMILY else dict(dt_loop=0.1)))


def algebra_inputs(cls, n, m, rows=64, seed=11):
    g = torch.Generator().manual_seed(seed + len(cls))
    z = torch.randn(rows, n, generator=g)
    v = torch.randn(rows, m, generator=g) * 1.5
    x = torch.randn(rows, 13, generator=g)
    x[:, 3:7] = x[:, 3:7] / torch.linalg.norm(x[:, 3:7], dim=1, keepdim=True)
    return z, v, x


def main():
    ref = H.import_reference()
    rd = ref.rom_dynamics
    out = dict(keep=np.array(KEEP))
    for cls in CLASSES:
        p = case_params(cls)
        tg, rom = H.make_reference_generator(p, seed=p.seed)
        z, v, x = algebra_inputs(cls, rom.n, rom.m)
        rom_rows = getattr(rd, cls)(p.rom_dt, rom.z_min, rom.z_max, rom.v_min, rom.v_max, n_robots=z.shape[0], backend="torch", device="cpu")
        pose, vel = rom_rows.des_pose_vel(z, v)
        lo, hi = rom_rows.compute_state_dependent_input_bounds(z)
        rom_np = getattr(rd, cls)(p.rom_dt, rom.z_min.numpy(), rom.z_max.numpy(), rom.v_min.numpy(), rom.v_max.numpy(), n_robots=z.shape[0],
                                  backend="numpy")
        a = dict(z=z, v=v, x=x, f=rom_rows.f(z, v), pose=pose, vel=vel, lo=lo, hi=hi, clip=rom_rows.clip_v_z(z, v))
        if cls != "ExtendedLateralUnicycle":
            # ExtendedLateralUnicycle.proj_z (:422-427) calls torch.squeeze on a numpy array (numpy backend) / multiplies a numpy matrix
            # into a torch tensor (torch backend): it raises on both, so there is no reference output to record for it.  The port
            # restates it as ExtendedUnicycle's (:360-365, np.squeeze) with both local velocity components.
            a["proj"] = np.asarray(rom_np.proj_z(x.numpy()))
        for k, t in a.items():
            out[f"{cls}_alg_{k}"] = t.numpy().copy() if torch.is_tensor(t) else np.asarray(t)
        out[f"{cls}_ramp_v_end0"] = tg.ramp_v_end.numpy().copy()
        g = torch.Generator().manual_seed(1)
        z0 = torch.randn(p.num_envs, rom.n, generator=g) * 0.3
        out[f"{cls}_z0"] = z0.numpy().copy()
        tg.reset(z0.clone())
        snap = lambda: dict(traj=tg.trajectory, vtraj=tg.v_trajectory, v=tg.v, t=tg.t, k=tg.k, t_final=tg.t_final, weights=tg.weights,
                            stationary=tg.stationary_inds, get_trajectory=tg.get_trajectory())
        for k, t in snap().items():
            out[f"{cls}_reset_{k}"] = t.detach().numpy().copy()
        for s in range(max(KEEP) + 1):
            tg.step()
            if s in KEEP:
                for k, t in snap().items():
                    out[f"{cls}_s{s}_{k}"] = t.detach().numpy().copy()
        out[f"{cls}_ctr"] = tg._shim.ctr.copy()
    path = os.path.join(GOLD, "romfam_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
