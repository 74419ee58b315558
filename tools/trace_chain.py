"""Per-step cycle budget of ppo_chain_kernel from its built-in event trace (first and last CTA of one minibatch launch at the cfg-5
size: 4096 envs x 24 steps, 4 minibatches).  Codes: 100+4s MMA warp: A operand ready / +1 chain MMAs issued / +2 weight-gradient MMAs
issued; 200+it loader: stage issued; 300+4j loader: drain j starts waiting / +1 accumulator ready / +2 drained; 399 epilogue: input
rows gathered; 400+4s epilogue: waiting / +1 accumulator ready / +2 step done; 999 role finished."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import glob
import subprocess

# traced build of the library (every source with -DB200GYM_CHAIN_TRACE; only ppo_chain.cu reads it), loaded through B200GYM_LIB
import __graft_entry__ as G
tdir = os.path.join(ROOT, "build", "trace")
os.makedirs(tdir, exist_ok=True)
tlib = os.path.join(tdir, "libb200gym_trace.so")
srcs = sorted(glob.glob(os.path.join(ROOT, "legged_gym_dev_b200", "csrc", "*.cu")))
procs = [subprocess.Popen([G._nvcc(), *G.NVCC_FLAGS, "-DB200GYM_CHAIN_TRACE", "-c", s, "-o", os.path.join(tdir, os.path.basename(s)[:-3] + ".o")])
         for s in srcs]
assert all(p.wait() == 0 for p in procs)
subprocess.run([G._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tlib,
                *[os.path.join(tdir, os.path.basename(s)[:-3] + ".o") for s in srcs]], check=True)
os.environ["B200GYM_LIB"] = tlib
import torch
from legged_gym_dev_b200 import _lib
from legged_gym_dev_b200.ppo import ActorCritic, PPO

N, T = 4096, 24
torch.manual_seed(0)
ac = ActorCritic(48, 48, 12, actor_hidden_dims=(128, 64, 32), critic_hidden_dims=(128, 64, 32))
alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", desired_kl=0.01, entropy_coef=0.01, device="cuda")
alg.init_storage(N, T, [48], [None], [12])
st = alg.storage
st.observations.normal_()
st.actions.normal_()
st.sigma.fill_(1.0)
st.rewards.normal_(0.02, 0.05)
st.compute_returns(torch.randn(N, 1, device="cuda"), 0.99, 0.95)
alg.use_graph = False
alg.update()
st.step = T
L = _lib.lib()
L.b200gym_debug_chain_trace.argtypes = [C.c_void_p]
buf = torch.zeros(6 * 256 * 2 + 4 * 4096, dtype=torch.int64, device="cuda")
idx = torch.randperm(N * T, device="cuda")[:N * T // 4]
lp = _lib.PpoLossParamsPOD()
lp.batch, lp.num_actions, lp.use_clipped_value_loss = idx.numel(), 12, 1
lp.clip_param, lp.value_loss_coef, lp.entropy_coef, lp.inv_global_batch = 0.2, 1.0, 0.01, 1.0 / idx.numel()
sc = torch.zeros(4, dtype=torch.double, device="cuda")
std_off, _ = ac._slices["std"]
run = lambda: ac._trainer.minibatch_forward_backward(st, idx, lp, ac.std, C.c_void_p(ac.flat_grad.data_ptr() + 4 * std_off), sc, True)
run()
torch.cuda.synchronize()
assert L.b200gym_debug_chain_trace(buf.data_ptr()) == 0
run()
torch.cuda.synchronize()
L.b200gym_debug_chain_trace(None)
tr = buf.cpu()[:6 * 256 * 2].view(6, 256, 2)
gantt = buf.cpu()[6 * 256 * 2:].view(-1, 4)
for base, label in ((0, "CTA 0 (first round, co-resident)"), (3, "last CTA (second round)")):
    ev = []
    for r, role in enumerate(("EPI", "LOAD", "MMA")):
        ev += [(int(t), role, int(c)) for c, t in tr[base + r].tolist() if t != 0]
    ev.sort()
    if not ev:
        continue
    t0 = ev[0][0]
    print(f"== {label}: span {ev[-1][0] - t0} cycles")
    for t, role, c in ev:
        print(f"   {t - t0:7d}  {role:4s} {c}")

# CTA Gantt chart (globaltimer, ns): start / end of every CTA relative to the first start
g = gantt[gantt[:, 0] != 0]
t0 = int(g[:, 0].min())
n = g.shape[0]
print(f"== {n} CTAs; kernel span {int(g[:, 1].max()) - t0} ns")
half = n // 2
for name, sel in (("actor", g[:half]), ("critic", g[half:])):
    st, en = sel[:, 0] - t0, sel[:, 1] - t0
    d = en - st
    print(f"   {name}: start min/median/max {int(st.min())}/{int(st.median())}/{int(st.max())} ns, end min/median/max "
          f"{int(en.min())}/{int(en.median())}/{int(en.max())} ns, lifetime min/median/max {int(d.min())}/{int(d.median())}/{int(d.max())} ns")
late = g[g[:, 0] - t0 > 2000]
if late.shape[0]:
    ls = late[:, 0] - t0
    print(f"   CTAs starting later than 2 us after the first: {late.shape[0]}; start min/median/max {int(ls.min())}/{int(ls.median())}/{int(ls.max())} ns")
per_sm = {}
for s_, e_, sm, _ in g.tolist():
    per_sm.setdefault(sm, []).append((s_ - t0, e_ - t0))
cnt = sorted(len(v) for v in per_sm.values())
print(f"   SMs used {len(per_sm)}; CTAs per SM min/median/max {cnt[0]}/{cnt[len(cnt) // 2]}/{cnt[-1]}")
setup = g[:, 3] - g[:, 0]
print(f"   setup (kernel entry -> roles start) min/median/max {int(setup.min())}/{int(setup.median())}/{int(setup.max())} ns")
out = os.environ.get("B200GYM_GANTT_OUT")
if out:
    with open(out, "w") as f:
        f.write("cta,start_ns,roles_start_ns,end_ns,sm\n")
        for i, (s_, e_, sm, r_) in enumerate(gantt[:n].tolist()):
            f.write(f"{i},{s_ - t0},{r_ - t0},{e_ - t0},{sm}\n")
