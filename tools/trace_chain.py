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
buf = torch.zeros(6 * 128 * 2, dtype=torch.int64, device="cuda")
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
tr = buf.cpu().view(6, 128, 2)
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
