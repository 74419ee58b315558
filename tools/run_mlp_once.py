"""One shape of the fused MLP forward, a few launches (target for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from bench_mlp import net
from legged_gym_dev_b200.mlp import FusedMLP
B = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
m = net(48, (128, 64, 32), 12)
f = FusedMLP(m)
x = torch.randn(B, 48, device="cuda")
for _ in range(4):
    f(x)
torch.cuda.synchronize()
print("ok")
