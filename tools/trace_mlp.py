"""Per-phase cycle budget of the pipelined MLP forward kernel from its built-in event trace (CTA 0)."""
import ctypes as C
import collections
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from bench_mlp import net
from legged_gym_dev_b200.mlp import FusedMLP
from legged_gym_dev_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
m = net(48, (128, 64, 32), 12)
f = FusedMLP(m)
x = torch.randn(B, 48, device="cuda")
for _ in range(3):
    f(x)
L = _lib.lib()
L.b200gym_debug_mlp_trace.argtypes = [C.c_void_p]
buf = torch.zeros(5 * 4096 * 2, dtype=torch.int64, device="cuda")
assert L.b200gym_debug_mlp_trace(buf.data_ptr()) == 0
f(x)
torch.cuda.synchronize()
L.b200gym_debug_mlp_trace(None)
tr = buf.cpu().view(5, 4096, 2)
names = {0: "pre-barrier", 1: "bar.sync", 2: "full wait", 3: "mma issue", 4: "commit", 5: "mma wait", 6: "epilogue"}
for region, label in ((0, "slot0 epilogue (l*10 + 0 start, 1 mma done, 2 epilogue done)"), (1, "slot1 epilogue"), (2, "slot0 loader"), (3, "slot1 loader"), (4, "mma warp (l*10 + slot: operands ready; +5: issued)")):
    ev = [(int(c), int(t)) for c, t in tr[region].tolist() if t != 0]
    if not ev:
        continue
    t0 = ev[0][1]
    print(f"== {label}: {len(ev)} events, span {ev[-1][1] - t0} cycles")
    agg = collections.defaultdict(list)
    for (c0, a), (c1, b) in zip(ev[:-1], ev[1:]):
        agg[(c0, c1)].append(b - a)
    for k, v in sorted(agg.items()):
        v2 = v[2:] if len(v) > 4 else v   # skip the cold first tiles
        print(f"   {k[0]:4d} -> {k[1]:4d}  n={len(v):3d}  mean {sum(v2)/len(v2):8.0f}  min {min(v2):6d}  max {max(v2):6d}")
    if region < 2 or region == 4:
        print("   first 40 events (code, dt):", [(c, t - t0) for c, t in ev[:40]])

# merged absolute timeline of one steady-state period (between the 5th and 7th tile start of slot 0)
allev = []
labels = ["E0", "E1", "L0", "L1", "MMA"]
for region in range(5):
    for c, t in tr[region].tolist():
        if t != 0:
            allev.append((int(t), labels[region], int(c)))
allev.sort()
starts = [t for t, lab, c in allev if lab == "E0" and c == 0]
if len(starts) > 7:
    lo, hi = starts[5], starts[7]
    print("== merged timeline (cycles since window start)")
    for t, lab, c in allev:
        if lo <= t <= hi:
            print(f"   {t - lo:7d}  {lab:4s} {c}")
