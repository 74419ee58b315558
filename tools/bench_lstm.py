"""Actuator-LSTM torque kernel A/B over env counts: packed FFMA2 (variant 3) vs scalar fma (variant 2); graph of 20 back-to-back launches,
us per launch.  JSON lines to stdout."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import legged_case as LC                                                         # noqa: E402
from legged_gym_dev_b200 import _lib                                             # noqa: E402

L = _lib.lib()
for N in (4096, 16384, 65536, 262144, 1048576):
    case = LC.build_case("flat_lstm_shipped", N, frames=2)
    out = dict(envs=N)
    for variant, name in ((3, "packed_us"), (2, "scalar_us")):
        L.b200gym_debug_set_lstm_variant(variant)
        env = LC.make_fused(case)
        a = case.tape.actions[0].cuda()
        env.step(a)
        env._stream = None          # launch on torch's current stream so that the capture sees it
        launch = lambda: env._compute_torques(a)
        launch()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                launch()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 20)
        out[name] = round(min(ts), 2)
        del env, g
    L.b200gym_debug_set_lstm_variant(0)
    print(json.dumps(out), flush=True)
