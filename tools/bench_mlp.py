"""Correctness + timing of the fused tcgen05 MLP forward (csrc/mlp.cu) against torch (cuBLAS) — one JSON line.
B200GYM_MLP_SERIAL=1 selects the serial (round-1 first version) kernel for an A/B in a second process."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn


def net(i, hs, o):
    layers, d = [], i
    for h in hs:
        layers += [nn.Linear(d, h), nn.ELU()]
        d = h
    return nn.Sequential(*layers, nn.Linear(d, o)).cuda()


def main():
    from legged_gym_dev_b200.mlp import FusedMLP
    torch.manual_seed(0)
    res = {"serial": os.environ.get("B200GYM_MLP_SERIAL", "0")}
    for name, (i, hs, o) in {"actor_flat": (48, (128, 64, 32), 12), "critic_flat": (48, (128, 64, 32), 1)}.items():
        m = net(i, hs, o)
        f = FusedMLP(m)
        for B in (7, 128, 4096, 65536, 393216, 1048576):
            x = torch.randn(B, i, device="cuda")
            with torch.no_grad():
                want = m(x)
            got = f(x)
            torch.cuda.synchronize()
            err = (got - want).abs().max().item()
            r = {"max_abs_err": err}
            if B >= 4096:
                for _ in range(5):
                    f(x)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(50):
                    f(x)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 50
                with torch.no_grad():
                    for _ in range(3):
                        m(x)
                    a.record()
                    for _ in range(20):
                        m(x)
                    b.record()
                    torch.cuda.synchronize()
                tms = a.elapsed_time(b) / 20
                flops = 2.0 * B * (48 * 128 + 128 * 64 + 64 * 32 + 32 * o)
                byt = B * (i + o) * 4
                r.update(fused_ms=ms, torch_ms=tms, tflops=flops / ms / 1e9, gbs=byt / ms / 1e6, frac_hbm=byt / ms / 1e6 / 6535.7)
            res[f"{name}_{B}"] = r
    print(json.dumps(res))


if __name__ == "__main__":
    main()
