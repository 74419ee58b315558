"""A/B of the launch shapes of the two step kernels at the per-GPU sizes of a sharded 1 M-env job (65 536 / 131 072 / 262 144 envs):
B200GYM_TILE (envs per CTA of post_physics) and B200GYM_PD_MAX_CTAS (grid cap of pd_torques).  Each configuration runs in its own
process (the knobs are read once); durations are averages over back-to-back launches replayed from a CUDA graph.

    python tools/tune_small_grids.py            # driver: prints one JSON line per configuration
    python tools/tune_small_grids.py --one N    # worker
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(n):
    import torch
    import bench
    env, tape = bench.build_env(n, 8, torch.device("cuda"), 0)
    acts = [tape.actions[f].cuda() for f in range(8)]
    out = bench.graph_kernel_times(env, acts)
    ab = bench.algorithmic_bytes(len(env.params.active_terms))
    peak, _ = bench.measured_peak()
    out.update(n=n, tile=os.environ.get("B200GYM_TILE", "32"), pd_cap=os.environ.get("B200GYM_PD_MAX_CTAS", "0"), pdl=os.environ.get("B200GYM_PDL", "auto"),
               pp_frac=ab["post_physics"] * n / out["post_physics_ms"] / 1e6 / peak, pd_frac=ab["pd_torques"] * n / out["pd_torques_ms"] / 1e6 / peak)
    print(json.dumps(out))


if __name__ == "__main__":
    if "--one" in sys.argv:
        one(int(sys.argv[sys.argv.index("--one") + 1]))
        sys.exit(0)
    for n in (65536, 131072, 262144):
        for tile, cap, pdl in (("32", "0", None), ("16", "0", None), ("64", "0", None), ("32", "296", None), ("32", "444", None), ("32", "592", None),
                               ("32", "888", None), ("32", "1184", None), ("32", "0", "0"), ("16", "592", "1")):
            envv = dict(os.environ, B200GYM_TILE=tile, B200GYM_PD_MAX_CTAS=cap)
            if pdl is not None:
                envv["B200GYM_PDL"] = pdl
            r = subprocess.run([sys.executable, __file__, "--one", str(n)], env=envv, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-300:], flush=True)
