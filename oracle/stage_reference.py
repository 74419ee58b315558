"""Recipe that stages the reference's own Python sources for the CPU arm of bench.py.  TEST / MEASUREMENT INFRASTRUCTURE.

The reference (wdc3iii/legged_gym_dev) is 100 % Python: there is nothing to compile into oracle/_ref/.  What the GPU box lacks is
the tree itself (/root/reference exists only in the build container), so `__graft_entry__.build()` runs this script where the
tree is present: it copies the package sources the hot path imports (legged_gym/, trajopt/, deep_tube_learning/ — *.py only) and
the actuator-net TorchScript file into oracle/_ref/, UNMODIFIED, byte for byte.  oracle/_ref/ is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it travels to the GPU box like the built .so files; there oracle/ref_harness.py
imports the reference from it and `bench.py --impl reference` / the `cpu_baseline` leg time the reference's OWN
LeggedRobot.step (kind "reference").  Nothing under legged_gym_dev_b200/ reads it.

    python oracle/stage_reference.py [SRC=/root/reference]
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("legged_gym", "trajopt", "deep_tube_learning")
EXTRA = ("resources/actuator_nets/anydrive_v3_lstm.pt",)


def stage(src="/root/reference", dst=DST):
    if not os.path.isdir(os.path.join(src, "legged_gym")):
        return None
    n, h = 0, hashlib.sha256()
    for pkg in PACKAGES:
        for root, _dirs, files in os.walk(os.path.join(src, pkg)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                a = os.path.join(root, f)
                b = os.path.join(dst, os.path.relpath(a, src))
                os.makedirs(os.path.dirname(b), exist_ok=True)
                shutil.copyfile(a, b)
                with open(a, "rb") as fh:
                    h.update(fh.read())
                n += 1
    for rel in EXTRA:
        a, b = os.path.join(src, rel), os.path.join(dst, rel)
        if os.path.exists(a):
            os.makedirs(os.path.dirname(b), exist_ok=True)
            shutil.copyfile(a, b)
            n += 1
    with open(os.path.join(dst, "STAGED.txt"), "w") as fh:
        fh.write(f"staged from {src}: {n} files, sha256 of the .py sources {h.hexdigest()}\n")
    return n


if __name__ == "__main__":
    k = stage(*(sys.argv[1:2]))
    print("reference tree not present: nothing staged" if k is None else f"staged {k} reference files into {DST}")
