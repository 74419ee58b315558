"""The oracle is test infrastructure: nothing under legged_gym_dev_b200/ may import it (or the tests' helpers), and the
product arm of bench.py / tools/bench_configs.py may touch it only inside the CPU-baseline functions."""
import ast
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORBIDDEN = ("oracle", "legged_case", "ref_harness")


def _imports(path):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            out += [(a.name, node.lineno) for a in node.names]
        elif isinstance(node, ast.ImportFrom) and node.module:
            out.append((node.module, node.lineno))
    return out


def test_package_never_imports_oracle_or_test_helpers():
    pkg = os.path.join(ROOT, "legged_gym_dev_b200")
    for fn in sorted(os.listdir(pkg)):
        if fn.endswith(".py"):
            for mod, line in _imports(os.path.join(pkg, fn)):
                assert not any(mod == f or mod.startswith(f + ".") for f in FORBIDDEN), f"{fn}:{line} imports {mod}"


def _function_spans(path):
    tree = ast.parse(open(path).read())
    return {n.name: (n.lineno, n.end_lineno) for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}


def test_bench_touches_the_oracle_only_in_its_cpu_baseline():
    for rel, allowed in (("bench.py", {"cpu_baseline"}), (os.path.join("tools", "bench_configs.py"), {"rom_per_call", "cpu_port_baselines"})):
        path = os.path.join(ROOT, rel)
        spans = _function_spans(path)
        for mod, line in _imports(path):
            if any(mod == f or mod.startswith(f + ".") for f in FORBIDDEN):
                owner = [k for k, (a, b) in spans.items() if a <= line <= b]
                assert owner and set(owner) <= allowed, f"{rel}:{line} imports {mod} outside {allowed}"


def test_importing_the_package_does_not_load_the_oracle():
    code = ("import sys; sys.path.insert(0, %r); import legged_gym_dev_b200, legged_gym_dev_b200.legged_robot, legged_gym_dev_b200.rom, "
            "legged_gym_dev_b200.ppo, legged_gym_dev_b200.mlp, legged_gym_dev_b200.datasets, legged_gym_dev_b200.task_registry, legged_gym_dev_b200.hopper, "
            "legged_gym_dev_b200.legged_robot_trajectory; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.') or m == 'legged_case']; print(bad); "
            "sys.exit(1 if bad else 0)") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
