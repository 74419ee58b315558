#!/bin/bash
# A/B of kernel variants in ONE gpurun call (same box, same clocks): tools/ab_bench.sh <envs> lib1.so lib2.so ...
E=$1; shift
for L in "$@"; do
  B200GYM_LIB=$PWD/$L python bench.py --envs $E --frames 4 --steps 100 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$L', d['config']['envs_per_gpu'], 'step_ms=%.4f'%d['ms_per_step'], 'pp_ms=%.4f'%r['avg_launch_ms'], 'frac=%.3f'%r['frac'], 'value=%.3e'%d['value'])"
done
