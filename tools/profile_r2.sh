#!/bin/bash
# Round-2 profile pass (run under gpurun, one GPU): launch lists + one `ncu --set full` capture per new kernel.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 16 --warmup 8 --no-sweep --no-extra --no-e2e --no-cpu-baseline --no-multi"
$B > $O/r2_bench_plain.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv $B > /dev/null 2>&1
U="python tools/bench_update.py --flat-only --no-profile --once"
$U > $O/r2_update_plain.json 2> /dev/null || exit 1
B200GYM_PPO_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_update.csv $U > /dev/null 2>&1
cap() {  # name regex driver [skip] [count]
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s ${4:-1} -c ${5:-1} -f -o $O/prof_r2_$1 $3 > $O/ncu_r2_$1.log 2>&1
}
export B200GYM_PPO_GRAPH=0
cap ppo_chain 'ppo_chain_kernel' "$U" 3
cap gemm_wgrad 'gemm_f16_kernel' "$U" 3
cap ppo_optimizer 'ppo_optimizer_step_kernel' "$U" 3
cap gemm_rough 'gemm_f16_kernel' "python tools/bench_update.py --rough-only --no-profile --once" 20 3
unset B200GYM_PPO_GRAPH
cap act_store 'ppo_act_store_kernel' "python tools/run_rollout_once.py 65536" 4
cap store_step 'ppo_store_step_kernel' "python tools/run_rollout_once.py 65536" 4
cap reset_idx 'reset_idx_kernel' "python tools/run_rollout_once.py 1048576" 0
cap hopper_env 'hopper_post_physics_kernel' "python tools/run_hopper_env_once.py" 2
cap hopper_prologue 'hopper_prologue_kernel' "python tools/run_hopper_env_once.py" 2
ls -la $O | grep r2_ | tail -30
