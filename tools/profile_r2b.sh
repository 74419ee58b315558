#!/bin/bash
# Round-2 second profile pass (after the one-launch minibatch): launch lists of the bench step and of one PPO update, and one
# `ncu --set full` capture each of the fused chain kernel, the optimiser kernel, the paired policy forward and act_store.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 16 --warmup 8 --no-sweep --no-extra --no-e2e --no-cpu-baseline --no-multi"
$B > $O/r2b_bench_plain.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2b_launches_bench.csv $B > /dev/null 2>&1
U="python tools/bench_update.py --flat-only --no-profile --once"
$U > $O/r2b_update_plain.json 2> /dev/null || exit 1
B200GYM_PPO_GRAPH=0 B200GYM_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2b_launches_update.csv $U > /dev/null 2>&1
cap() {  # name regex driver [skip] [count]
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s ${4:-1} -c ${5:-1} -f -o $O/prof_r2b_$1 $3 > $O/ncu_r2b_$1.log 2>&1
}
export B200GYM_PPO_GRAPH=0 B200GYM_PDL=0
cap ppo_chain 'ppo_chain_kernel' "$U" 3
cap ppo_optimizer 'ppo_optimizer_step_kernel' "$U" 3
unset B200GYM_PPO_GRAPH
cap pair_forward 'mlp_forward_h4_pair_kernel' "python tools/run_rollout_once.py 4096" 4
cap act_store 'ppo_act_store_kernel' "python tools/run_rollout_once.py 4096" 4
cap post_physics_131072 'post_physics_kernel' "python bench.py --envs 131072 --frames 4 --steps 8 --warmup 4 --no-e2e --no-cpu-baseline --no-sweep --no-extra --no-multi --no-graph" 6
ls -la $O | grep r2b_ | tail -30
