#!/bin/bash
# Round profile pass (run under gpurun, one GPU): launch list of the bench command + one `ncu --set full` capture per kernel.
# Outputs go to gpurun_out/; tools/ncu_table.py turns them into profiles/<round>_kernels_ncu.md.
set -u
R=${1:-r1}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 16 --warmup 8 --no-sweep --no-extra --no-e2e --no-cpu-baseline"
$B > $O/${R}_bench_plain.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_bench.csv $B > /dev/null 2>&1
K="python tools/run_kernels_once.py"
$K > $O/${R}_kernels_plain.log 2>&1 || exit 1
cap() {  # name regex [skip]
  timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s ${3:-1} -c 1 -f -o $O/prof_${R}_$1 $K > $O/ncu_${R}_$1.log 2>&1
}
cap post_physics_rough 'post_physics_kernel<\(int\)32, \(bool\)1, \(bool\)0>' 2
cap post_physics_traj 'post_physics_kernel<\(int\)32, \(bool\)0, \(bool\)1>' 2
cap rom_reset_root 'rom_reset_root_kernel' 2
cap lstm_torques 'lstm_torques_kernel' 4
cap rom_rollout 'rom_rollout_kernel' 1
cap rom_step 'rom_step_kernel' 5
cap gae_returns 'gae_returns_kernel' 2
cap gather_rows 'gather_rows_kernel' 1
cap ppo_loss 'ppo_loss_kernel' 1
cap clip_adam 'clip_adam' 1
cap sliding_window 'sliding_window_kernel' 1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:post_physics_kernel" -s 4 -c 1 -f -o $O/prof_${R}_post_physics_flat $B > $O/ncu_${R}_pp_flat.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:pd_torques_kernel" -s 8 -c 1 -f -o $O/prof_${R}_pd_torques $B > $O/ncu_${R}_pd.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:mlp_forward_h4" -s 2 -c 1 -f -o $O/prof_${R}_mlp_forward python tools/run_mlp_once.py 1048576 > $O/ncu_${R}_mlp.log 2>&1
# the generic rom-family kernels and the Hopper kernels (own drivers, 1 M envs)
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:fam_step_tile" -c 6 -f -o $O/prof_${R}_romfam_step python tools/run_romfam_once.py > $O/ncu_${R}_romfam_step.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:fam_reset" -c 1 -f -o $O/prof_${R}_romfam_reset python tools/run_romfam_once.py > $O/ncu_${R}_romfam_reset.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:hopper" -c 6 -f -o $O/prof_${R}_hopper python tools/run_hopper_once.py > $O/ncu_${R}_hopper.log 2>&1
ls -la $O | tail -20
