#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_step.py > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
cat gpurun_out/plain.log | tail -1
SKIP=$(grep -o "NSKIP=[0-9]*" gpurun_out/plain.log | cut -d= -f2)
CNT=$(grep -o "NCOUNT=[0-9]*" gpurun_out/plain.log | cut -d= -f2)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(gemm_tc_kernel|attn_tc_kernel|attn_ts_kernel|layernorm_kernel|adaln_batched_kernel|gn_stats_kernel|gn_apply_kernel|linear_small_kernel|conv3x3_direct_kernel|conv3x3_small_cin_kernel|conv3x3_small_cout_kernel|cast2d_kernel|concat_inject_kernel|silu_kernel|add_kernel|cfg_ddpm_kernel|im2col3x3_s2_kernel|upsample2x_kernel|timestep_embedding_kernel|lcm_step_kernel|add_noise_kernel)$" -s $SKIP -c $CNT --csv --log-file gpurun_out/launches_r01b.csv python tools/ncu_step.py > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r01b.csv
