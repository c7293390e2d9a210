#!/bin/bash
# full ncu capture of the four GEMM launches of tools/ncu_gemm.py + launch list of one eager config-2 step
mkdir -p gpurun_out
timeout 300 python tools/ncu_gemm.py > gpurun_out/plain4.log 2>&1 || { tail -3 gpurun_out/plain4.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 4 -f -o gpurun_out/prof_gemm4_r01d python tools/ncu_gemm.py > gpurun_out/ncu4.log 2>&1
echo "ncu gemm4 rc=$?"; tail -1 gpurun_out/ncu4.log
python tools/ncu_step.py > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
SKIP=$(grep -o "NSKIP=[0-9]*" gpurun_out/plain.log | cut -d= -f2)
CNT=$(grep -o "NCOUNT=[0-9]*" gpurun_out/plain.log | cut -d= -f2)
KRE='^(gemm_tc_kernel|attn_tc_kernel|attn_ts_kernel|layernorm_kernel|adaln_batched_kernel|gn_stats_kernel|gn_apply_kernel|linear_small_kernel|conv3x3_direct_kernel|conv3x3_small_cin_kernel|conv3x3_small_cout_kernel|cast2d_kernel|concat_inject_kernel|silu_kernel|add_kernel|cfg_ddpm_kernel|im2col3x3_s2_kernel|upsample2x_kernel|timestep_embedding_kernel|lcm_step_kernel|add_noise_kernel)$'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KRE" -s $SKIP -c $CNT --csv --log-file gpurun_out/launches_r01d.csv python tools/ncu_step.py > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_r01d.csv
