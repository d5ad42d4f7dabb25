#!/bin/bash
# ncu evidence for round 2 (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/)
set -x
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/r02_launches_step_2steps.csv python tools/one_step.py 2 > $O/ncu_a.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_bf16_tcgen05 -c 400 --csv --log-file $O/r02_ncu_gemm_dram_bytes_per_launch_one_step.csv python tools/one_step.py 1 > $O/ncu_b.log 2>&1
$NCU --set full --import-source on -k regex:mqa_attn --launch-skip 5 -c 3 -f -o $O/r02_ncu_attn python tools/one_step.py 1 > $O/ncu_c.log 2>&1
$NCU --set full --import-source on -k regex:"ru_tc_kernel|conv_tc_kernel|select_kernel|first_conv" --launch-skip 78 -c 26 -f -o $O/r02_ncu_codec python tools/profile_codec.py > $O/ncu_d.log 2>&1
$NCU --set full --import-source on -k regex:gemm_bf16_tcgen05 --launch-skip 30 -c 4 -f -o $O/r02_ncu_gemm python tools/one_step.py 1 > $O/ncu_e.log 2>&1
tail -2 $O/ncu_*.log
