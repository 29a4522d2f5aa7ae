#!/bin/bash
# Runs on the GPU box (under gpurun): one `ncu --set full` capture per hot kernel at the benchmark's size, exported to CSV on the box
# (raw metrics page + source page with per-instruction stall samples); the .ncu-rep files are deleted so gpurun_out stays small.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
cap() {  # name, kernel regex, skip, count, target mode
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c "$4" -f -o $O/$1 python tools/ncu_targets.py "$5" > /dev/null 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | head -c 6000000 > $O/$1_source.csv
  rm -f $O/$1.ncu-rep
}
cap r2_ncu_msm_accumulate_g1 msm_accumulate_kernel 1 1 msm_g1
cap r2_ncu_msm_accumulate_g2 msm_accumulate_kernel 1 1 msm_g2
cap r2_ncu_msm_marginals_multi msm_marginals 2 2 msm_multi
cap r2_ncu_ntt_pass ntt_pass 3 3 ntt
cap r2_ncu_plonk_quotient_l1 plonk_quotient_l1 0 1 plonk
cap r2_ncu_plonk_quotient_l2 plonk_quotient_l2 0 1 plonk
ls -la $O | tail -20
