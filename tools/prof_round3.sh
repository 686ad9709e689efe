#!/bin/bash
# final batch of round 1 (run under gpurun): launch lists of both steps + full captures of the last reworked kernels
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
bash tools/prof_launches.sh
for m in cnn vit; do mv gpurun_out/r01b_${m}_train_launches.csv gpurun_out/r01c_${m}_train_launches.csv; done
cap() {  # name regex model batch skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $5 -c $6 -o gpurun_out/r01c_$1 \
      python tools/bench_train.py --model $3 --batch $4 --steps 1 --warmup 1 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/r01c_$1.ncu-rep --page raw --csv > gpurun_out/r01c_$1_full_raw.csv 2>/dev/null
  ncu -i gpurun_out/r01c_$1.ncu-rep --page details --csv > gpurun_out/r01c_$1_details.csv 2>/dev/null
}
cap dwconv_fwd 'dwconv3x3_kernel<.int.1' cnn 128 30 1
cap attn_fwd 'attn_fwd_tc_kernel<.int.64' vit 64 16 1
cap gemm_gelu_wide 'gemm_bf16_tn_kernel<.int.256, .int.3, .int.64, .int.0, .int.0, .int.0, .int.1' vit 64 20 1
rm -f gpurun_out/r01c_gemm_gelu_wide.ncu-rep gpurun_out/r01c_attn_fwd.ncu-rep
ls -la gpurun_out/r01c_*
