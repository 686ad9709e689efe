#!/bin/bash
# ncu full captures (run under gpurun) of the kernels reworked late in round 1; PREFIX names the batch
mkdir -p gpurun_out
cap() {  # name regex model batch skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $5 -c $6 -o gpurun_out/${PREFIX:-r01b}_$1 \
      python tools/bench_train.py --model $3 --batch $4 --steps 1 --warmup 1 > gpurun_out/ncu_$1.log 2>&1
}
cap dwconv_fwd 'dwconv3x3_kernel<.int.1' cnn 128 30 2
cap dwconv_wgrad 'dwconv_bwd_weight_kernel<.int.1' cnn 128 12 1
cap gemm_gelu 'gemm_bf16_tn_kernel<.int.128, .int.5, .int.64, .int.0, .int.0, .int.0, .int.1' vit 64 30 1
cap gemm_wide 'gemm_bf16_tn_kernel<.int.256' vit 64 60 1
cap attn_fwd 'attn_fwd_tc_kernel<.int.64' vit 64 16 1
cap attn_dkv 'attn_bwd_dkv_tc_kernel<.int.64' vit 64 6 1
cap attn_dq 'attn_bwd_dq_tc_kernel<.int.64' vit 64 6 1
# the GEMM reports carry ~25 MB of imported source each: keep their raw pages only (gpurun copies back at most 64 MiB)
for n in gemm_gelu gemm_wide; do
  ncu -i gpurun_out/${PREFIX:-r01b}_$n.ncu-rep --page raw --csv > gpurun_out/${PREFIX:-r01b}_${n}_full_raw.csv 2>/dev/null
  ncu -i gpurun_out/${PREFIX:-r01b}_$n.ncu-rep --page details --csv > gpurun_out/${PREFIX:-r01b}_${n}_details.csv 2>/dev/null
  rm -f gpurun_out/${PREFIX:-r01b}_$n.ncu-rep
done
ls -la gpurun_out/${PREFIX:-r01b}_*
