#!/bin/bash
# ncu evidence for the training steps (run under gpurun): launch lists of one step + full captures of the top kernels.
set -e
mkdir -p gpurun_out
for m in vit cnn; do
  b=64; [ $m = cnn ] && b=128
  python tools/bench_train.py --model $m --batch $b --steps 1 --warmup 2 > gpurun_out/plain_$m.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file gpurun_out/r01_${m}_train_launches.csv \
      python tools/bench_train.py --model $m --batch $b --steps 1 --warmup 2 > gpurun_out/ncu_$m.log 2>&1
done
for k in attn_fwd_tc_kernel attn_bwd_dkv_tc_kernel attn_bwd_dq_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -o gpurun_out/r01_$k \
      python tools/bench_train.py --model vit --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 40 -c 3 -o gpurun_out/r01_gemm_vit \
    python tools/bench_train.py --model vit --batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
for k in bn_bwd_reduce_kernel dwconv_bwd_weight_kernel bn_apply_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r01_$k \
      python tools/bench_train.py --model cnn --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
