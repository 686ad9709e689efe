#!/bin/bash
# round 2 evidence (run under gpurun): launch lists + step shares of both training steps, full captures of the kernels that
# changed this round (lean-epilogue 1x1 GEMM with BatchNorm statistics, conv1.1 implicit GEMM after the producer fix,
# conv1.0 weight gradient on 32-channel pixels)
TAG=${1:-r02f}
mkdir -p gpurun_out
bash tools/prof_launches.sh $TAG
cap() {  # name regex skip count
  POSE_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      --profile-from-start off -o gpurun_out/${TAG}_$1 \
      python tools/bench_train.py --model cnn --batch 128 --steps 1 --warmup 3 --cuda-profiler > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_full_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap gemm_lean_stats 'gemm_bf16_tn_kernel<.int.128, .int.5, .int.64, .int.0, .int.0, .int.0, .int.2' 3 1
cap conv_fwd64 'gemm_bf16_tn_kernel<.int.64, .int.6, .int.64, .int.1' 0 1
cap conv_wgrad 'gemm_bf16_tn_kernel<.int.128, .int.4, .int.64, .int.2' 0 2
cap bn_bwd_apply 'bn_bwd_apply_kernel<.int.2' 20 1
ls -la gpurun_out/${TAG}_*
