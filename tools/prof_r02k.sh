#!/bin/bash
# round 2, late evidence (run under gpurun): launch lists + step shares of both training steps, full captures of the kernels
# that changed after r02g: BatchNorm backward passes on the cp.async row ring, the SE-gate backward carrying the BatchNorm
# reduction, the conv1.1 implicit GEMM with the lean TMA-store epilogue
TAG=${1:-r02k}
mkdir -p gpurun_out
bash tools/prof_launches.sh $TAG
cap() {  # name regex skip count
  POSE_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      --profile-from-start off -o gpurun_out/${TAG}_$1 \
      python tools/bench_train.py --model cnn --batch 128 --steps 1 --warmup 3 --cuda-profiler > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_full_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap bn_bwd_reduce 'bn_bwd_reduce_kernel<.int.2' 20 1
cap bn_bwd_apply 'bn_bwd_apply_kernel<.int.2' 20 1
cap gate_bwd_apply_bn 'gate_bwd_apply_bn_kernel' 0 1
cap conv_fwd64_lean 'gemm_bf16_tn_kernel<.int.64, .int.6, .int.64, .int.1, .int.0, .int.0, .int.2' 0 1
ls -la gpurun_out/${TAG}_*
