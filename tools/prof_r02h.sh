#!/bin/bash
# round 2 (late): full captures of the ViT's bandwidth kernels (LayerNorm backward / forward, bias-gradient column sums)
TAG=${1:-r02h}
mkdir -p gpurun_out
cap() {  # name regex skip count model batch
  POSE_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      --profile-from-start off -o gpurun_out/${TAG}_$1 \
      python tools/bench_train.py --model $5 --batch $6 --steps 1 --warmup 3 --cuda-profiler > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_full_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_$1_source.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap ln_bwd 'layernorm_bwd_kernel' 10 1 vit 64
cap ln_fwd 'layernorm_kernel' 10 1 vit 64
cap colsum 'colsum_kernel' 10 2 vit 64
ls -la gpurun_out/${TAG}_*
